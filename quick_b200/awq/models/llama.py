"""Llama adapter — the hooks the quantizer and the loader need (reference quick/awq/models/llama.py:13-76) and the
fuser that swaps the HF decoder for the library's fused runner (reference LlamaFuser, llama.py:79-126, which builds
LlamaLikeModel out of QuantAttentionFused / fused MLP blocks)."""
from .base import BaseAWQForCausalLM


class LlamaAWQForCausalLM(BaseAWQForCausalLM):
    layer_type = "LlamaDecoderLayer"
    max_new_tokens_key = "max_position_embeddings"

    @staticmethod
    def fuse_layers(model, batch_size: int = 1):
        from .llama_like import fuse_hf_model
        fuse_hf_model(model, batch_size=batch_size)

    @staticmethod
    def get_model_layers(model):
        return model.model.layers

    @staticmethod
    def get_act_for_scaling(module):
        return dict(is_scalable=False)

    @staticmethod
    def move_embed(model, device):
        model.model.embed_tokens = model.model.embed_tokens.to(device)
        if hasattr(model.model, "rotary_emb"):
            model.model.rotary_emb = model.model.rotary_emb.to(device)

    @staticmethod
    def get_layers_for_scaling(module, input_feat, module_kwargs):
        attn, mlp = module.self_attn, module.mlp
        groups = [dict(prev_op=module.input_layernorm, layers=[attn.q_proj, attn.k_proj, attn.v_proj],
                       inp=input_feat["self_attn.q_proj"], module2inspect=attn, kwargs=module_kwargs)]
        # v_proj -> o_proj only when the shapes agree: with grouped-query attention v is narrower than o's input
        # (reference llama.py:50-57)
        if attn.v_proj.weight.shape == attn.o_proj.weight.shape:
            groups.append(dict(prev_op=attn.v_proj, layers=[attn.o_proj], inp=input_feat["self_attn.o_proj"]))
        groups.append(dict(prev_op=module.post_attention_layernorm, layers=[mlp.gate_proj, mlp.up_proj],
                           inp=input_feat["mlp.gate_proj"], module2inspect=mlp))
        groups.append(dict(prev_op=mlp.up_proj, layers=[mlp.down_proj], inp=input_feat["mlp.down_proj"]))
        return groups
