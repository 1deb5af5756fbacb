"""Qwen2 adapter: the Llama decoder layout with biased q / k / v projections — exercises the bias path of the fused
q‖k‖v module (the reference's fuse_qkv_quick concatenates biases the same way, fused_utils.py:97-117; its own `qwen`
adapter targets the older remote-code Qwen, quick/awq/models/qwen.py)."""
from .llama import LlamaAWQForCausalLM


class Qwen2AWQForCausalLM(LlamaAWQForCausalLM):
    layer_type = "Qwen2DecoderLayer"
    max_new_tokens_key = "max_position_embeddings"
