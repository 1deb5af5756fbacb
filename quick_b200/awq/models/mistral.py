"""Mistral adapter (reference quick/awq/models/mistral.py:13-129): the same decoder structure as Llama with
grouped-query attention — which the reference's QUICK fuser cannot fuse (QUICK_cat rejects unequal widths,
fused_utils.py:139-142) and this library can.  Sliding-window attention is not applied by the fused runner; it only
matters beyond ``sliding_window`` tokens of context, and the fuser refuses a cache longer than the window."""
from .llama import LlamaAWQForCausalLM


class MistralAWQForCausalLM(LlamaAWQForCausalLM):
    layer_type = "MistralDecoderLayer"
    max_new_tokens_key = "max_position_embeddings"
