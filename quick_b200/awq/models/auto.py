"""``AutoAWQForCausalLM`` — the model-level entry point with the reference's call signatures
(quick/awq/models/auto.py:27-99: ``from_pretrained(model_path, …)``, ``from_quantized(quant_path, quant_filename, …)``).

Family registry: Llama and Mistral (the configs BASELINE.json names) plus Qwen2, which shares the Llama decoder layout
with biased q / k / v.  The reference's other adapters (mpt, opt, falcon, bloom, gptj, …) wrap model code outside the
W4A16 hot path and are not rebuilt; an unknown ``config.model_type`` raises the same ``TypeError``.
"""
import os

from .base import BaseAWQForCausalLM
from .llama import LlamaAWQForCausalLM
from .mistral import MistralAWQForCausalLM
from .qwen2 import Qwen2AWQForCausalLM

AWQ_CAUSAL_LM_MODEL_MAP = dict(llama=LlamaAWQForCausalLM, mistral=MistralAWQForCausalLM, qwen2=Qwen2AWQForCausalLM)


def check_and_get_model_type(model_dir, trust_remote_code=True, **model_init_kwargs):
    """model_type of a LOCAL checkpoint directory, validated against the registry."""
    if not os.path.isdir(model_dir):
        raise FileNotFoundError(f"{model_dir} is not a local directory (no hub access; download the checkpoint first)")
    from transformers import AutoConfig
    model_type = AutoConfig.from_pretrained(model_dir, trust_remote_code=trust_remote_code, **model_init_kwargs).model_type
    if model_type not in AWQ_CAUSAL_LM_MODEL_MAP:
        raise TypeError(f"{model_type} isn't supported yet.")
    return model_type


def _adapter_for(path, trust_remote_code, **kw):
    model_type = check_and_get_model_type(path, trust_remote_code, **kw)
    return model_type, AWQ_CAUSAL_LM_MODEL_MAP[model_type]


class AutoAWQForCausalLM:
    """Not instantiable: use the two constructors."""

    def __init__(self):
        raise EnvironmentError("You must instantiate AutoAWQForCausalLM with\n"
                               "AutoAWQForCausalLM.from_quantized or AutoAWQForCausalLM.from_pretrained")

    @classmethod
    def from_pretrained(cls, model_path, trust_remote_code=True, safetensors=False, device_map=None,
                        **model_init_kwargs) -> BaseAWQForCausalLM:
        """The fp16 model that ``quantize`` will turn into an AWQ-QUICK model."""
        model_type, adapter = _adapter_for(model_path, trust_remote_code, **model_init_kwargs)
        options = dict(model_init_kwargs, trust_remote_code=trust_remote_code, safetensors=safetensors, device_map=device_map)
        return adapter.from_pretrained(model_path, model_type, **options)

    @classmethod
    def from_quantized(cls, quant_path, quant_filename="", max_new_tokens=None, trust_remote_code=True, use_exllama=False,
                       use_exllama_v2=False, batch_size=1, safetensors=True, device_map="balanced", offload_folder=None,
                       use_quick=False, fuse_layers=True, **config_kwargs) -> BaseAWQForCausalLM:
        """A saved AWQ checkpoint (QUICK layout, or AWQ-GEMM layout converted on the fly).  ``batch_size`` sizes the static
        KV cache of the fused model (the reference passes it through AWQ_BATCH_SIZE, auto.py:83, kept for code that reads
        it); ``fuse_layers`` is an addition (the reference always fuses, auto.py:91); ``use_quick`` is accepted and
        ignored — every quantized linear is a QUICK module here."""
        os.environ["AWQ_BATCH_SIZE"] = str(batch_size)
        model_type, adapter = _adapter_for(quant_path, trust_remote_code)
        options = dict(config_kwargs, trust_remote_code=trust_remote_code, fuse_layers=fuse_layers, use_exllama=use_exllama,
                       use_exllama_v2=use_exllama_v2, safetensors=safetensors, device_map=device_map,
                       offload_folder=offload_folder, batch_size=batch_size)
        return adapter.from_quantized(quant_path, model_type, quant_filename, max_new_tokens, **options)
