"""AutoAWQForCausalLM — same entry points and argument names as the reference (quick/awq/models/auto.py:27-99).
Families: the ones BASELINE.json's configs name (Llama-2, Mistral) plus Qwen2, which shares the Llama decoder layout
(with biased q / k / v); the
reference's other adapters (mpt, opt, falcon, bloom, gptj, …) wrap model code that is outside the W4A16 hot path."""
import os

from .base import BaseAWQForCausalLM
from .llama import LlamaAWQForCausalLM
from .mistral import MistralAWQForCausalLM
from .qwen2 import Qwen2AWQForCausalLM

AWQ_CAUSAL_LM_MODEL_MAP = {
    "llama": LlamaAWQForCausalLM,
    "mistral": MistralAWQForCausalLM,
    "qwen2": Qwen2AWQForCausalLM,
}


def check_and_get_model_type(model_dir, trust_remote_code=True, **model_init_kwargs):
    from transformers import AutoConfig
    if not os.path.isdir(model_dir):
        raise FileNotFoundError(f"{model_dir} is not a local directory (no hub access; download the checkpoint first)")
    config = AutoConfig.from_pretrained(model_dir, trust_remote_code=trust_remote_code, **model_init_kwargs)
    if config.model_type not in AWQ_CAUSAL_LM_MODEL_MAP:
        raise TypeError(f"{config.model_type} isn't supported yet.")
    return config.model_type


class AutoAWQForCausalLM:
    def __init__(self):
        raise EnvironmentError("You must instantiate AutoAWQForCausalLM with\n"
                               "AutoAWQForCausalLM.from_quantized or AutoAWQForCausalLM.from_pretrained")

    @classmethod
    def from_pretrained(cls, model_path, trust_remote_code=True, safetensors=False, device_map=None,
                        **model_init_kwargs) -> BaseAWQForCausalLM:
        model_type = check_and_get_model_type(model_path, trust_remote_code, **model_init_kwargs)
        return AWQ_CAUSAL_LM_MODEL_MAP[model_type].from_pretrained(
            model_path, model_type, trust_remote_code=trust_remote_code, safetensors=safetensors, device_map=device_map,
            **model_init_kwargs)

    @classmethod
    def from_quantized(cls, quant_path, quant_filename="", max_new_tokens=None, trust_remote_code=True, use_exllama=False,
                       use_exllama_v2=False, batch_size=1, safetensors=True, device_map="balanced", offload_folder=None,
                       use_quick=False, fuse_layers=True, **config_kwargs) -> BaseAWQForCausalLM:
        """``fuse_layers`` is an addition (the reference always fuses, auto.py:91); ``use_quick`` is accepted and
        ignored — every module is a QUICK module here."""
        os.environ["AWQ_BATCH_SIZE"] = str(batch_size)
        model_type = check_and_get_model_type(quant_path, trust_remote_code)
        return AWQ_CAUSAL_LM_MODEL_MAP[model_type].from_quantized(
            quant_path, model_type, quant_filename, max_new_tokens, trust_remote_code=trust_remote_code,
            fuse_layers=fuse_layers, use_exllama=use_exllama, use_exllama_v2=use_exllama_v2, safetensors=safetensors,
            device_map=device_map, offload_folder=offload_folder, batch_size=batch_size, **config_kwargs)
