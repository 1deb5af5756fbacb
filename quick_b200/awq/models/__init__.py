from .llama import LlamaAWQForCausalLM  # noqa: F401
from .mistral import MistralAWQForCausalLM  # noqa: F401
from .qwen2 import Qwen2AWQForCausalLM  # noqa: F401
