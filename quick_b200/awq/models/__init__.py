from .llama import LlamaAWQForCausalLM  # noqa: F401
from .mistral import MistralAWQForCausalLM  # noqa: F401
