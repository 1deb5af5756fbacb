"""``from quick_b200.awq import AutoAWQForCausalLM`` — the counterpart of ``from quick.awq import AutoAWQForCausalLM``
(reference quick/awq/__init__.py:2)."""
from .models.auto import AutoAWQForCausalLM  # noqa: F401
