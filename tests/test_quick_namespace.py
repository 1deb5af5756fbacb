"""The reference's import path: ``from quick.awq import AutoAWQForCausalLM`` (reference examples/benchmark.py:6-7,
quick/awq/__init__.py:2) must resolve to this framework, one module object under both names."""
import ast
import importlib
import os

import pytest

REF_EXAMPLES = "/root/reference/examples"
# what the reference's scripts import from the hot path's plugin surface (checked against the real files below when
# the reference tree is present; evaluation harnesses — lm-eval / HumanEval / KL-div — are out of scope)
SURFACE = [
    ("quick.awq", "AutoAWQForCausalLM"),
    ("quick.awq.models.base", "BaseAWQForCausalLM"),
    ("quick.awq.models.auto", "AutoAWQForCausalLM"),
    ("quick.awq.modules.linear.quick", "WQLinear_QUICK"),
    ("quick.awq.utils.fused_utils", "fuse_qkv_quick"),
    ("quick.awq.utils.fused_utils", "QUICK_cat"),
    ("quick.awq.quantize.quantizer", "AwqQuantizer"),
    ("quick.awq.models.llama", "LlamaAWQForCausalLM"),
    ("quick.awq.models.mistral", "MistralAWQForCausalLM"),
]
OUT_OF_SCOPE = {"quick.awq.evaluation"}


def test_quick_namespace_is_quick_b200(built):
    for mod, name in SURFACE:
        m = importlib.import_module(mod)
        real = importlib.import_module(mod.replace("quick.", "quick_b200.", 1))
        assert m is real, f"{mod} is not the quick_b200 module object"
        assert hasattr(m, name), f"{mod}.{name} missing"
    import quick
    import quick_b200
    assert quick.__version__ == quick_b200.__version__
    with pytest.raises(ImportError):
        importlib.import_module("quick.awq.no_such_module")


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference tree not present (GPU box)")
def test_every_import_of_the_reference_scripts_resolves(built):
    """Parse the reference's own example scripts and import what they import from ``quick``."""
    seen = 0
    for fn in sorted(os.listdir(REF_EXAMPLES)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(REF_EXAMPLES, fn)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] == "quick":
                if node.module in OUT_OF_SCOPE:
                    continue
                m = importlib.import_module(node.module)
                for alias in node.names:
                    assert hasattr(m, alias.name), f"{fn}: from {node.module} import {alias.name}"
                    seen += 1
    assert seen >= 3
