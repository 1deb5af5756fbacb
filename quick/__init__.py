"""``quick`` — the reference's import path, served by quick_b200.

The reference's scripts and downstream users write ``from quick.awq import AutoAWQForCausalLM``
(/root/reference/examples/benchmark.py:6-7, quick/awq/__init__.py:2) and reach into
``quick.awq.modules.linear.quick.WQLinear_QUICK``, ``quick.awq.models.base.BaseAWQForCausalLM``,
``quick.awq.utils.fused_utils`` ...  This package owns no code: every ``quick.awq[.x.y]`` import is resolved to
the module object of ``quick_b200.awq[.x.y]`` (one object under two names, so ``isinstance`` checks and module
state agree whichever path a caller used).
"""
import importlib
import importlib.abc
import importlib.util
import sys

_ALIAS = {"quick.awq": "quick_b200.awq"}


def _target(fullname):
    for src, dst in _ALIAS.items():
        if fullname == src or fullname.startswith(src + "."):
            return dst + fullname[len(src):]
    return None


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)      # the real module object, shared by both names

    def exec_module(self, module):
        pass                                             # already executed under its own name


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        real = _target(fullname)
        if real is None:
            return None
        try:
            real_spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            real_spec = None
        if real_spec is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real), is_package=real_spec.submodule_search_locations is not None)
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

from quick_b200 import __version__  # noqa: E402,F401
