"""Module-tree helpers under the names the reference's adapters and quantizer use (quick/awq/utils/module.py:3-53),
implemented on ``nn.Module.get_submodule`` / ``named_modules``."""
import torch.nn as nn


def get_named_linears(module: nn.Module) -> dict:
    """{relative name: nn.Linear} for every plain linear below ``module`` (packed linears are not nn.Linear)."""
    found = {}
    for name, sub in module.named_modules():
        if isinstance(sub, nn.Linear):
            found[name] = sub
    return found


def get_op_by_name(module: nn.Module, op_name: str) -> nn.Module:
    try:
        return module.get_submodule(op_name)
    except AttributeError:
        raise ValueError(f"Cannot find op {op_name} in module {module}") from None


def set_op_by_name(layer: nn.Module, name: str, new_module: nn.Module) -> None:
    """Replace the submodule ``name`` (dotted path, numeric parts index containers) by ``new_module``."""
    parent_name, _, leaf = name.rpartition(".")
    owner = layer.get_submodule(parent_name) if parent_name else layer
    setattr(owner, leaf, new_module)


def get_op_name(module: nn.Module, op: nn.Module) -> str:
    """Dotted path of ``op`` below ``module`` (identity match)."""
    match = next((name for name, sub in module.named_modules() if sub is op), None)
    if match is None:
        raise ValueError(f"Cannot find op {op} in module {module}")
    return match


def append_str_prefix(x, prefix: str):
    """Prefix every string inside (nested) tuples / lists; other leaves pass through."""
    if isinstance(x, str):
        return prefix + x
    if isinstance(x, (tuple, list)):
        return type(x)(append_str_prefix(item, prefix) for item in x)
    return x


def exclude_layers_to_not_quantize(linear_layers: dict, modules_to_not_convert) -> dict:
    """Drop every linear whose name contains one of the ``modules_to_not_convert`` keys."""
    skip = tuple(modules_to_not_convert or ())
    return {name: lin for name, lin in linear_layers.items() if not any(key in name for key in skip)}
