"""Module-tree helpers with the names the reference uses (quick/awq/utils/module.py:3-53) so that model adapters
written for the reference read the same here."""
import torch.nn as nn


def get_named_linears(module):
    return {name: m for name, m in module.named_modules() if isinstance(m, nn.Linear)}


def get_op_by_name(module, op_name):
    try:
        return module.get_submodule(op_name)
    except AttributeError:
        raise ValueError(f"Cannot find op {op_name} in module {module}") from None


def set_op_by_name(layer, name, new_module):
    parent_name, _, leaf = name.rpartition(".")
    parent = layer.get_submodule(parent_name) if parent_name else layer
    setattr(parent, leaf, new_module)


def get_op_name(module, op):
    for name, m in module.named_modules():
        if m is op:
            return name
    raise ValueError(f"Cannot find op {op} in module {module}")


def append_str_prefix(x, prefix):
    if isinstance(x, str):
        return prefix + x
    if isinstance(x, (tuple, list)):
        return type(x)(append_str_prefix(y, prefix) for y in x)
    return x


def exclude_layers_to_not_quantize(linear_layers, modules_to_not_convert):
    if not modules_to_not_convert:
        return linear_layers
    return {n: m for n, m in linear_layers.items() if not any(key in n for key in modules_to_not_convert)}
