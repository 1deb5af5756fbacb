"""QUICK-layout concatenation helpers — mirror of the reference's
quick/awq/utils/fused_utils.py:97-159 (``fuse_qkv_quick``, ``QUICK_cat``), generalised to unequal
widths so grouped-query models (k/v narrower than q) can be fused; the reference raises for those
(fused_utils.py:139-142) although the view algebra is exact."""
from typing import Optional, Tuple

import torch

from ..modules.linear.quick import WQLinear_QUICK


def QUICK_cat(*input_layers: torch.Tensor, options: str, reshape_dims: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    if len(input_layers) < 2:
        raise ValueError("At least two input layers are required")
    H = input_layers[0].shape[0]
    for layer in input_layers[1:]:
        if layer.shape[0] != H:
            raise ValueError("All input layers must have the same number of rows")
    if not reshape_dims:
        rows = {"qweight": H // 2, "qzeros": H * 4, "scales": H * 4}.get(options)
        if rows is None:
            raise ValueError("Unknown options provided or invalid reshape dimensions")
        layers_to_cat = [layer.reshape(rows, -1) for layer in input_layers]
    else:
        layers_to_cat = [layer.reshape(*reshape_dims) for layer in input_layers]
    return torch.cat(layers_to_cat, dim=1).reshape(H, -1)


def fuse_quick_linears(*mods: WQLinear_QUICK) -> WQLinear_QUICK:
    """One WQLinear_QUICK computing the N-concatenation of the given ones (same K and group size, any widths):
    q‖k‖v, gate‖up."""
    first = mods[0]
    if any((m.in_features, m.group_size, m.w_bit) != (first.in_features, first.group_size, first.w_bit) for m in mods):
        raise ValueError("fused linears must share in_features, group size and bit width")
    has_bias = [m.bias is not None for m in mods]
    if any(has_bias) and not all(has_bias):
        raise ValueError("either all or none of the fused linears may have a bias")
    out = WQLinear_QUICK(first.w_bit, first.group_size, first.in_features, sum(m.out_features for m in mods), False,
                         "meta", first.k_split_1, first.k_split_2)
    out.qweight = QUICK_cat(*(m.qweight for m in mods), options="qweight")
    out.qzeros = QUICK_cat(*(m.qzeros for m in mods), options="qzeros")
    out.scales = QUICK_cat(*(m.scales for m in mods), options="scales")
    out.bias = torch.cat([m.bias for m in mods], dim=0) if all(has_bias) else None
    return out


def fuse_qkv_quick(module, q_proj, k_proj, v_proj):
    qkv_layer = WQLinear_QUICK(
        q_proj.w_bit,
        q_proj.group_size,
        q_proj.in_features,
        q_proj.out_features + k_proj.out_features + v_proj.out_features,
        q_proj.bias is not None,
        next(iter(module.state_dict().values())).device,
    )
    bias = torch.cat([q_proj.bias, k_proj.bias, v_proj.bias], dim=0) if q_proj.bias is not None else None
    qkv_layer.qweight = QUICK_cat(q_proj.qweight, k_proj.qweight, v_proj.qweight, options="qweight")
    qkv_layer.qzeros = QUICK_cat(q_proj.qzeros, k_proj.qzeros, v_proj.qzeros, options="qzeros")
    qkv_layer.scales = QUICK_cat(q_proj.scales, k_proj.scales, v_proj.scales, options="scales")
    qkv_layer.bias = bias
    return qkv_layer
