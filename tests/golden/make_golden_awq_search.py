"""Golden vectors for the AWQ search arithmetic, produced by the reference's OWN AwqQuantizer methods (run in the
build container only: needs /root/reference; tests and the GPU box never run this).

  python tests/golden/make_golden_awq_search.py

The reference package does not import here (transformers 5.x / accelerate / awq_ext), so the class body of
``AwqQuantizer`` (quick/awq/quantize/quantizer.py:27-466) is extracted with ``ast`` and exec'd with the handful of
names it needs; ``get_op_name`` comes from quick/awq/utils/module.py.  Methods exercised, unmodified:
``pseudo_quantize_tensor`` (:47-70), ``_search_best_scale`` + ``_compute_best_scale`` (:196-296),
``_compute_best_clip`` (:312-362).  Inputs are fp32 CPU tensors of a small SwiGLU MLP.

One harness detail: ``_compute_best_scale`` keeps the original weights as ``{k: v.cpu() for … state_dict()}``
(:247) and restores them after every candidate.  For CUDA modules ``.cpu()`` copies; for CPU modules it ALIASES the
live weights, which the next in-place ``fc.weight.mul_`` (:262) then corrupts.  The inspected modules therefore get a
``state_dict`` that returns clones, which reproduces the behaviour the reference has on its intended device.
"""
import ast
import functools
import inspect
import logging
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref_quantizer_class():
    src = open(f"{REF}/quick/awq/quantize/quantizer.py").read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "AwqQuantizer"][0]
    mod_src = open(f"{REF}/quick/awq/utils/module.py").read()
    ns = {}
    exec(compile(mod_src, "ref_module.py", "exec"), ns)
    from typing import Dict, List
    env = {"torch": torch, "nn": nn, "inspect": inspect, "logging": logging, "functools": functools, "Dict": Dict, "List": List,
           "get_op_name": ns["get_op_name"], "clear_memory": lambda *a: None, "get_best_device": lambda: "cpu", "tqdm": lambda x, **k: x}
    exec(compile(ast.get_source_segment(src, cls), "ref_quantizer.py", "exec"), env)
    return env["AwqQuantizer"]


class MLP(nn.Module):
    def __init__(self, H, I):
        super().__init__()
        self.gate_proj, self.up_proj, self.down_proj = nn.Linear(H, I, bias=False), nn.Linear(H, I, bias=False), nn.Linear(I, H, bias=False)

    def forward(self, x):
        return self.down_proj(nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))


class Holder(nn.Module):          # the "decoder layer" _search_best_scale resolves op names in
    def __init__(self, H, I):
        super().__init__()
        self.norm = nn.LayerNorm(H)
        self.mlp = MLP(H, I)


def load_ref_scale_functions():
    """scale_ln_fcs / scale_fc_fc / apply_clip from quick/awq/quantize/scale.py (:16-26, :63-101), extracted one by one:
    the module's own imports (BloomGelu, PytorchGELUTanh, …) do not resolve on transformers 5.x."""
    src = open(f"{REF}/quick/awq/quantize/scale.py").read()
    mod_src = open(f"{REF}/quick/awq/utils/module.py").read()
    ns = {}
    exec(compile(mod_src, "ref_module.py", "exec"), ns)
    from typing import List, Tuple
    env = {"torch": torch, "nn": nn, "List": List, "Tuple": Tuple, "get_op_by_name": ns["get_op_by_name"],
           "set_op_by_name": ns["set_op_by_name"], "get_best_device": lambda: "cpu", "allowed_act_fns": []}
    for fn in ast.parse(src).body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("scale_ln_fcs", "scale_fc_fc", "apply_clip"):
            exec(compile(ast.get_source_segment(src, fn), f"ref_scale_{fn.name}.py", "exec"), env)
    return env["scale_ln_fcs"], env["scale_fc_fc"], env["apply_clip"]


def main():
    Ref = load_ref_quantizer_class()
    ref_scale_ln_fcs, ref_scale_fc_fc, ref_apply_clip = load_ref_scale_functions()
    for (H, I, G, T, duo, seed) in [(128, 256, 128, 512, True, 0), (128, 256, 64, 640, True, 1), (128, 384, 32, 768, False, 2)]:
        torch.manual_seed(seed)
        layer = Holder(H, I)
        layer.norm.weight.data.normal_(1.0, 0.1)
        layer.norm.bias.data.normal_(0.0, 0.1)
        # outlier input channels and heavy-tailed weights so that the searches have something to find
        x = torch.randn(T, H) * (1 + 8 * (torch.rand(H) < 0.05).float())
        for lin in (layer.mlp.gate_proj, layer.mlp.up_proj, layer.mlp.down_proj):
            lin.weight.data = lin.weight.data * (1 + 3 * (torch.rand_like(lin.weight) < 0.01).float())
        for m in (layer.mlp, layer.mlp.down_proj):
            m.state_dict = functools.partial(lambda mod, *a, **k: {n: v.clone() for n, v in nn.Module.state_dict(mod, *a, **k).items()}, m)
        q = Ref.__new__(Ref)
        q.w_bit, q.group_size, q.duo_scaling = 4, G, duo
        mlp = layer.mlp
        with torch.no_grad():
            dq, s, z = q.pseudo_quantize_tensor(mlp.down_proj.weight.data.clone(), get_scale_zp=True)
            _, _, s_gate_up = q._search_best_scale(layer, layer.norm, [mlp.gate_proj, mlp.up_proj], x.clone(), module2inspect=mlp, kwargs={})
            h = (nn.functional.silu(mlp.gate_proj(x)) * mlp.up_proj(x)).detach()
            _, _, s_down = q._search_best_scale(layer, mlp.up_proj, [mlp.down_proj], h.clone())
            clip = q._compute_best_clip(mlp.down_proj.weight, h.clone())
        # folding the found scales and clips into the layer with the reference's own functions (on a copy)
        import copy
        folded = copy.deepcopy(layer)
        with torch.no_grad():
            ref_scale_ln_fcs(folded.norm, [folded.mlp.gate_proj, folded.mlp.up_proj], s_gate_up.clone())
            ref_scale_fc_fc(folded.mlp.up_proj, folded.mlp.down_proj, s_down.clone())
            after_scale = {k: v.clone() for k, v in nn.Module.state_dict(folded).items()}
            ref_apply_clip(folded, [("mlp.down_proj", clip.clone())])
            after_clip_down = folded.mlp.down_proj.weight.data.clone()
        out = dict(norm_w=layer.norm.weight.data.numpy(), norm_b=layer.norm.bias.data.numpy(),
                   fold_norm_w=after_scale["norm.weight"].numpy(), fold_norm_b=after_scale["norm.bias"].numpy(),
                   fold_gate=after_scale["mlp.gate_proj.weight"].numpy(), fold_up=after_scale["mlp.up_proj.weight"].numpy(),
                   fold_down=after_scale["mlp.down_proj.weight"].numpy(), clip_applied_down=after_clip_down.numpy(),
                   H=np.int32(H), I=np.int32(I), G=np.int32(G), duo=np.bool_(duo), x=x.numpy(), gate=mlp.gate_proj.weight.data.numpy(),
                   up=mlp.up_proj.weight.data.numpy(), down=mlp.down_proj.weight.data.numpy(), pq_dq=dq.numpy(), pq_scales=s.numpy(),
                   pq_zeros=z.numpy(), scales_gate_up=s_gate_up.numpy(), scales_down=s_down.numpy(), clip_down=clip.numpy())
        name = f"awqsearch_H{H}_I{I}_G{G}_T{T}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()}, os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    main()
