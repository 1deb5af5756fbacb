"""WQLinear_QUICK — host-side mirror of the reference module
(/root/reference/quick/awq/modules/linear/quick.py:35-171): same constructor, same buffer names,
shapes and dtypes (checkpoints load unchanged), same ``from_linear`` arguments, same ``forward``
semantics.  Differences, all behind the same interface:

* the packer is a vectorised closed form (quick_b200.layout.pack_quick / the GPU packer) instead of
  python loops of tiny kernels; it has no ``N == 128 or N % 256 == 0`` restriction and no hard-coded
  'cuda' (reference quick.py:95,101,110-115,147);
* ``forward`` converts the weight once to the B200 layout (non-persistent buffers, re-done if the
  packed buffers are replaced or modified) and runs the tcgen05 kernel with the bias fused into the
  epilogue: one launch instead of GEMM + ``sum`` + bias add (reference quick.py:161-165,
  gemm_cuda_quick.cu:1515).  ``k_split_1/2`` are accepted and forwarded as hints only.
"""
import torch
import torch.nn as nn

import quick_kernels  # the drop-in extension; import failure must break this module like the reference (quick.py:4)
from quick_kernels import gemm_forward_cuda_quick  # noqa: F401  (re-exported, same name as the reference)

from ....layout import pack_quick


class _FastLinear:
    """quick_kernels.B200Linear plus the identities of what it was built from (rebuilt when the copy or the bias changes)."""
    __slots__ = ("obj", "wq_id", "bias_id")

    def __init__(self, wq, sz, bias, K, N, G):
        self.obj = quick_kernels.B200Linear(wq, sz, bias, K, N, G)
        self.wq_id, self.bias_id = id(wq), id(bias)


class WQLinear_QUICK(nn.Module):
    SILU_FUSED_MAX_ROWS = 256     # forward_silu_mul: rows up to which SiLU·up runs inside the GEMM epilogue

    def __init__(self, w_bit, group_size, in_features, out_features, bias, dev, k_split_1=2, k_split_2=8):
        super().__init__()
        if w_bit not in [4]:
            raise NotImplementedError("Only 4-bit are supported for now.")
        self.in_features = in_features
        self.out_features = out_features
        self.w_bit = w_bit
        self.group_size = group_size if group_size != -1 else in_features
        self.k_split_1 = k_split_1
        self.k_split_2 = k_split_2
        assert self.in_features % self.group_size == 0
        assert out_features % (32 // self.w_bit) == 0
        pack = 32 // self.w_bit
        self.register_buffer("qweight", torch.zeros((in_features // 4, out_features // pack * 4), dtype=torch.int32, device=dev))
        self.register_buffer("qzeros", torch.zeros((in_features // self.group_size, out_features * 2 // pack), dtype=torch.int32, device=dev))
        self.register_buffer("scales", torch.zeros((in_features // self.group_size, out_features * 2), dtype=torch.float16, device=dev))
        if bias:
            self.register_buffer("bias", torch.zeros((out_features), dtype=torch.float16, device=dev))
        else:
            self.bias = None
        self._b200 = None       # (wq, sz) in the B200 layout — derived, never saved
        self._b200_src = None   # the packed tensors it was derived from (strong references: their storage cannot be
                                # freed and re-used at the same address while the copy is cached) and their versions
        self._b200_frozen = False
        self._fast = None       # quick_kernels.B200Linear bound to the current B200 copy and bias (forward's fast path)
        self.gated_pairs = False    # True: the B200 copy has interleaved gate / up channels (enable_silu_mul)

    @classmethod
    def from_linear(cls, linear, w_bit, group_size, init_only=False, scales=None, zeros=None, k_split_1=2, k_split_2=8):
        awq_linear = cls(w_bit, group_size, linear.in_features, linear.out_features, linear.bias is not None,
                         linear.weight.device, k_split_1, k_split_2)
        if init_only:  # just prepare for loading sd
            return awq_linear
        assert scales is not None and zeros is not None
        G = awq_linear.group_size
        # intweight[k, n] = round((W[n, k] + z*s) / s)   (reference quick.py:76-81), scales/zeros are (N, K/G)
        # (computed in fp32: with fp16 weights the reference's own-dtype division can land on .5 ties)
        s16 = scales.clone().half()
        s_rep = s16.float().repeat_interleave(G, dim=1)
        zs_rep = (zeros.float() * scales.float()).repeat_interleave(G, dim=1)
        intweight = torch.round((linear.weight.data.float() + zs_rep) / s_rep).clamp_(0, 15).to(torch.int32).t().contiguous()
        qweight, qzeros, qscales = pack_quick(intweight, zeros.t().contiguous().to(torch.int32), s16.t().contiguous())
        awq_linear.qweight = qweight
        awq_linear.qzeros = qzeros
        awq_linear.scales = qscales
        if linear.bias is not None:
            awq_linear.bias = linear.bias.clone().half()
        return awq_linear

    @classmethod
    def from_awq_gemm(cls, qweight, qzeros, scales, bias=None, w_bit=4, group_size=None, k_split_1=2, k_split_2=8):
        """Build the module from AWQ "GEMM" checkpoint tensors (the layout of WQLinear_GEMM, reference
        quick/awq/modules/linear/gemm.py:37-58: qweight int32 [K, N/8], qzeros int32 [K/G, N/8], scales fp16 [K/G, N])
        without re-quantizing: a bit-exact nibble permutation into the QUICK layout (GPU kernel on CUDA tensors,
        quick_b200.layout on CPU tensors).  The reference can only produce QUICK modules from fp16 weights."""
        K, N = int(qweight.shape[0]), int(qweight.shape[1]) * 8
        G = K // int(scales.shape[0])
        if group_size not in (None, -1, G):
            raise ValueError(f"group_size={group_size} does not match the tensors (K/G rows of scales -> G={G})")
        m = cls(w_bit, G, K, N, bias is not None, qweight.device, k_split_1, k_split_2)
        if qweight.is_cuda:
            from .... import ops
            qw, qz, sc = ops.awq_gemm_to_quick(qweight, qzeros, scales)
        else:
            from ....layout import awq_gemm_to_quick
            qw, qz, sc = awq_gemm_to_quick(qweight, qzeros, scales)
        m.qweight, m.qzeros, m.scales = qw, qz, sc
        if bias is not None:
            m.bias = bias.clone().half()
        return m

    def _prepacked(self):
        """(wq, sz): the B200-layout copy the kernel streams, rebuilt when a packed buffer is replaced or modified
        in place (ordinary tensors: version counter; inference tensors have none and count as immutable)."""
        if self._b200_frozen:
            return self._b200
        src = (self.qweight, self.qzeros, self.scales)
        ver = tuple(0 if t.is_inference() else t._version for t in src)
        cur = self._b200_src
        if self._b200 is None or cur is None or any(a is not b for a, b in zip(cur[0], src)) or cur[1] != ver:
            wq, sz = quick_kernels.prepack_quick(self.qweight, self.scales, self.qzeros, self.in_features)
            self._b200, self._b200_src, self._fast = (wq, sz), (src, ver), None
        return self._b200

    def enable_silu_mul(self):
        """This module is a fused gate|up projection ([gate | up] along N): switch its B200 copy to interleaved output
        channels so that ``forward_silu_mul`` computes silu(gate(x)) * up(x) in the GEMM epilogue (QB200_GEMM_SILU_MUL,
        the reference's QuantFusedMLP.our_llama_mlp, modules/fused/mlp.py:52-76, in one launch).  The plain ``forward``
        is no longer available on this module afterwards."""
        if getattr(self, "gated_pairs", False):
            return self
        from .... import ops
        wq, sz = self._prepacked()
        wq, sz, bias = ops.interleave_pairs(wq, sz, self.in_features, self.out_features, self.group_size, self.bias)
        self._b200, self._pair_bias, self.gated_pairs, self._b200_frozen = (wq, sz), bias, True, True
        return self

    @torch.no_grad()
    def forward_silu_mul(self, x, ssq=None, eps=1e-6):
        """silu(gate(x)) * up(x) -> (..., out_features / 2), bit-identical to forward + qb200_silu_mul.
        ssq (fp32 [K/128, rows]): x is the gamma-scaled copy a ``forward_norm_out`` producer wrote — the RMSNorm in front
        of this projection is applied inside the kernel (rows scaled by 1/rms before the SiLU)."""
        assert getattr(self, "gated_pairs", False), "call enable_silu_mul() first"
        wq, sz = self._b200
        x2d = x.reshape(-1, x.shape[-1])
        fused = x2d.shape[0] <= self.SILU_FUSED_MAX_ROWS
        # large token tiles: the epilogue's SiLU math would idle the tensor pipe (prefill -4 % measured on 7B shapes);
        # run the GEMM plain (its output columns are the interleaved pairs) and one elementwise kernel behind it
        if ssq is None:
            out = quick_kernels.gemm_forward_b200(x2d, wq, sz, self._pair_bias, self.out_features, self.group_size, False, None, fused)
        else:
            out = quick_kernels.gemm_forward_b200_norm(x2d, wq, sz, self._pair_bias, self.out_features, self.group_size, None, fused,
                                                       None, ssq, eps)[0]
        if not fused:
            from .... import ops
            out = ops.silu_mul_interleaved(out)
        return out.reshape(x.shape[:-1] + (self.out_features // 2,))

    @torch.no_grad()
    def forward_norm_out(self, x, residual, gamma):
        """RMSNorm folded around the GEMMs (qb200_gemm_w4a16_norm, producer side): h = residual + linear(x) as ``forward``
        does, plus what the RMSNorm with weight ``gamma`` that reads h next needs — returns (h, h * gamma in fp16,
        per-128-column-tile sums of squares fp32 [out_features / 128, rows]).  Feed the last two to ``forward_normed`` /
        ``forward_silu_mul(ssq=…)`` of the projection behind that norm."""
        if getattr(self, "gated_pairs", False):
            raise RuntimeError("forward_norm_out is not available on an interleaved gate|up module")
        wq, sz = self._prepacked()
        res2d = None if residual is None else residual.reshape(-1, self.out_features)
        out, normed, ssq = quick_kernels.gemm_forward_b200_norm(x.reshape(-1, x.shape[-1]), wq, sz, self.bias, self.out_features,
                                                                self.group_size, res2d, False, gamma, None, 0.0)
        shape = x.shape[:-1] + (self.out_features,)
        return out.reshape(shape), normed.reshape(shape), ssq

    @torch.no_grad()
    def forward_normed(self, x_gamma, ssq, eps, residual=None):
        """Consumer side: linear(rmsnorm(h)) from a producer's (h * gamma, ssq) — rows scaled by 1/rms inside the kernel."""
        if getattr(self, "gated_pairs", False):
            raise RuntimeError("this gate|up module has interleaved output channels: use forward_silu_mul(ssq=...)")
        wq, sz = self._prepacked()
        res2d = None if residual is None else residual.reshape(-1, self.out_features)
        out = quick_kernels.gemm_forward_b200_norm(x_gamma.reshape(-1, x_gamma.shape[-1]), wq, sz, self.bias, self.out_features,
                                                   self.group_size, res2d, False, None, ssq, eps)[0]
        return out.reshape(x_gamma.shape[:-1] + (self.out_features,))

    def release_quick_buffers(self, drop=False):
        """Inference-only deployments: keep the B200 copy on the GPU and move the QUICK-layout buffers (qweight /
        qzeros / scales — the checkpoint format, no longer read by the kernel) to host memory, halving the weight
        footprint on the device.  ``state_dict`` / ``save_quantized`` keep working from the host copies;
        ``restore_quick_buffers`` (or loading a state dict) brings them back.  drop=True discards them instead
        (throw-away random-init benchmark models: nothing to save)."""
        if not self.qweight.is_cuda:
            return self
        self._prepacked()
        self._b200_device = self.qweight.device
        for name in ("qweight", "qzeros", "scales"):
            t = getattr(self, name)
            setattr(self, name, torch.empty((0,) * t.dim(), dtype=t.dtype) if drop else t.to("cpu"))
        self._b200_src = None
        self._b200_frozen = True
        return self

    def restore_quick_buffers(self):
        if self._b200_frozen and getattr(self, "_b200_device", None) is not None:
            dev = self._b200_device
            for name in ("qweight", "qzeros", "scales"):
                setattr(self, name, getattr(self, name).to(dev))
            self._b200_frozen = False
        return self

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        if self._b200_frozen:           # new packed tensors arrived: the frozen copy is stale
            self.restore_quick_buffers()
        self._b200 = self._b200_src = None
        self._b200_frozen = self.gated_pairs = False

    @torch.no_grad()
    def forward(self, x, residual=None):
        """residual (same shape as the output): returns residual + linear(x), the add fused into the GEMM epilogue."""
        if getattr(self, "gated_pairs", False):
            raise RuntimeError("this gate|up module has interleaved output channels (enable_silu_mul): use forward_silu_mul")
        wq, sz = self._prepacked()
        fast = self._fast
        if fast is None or fast.wq_id != id(wq) or fast.bias_id != id(self.bias):
            # one C++ object per B200 copy does the per-call work (flatten, launch, reshape): the eager module surface is
            # host-bound below 64 rows, and most of that was this method
            fast = self._fast = _FastLinear(wq, sz, self.bias, self.in_features, self.out_features, self.group_size)
        return fast.obj.forward(x, residual)

    @torch.no_grad()
    def forward_reference_call(self, x):
        """The reference's exact call sequence (quick.py:158-166) through the drop-in symbol."""
        out_shape = x.shape[:-1] + (self.out_features,)
        split = self.k_split_1 if self.out_features > self.in_features else self.k_split_2
        out = gemm_forward_cuda_quick(x.reshape(-1, x.shape[-1]), self.qweight, self.scales, self.qzeros, split)
        out = out + self.bias if self.bias is not None else out
        return out.reshape(out_shape)

    def extra_repr(self) -> str:
        return "in_features={}, out_features={}, bias={}, w_bit={}, group_size={}".format(
            self.in_features, self.out_features, self.bias is not None, self.w_bit, self.group_size)
