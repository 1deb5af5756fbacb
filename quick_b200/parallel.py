"""Column-parallel (N-sharded) QUICK linear — the only place the hot path has an exchange step.

The reference has no distributed code at all (accelerate layer placement only).  The W4A16 GEMM
shards naturally along N — independent 128-column tiles, no K reduction across ranks — so rank r of R holds
`layout.shard_columns(qweight, qzeros, scales, r, R)` and computes its (M, N/R) slab with the same kernel;
one `all_gather_into_tensor` (NCCL over NVLink on GPUs, gloo in the CPU tests) rebuilds (M, N) where the
consumer needs the full width.  One process per GPU; torch.distributed is plumbing.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import layout


def default_gemm(x2d: torch.Tensor, shard: "ColumnShard") -> torch.Tensor:
    """The product path: tcgen05 kernel through the C-ABI (CUDA only; raises without a GPU)."""
    from . import ops
    if shard.b200 is None:
        wq, sz, *_ = ops.prepack(shard.qweight, shard.qzeros, shard.scales)
        shard.b200 = (wq, sz)
    return ops.gemm(x2d, shard.b200[0], shard.b200[1], shard.n_local, shard.group_size, bias=shard.bias)


class ColumnShard:
    """This rank's slice of a packed weight (QUICK layout) plus its lazily built B200-layout copy."""

    def __init__(self, qweight, qzeros, scales, bias, rank: int, world: int):
        self.rank, self.world = rank, world
        self.qweight, self.qzeros, self.scales = layout.shard_columns(qweight, qzeros, scales, rank, world)
        self.in_features = qweight.shape[0] * 4
        self.n_total = qweight.shape[1] * 2
        self.n_local = self.n_total // world
        self.group_size = self.in_features // qzeros.shape[0]
        self.bias = None if bias is None else bias[rank * self.n_local:(rank + 1) * self.n_local].contiguous()
        self.b200 = None


class ColumnParallelQuickLinear(torch.nn.Module):
    """y = x · W with W's output columns split across the process group; `gather_output` all-gathers the slabs.

    `gemm_fn(x2d, shard) -> (M, N/R)` is injectable so the host logic (sharding, collective, reassembly) is
    testable on CPU with the oracle standing in for the kernel; the default is the CUDA kernel."""

    def __init__(self, qweight, qzeros, scales, bias=None, group: Optional[dist.ProcessGroup] = None,
                 gather_output: bool = True, gemm_fn: Callable = default_gemm):
        super().__init__()
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.shard = ColumnShard(qweight, qzeros, scales, bias, self.rank, self.world)
        self.gather_output = gather_output
        self.gemm_fn = gemm_fn

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x2d = x.reshape(-1, x.shape[-1])
        local = self.gemm_fn(x2d, self.shard)                       # (M, N/R)
        if not self.gather_output or self.world == 1:
            return local.reshape(x.shape[:-1] + (local.shape[-1],))
        M = local.shape[0]
        gathered = torch.empty((self.world * M, self.shard.n_local), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(gathered, local.contiguous(), group=self.group)
        # rank-major slabs -> natural column order
        out = gathered.view(self.world, M, self.shard.n_local).permute(1, 0, 2).reshape(M, self.shard.n_total)
        return out.reshape(x.shape[:-1] + (self.shard.n_total,))


class PeerGatherWorkspace:
    """Fused GEMM + all-gather over peer memory (C-ABI qb200_gemm_w4a16_allgather + qb200_peer_barrier).

    One full-width [max_rows, n_total] fp16 output buffer per rank in torch symmetric memory (every rank's buffer is
    mapped into every process), a small symmetric flag array and a device epoch counter.  `gemm()` makes this rank's
    GEMM store its column slab into all ranks' buffers and then meets the peers; the returned tensor is a view of the
    local buffer holding the complete rows in natural column order — no NCCL call, no re-layout copy.
    A workspace may be shared by calls that are separated by at least one other peer barrier (e.g. one workspace
    per projection type of a decoder layer): the barrier in between proves every rank is done reading the old rows."""

    def __init__(self, max_rows: int, n_total: int, group=None, device=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if not (1 <= self.world <= 8):
            raise ValueError("peer gather supports 1..8 ranks (one NVLink domain)")
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.max_rows, self.n_total = max_rows, n_total
        self.buf = symm_mem.empty((max_rows, n_total), dtype=torch.float16, device=device)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)
        self.flags.zero_()
        hb = symm_mem.rendezvous(self.buf, self.group)
        hf = symm_mem.rendezvous(self.flags, self.group)
        self.buf_ptrs = [int(p) for p in hb.buffer_ptrs]
        # NVSwitch multicast mapping of the same buffers (one multimem.st reaches every rank); QB200_TP_MULTICAST=0 keeps
        # the loop of peer stores
        import os
        self.multicast_ptr = None
        if os.environ.get("QB200_TP_MULTICAST", "1") == "1" and self.world > 1:
            try:
                if hb.has_multicast_support:
                    mp_ = int(hb.multicast_ptr)
                    self.multicast_ptr = mp_ if mp_ != 0 else None
            except Exception:
                self.multicast_ptr = None
        self.flag_ptrs = [int(p) for p in hf.buffer_ptrs]
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self._handles = (hb, hf)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)          # every rank's flags are zero before anybody signals

    def gemm(self, x2d: torch.Tensor, wq: torch.Tensor, sz: torch.Tensor, n_local: int, G: int, bias=None, residual=None):
        """x2d [M, K] fp16 -> view [M, n_total] of the local buffer with every rank's slab in place.
        residual: local full-width [M, n_total] tensor added to this rank's columns before they are sent."""
        from . import ops
        M = x2d.shape[0]
        if M > self.max_rows:
            raise ValueError(f"M={M} exceeds the workspace ({self.max_rows} rows)")
        ops.gemm_allgather(x2d, wq, sz, n_local, G, self.buf_ptrs, self.n_total, self.rank * n_local, bias=bias, residual=residual,
                           multicast_ptr=self.multicast_ptr)
        ops.peer_barrier(self.epoch, self.flag_ptrs, self.rank)
        return self.buf[:M]

    def release(self):
        """Extra meeting point: call after the last read of the gathered rows when the NEXT use of this workspace
        is not separated from this one by another peer barrier (back-to-back calls on one workspace)."""
        from . import ops
        ops.peer_barrier(self.epoch, self.flag_ptrs, self.rank)


class GatheredBuffer:
    """A [max_rows, width] fp16 buffer that exists on every rank (torch symmetric memory: every rank's copy is mapped
    into every process, plus the NVSwitch multicast mapping when the fabric has one) with the hand-over state of
    include/quick_b200.h (qb200_peer_signal / qb200_peer_wait): a local epoch counter and a symmetric flag array.
    Producers (a column-parallel GEMM epilogue, silu_mul_tp, scatter_cols) store this rank's column slab into ALL
    copies; the first consumer of a fill (GEMM activations, RMSNorm rows) announces this rank's slab and meets the
    other ranks inside its own prologue — no barrier kernel, no NCCL call, CUDA-graph capturable."""

    def __init__(self, max_rows: int, width: int, group=None, device=None):
        import ctypes as C
        import os
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if not (1 <= self.world <= 8):
            raise ValueError("peer gather supports 1..8 ranks (one NVLink domain)")
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.max_rows, self.width = max_rows, width
        self.buf = symm_mem.empty((max_rows, width), dtype=torch.float16, device=device)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)
        self.flags.zero_()
        hb = symm_mem.rendezvous(self.buf, self.group)
        hf = symm_mem.rendezvous(self.flags, self.group)
        self.buf_ptrs = (C.c_void_p * self.world)(*[int(p) for p in hb.buffer_ptrs])
        self.flag_ptrs = (C.c_void_p * self.world)(*[int(p) for p in hf.buffer_ptrs])
        self.multicast_ptr = None
        if os.environ.get("QB200_TP_MULTICAST", "1") == "1" and self.world > 1:
            try:
                if hb.has_multicast_support:
                    mp_ = int(hb.multicast_ptr)
                    self.multicast_ptr = mp_ if mp_ != 0 else None
            except Exception:
                self.multicast_ptr = None
        self.state = torch.zeros(1, dtype=torch.int32, device=device)      # the local epoch counter
        self.wait = _lib.PeerWait(self.state.data_ptr(), self.flag_ptrs, self.rank, self.world)
        self.signal = _lib.PeerSignal(self.state.data_ptr())
        self._handles = (hb, hf)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)          # every rank's flags are zero before anybody publishes

    def rows(self, M: int) -> torch.Tensor:
        if M > self.max_rows:
            raise ValueError(f"M={M} exceeds the gathered buffer ({self.max_rows} rows)")
        return self.buf[:M]
