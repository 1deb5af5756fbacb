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
