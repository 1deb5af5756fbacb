"""CPU, world_size 2, gloo: the N-sharded linear's host logic (shard slicing, all-gather, reassembly).
The kernel cannot run on CPU, so the oracle stands in for the local GEMM through the injectable gemm_fn —
the oracle is the checker here as everywhere else in tests/."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import quick_oracle as qo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_gemm(x2d, shard):
    out = qo.forward_oracle(x2d.numpy(), shard.qweight.numpy(), shard.qzeros.numpy(), shard.scales.numpy(),
                            None if shard.bias is None else shard.bias.numpy())
    return torch.from_numpy(out)


def _worker(rank, world, port, K, N, G, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quick_b200 import layout
    from quick_b200.parallel import ColumnParallelQuickLinear
    q, z, s = qo.make_case(K, N, G, seed=5)
    qw, qz, sc = layout.pack_quick(torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s))
    bias = torch.from_numpy(np.random.default_rng(1).standard_normal(N).astype(np.float16))
    lin = ColumnParallelQuickLinear(qw, qz, sc, bias, gemm_fn=_oracle_gemm)
    assert lin.shard.n_local == N // world and lin.shard.qweight.shape == (K // 4, N // world // 2)
    x = torch.from_numpy(qo.make_activations(6, K, seed=9)).reshape(2, 3, K)
    y = lin(x)
    full = torch.from_numpy(qo.forward_oracle(x.numpy(), qw.numpy(), qz.numpy(), sc.numpy(), bias.numpy()))
    ok = y.shape == (2, 3, N) and torch.equal(y, full)
    lin.gather_output = False
    y_local = lin(x)
    ok = ok and torch.equal(y_local, full[..., rank * N // world:(rank + 1) * N // world])
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_column_parallel_allgather_world2_gloo():
    world, K, N, G = 2, 256, 512, 128
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), K, N, G, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _runner_gather_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quick_b200.awq.models.llama_like import PRESETS, TensorParallel, tp_pad_intermediate, tp_world
    tp = TensorParallel(PRESETS["tiny"], 1, "cpu")          # CPU: the all-gather ("nccl") mode, no peer buffers
    local = (torch.arange(6 * 4, dtype=torch.float32).reshape(6, 4) + 100 * rank).half()
    full = tp.all_gather_cols(local)
    want = torch.cat([(torch.arange(24, dtype=torch.float32).reshape(6, 4) + 100 * r).half() for r in range(world)], dim=-1)
    ok = tp_world() == world and tp.mode == "nccl" and full.shape == (6, 4 * world) and torch.equal(full, want)
    ok = ok and tp_pad_intermediate(11008, 8) == 11264 and tp_pad_intermediate(11008, 2) == 11008 and tp.I_pad == 1024
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_runner_tensor_parallel_gather_world2_gloo():
    """The runner's tensor-parallel reassembly (column slabs -> full width, rank-major order) on CPU."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_runner_gather_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
