"""CPU: host-side layout algebra, the WQLinear_QUICK mirror, the C-ABI surface (symbols only — no
compute without a GPU) and the loud failure when no CUDA device is present."""
import glob
import os
import re

import numpy as np
import pytest
import torch

from oracle import quick_oracle as qo
from quick_b200 import layout

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACK_FILES = sorted(glob.glob(os.path.join(GOLD, "pack_*.npz")))


@pytest.mark.parametrize("path", PACK_FILES, ids=os.path.basename)
def test_layout_pack_matches_golden(path):
    d = np.load(path)
    q, z, s = (torch.from_numpy(d[k].astype(np.int32)) for k in ("q", "z")), None, None
    q = torch.from_numpy(d["q"].astype(np.int32)); z = torch.from_numpy(d["z"].astype(np.int32)); s = torch.from_numpy(d["s"])
    qw, qz, sc = layout.pack_quick(q, z, s)
    assert np.array_equal(qw.numpy(), d["qweight"]) and np.array_equal(qz.numpy(), d["qzeros"])
    assert np.array_equal(sc.numpy().view(np.uint16), d["scales"].view(np.uint16))
    q2, z2, s2 = layout.unpack_quick(qw, qz, sc)
    assert torch.equal(q2, q) and torch.equal(z2, z) and torch.equal(s2, s)


def test_layout_n384_works_unlike_reference():
    """The reference packer only supports N == 128 or N % 256 == 0 (quick.py:110-115); the closed
    form only needs N % 128 == 0, like the kernel (gemm_cuda_quick.cu:1479)."""
    q, z, s = qo.make_case(128, 384, 64)
    qw, qz, sc = layout.pack_quick(torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s))
    eq, ez, es = qo.pack_quick(q, z, s)
    assert np.array_equal(qw.numpy(), eq) and np.array_equal(qz.numpy(), ez)
    with pytest.raises(ValueError, match="cta_N = 128"):
        layout.pack_quick(torch.zeros(128, 192, dtype=torch.int32), torch.zeros(1, 192, dtype=torch.int32), torch.zeros(1, 192))


def test_cat_and_shard_roundtrip():
    K, G = 256, 128
    parts = [qo.make_case(K, n, G, seed=i) for i, n in enumerate((512, 128, 128))]
    packed = [layout.pack_quick(*(torch.from_numpy(a) for a in p)) for p in parts]
    cat = [layout.quick_cat([p[i] for p in packed], name) for i, name in enumerate(("qweight", "qzeros", "scales"))]
    whole = layout.pack_quick(*(torch.from_numpy(np.concatenate([p[i] for p in parts], 1)) for i in range(3)))
    for a, b in zip(cat, whole):
        assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    # column-parallel shards re-concatenate to the whole (the all-gather identity of SURVEY §8e)
    for world in (2, 3, 6):
        shards = [layout.shard_columns(*whole, r, world) for r in range(world)]
        for i, name in enumerate(("qweight", "qzeros", "scales")):
            re_cat = layout.quick_cat([sh[i] for sh in shards], name)
            assert torch.equal(re_cat.view(torch.uint8), whole[i].view(torch.uint8))
        # each shard is itself a valid packed weight of the corresponding logical columns
        N = whole[0].shape[1] * 2
        q_all = np.concatenate([p[0] for p in parts], 1)
        q_sh, _, _ = layout.unpack_quick(*shards[1])
        assert np.array_equal(q_sh.numpy(), q_all[:, N // world:2 * N // world])
    with pytest.raises(ValueError):
        layout.shard_columns(*whole, 0, 4)   # 768 / 4 = 192 is not a multiple of 128


def test_quantize_rtn_roundtrip():
    torch.manual_seed(0)
    W = torch.randn(256, 512) * 0.02
    q, z, s = layout.quantize_rtn(W, 128)
    assert q.min() >= 0 and q.max() <= 15 and z.min() >= 0 and z.max() <= 15
    W_hat = ((q - z.repeat_interleave(128, 0)).float() * s.float().repeat_interleave(128, 0)).t()
    assert (W - W_hat).abs().max() <= 0.6 * s.float().max()


def test_c_abi_exports_every_declared_symbol(built):
    """The library loads on a CPU-only box and exports exactly what include/quick_b200.h declares."""
    from quick_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "quick_b200.h")).read()
    declared = set(re.findall(r"\b(qb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.qb200_version()
    assert lib.qb200_wq_bytes(4096, 4096) == 4096 * 4096 // 2
    assert lib.qb200_sz_bytes(4096, 4096, 128) == 32 * 4096 * 4


def test_c_abi_shape_errors_match_reference_messages(built):
    from quick_b200 import _lib
    lib = _lib.load()
    assert lib.qb200_check_shape(1, 512, 512, 128) == 0
    for (K, N, G, msg) in [(512, 192, 128, "OC is not multiple of cta_N = 128"),
                           (512, 256, 48, "Group size should be a multiple of 32"),
                           (96, 256, 32, "IC is not a multiple of 64"),
                           (512, 256, 384, "IC is not a multiple of the group size")]:
        assert lib.qb200_check_shape(1, K, N, G) == _lib.QB200_EINVAL
        assert msg in lib.qb200_last_error().decode()
        with pytest.raises(ValueError, match=re.escape(msg)):
            _lib.check(lib.qb200_check_shape(1, K, N, G))


def test_gemm_plan_fills_the_machine(built):
    from quick_b200 import ops
    for M in (1, 8, 16, 64, 128, 256, 512, 4096):
        tok, split, ctas = ops.plan(M, 4096, 4096, 128)
        assert tok in (16, 32, 64, 128, 256) and split in (1, 2, 4, 8) and ctas >= 32
        assert ctas == (4096 // 128) * -(-M // tok) * split
        if M <= 512:
            # one ordered GEMM fills the 148 SMs; small tiles are sized for two co-resident CTAs per SM
            assert 100 <= ctas <= (2 * 148 if tok <= 64 else 148 + 148 // 4)
        # independent launches overlap each other: never split K, largest token tile
        itok, isplit, ictas = ops.plan(M, 4096, 4096, 128, independent=True)
        assert isplit == 1 and itok >= min(M, 256) and ictas == (4096 // 128) * -(-M // itok)


def test_module_fast_path_object_refuses_cpu_tensors(built):
    """quick_kernels.B200Linear (the per-call work of WQLinear_QUICK.forward in one C++ object) has no CPU path either: it
    refuses CPU weights at construction, and the module's forward on CPU tensors fails loudly instead of falling back."""
    import quick_kernels
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
    with pytest.raises(RuntimeError, match="no CPU path"):
        quick_kernels.B200Linear(torch.zeros(512 * 512 // 8, dtype=torch.int32), torch.zeros(4 * 512, dtype=torch.int32), None, 512, 512, 128)
    m = WQLinear_QUICK(4, 128, 512, 512, False, "cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 512, dtype=torch.float16))


def test_gemm_plan_picks_the_measured_optimum_for_the_layer_shapes(built):
    """Planner vs the split sweeps on B200 (tools/tune_shapes.py, profiles/r2b_tune_shapes_70b.log, r2g_tune_*): the
    choice for decode-sized M on the Llama-2-7B, Mistral-7B and Llama-2-70B layer shapes is the measured optimum
    (or within its noise).  Without a GPU the planner assumes 148 SMs."""
    from quick_b200 import ops
    want = {  # (K, N): split at M <= 16
        (4096, 12288): 2, (4096, 4096): 8, (4096, 22016): 2, (11008, 4096): 8,            # 7B q|k|v, o, gate|up, down
        (4096, 6144): 4, (4096, 28672): 1, (14336, 4096): 8,                              # Mistral-7B
        (8192, 10240): 4, (8192, 8192): 4, (8192, 57344): 2, (28672, 8192): 4,            # 70B on one GPU
        (8192, 5120): 4, (8192, 4096): 8, (8192, 28672): 2, (28672, 4096): 8,             # 70B on two GPUs
    }
    for (K, N), split in want.items():
        for M in (1, 16):
            tok, s, ctas = ops.plan(M, K, N, 128)
            assert (tok, s) == (16, split), (K, N, M, tok, s)
            assert ctas == N // 128 * s


def test_quick_kernels_module_surface(built):
    """Drop-in boundary: module name and symbol of csrc/pybind.cpp:5-8."""
    import quick_kernels
    assert callable(quick_kernels.gemm_forward_cuda_quick)
    assert "QUICK AWQ GEMM kernel." in quick_kernels.gemm_forward_cuda_quick.__doc__
    x = torch.zeros(1, 512, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        quick_kernels.gemm_forward_cuda_quick(x, torch.zeros(128, 256, dtype=torch.int32),
                                              torch.zeros(4, 1024, dtype=torch.float16),
                                              torch.zeros(4, 128, dtype=torch.int32), 8)


def test_product_path_fails_loudly_without_cuda(built):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from quick_b200 import _lib, ops
    with pytest.raises(_lib.QuickB200Error, match="no CPU fallback"):
        ops.prepack(torch.zeros(128, 256, dtype=torch.int32), torch.zeros(4, 128, dtype=torch.int32),
                    torch.zeros(4, 1024, dtype=torch.float16))


def test_wqlinear_quick_mirror_cpu(built):
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
    d = np.load(os.path.join(GOLD, "pack_K256_N768_G64.npz"))
    G, K, N = 64, 256, 768
    q = torch.from_numpy(d["q"].astype(np.int32)); z = torch.from_numpy(d["z"].astype(np.int32)); s = torch.from_numpy(d["s"])
    W = ((q - z.repeat_interleave(G, 0)).float() * s.float().repeat_interleave(G, 0)).t().contiguous()
    lin = torch.nn.Linear(K, N, bias=True)
    lin.weight.data = W
    m = WQLinear_QUICK.from_linear(lin, 4, G, False, scales=s.float().t().contiguous(), zeros=z.float().t().contiguous())
    assert np.array_equal(m.qweight.numpy(), d["qweight"]) and np.array_equal(m.qzeros.numpy(), d["qzeros"])
    assert np.array_equal(m.scales.numpy().view(np.uint16), d["scales"].view(np.uint16))
    # buffer names / shapes / dtypes of the reference (quick.py:52-58): checkpoints load unchanged
    sd = m.state_dict()
    assert set(sd) == {"qweight", "qzeros", "scales", "bias"}
    assert sd["qweight"].shape == (K // 4, N // 2) and sd["qweight"].dtype == torch.int32
    assert sd["qzeros"].shape == (K // G, N // 4) and sd["scales"].shape == (K // G, 2 * N)
    empty = WQLinear_QUICK.from_linear(lin, 4, G, init_only=True)
    empty.load_state_dict(sd)
    assert torch.equal(empty.qweight, m.qweight)
    assert "in_features=256, out_features=768, bias=True, w_bit=4, group_size=64" in repr(m)
    with pytest.raises(NotImplementedError):
        WQLinear_QUICK(8, 64, 256, 768, False, "cpu")


def test_fuse_qkv_quick_gqa_cpu(built):
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
    from quick_b200.awq.utils.fused_utils import QUICK_cat, fuse_qkv_quick
    K, G = 128, 128
    mods, parts = [], []
    for i, n in enumerate((256, 128, 128)):
        q, z, s = qo.make_case(K, n, G, seed=20 + i)
        parts.append((q, z, s))
        m = WQLinear_QUICK(4, G, K, n, False, "cpu")
        m.qweight, m.qzeros, m.scales = layout.pack_quick(torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s))
        mods.append(m)
    holder = torch.nn.ModuleList(mods)
    fused = fuse_qkv_quick(holder, *mods)
    whole = qo.pack_quick(*[np.concatenate([p[i] for p in parts], 1) for i in range(3)])
    assert fused.out_features == 512
    assert np.array_equal(fused.qweight.numpy(), whole[0]) and np.array_equal(fused.qzeros.numpy(), whole[1])
    with pytest.raises(ValueError):
        QUICK_cat(mods[0].qweight, options="qweight")
    with pytest.raises(ValueError):
        QUICK_cat(mods[0].qweight, mods[1].qweight, options="bogus")
