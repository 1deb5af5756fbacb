"""GPU parity tests (run with -m gpu on a B200).  Every call goes through the C-ABI
(libquick_b200.so via ctypes) or the drop-in `quick_kernels` module; the oracle
(oracle/quick_oracle.py, and the unmodified reference kernel in oracle/_ref when present) is only
the checker.

Tolerances (north_star): bit-exact for the integer unpack / index work (pack, relayout, W16);
GEMM outputs within rtol = 1e-2, atol = 1e-2 * rms(reference) of the fp64 product of the same fp16
operands (tensor-core accumulation order is not IEEE-sequential).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import quick_oracle as qo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-2


def assert_close(out, exact, what=""):
    """out: torch fp16 (GPU); exact: torch fp64 or fp32 (GPU)."""
    exact = exact.double()
    rms = exact.pow(2).mean().sqrt().item()
    ok = torch.allclose(out.double(), exact, rtol=RTOL, atol=RTOL * rms)
    if not ok:
        err = (out.double() - exact).abs().max().item()
        raise AssertionError(f"{what}: max_abs_err={err:.4g} rms={rms:.4g}")


@pytest.fixture(scope="module")
def ops(built):
    from quick_b200 import ops as _ops
    assert torch.cuda.get_device_capability()[0] == 10, "these tests need an sm_100 device"
    return _ops


def make_gpu_case(ops, K, N, G, seed=1234):
    q, z, s = qo.make_case(K, N, G, seed)
    tq, tz, ts = torch.from_numpy(q).cuda(), torch.from_numpy(z).cuda(), torch.from_numpy(s).cuda()
    qw, qz, sc = ops.pack_quick(tq, tz, ts, G)
    W16 = ((tq - tz.repeat_interleave(G, 0)).half() * ts.repeat_interleave(G, 0))   # == oracle dequant_w16 (checked below)
    return q, z, s, qw, qz, sc, W16


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pack_*.npz"))), ids=os.path.basename)
def test_gpu_packer_and_relayout_bitexact_vs_golden(ops, path):
    """GPU packer == reference packer output; relayout + dequantize == oracle W16, all bit-exact."""
    d = np.load(path)
    G = int(d["G"])
    q, z, s = d["q"].astype(np.int32), d["z"].astype(np.int32), d["s"]
    qw, qz, sc = ops.pack_quick(torch.from_numpy(q).cuda(), torch.from_numpy(z).cuda(), torch.from_numpy(s).cuda(), G)
    assert np.array_equal(qw.cpu().numpy(), d["qweight"])
    assert np.array_equal(qz.cpu().numpy(), d["qzeros"])
    assert np.array_equal(sc.cpu().numpy().view(np.uint16), d["scales"].view(np.uint16))
    # relayout of the GOLDEN tensors (identical packed inputs as the reference would see)
    wq, sz, K, N, G2 = ops.prepack(torch.from_numpy(d["qweight"]).cuda(), torch.from_numpy(d["qzeros"]).cuda(),
                                   torch.from_numpy(d["scales"]).cuda())
    assert G2 == G
    W = ops.dequantize(wq, sz, K, N, G).cpu().numpy()
    assert np.array_equal(W.view(np.uint16), qo.dequant_w16(q, z, s, G).view(np.uint16))
    assert np.array_equal(W.view(np.uint16), qo.kernel_view_w16(d["qweight"], d["qzeros"], d["scales"], G).view(np.uint16))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "awqgemm_*.npz"))), ids=os.path.basename)
def test_awq_gemm_checkpoint_converters_bitexact_vs_golden(ops, path):
    """AWQ-GEMM checkpoint tensors -> QUICK layout and -> B200 layout on the GPU, against the
    integers the reference's own unpacker produced (tests/golden/make_golden_awq_gemm.py), then through the GEMM."""
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
    d = np.load(path)
    G = int(d["G"])
    gq, gz, gs = (torch.from_numpy(d[k]).cuda() for k in ("qweight", "qzeros", "scales"))
    qw, qz, sc = ops.awq_gemm_to_quick(gq, gz, gs)
    want = qo.pack_quick(d["q"].astype(np.int32), d["z"].astype(np.int32), d["scales"])
    assert np.array_equal(qw.cpu().numpy(), want[0]) and np.array_equal(qz.cpu().numpy(), want[1])
    assert np.array_equal(sc.cpu().numpy().view(np.uint16), want[2].view(np.uint16))
    wq, sz, K, N, G2 = ops.prepack_awq_gemm(gq, gz, gs)
    assert G2 == G
    wq2, sz2, *_ = ops.prepack(qw, qz, sc)
    assert torch.equal(wq, wq2) and torch.equal(sz, sz2)
    assert np.array_equal(ops.dequantize(wq, sz, K, N, G).cpu().numpy().view(np.uint16), d["W16"].view(np.uint16))
    m = WQLinear_QUICK.from_awq_gemm(gq, gz, gs)
    x = torch.from_numpy(qo.make_activations(33, K, seed=5)).cuda()
    assert_close(m(x), x.double() @ torch.from_numpy(d["W16"]).cuda().double(), "from_awq_gemm forward")
    with pytest.raises(ValueError):
        ops.awq_gemm_to_quick(gq, gz[:, :-1], gs)


def test_identity_probe_reads_w16_through_the_tensor_cores(ops):
    """A = rows of the identity: the GEMM output must equal W16 rows exactly (products by 1.0 and
    sums with exact zeros are exact) — the 'bit-exact integer unpack/index' pin through the hot path."""
    K, N, G = 512, 256, 64
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
    assert np.array_equal(W16.cpu().numpy().view(np.uint16), qo.dequant_w16(q, z, s, G).view(np.uint16))
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    for k0 in (0, 64, K - 128):
        A = torch.zeros(128, K, dtype=torch.float16, device="cuda")
        A[torch.arange(128), k0 + torch.arange(128)] = 1
        for tok, split in ((16, 1), (128, 1), (128, 2), (256, 4), (None, None)):
            out = ops.gemm(A, wq, sz, N, G, tok=tok, split=split)
            assert torch.equal(out, W16[k0:k0 + 128]), (k0, tok, split)


CASES = [
    # (K, N, G, Ms)
    (512, 512, 128, (1, 2, 8, 16, 17, 33, 100, 256)),          # BASELINE config 1 shape (M=1) + ragged M
    (256, 768, 64, (1, 7, 64)),
    (128, 512, 32, (3, 16, 130)),
    (4096, 4096, 128, (1, 8, 16, 64, 128, 256, 512)),           # BASELINE config 2: the full sweep
    (4096, 11008, 128, (1, 64)),                                # Llama-2-7B gate/up
    (11008, 4096, 128, (1, 32, 300)),                           # Llama-2-7B down (K = 172 k-blocks, odd splits)
    (8192, 1280, 128, (5,)),                                    # Llama-2-70B qkv shard on 8 ranks
    # full-size shapes of BASELINE configs 4 / 5 and the other group sizes at the headline size
    (14336, 4096, 128, (1, 64)),                                # Mistral-7B down
    (4096, 28672, 128, (1, 16)),                                # Mistral-7B gate|up
    (8192, 10240, 128, (1, 8)),                                 # Llama-2-70B q|k|v
    (28672, 8192, 128, (1, 8)),                                 # Llama-2-70B down
    (8192, 3584, 128, (1, 64, 200)),                            # Llama-2-70B gate (or up) shard on 8 ranks
    (28672, 1024, 128, (1, 33)),                                # Llama-2-70B down shard on 8 ranks
    (4096, 4096, 64, (1, 16, 256)),
    (4096, 4096, 32, (1, 16, 256)),
]


@pytest.mark.parametrize("K,N,G,Ms", CASES, ids=[f"K{c[0]}_N{c[1]}_G{c[2]}" for c in CASES])
def test_gemm_parity_auto_config(ops, K, N, G, Ms):
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    Wd = W16.double()
    for M in Ms:
        A = torch.from_numpy(qo.make_activations(M, K, seed=M)).cuda()
        out = ops.gemm(A, wq, sz, N, G)
        assert out.shape == (M, N) and out.dtype == torch.float16
        assert_close(out, A.double() @ Wd, f"K={K} N={N} G={G} M={M}")


@pytest.mark.parametrize("tok", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("split", [1, 2, 4, 8])
def test_gemm_every_tile_config_matches_oracle_and_simt(ops, tok, split):
    """All (token tile, cluster split-K) instantiations against the numpy oracle (small) and the
    CUDA-core cross-check kernel."""
    if split == 8 and tok > 32:
        pytest.skip("split 8 only instantiated for tok <= 32")
    K, N, G = 1024, 256, 128
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G, seed=tok + split)
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    for M in sorted({1, tok - 1, tok, tok + 3}):
        A_np = qo.make_activations(M, K, seed=M)
        A = torch.from_numpy(A_np).cuda()
        out = ops.gemm(A, wq, sz, N, G, tok=tok, split=split)
        exact = torch.from_numpy(qo.gemm_exact(A_np, W16.cpu().numpy())).cuda()
        assert_close(out, exact, f"tok={tok} split={split} M={M}")
        assert_close(ops.gemm_simt(A, wq, sz, N, G), exact, "simt cross-check")


def test_bias_fused_in_epilogue(ops):
    K, N, G, M = 512, 256, 128, 9
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    A = torch.from_numpy(qo.make_activations(M, K, seed=3)).cuda()
    bias = (torch.randn(N, device="cuda") * 0.5).half()
    for tok, split in ((16, 1), (16, 4), (None, None)):
        out = ops.gemm(A, wq, sz, N, G, bias=bias, tok=tok, split=split)
        assert_close(out, A.double() @ W16.double() + bias.double(), "bias")


def test_drop_in_symbol_matches_reference_contract(ops):
    """quick_kernels.gemm_forward_cuda_quick: positional (x2d, qweight, scales, qzeros, split_k),
    (1,M,N) when split_k == 1 (gemm_cuda_quick.cu:1515-1516), ValueError for the reference's checks."""
    import quick_kernels
    K, N, G, M = 512, 512, 128, 5
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
    A = torch.from_numpy(qo.make_activations(M, K, seed=5)).cuda()
    exact = A.double() @ W16.double()
    o8 = quick_kernels.gemm_forward_cuda_quick(A, qw, sc, qz, 8)
    o1 = quick_kernels.gemm_forward_cuda_quick(A, qw, sc, qz, 1)
    assert o8.shape == (M, N) and o1.shape == (1, M, N)
    assert_close(o8, exact, "drop-in sk=8")
    assert torch.equal(o1[0], o8)
    # stateless C-ABI entry (relayout + GEMM in one call) agrees bit-for-bit with the cached path
    assert torch.equal(ops.gemm_forward_quick_stateless(A, qw, sc, qz, 8), o8)
    with pytest.raises(ValueError, match="cta_N = 128"):
        quick_kernels.gemm_forward_cuda_quick(A, qw[:, :96].contiguous(), sc[:, :384].contiguous(), qz[:, :48].contiguous(), 8)
    with pytest.raises(RuntimeError):   # wrong dtype, like the reference's data_ptr<at::Half>()
        quick_kernels.gemm_forward_cuda_quick(A.float(), qw, sc, qz, 8)


def test_prepack_cache_is_never_stale(ops):
    """The binding caches the B200 relayout per weight storage; freeing / re-allocating / mutating the
    packed tensors must never serve a stale copy."""
    import quick_kernels
    K, N, G, M = 256, 256, 128, 4
    A = torch.from_numpy(qo.make_activations(M, K, seed=1)).cuda()
    for seed in range(6):   # same shapes -> the caching allocator hands back the same addresses
        q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G, seed=seed)
        out = quick_kernels.gemm_forward_cuda_quick(A, qw, sc, qz, 8)
        assert_close(out, A.double() @ W16.double(), f"fresh weight {seed}")
        out2 = quick_kernels.gemm_forward_cuda_quick(A, qw, sc, qz, 8)
        assert torch.equal(out, out2)
        del q, z, s, qw, qz, sc, W16
    # in-place update of a live weight bumps its version counter
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G, seed=100)
    quick_kernels.gemm_forward_cuda_quick(A, qw, sc, qz, 8)
    q2, z2, s2, qw2, qz2, sc2, W16b = make_gpu_case(ops, K, N, G, seed=101)
    qw.copy_(qw2); qz.copy_(qz2); sc.copy_(sc2)
    out = quick_kernels.gemm_forward_cuda_quick(A, qw, sc, qz, 8)
    assert_close(out, A.double() @ W16b.double(), "after in-place update")
    size, hits, misses = quick_kernels.cache_stats()
    assert hits >= 6 and misses >= 8


def test_wqlinear_quick_forward_and_fused_qkv(ops):
    from quick_b200 import layout
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
    from quick_b200.awq.utils.fused_utils import fuse_qkv_quick
    K, G = 512, 128
    torch.manual_seed(0)
    mods, Ws = [], []
    for n in (512, 128, 128):   # GQA-style widths
        lin = torch.nn.Linear(K, n, bias=True).half().cuda()
        q, z, s = layout.quantize_rtn(lin.weight.data.float(), G)
        # AWQ hands from_linear the pseudo-quantised weight (quantizer.py:154-157), i.e. (q - z) * s
        lin.weight.data = ((q - z.repeat_interleave(G, 0)).float() * s.float().repeat_interleave(G, 0)).t().contiguous().half()
        m = WQLinear_QUICK.from_linear(lin, 4, G, False, scales=s.t().contiguous().float(), zeros=z.t().contiguous().float())
        q2, z2, s2 = layout.unpack_quick(m.qweight, m.qzeros, m.scales)
        assert torch.equal(q2, q) and torch.equal(z2, z)
        mods.append(m)
        Ws.append(((q - z.repeat_interleave(G, 0)).half() * s.repeat_interleave(G, 0)))
    x = torch.randn(2, 3, K, device="cuda").half()
    for m, W in zip(mods, Ws):
        y = m(x)
        assert y.shape == (2, 3, m.out_features)
        exact = x.reshape(-1, K).double() @ W.double() + m.bias.double()
        assert_close(y.reshape(-1, m.out_features), exact, "WQLinear_QUICK.forward")
        assert_close(m.forward_reference_call(x).reshape(-1, m.out_features), exact, "reference call sequence")
    fused = fuse_qkv_quick(torch.nn.ModuleList(mods), *mods)
    y = fused(x).reshape(-1, 768)
    exact = x.reshape(-1, K).double() @ torch.cat(Ws, 1).double() + torch.cat([m.bias for m in mods]).double()
    assert_close(y, exact, "fused qkv")


def test_host_buffer_handle_end_to_end(ops):
    K, N, G = 512, 512, 128
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
    h = ops.HostLinear(qw, qz, sc, max_m=64)
    for M in (1, 33, 64):
        x = torch.from_numpy(qo.make_activations(M, K, seed=M)).pin_memory()
        y = h.forward_host(x)
        assert not y.is_cuda
        assert_close(y.cuda(), x.cuda().double() @ W16.double(), f"host handle M={M}")
    with pytest.raises(ValueError):
        h.forward_host(torch.zeros(65, K, dtype=torch.float16))
    h.close()


def test_host_buffer_handle_async_calls_overlap_and_match(ops):
    """qb200_linear_forward_host_async on several handles (= several streams), one synchronize per handle:
    every result must equal the synchronous call's."""
    K, N, G = 512, 768, 128
    cases = [make_gpu_case(ops, K, N, G, seed=100 + i) for i in range(3)]
    hs = [ops.HostLinear(c[3], c[4], c[5], max_m=128) for c in cases]
    xs = {M: torch.from_numpy(qo.make_activations(M, K, seed=M)).pin_memory() for M in (1, 17, 128)}
    ys = {(M, i): torch.empty(M, N, dtype=torch.float16).pin_memory() for M in xs for i in range(3)}
    for M, x in xs.items():
        for i, h in enumerate(hs):
            h.forward_host_async(x, ys[(M, i)])
    for h in hs:
        h.synchronize()
    for M, x in xs.items():
        for i, h in enumerate(hs):
            assert torch.equal(ys[(M, i)], h.forward_host(x)), (M, i)
            assert_close(ys[(M, i)].cuda(), x.cuda().double() @ cases[i][6].double(), f"async host handle M={M} #{i}")
    with pytest.raises(AssertionError):
        hs[0].forward_host_async(torch.zeros(1, K, dtype=torch.float16), ys[(1, 0)])   # unpinned x
    for h in hs:
        h.close()


def test_independent_launches_overlap_but_complete_in_stream_order(ops):
    """QB200_GEMM_INDEPENDENT: consecutive GEMMs (distinct weights and outputs) overlap under programmatic
    dependent launch.  Results must be bit-identical to ordered launches of the same configuration, and an
    ordinary kernel enqueued afterwards must see every result (completion stays transitive)."""
    K = N = 4096
    G = 128
    sets = []
    for i in range(8):
        g = torch.Generator(device="cuda"); g.manual_seed(50 + i)
        wq = torch.randint(-2 ** 31, 2 ** 31 - 1, (K * N // 8,), device="cuda", dtype=torch.int32, generator=g)
        s = (torch.rand(K // G * N, device="cuda", generator=g) * 0.01 + 0.002).half().view(torch.int16).to(torch.int32) & 0xFFFF
        z = torch.randint(0, 16, (K // G * N,), device="cuda", generator=g, dtype=torch.int32)
        sets.append((wq, (s | ((0x6400 + z) << 16)).to(torch.int32)))
    for M in (1, 16, 64, 128, 256, 300):
        x = torch.randn(M, K, device="cuda").half()
        tok, split, _ = ops.plan(M, K, N, G, independent=True)
        assert split == 1
        want = [ops.gemm(x, w, z_, N, G, tok=tok, split=split) for (w, z_) in sets]
        torch.cuda.synchronize()
        outs = [torch.zeros(M, N, device="cuda", dtype=torch.float16) for _ in sets]
        for rep in range(20):
            for o in outs:
                o.zero_()
            for (w, z_), o in zip(sets, outs):
                ops.gemm(x, w, z_, N, G, out=o, independent=True)
            total = torch.stack(outs).float().sum(0)        # ordinary kernel right behind the last GEMM
            torch.cuda.synchronize()
            for o, w_ in zip(outs, want):
                assert torch.equal(o, w_), (M, rep)
            assert torch.equal(total, torch.stack(want).float().sum(0)), (M, rep)
        # and against the oracle definition for one of them
        W16 = ops.dequantize(sets[0][0], sets[0][1], K, N, G)
        assert_close(outs[0], x.double() @ W16.double(), f"independent M={M}")


def test_full_size_properties(ops):
    """At BASELINE's full sizes: size-independent properties instead of a CPU oracle —
    linearity in A, row independence (ragged M == prefix of padded M), and agreement of every
    split-K factor (the cluster reduction) with split 1."""
    K = N = 4096
    G = 128
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    A1 = torch.from_numpy(qo.make_activations(256, K, seed=11)).cuda()
    A2 = torch.from_numpy(qo.make_activations(256, K, seed=12)).cuda()
    y1, y2 = ops.gemm(A1, wq, sz, N, G), ops.gemm(A2, wq, sz, N, G)
    y12 = ops.gemm((A1.float() + A2.float()).half(), wq, sz, N, G)
    assert_close(y12, y1.double() + y2.double(), "linearity")
    # row independence / ragged M
    for M in (1, 100, 255):
        yM = ops.gemm(A1[:M].contiguous(), wq, sz, N, G, tok=256, split=1)
        assert torch.equal(yM, ops.gemm(A1, wq, sz, N, G, tok=256, split=1)[:M])
    # every split factor against split 1 (same tile): differences only from fp32 summation order
    base = ops.gemm(A1, wq, sz, N, G, tok=128, split=1)
    for split in (2, 4):
        assert_close(ops.gemm(A1, wq, sz, N, G, tok=128, split=split), base.double(), f"split {split}")


def test_against_unmodified_reference_kernel(ops):
    """The real reference (oracle/_ref, built from /root/reference/csrc unmodified) on identical packed
    inputs: outputs within 1e-2 rel of each other, W16 bit-exact through identity probes."""
    from oracle.build_ref import load_ref
    ref = load_ref()
    if ref is None:
        pytest.skip("oracle/_ref/quick_kernels_ref.so not built (needs /root/reference at build time)")
    import quick_kernels
    shapes = [(512, 512, 128, 8), (4096, 4096, 128, 8), (4096, 11008, 128, 2), (256, 768, 64, 2),
              (14336, 4096, 128, 8), (4096, 28672, 128, 2), (8192, 3584, 128, 2), (28672, 1024, 128, 8), (4096, 4096, 32, 8)]
    for (K, N, G, sk) in shapes:
        q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
        for M in ((1, 8, 16, 17, 64, 100, 256, 512) if K * N <= 4096 * 11008 else (1, 64, 512)):
            A = torch.from_numpy(qo.make_activations(M, K, seed=M)).cuda()
            r = ref.gemm_forward_cuda_quick(A, qw, sc, qz, sk).reshape(M, N)
            mine = quick_kernels.gemm_forward_cuda_quick(A, qw, sc, qz, sk).reshape(M, N)
            assert_close(mine, r.double(), f"vs reference K={K} N={N} M={M}")
        A = torch.zeros(64, K, dtype=torch.float16, device="cuda")
        A[torch.arange(64), torch.arange(64)] = 1
        r = ref.gemm_forward_cuda_quick(A, qw, sc, qz, 1).reshape(64, N)
        mine = quick_kernels.gemm_forward_cuda_quick(A, qw, sc, qz, 1).reshape(64, N)
        assert torch.equal(mine, r), "identity probe: W16 must be bit-identical to the reference kernel's"


def test_repeated_launches_are_deterministic_and_never_hang(ops):
    """Back-to-back launches with rotating weights (L2 partly warm -> TMA completions out of order):
    regression test for the mbarrier parity-aliasing hang; every result must be bit-identical to the
    first launch on the same operands."""
    K = N = 4096
    G = 128
    sets = []
    for i in range(6):
        g = torch.Generator(device="cuda"); g.manual_seed(i)
        wq = torch.randint(-2 ** 31, 2 ** 31 - 1, (K * N // 8,), device="cuda", dtype=torch.int32, generator=g)
        s = (torch.rand(K // G * N, device="cuda", generator=g) * 0.01 + 0.002).half().view(torch.int16).to(torch.int32) & 0xFFFF
        z = torch.randint(0, 16, (K // G * N,), device="cuda", generator=g, dtype=torch.int32)
        sets.append((wq, (s | ((0x6400 + z) << 16)).to(torch.int32)))
    for (M, tok, split) in ((1, 16, 1), (16, 16, 4), (64, 64, 4), (256, 256, 4)):
        x = torch.randn(M, K, device="cuda").half()
        first = [ops.gemm(x, w, z_, N, G, tok=tok, split=split).clone() for (w, z_) in sets]
        torch.cuda.synchronize()
        for it in range(3000):
            out = ops.gemm(x, *sets[it % 6], N, G, tok=tok, split=split)
            if it % 500 == 499:
                torch.cuda.synchronize()
                assert torch.equal(out, first[it % 6]), (M, tok, split, it)


def test_decoder_glue_kernels_match_the_torch_expressions(ops):
    """RMSNorm, rotary + KV-cache update, SiLU*up and the residual fused into the GEMM epilogue
    against the plain torch expressions of the runner (FUSED_GLUE = False path).  The residual add and the rotary /
    cache update replicate torch's fp16 roundings exactly; RMSNorm and SiLU may differ by one fp16 ulp (reduction
    order, expf)."""
    import quick_kernels
    from quick_b200.awq.models.llama_like import RMSNorm, _rope
    import torch.nn.functional as F
    torch.manual_seed(3)
    # RMSNorm
    for (rows, H) in ((1, 4096), (67, 4096), (5, 512), (3, 11008)):
        x = (torch.randn(rows, H, device="cuda") * 2).half()
        n = RMSNorm(H, 1e-5, "cuda"); n.weight.data = (1 + 0.1 * torch.randn(H, device="cuda")).half()
        got, want = quick_kernels.rmsnorm(x, n.weight, 1e-5), n.forward_torch(x)
        assert torch.allclose(got.float(), want.float(), rtol=2e-3, atol=1e-4), (rows, H)
        assert (got != want).float().mean().item() < 0.02          # almost everywhere bit-identical
    # SiLU * up
    for (rows, I) in ((1, 11008), (33, 1024), (128, 14336)):
        gu = (torch.randn(rows, 2 * I, device="cuda") * 3).half()
        want = F.silu(gu[:, :I]) * gu[:, I:]
        got = quick_kernels.silu_mul(gu)
        assert got.shape == want.shape and torch.allclose(got.float(), want.float(), rtol=2e-3, atol=1e-4)
        assert (got != want).float().mean().item() < 0.01
    # rotary + cache update (GQA and MHA, prefill and single-token decode)
    for (B, T, nh, nkv, hd, S) in ((2, 5, 8, 2, 64, 32), (1, 1, 32, 32, 128, 256), (3, 16, 4, 4, 128, 64)):
        qkv = torch.randn(B, T, (nh + 2 * nkv) * hd, device="cuda").half()
        inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device="cuda").float() / hd))
        ang = torch.outer(torch.arange(S, device="cuda").float(), inv); ang = torch.cat((ang, ang), -1)
        cos_t, sin_t = ang.cos().half(), ang.sin().half()
        pos = torch.arange(3, 3 + T, device="cuda")
        ck, cv = (torch.randn(B, nkv, S, hd, device="cuda").half() for _ in range(2))
        ck2, cv2 = ck.clone(), cv.clone()
        q, k, v = qkv.split([nh * hd, nkv * hd, nkv * hd], dim=-1)
        q = q.view(B, T, nh, hd).transpose(1, 2); k = k.view(B, T, nkv, hd).transpose(1, 2); v = v.view(B, T, nkv, hd).transpose(1, 2)
        cos, sin = cos_t.index_select(0, pos)[None, None], sin_t.index_select(0, pos)[None, None]
        q_want, k_want = _rope(q, cos, sin), _rope(k, cos, sin)
        ck.index_copy_(2, pos, k_want); cv.index_copy_(2, pos, v)
        q_got = quick_kernels.rope_kv_update(qkv, cos_t, sin_t, pos, ck2, cv2, nh, nkv)
        assert torch.equal(q_got, q_want) and torch.equal(ck2, ck) and torch.equal(cv2, cv), (B, T, nh, nkv, hd)
    # residual fused into the GEMM epilogue: every epilogue path (direct store, staged store, split-K, large tiles)
    K, N, G = 1024, 512, 128
    q_, z_, s_, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    bias = torch.randn(N, device="cuda").half()
    for M in (1, 16, 40, 100, 200, 300):
        x = torch.from_numpy(qo.make_activations(M, K, seed=M)).cuda()
        res = torch.randn(M, N, device="cuda").half()
        plain = quick_kernels.gemm_forward_b200(x, wq, sz, bias, N, G)
        fused = quick_kernels.gemm_forward_b200(x, wq, sz, bias, N, G, False, res)
        assert torch.equal(fused, res + plain), M


def test_fused_allgather_stores_on_one_gpu(ops):
    """qb200_gemm_w4a16_allgather with the 'peers' being several local buffers: every destination receives the slab at
    column col0 of ld_c-wide rows (all epilogue paths), the rest of the rows is untouched, the fused residual is read at
    the same columns; qb200_peer_barrier with a single rank returns."""
    K, N, G = 1024, 512, 128
    q_, z_, s_, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G)
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    ld, col0 = 3 * N, N
    for M in (1, 16, 40, 100, 300):
        x = torch.from_numpy(qo.make_activations(M, K, seed=M)).cuda()
        plain = ops.gemm(x, wq, sz, N, G)
        bufs = [torch.full((M, ld), 7.0, device="cuda", dtype=torch.float16) for _ in range(3)]
        ops.gemm_allgather(x, wq, sz, N, G, [b.data_ptr() for b in bufs], ld, col0)
        for b in bufs:
            assert torch.equal(b[:, col0:col0 + N], plain), M
            assert (b[:, :col0] == 7).all() and (b[:, col0 + N:] == 7).all()
        res = torch.randn(M, ld, device="cuda").half()
        out = torch.zeros(M, ld, device="cuda", dtype=torch.float16)
        ops.gemm_allgather(x, wq, sz, N, G, [out.data_ptr()], ld, col0, residual=res)
        assert torch.equal(out[:, col0:col0 + N], res[:, col0:col0 + N] + plain), M
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    epoch = torch.zeros(1, dtype=torch.int32, device="cuda")
    for i in range(3):
        ops.peer_barrier(epoch, [flags.data_ptr()], 0)
    torch.cuda.synchronize()
    assert epoch.item() == 3 and flags[0].item() == 3
    with pytest.raises(ValueError):
        ops.gemm_allgather(x, wq, sz, N, G, [out.data_ptr()], N - 8, 0)      # ld_c < col0 + N


def test_fused_allgather_two_gpus_matches_nccl():
    """2 ranks (torchrun): fused GEMM + all-gather over peer memory == kernel + NCCL all-gather, eager and under
    CUDA-graph replay (tools/tp_check.py).  Skipped on a single-GPU box."""
    import subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", os.path.join(root, "tools", "tp_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "TP_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_llama_like_runner_matches_dense_fp16_model(ops):
    """the minimal runner (all linears through the tcgen05 kernel, CUDA-graph-free here) against
    the same network with every WQLinear_QUICK replaced by its dequantised fp16 weight and torch.matmul."""
    from quick_b200.awq.models.llama_like import PRESETS, LlamaLikeQuickModel
    import copy
    cfg = copy.deepcopy(PRESETS["tiny"])
    model = LlamaLikeQuickModel(cfg, batch=2)
    ids = torch.randint(0, cfg.vocab_size, (2, 24), device="cuda")
    pos = torch.arange(24, device="cuda")
    logits = model(ids, pos)
    # decode one more token on top of the cache
    nxt = model(logits.argmax(-1).view(2, 1), torch.tensor([24], device="cuda"))

    dense = {}
    for blk in model.blocks:
        for m in (blk.qkv_proj, blk.o_proj, blk.gate_up_proj, blk.down_proj):
            wq, sz, K, N, G = ops.prepack(m.qweight, m.qzeros, m.scales)
            dense[id(m)] = ops.dequantize(wq, sz, K, N, G)
    import quick_b200.awq.models.llama_like as ll
    orig = ll._linear
    def dense_linear(m, x, ref_mod=None, residual=None):
        y = (x.reshape(-1, x.shape[-1]).float() @ dense[id(m)].float()).half().reshape(x.shape[:-1] + (m.out_features,))
        return y if residual is None else residual + y

    ll._linear = dense_linear
    fused = ll.FUSED_GLUE
    ll.FUSED_GLUE = False            # the dense reference uses the plain torch expressions for norm / rope / silu / residual
    try:
        for blk in model.blocks:
            blk.cache_k.zero_(); blk.cache_v.zero_()
        ref_logits = model(ids, pos)
        ref_nxt = model(ref_logits.argmax(-1).view(2, 1), torch.tensor([24], device="cuda"))
    finally:
        ll._linear = orig
        ll.FUSED_GLUE = fused
    for a, b in ((logits, ref_logits), (nxt, ref_nxt)):
        rms = b.float().pow(2).mean().sqrt().item()
        assert (a.float() - b.float()).abs().max().item() <= 3e-2 * rms + 1e-3, "runner logits drifted from the dense model"


@pytest.mark.parametrize("nh,nkv,hd", [(8, 4, 64), (32, 32, 128), (32, 8, 128), (64, 8, 128)],
                         ids=["tiny_gqa2_hd64", "mha_7b", "gqa4_mistral", "gqa8_70b"])
def test_attn_decode_matches_rope_cache_update_plus_sdpa(ops, nh, nkv, hd):
    """qb200_attn_decode (rotary + KV-cache update + one-token attention in one kernel) against the path it replaces
    (qb200_rope_kv_update + torch SDPA over the masked static cache) and an fp32 softmax(q·kᵀ/√hd)·v of the same fp16
    operands.  Cache writes must be bit-identical; outputs within 4e-3·rms + 1e-4 of the fp32 result (one fp16 rounding of
    values up to ~4 rms).  Positions
    beyond the current one hold NaN: they must never be read."""
    import quick_kernels
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(nh * 131 + nkv)
    # cache lengths / forced cluster sizes: 96 positions -> the launcher's own choice (2 CTAs per kv head), 600 -> 8 CTAs per
    # kv head, and every cluster size forced on the short cache (CTAs without positions, p = 0 included)
    for S, split in ((96, None), (600, None), (96, 1), (96, 4), (96, 8)):
        _attn_decode_case(nh, nkv, hd, S, split, g)
    # unsupported shapes report so (the runner then keeps rope_kv_update + SDPA): group 16, group 3, head dim 96
    assert not any(quick_kernels.attn_decode_supported(*c) for c in ((32, 2, 128, 96), (24, 8, 128, 96), (32, 8, 96, 96)))
    assert quick_kernels.attn_decode_supported(64, 8, 128, 8192)        # no limit on the cache length
    S = 96
    cos = torch.zeros(S, hd, device="cuda", dtype=torch.float16)
    with pytest.raises(Exception):
        quick_kernels.attn_decode(torch.zeros(1, 2, (nh + 2 * nkv) * hd, device="cuda", dtype=torch.float16), cos, cos,
                                  torch.tensor([0], device="cuda"), torch.zeros(1, nkv, S, hd, device="cuda", dtype=torch.float16),
                                  torch.zeros(1, nkv, S, hd, device="cuda", dtype=torch.float16), nh, nkv)


def _attn_decode_case(nh, nkv, hd, S, split, g):
    import os
    import quick_kernels
    import torch.nn.functional as F
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device="cuda").float() / hd))
    ang = torch.outer(torch.arange(S, device="cuda").float(), inv)
    ang = torch.cat((ang, ang), dim=-1)
    cos, sin = ang.cos().half(), ang.sin().half()
    assert quick_kernels.attn_decode_supported(nh, nkv, hd, S)
    if split is not None:
        os.environ["QB200_ATTN_SPLIT"] = str(split)
    try:
        _attn_decode_positions(nh, nkv, hd, S, g, cos, sin)
    finally:
        os.environ.pop("QB200_ATTN_SPLIT", None)


def _attn_decode_positions(nh, nkv, hd, S, g, cos, sin):
    import quick_kernels
    import torch.nn.functional as F
    for B in (1, 3):
        for p in (0, 1, 37, S - 1):
            qkv = torch.randn(B, 1, (nh + 2 * nkv) * hd, device="cuda", generator=g).half()
            ck = torch.randn(B, nkv, S, hd, device="cuda", generator=g).half()
            cv = torch.randn(B, nkv, S, hd, device="cuda", generator=g).half()
            ck[:, :, p + 1:] = float("nan"); cv[:, :, p + 1:] = float("nan")
            pos = torch.tensor([p], device="cuda")
            ck_ref, cv_ref = ck.clone(), cv.clone()
            q = quick_kernels.rope_kv_update(qkv, cos, sin, pos, ck_ref, cv_ref, nh, nkv)          # [B, nh, 1, hd]
            ck_new, cv_new = ck.clone(), cv.clone()
            out = quick_kernels.attn_decode(qkv, cos, sin, pos, ck_new, cv_new, nh, nkv)
            assert out.shape == (B, 1, nh * hd) and out.dtype == torch.float16
            assert torch.equal(ck_new[:, :, :p + 1], ck_ref[:, :, :p + 1]) and torch.equal(cv_new[:, :, :p + 1], cv_ref[:, :, :p + 1])
            assert torch.isnan(ck_new[:, :, p + 1:]).all() and torch.isnan(cv_new[:, :, p + 1:]).all()   # nothing else written
            kf = ck_ref[:, :, :p + 1].float().repeat_interleave(nh // nkv, dim=1)                   # [B, nh, L, hd]
            vf = cv_ref[:, :, :p + 1].float().repeat_interleave(nh // nkv, dim=1)
            w = torch.softmax((q.float() @ kf.transpose(-1, -2)) / hd ** 0.5, dim=-1)
            exact = (w @ vf).transpose(1, 2).reshape(B, 1, nh * hd)
            rms = exact.pow(2).mean().sqrt().item()
            err = (out.float() - exact).abs().max().item()
            assert not torch.isnan(out).any() and err <= 4e-3 * rms + 1e-4, f"B={B} p={p}: max|err|={err:.3g} rms={rms:.3g}"
            # the path it replaces, on the same cache
            keys = torch.arange(S, device="cuda")
            mask = keys[None, :] <= pos[:, None]
            sd = F.scaled_dot_product_attention(q, torch.nan_to_num(ck_ref), torch.nan_to_num(cv_ref), attn_mask=mask, enable_gqa=(nkv != nh))
            sd = sd.transpose(1, 2).reshape(B, 1, nh * hd)
            assert (out.float() - sd.float()).abs().max().item() <= 1e-2 * rms + 5e-4, f"B={B} p={p}: differs from the SDPA path"


@pytest.mark.parametrize("K,I,G", [(512, 256, 128), (4096, 11008, 128), (1024, 1408 - 128, 64)], ids=["small", "llama7b_gate_up", "g64"])
def test_silu_mul_fused_into_the_gate_up_epilogue_is_bit_identical(ops, K, I, G):
    """SURVEY §8 f4 / reference modules/fused/mlp.py:52-76: silu(gate(x)) * up(x) computed in the epilogue of the fused gate|up
    GEMM (QB200_GEMM_SILU_MUL, interleaved output channels) == the unfused GEMM followed by qb200_silu_mul, bit for bit, for
    every tile configuration the planner picks and for forced ones; also through WQLinear_QUICK.forward_silu_mul."""
    import quick_kernels
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
    N = 2 * I
    q, z, s, qw, qz, sc, W16 = make_gpu_case(ops, K, N, G, seed=I)
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    bias = (torch.randn(N, device="cuda") * 0.1).half()
    wq_i, sz_i, bias_i = ops.interleave_pairs(wq, sz, K, N, G, bias)
    # the interleaved weight is a column permutation of the original one
    Wd = ops.dequantize(wq_i, sz_i, K, N, G)
    assert torch.equal(Wd[:, 0::2], W16[:, :I]) and torch.equal(Wd[:, 1::2], W16[:, I:])
    mod = WQLinear_QUICK(4, G, K, N, True, "cuda")
    mod.qweight, mod.qzeros, mod.scales, mod.bias = qw, qz, sc, bias
    for M in (1, 8, 16, 33, 64, 100, 130, 300, 512):
        A = torch.from_numpy(qo.make_activations(M, K, seed=M)).cuda()
        for use_bias in (False, True):
            b, bi = (bias, bias_i) if use_bias else (None, None)
            want = quick_kernels.silu_mul(ops.gemm(A, wq, sz, N, G, bias=b))
            got = ops.gemm_tp(A, wq_i, sz_i, N, G, bias=bi, silu_mul=True)
            assert got.shape == (M, I) and torch.equal(got, want), (M, use_bias)
        assert torch.equal(mod.enable_silu_mul().forward_silu_mul(A), quick_kernels.silu_mul(ops.gemm(A, wq, sz, N, G, bias=bias)))
    with pytest.raises(RuntimeError, match="forward_silu_mul"):
        mod(A)
    # against the fp64 definition as well (not only against our own unfused kernels)
    A = torch.from_numpy(qo.make_activations(64, K, seed=3)).cuda()
    y = A.double() @ W16.double()
    g, u = y[:, :I].half().double(), y[:, I:].half().double()
    assert_close(ops.gemm_tp(A, wq_i, sz_i, N, G, silu_mul=True), (g / (1 + torch.exp(-g))) * u, "fused SwiGLU vs fp64")


@pytest.mark.parametrize("H,I,G", [(512, 384, 128), (4096, 11008, 128), (1024, 640, 64)], ids=["small", "llama7b", "g64"])
def test_rmsnorm_folded_around_the_gemms(ops, H, I, G):
    """SURVEY §8 f4 / reference modules/fused/block.py:61-74 + norm.py:16-19 (x + o_proj(...) -> RMSNorm -> gate|up):
    qb200_gemm_w4a16_norm — the producing GEMM (residual fused) also emits h * gamma and per-tile sums of squares, the
    consuming GEMM scales its rows by 1/rms — against the unfused sequence GEMM -> qb200_rmsnorm -> GEMM (+ SiLU·up) and
    against the fp64 definition.  h must be bit-identical, h * gamma must be torch's fp16 product, the sums of squares the
    fp32 sums of h² per 128-column tile, and the consumer's output within fp16 rounding of both references, for every
    tile configuration the planner picks (direct stores, staged tiles, one and several token tiles)."""
    import quick_kernels
    from quick_b200.awq.modules.linear.quick import WQLinear_QUICK
    eps = 1e-5
    gen = torch.Generator(device="cuda").manual_seed(H + I)
    _, _, _, qw1, qz1, sc1, Wp = make_gpu_case(ops, H, H, G, seed=11)             # producer: an o_proj-like H -> H linear
    _, _, _, qw2, qz2, sc2, Wc = make_gpu_case(ops, H, 2 * I, G, seed=12)         # consumer: gate|up, H -> 2 I
    prod, cons, cons_silu = (WQLinear_QUICK(4, G, H, n, False, "cuda") for n in (H, 2 * I, 2 * I))
    prod.qweight, prod.qzeros, prod.scales = qw1, qz1, sc1
    for m in (cons, cons_silu):
        m.qweight, m.qzeros, m.scales = qw2, qz2, sc2
    cons_silu.enable_silu_mul()
    gamma = (1.0 + 0.2 * torch.randn(H, device="cuda", generator=gen)).half()
    for M in (1, 3, 8, 16, 17, 64, 100, 300):
        a = torch.from_numpy(qo.make_activations(M, H, seed=M)).cuda()
        res = torch.randn(M, H, device="cuda", generator=gen).half()
        h_ref = prod(a, res)
        h, hg, ssq = prod.forward_norm_out(a, res, gamma)
        assert torch.equal(h, h_ref), M
        assert torch.equal(hg, h_ref * gamma), M
        assert ssq.shape == (H // 128, M) and ssq.dtype == torch.float32
        want_ssq = h_ref.float().pow(2).view(M, H // 128, 128).sum(-1).t()
        assert torch.allclose(ssq, want_ssq, rtol=1e-5, atol=0), M
        # the unfused sequence and the fp64 definition of norm -> linear
        xn = quick_kernels.rmsnorm(h_ref, gamma, eps)
        hd = h_ref.double()
        xn64 = hd * torch.rsqrt(hd.pow(2).mean(-1, keepdim=True) + eps) * gamma.double()
        y_unfused, y64 = cons(xn), xn64 @ Wc.double()
        y = cons.forward_normed(hg, ssq, eps)
        assert_close(y, y64, f"M={M} fused norm -> linear vs fp64")
        rms = y64.pow(2).mean().sqrt().item()
        # against the unfused kernels: the fp16 roundings of the normed inputs sit in different places (h * gamma vs
        # h * rstd, then * gamma) and the outputs are rounded independently -> half a percent, relative + rms floor
        diff = (y.double() - y_unfused.double()).abs()
        assert bool((diff <= 5e-3 * y_unfused.double().abs() + 5e-3 * rms).all()), f"M={M}: differs from rmsnorm kernel + GEMM by {diff.max().item():.3g}"
        # consumer with SiLU·up in the epilogue
        g64, u64 = y64[:, :I].half().double(), y64[:, I:].half().double()
        act = cons_silu.forward_silu_mul(hg, ssq, eps)
        assert_close(act, (g64 / (1 + torch.exp(-g64))) * u64, f"M={M} fused norm -> gate|up -> SiLU·up vs fp64")
        assert torch.equal(act, quick_kernels.silu_mul(y)), M      # same rows, same rounding as the unfused SiLU kernel
        # determinism (fixed summation order everywhere)
        h2, hg2, ssq2 = prod.forward_norm_out(a, res, gamma)
        assert torch.equal(ssq2, ssq) and torch.equal(cons.forward_normed(hg2, ssq2, eps), y)
    # argument checks: producer side refuses SiLU outputs
    with pytest.raises(Exception):
        quick_kernels.gemm_forward_b200_norm(a, *cons_silu._b200, None, 2 * I, G, None, True, gamma, None, eps)


def test_attn_decode_tp_stores_the_heads_into_every_peer_buffer(ops):
    """qb200_attn_decode_tp (tensor parallel: this rank's heads, output stored straight into every rank's attention
    buffer) with local buffers standing in for the peers: every destination receives exactly what qb200_attn_decode
    returns at column col0 of its ld-wide rows, nothing else is written, the epoch of the gathered buffer advances and
    the KV-cache update is the same."""
    import ctypes as C
    import quick_kernels
    from quick_b200 import _lib
    nh, nkv, hd, S, p = 8, 4, 128, 160, 77
    gen = torch.Generator(device="cuda").manual_seed(5)
    ang = torch.outer(torch.arange(S, device="cuda").float(), 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device="cuda").float() / hd)))
    ang = torch.cat((ang, ang), dim=-1)
    cos, sin = ang.cos().half(), ang.sin().half()

    class FakeGathered:          # the fields ops.attn_decode_tp reads of a parallel.GatheredBuffer
        def __init__(self, bufs, width):
            self.world, self.width = len(bufs), width
            self.buf_ptrs = (C.c_void_p * len(bufs))(*[b.data_ptr() for b in bufs])
            self.state = torch.zeros(1, dtype=torch.int32, device="cuda")
            self.signal = _lib.PeerSignal(self.state.data_ptr())

    for B in (1, 5):
        qkv = torch.randn(B, 1, (nh + 2 * nkv) * hd, device="cuda", generator=gen).half()
        ck = torch.randn(B, nkv, S, hd, device="cuda", generator=gen).half()
        cv = torch.randn(B, nkv, S, hd, device="cuda", generator=gen).half()
        pos = torch.tensor([p], device="cuda")
        ck1, cv1, ck2, cv2 = ck.clone(), cv.clone(), ck.clone(), cv.clone()
        want = quick_kernels.attn_decode(qkv, cos, sin, pos, ck1, cv1, nh, nkv).view(B, nh * hd)
        ld, col0 = 3 * nh * hd, nh * hd
        bufs = [torch.full((B, ld), 3.0, device="cuda", dtype=torch.float16) for _ in range(3)]
        dst = FakeGathered(bufs, ld)
        ops.attn_decode_tp(qkv, cos, sin, pos, ck2, cv2, nh, nkv, dst, col0)
        torch.cuda.synchronize()
        for b in bufs:
            assert torch.equal(b[:, col0:col0 + nh * hd], want), B
            assert (b[:, :col0] == 3).all() and (b[:, col0 + nh * hd:] == 3).all()
        assert torch.equal(ck1, ck2) and torch.equal(cv1, cv2) and dst.state.item() == 1
