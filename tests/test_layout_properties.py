"""CPU: size-independent properties of the packed-operand algebra over randomly drawn shapes (hypothesis) — the host
restatement (quick_b200.layout) against the oracle (oracle/quick_oracle.py) and against itself: pack ↔ unpack is a
bijection on the de-duplicated data, AWQ-GEMM → QUICK is the same permutation whichever way it is reached, column
shards re-concatenate to the whole, and concatenation commutes with packing."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import quick_oracle as qo
from quick_b200 import layout

shapes = st.tuples(st.integers(1, 6), st.integers(1, 5), st.sampled_from([32, 64, 128]), st.integers(0, 2 ** 16)).map(
    lambda t: (max(t[2], 64) * t[0], 128 * t[1], t[2], t[3]))        # K multiple of max(G, 64), N multiple of 128


def _case(K, N, G, seed):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, size=(K, N), dtype=np.int32)
    z = rng.integers(0, 16, size=(K // G, N), dtype=np.int32)
    s = (0.001 + 0.02 * rng.random((K // G, N))).astype(np.float16)
    return q, z, s


@settings(max_examples=25, deadline=None)
@given(shapes)
def test_pack_unpack_bijection_and_oracle_agreement(shape):
    K, N, G, seed = shape
    q, z, s = _case(K, N, G, seed)
    tq, tz, ts = torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s)
    qw, qz, sc = layout.pack_quick(tq, tz, ts)
    assert qw.shape == (K // 4, N // 2) and qz.shape == (K // G, N // 4) and sc.shape == (K // G, 2 * N)
    oq, oz, os_ = qo.pack_quick(q, z, s)
    assert np.array_equal(qw.numpy(), oq) and np.array_equal(qz.numpy(), oz)
    assert np.array_equal(sc.numpy().view(np.uint16), os_.view(np.uint16))
    q2, z2, s2 = layout.unpack_quick(qw, qz, sc)
    assert torch.equal(q2, tq) and torch.equal(z2, tz) and torch.equal(s2, ts)
    # every zero / scale is stored exactly twice (quick.py:129-130,141-150)
    assert torch.equal(sc[:, 0::2], sc[:, 1::2])
    assert torch.equal(qz & 0xFFFF, (qz >> 16) & 0xFFFF)
    # the kernel's own pointer math reads back W16 = fp16(q - z) * s from the packed tensors
    if K * N <= 256 * 256:
        w16 = qo.kernel_view_w16(oq, oz, os_, G)
        assert np.array_equal(w16.view(np.uint16), qo.dequant_w16(q, z, s, G).view(np.uint16))


@settings(max_examples=15, deadline=None)
@given(shapes)
def test_awq_gemm_layout_converts_to_the_same_quick_tensors(shape):
    K, N, G, seed = shape
    q, z, s = _case(K, N, G, seed)
    gq, gz = qo.pack_awq_gemm(q, z)
    uq, uz = qo.unpack_awq_gemm(gq, gz)
    assert np.array_equal(uq, q) and np.array_equal(uz, z)
    conv = layout.awq_gemm_to_quick(torch.from_numpy(gq), torch.from_numpy(gz), torch.from_numpy(s))
    direct = layout.pack_quick(torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s))
    for a, b in zip(conv, direct):
        assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))


@settings(max_examples=15, deadline=None)
@given(shapes, st.integers(1, 5))
def test_shards_and_concatenation_commute_with_packing(shape, world):
    K, N, G, seed = shape
    tiles = N // 128
    world = max(w for w in range(1, world + 1) if tiles % w == 0)
    q, z, s = _case(K, N, G, seed)
    whole = layout.pack_quick(torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s))
    n = N // world
    shards = [layout.shard_columns(*whole, r, world) for r in range(world)]
    for r, sh in enumerate(shards):        # a shard is the packing of the corresponding logical columns
        want = layout.pack_quick(torch.from_numpy(q[:, r * n:(r + 1) * n].copy()), torch.from_numpy(z[:, r * n:(r + 1) * n].copy()),
                                 torch.from_numpy(s[:, r * n:(r + 1) * n].copy()))
        for a, b in zip(sh, want):
            assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    if world > 1:                          # ... and the shards concatenate back to the whole (the all-gather identity)
        for i, name in enumerate(("qweight", "qzeros", "scales")):
            assert torch.equal(layout.quick_cat([sh[i] for sh in shards], name).view(torch.uint8), whole[i].view(torch.uint8))
