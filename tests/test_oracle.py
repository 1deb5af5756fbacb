"""CPU: the oracle against every golden vector produced by the reference's own python code
(tests/golden/make_golden.py), plus its internal consistency.  Bit-exact everywhere: the data are
integers, nibbles and fp16 bit patterns."""
import glob
import os

import numpy as np
import pytest

from oracle import quick_oracle as qo

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PACK_FILES = sorted(glob.glob(os.path.join(GOLD, "pack_*.npz")))


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint16)


def test_golden_files_present():
    assert len(PACK_FILES) >= 6
    assert os.path.exists(os.path.join(GOLD, "cat_K256_3xN256_G128.npz"))


@pytest.mark.parametrize("path", PACK_FILES, ids=os.path.basename)
def test_pack_matches_reference_packer(path):
    d = np.load(path)
    q, z, s = d["q"].astype(np.int32), d["z"].astype(np.int32), d["s"]
    qw, qz, sc = qo.pack_quick(q, z, s)
    assert qw.dtype == np.int32 and qw.shape == d["qweight"].shape
    assert np.array_equal(qw, d["qweight"])
    assert np.array_equal(qz, d["qzeros"])
    assert np.array_equal(_bits(sc), _bits(d["scales"]))


@pytest.mark.parametrize("path", PACK_FILES, ids=os.path.basename)
def test_unpack_inverts_reference_packer(path):
    d = np.load(path)
    q, z, s = qo.unpack_quick(d["qweight"], d["qzeros"], d["scales"])
    assert np.array_equal(q, d["q"].astype(np.int32))
    assert np.array_equal(z, d["z"].astype(np.int32))
    assert np.array_equal(_bits(s), _bits(d["s"]))


@pytest.mark.parametrize("path", PACK_FILES, ids=os.path.basename)
def test_kernel_view_equals_dequant(path):
    """W16 derived from the kernel's pointer math + bit-level lop3/fma emulation == fp16(q-z)*s."""
    d = np.load(path)
    G = int(d["G"])
    W_kernel = qo.kernel_view_w16(d["qweight"], d["qzeros"], d["scales"], G)
    W_plain = qo.dequant_w16(d["q"].astype(np.int32), d["z"].astype(np.int32), d["s"], G)
    assert np.array_equal(_bits(W_kernel), _bits(W_plain))


def test_s4_to_fp16_magic_numbers():
    w = np.array([0x76543210, 0xFEDCBA98, 0x00000000, 0xFFFFFFFF], dtype=np.uint32)
    out = qo.s4_to_fp16x2_fused(w).astype(np.float32) - 1024.0
    # register j = (nibble j, nibble j+4)
    assert out[0].tolist() == [0, 4, 1, 5, 2, 6, 3, 7]
    assert out[1].tolist() == [8, 12, 9, 13, 10, 14, 11, 15]
    assert out[2].tolist() == [0] * 8 and out[3].tolist() == [15] * 8


def test_quick_cat_golden():
    d = np.load(os.path.join(GOLD, "cat_K256_3xN256_G128.npz"))
    for name in ("qweight", "qzeros", "scales"):
        c = qo.quick_cat([d[f"{name}{i}"] for i in range(3)], name)
        assert np.array_equal(np.ascontiguousarray(c).view(np.uint8), np.ascontiguousarray(d["cat_" + name]).view(np.uint8))
    # concatenating packed tensors == packing the concatenated logical tensors
    q = np.concatenate([d[f"q{i}"] for i in range(3)], 1).astype(np.int32)
    z = np.concatenate([d[f"z{i}"] for i in range(3)], 1).astype(np.int32)
    s = np.concatenate([d[f"s{i}"] for i in range(3)], 1)
    qw, qz, sc = qo.pack_quick(q, z, s)
    assert np.array_equal(qw, d["cat_qweight"]) and np.array_equal(qz, d["cat_qzeros"])
    assert np.array_equal(_bits(sc), _bits(d["cat_scales"]))
    for r in range(3):
        a, b, c = qo.shard_columns(d["cat_qweight"], d["cat_qzeros"], d["cat_scales"], r, 3)
        assert np.array_equal(a, d[f"qweight{r}"]) and np.array_equal(b, d[f"qzeros{r}"])
        assert np.array_equal(_bits(c), _bits(d[f"scales{r}"]))


def test_quick_cat_unequal_widths():
    """GQA-style concat (q wide, k/v narrow) — rejected by the reference, exact in the layout algebra."""
    K, G = 128, 64
    parts = [qo.make_case(K, n, G, seed=7 + i) for i, n in enumerate((256, 128, 128))]
    packed = [qo.pack_quick(*p) for p in parts]
    cat = [qo.quick_cat([p[i] for p in packed], name) for i, name in enumerate(("qweight", "qzeros", "scales"))]
    whole = qo.pack_quick(*[np.concatenate([p[i] for p in parts], 1) for i in range(3)])
    assert np.array_equal(cat[0], whole[0]) and np.array_equal(cat[1], whole[1])
    assert np.array_equal(_bits(cat[2]), _bits(whole[2]))


def test_config1_m1_k512_n512_cpu():
    """BASELINE.json configs[0]: M=1 K=512 N=512 g=128 vs dequant + matmul on CPU."""
    K = N = 512
    G = 128
    q, z, s = qo.make_case(K, N, G)
    qw, qz, sc = qo.pack_quick(q, z, s)
    A = qo.make_activations(1, K, seed=1)
    W16 = qo.dequant_w16(q, z, s, G)
    out = qo.forward_oracle(A, qw, qz, sc)
    exact = qo.gemm_exact(A, W16)
    rms = float(np.sqrt(np.mean(exact ** 2)))
    assert out.shape == (1, N) and out.dtype == np.float16
    np.testing.assert_allclose(out.astype(np.float64), exact, rtol=1e-2, atol=1e-2 * rms)
    # the reference's split-K rounding chain stays within the same tolerance of the exact product
    sk = qo.gemm_oracle_splitk(A, W16, 8)
    np.testing.assert_allclose(sk.astype(np.float64), exact, rtol=1e-2, atol=1e-2 * rms)


def test_reference_arg_checks_and_shape_quirk():
    with pytest.raises(ValueError, match="cta_N = 128"):
        qo.check_args(512, 192, 128)
    with pytest.raises(ValueError, match="multiple of 32"):
        qo.check_args(512, 256, 48)
    qo.check_args(512, 256, 64)
    assert qo.reference_output_shape(5, 256, 1) == (1, 5, 256)
    assert qo.reference_output_shape(5, 256, 8) == (5, 256)


def test_linearity_property():
    """Size-independent property used at full size on the GPU: GEMM is linear in A."""
    K, N, G = 256, 128, 128
    q, z, s = qo.make_case(K, N, G)
    W16 = qo.dequant_w16(q, z, s, G)
    A1, A2 = qo.make_activations(4, K, 1), qo.make_activations(4, K, 2)
    lhs = qo.gemm_exact((A1.astype(np.float32) + A2.astype(np.float32)).astype(np.float16), W16)
    rhs = qo.gemm_exact(A1, W16) + qo.gemm_exact(A2, W16)
    # A1 + A2 is rounded to fp16 once, so compare against that rounding explicitly
    assert np.max(np.abs(lhs - rhs)) <= 2e-2 * np.sqrt(np.mean(rhs ** 2))


# ---- AWQ "GEMM" checkpoint layout: golden vectors from the reference's packing_utils.py ----
AWQ_FILES = sorted(glob.glob(os.path.join(GOLD, "awqgemm_*.npz")))


def test_awq_gemm_golden_files_present():
    assert len(AWQ_FILES) >= 3


@pytest.mark.parametrize("path", AWQ_FILES, ids=os.path.basename)
def test_awq_gemm_unpack_matches_reference_unpacker(path):
    d = np.load(path)
    q, z = qo.unpack_awq_gemm(d["qweight"], d["qzeros"])
    assert np.array_equal(q, d["q"]) and np.array_equal(z, d["z"])
    # the packer is its inverse on arbitrary 32-bit words (every nibble is used)
    qw, qz = qo.pack_awq_gemm(q, z)
    assert np.array_equal(qw, d["qweight"]) and np.array_equal(qz, d["qzeros"])
    # and the oracle's W16 definition is the reference's CPU dequantize_gemm, bit for bit
    W16 = qo.dequant_w16(q.astype(np.int32), z.astype(np.int32), d["scales"], int(d["G"]))
    assert np.array_equal(_bits(W16), _bits(d["W16"]))


@pytest.mark.parametrize("path", AWQ_FILES, ids=os.path.basename)
def test_awq_gemm_to_quick_host_converter(path):
    """quick_b200.layout.awq_gemm_to_quick (torch, CPU): QUICK tensors that unpack to the reference's integers."""
    import torch
    from quick_b200 import layout
    d = np.load(path)
    tq, tz, ts = (torch.from_numpy(d[k]) for k in ("qweight", "qzeros", "scales"))
    q, z = layout.unpack_awq_gemm(tq, tz)
    assert np.array_equal(q.numpy(), d["q"]) and np.array_equal(z.numpy(), d["z"])
    pq, pz = layout.pack_awq_gemm(q, z)
    assert torch.equal(pq, tq) and torch.equal(pz, tz)
    qw, qz, sc = layout.awq_gemm_to_quick(tq, tz, ts)
    want = qo.pack_quick(d["q"].astype(np.int32), d["z"].astype(np.int32), d["scales"])
    assert np.array_equal(qw.numpy(), want[0]) and np.array_equal(qz.numpy(), want[1])
    assert np.array_equal(_bits(sc.numpy()), _bits(want[2]))
