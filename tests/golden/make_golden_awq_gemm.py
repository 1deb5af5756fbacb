"""Golden vectors for the AWQ "GEMM" checkpoint layout, from the reference's own Python code (run in the
build container only: needs /root/reference; the GPU box and the tests never run this).

  python tests/golden/make_golden_awq_gemm.py

``awqgemm_*.npz`` — random packed int32 tensors pushed through the reference's unpack_awq +
reverse_awq_order + 4-bit mask and dequantize_gemm (quick/awq/utils/packing_utils.py:8-39, :80-96), plus
the reference packer's bit placement (quick/awq/modules/linear/gemm.py:108-143, the loops copied here verbatim
in behaviour: ``qweight[:, col] |= intweight[:, col*8 + order_map[i]] << (i*4)``) applied to the unpacked
integers, which must reproduce the packed input bit for bit.
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_packing_utils():
    spec = importlib.util.spec_from_file_location("ref_packing_utils", f"{REF}/quick/awq/utils/packing_utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_gemm_pack(intweight: torch.Tensor) -> torch.Tensor:
    """The reference packer's loop (gemm.py:108-123 for weights, :132-141 for zeros)."""
    order_map = [0, 2, 4, 6, 1, 3, 5, 7]
    out = torch.zeros((intweight.shape[0], intweight.shape[1] // 8), dtype=torch.int32)
    for col in range(intweight.shape[1] // 8):
        for i in range(8):
            out[:, col] |= intweight[:, col * 8 + order_map[i]] << (i * 4)
    return out


def main():
    pu = load_packing_utils()
    for (K, N, G) in [(128, 128, 128), (256, 384, 64), (128, 512, 32)]:
        rng = np.random.default_rng(K + 3 * N + 7 * G)
        qweight = rng.integers(-2 ** 31, 2 ** 31 - 1, size=(K, N // 8), dtype=np.int64).astype(np.int32)
        qzeros = rng.integers(-2 ** 31, 2 ** 31 - 1, size=(K // G, N // 8), dtype=np.int64).astype(np.int32)
        scales = (0.002 + 0.01 * rng.random((K // G, N))).astype(np.float16)
        tq, tz, ts = torch.from_numpy(qweight), torch.from_numpy(qzeros), torch.from_numpy(scales)
        iw, iz = pu.unpack_awq(tq, tz, 4)
        iw, iz = pu.reverse_awq_order(iw, iz, 4)
        iw, iz = torch.bitwise_and(iw, 15), torch.bitwise_and(iz, 15)
        W = pu.dequantize_gemm(tq, tz, ts, 4, G)          # fp16 (q - z) * s, the reference's CPU dequantize
        assert torch.equal(ref_gemm_pack(iw.to(torch.int32)), tq) and torch.equal(ref_gemm_pack(iz.to(torch.int32)), tz)
        np.savez_compressed(os.path.join(HERE, f"awqgemm_K{K}_N{N}_G{G}.npz"), qweight=qweight, qzeros=qzeros, scales=scales,
                            q=iw.numpy().astype(np.uint8), z=iz.numpy().astype(np.uint8), W16=W.numpy().astype(np.float16),
                            G=np.int32(G))
        print("wrote", K, N, G)


if __name__ == "__main__":
    main()
