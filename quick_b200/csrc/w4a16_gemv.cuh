// quick_b200 — W4A16 GEMV for decode-sized inputs (M <= 4 token rows) on sm_100a.
//
// Below ~8 tokens the W4A16 product is a pure weight stream: K·N/2 bytes must cross HBM once and the arithmetic
// (2·M·K·N flops) is far under what the CUDA cores do in that time, so this path does not go through the tensor
// cores at all (reference: the M = 1 kernel gemm_forward_4bit_cuda_quick_m1n128k32, gemm_cuda_quick.cu:1199-1242,
// runs mma.sync with 15/16 of the rows wasted and every k-step as a dependent LDG → dequant → HMMA → 2 barriers).
//
// What bounds a chain of such GEMVs in stream order is not bandwidth but the hand-over between dependent kernels
// (profiles/README.md: ≈1.1 µs from the last CTA's exit to the successor's griddepcontrol.wait returning).  The
// weights do not depend on the predecessor — only the activations do — so every CTA
//   1. signals griddepcontrol.launch_dependents at once,
//   2. issues cp.async copies of its WHOLE weight slice (up to 64 KB of nibbles + the scale/zero words) into a
//      per-thread ring in shared memory — while the predecessor is still running,
//   3. only then executes griddepcontrol.wait, stages the M activation rows in shared memory,
//   4. dequantises from shared memory and accumulates.
// In steady state the HBM stream of GEMV i+1 overlaps the arithmetic and the exit/hand-over of GEMV i.
//
// Work split: one CTA = 32 output channels (one warp width: a 512-byte contiguous run per k32 half in the B200
// layout wq[N/128][K/64][2][128][4]) × the full K.  No split-K across CTAs, no cluster, no exchange: 16 warps take the
// k32 halves round-robin (lane = channel) and meet once in shared memory.  N = 4096 → 128 CTAs of 64 KB.
// Each thread only ever reads ring slots it filled itself, so ring reuse (K > 4096: more than 8 halves per warp)
// needs no barrier: cp.async.wait_group orders a thread's own copies, and a slot is refilled after its LDS.
//
// Arithmetic = the reference's: q − z exact in fp16, one mul.rn by the scale (W16 bit-identical),
// products accumulated by HFMA2 in chains of 8 terms per lane, chains summed in fp32; bias added in fp32 before
// the single rounding of the output; residual added in fp16 (like torch's x + linear(x)).
#pragma once
#include "w4a16_umma.cuh"

namespace qb200 {

constexpr int kGemvCh = 32;          // output channels per CTA
constexpr int kGemvWarps = 16;
constexpr int kGemvThreads = kGemvWarps * 32;
constexpr int kGemvRing = 8;         // per-thread ring slots: 16 B of nibbles (32 k of one channel) + 4 B scale/zero
constexpr int kGemvGroup = 4;        // slots per cp.async commit group
constexpr int kGemvMaxRows = 4;
constexpr int kGemvRingBytes = kGemvRing * kGemvThreads * (16 + 4);   // 80 KB

__host__ __device__ constexpr int gemv_smem_bytes(int mrows, int K) {
  return kGemvRingBytes + mrows * K * 2 + kGemvWarps * mrows * kGemvCh * 4;
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int MROWS>
__global__ void __launch_bounds__(kGemvThreads, 2) w4a16_gemv_kernel(const __half* __restrict__ A, const GemmArgs args) {
  extern __shared__ __align__(16) uint8_t gemv_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool independent = (args.flags & kFlagIndependent) != 0;
  const int K = args.K;
  const int H = K >> 5;                        // k32 halves
  const int hpg_shift = __ffs(args.G >> 5) - 1;   // halves per quantisation group = G / 32, a power of two (host check)
  const int ch0 = blockIdx.x * kGemvCh;
  const int nt = ch0 >> 7;
  const int c = (ch0 & 127) + lane;            // channel inside the 128-channel tile
  const uint32_t* wbase = args.wq + static_cast<size_t>(nt) * (K >> 6) * 1024 + c * 4;    // + h * 512 words
  const uint32_t* szbase = args.sz + static_cast<size_t>(nt) * (K / args.G) * 128 + c;   // + group * 128 words
  const int nh = (H - warp + kGemvWarps - 1) / kGemvWarps;      // halves of this warp: h = warp + 16 i
  const int ngroups = (nh + kGemvGroup - 1) / kGemvGroup;

  const uint32_t w_s = smem_u32(gemv_smem);
  const uint32_t sz_s = w_s + kGemvRing * kGemvThreads * 16;
  __half* xs = reinterpret_cast<__half*>(gemv_smem + kGemvRingBytes);
  float* red = reinterpret_cast<float*>(gemv_smem + kGemvRingBytes + MROWS * K * 2);

  auto issue = [&](int grp) {      // always commits, also when nothing is left (keeps the group count uniform)
#pragma unroll
    for (int j = 0; j < kGemvGroup; ++j) {
      const int i = grp * kGemvGroup + j;
      if (i < nh) {
        const int h = warp + kGemvWarps * i;
        const int slot = i % kGemvRing;
        cp_async_16(w_s + (slot * kGemvThreads + tid) * 16, wbase + static_cast<size_t>(h) * 512);
        cp_async_4(sz_s + (slot * kGemvThreads + tid) * 4, szbase + static_cast<size_t>(h >> hpg_shift) * 128);
      }
    }
    cp_async_commit();
  };

  pdl_launch_dependents();
  issue(0);                      // the weight stream starts while the previous kernel of the stream is still running
  issue(1);
  if (!independent) pdl_wait_prior_grid();     // the activations (and C) belong to the previous kernel until here

  {   // activation rows -> shared memory (rows beyond M are zero)
    const int vec_per_row = K >> 3;
#pragma unroll
    for (int m = 0; m < MROWS; ++m) {
      for (int kv = tid; kv < vec_per_row; kv += kGemvThreads) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (m < args.M) v = *reinterpret_cast<const uint4*>(A + static_cast<size_t>(m) * K + kv * 8);
        *reinterpret_cast<uint4*>(xs + static_cast<size_t>(m) * K + kv * 8) = v;
      }
    }
  }
  __syncthreads();

  float acc[MROWS];
#pragma unroll
  for (int m = 0; m < MROWS; ++m) acc[m] = 0.f;

  for (int grp = 0; grp < ngroups; ++grp) {
    cp_async_wait<1>();          // every group but the most recently committed one has landed -> group grp is in
#pragma unroll
    for (int j = 0; j < kGemvGroup; ++j) {
      const int i = grp * kGemvGroup + j;
      if (i < nh) {              // warp-uniform
        const int h = warp + kGemvWarps * i;
        const int slot = i % kGemvRing;
        const uint4 wv = lds128(w_s + (slot * kGemvThreads + tid) * 16);
        const GroupConsts g = make_group_consts(lds_u32(sz_s + (slot * kGemvThreads + tid) * 4));
        uint32_t o[16];          // 32 consecutive k of this channel as fp16 pairs, W16 = fp16((q − z)·s)
        dequant_word(wv.x, g, o);
        dequant_word(wv.y, g, o + 4);
        dequant_word(wv.z, g, o + 8);
        dequant_word(wv.w, g, o + 12);
#pragma unroll
        for (int m = 0; m < MROWS; ++m) {
          const uint32_t xa = smem_u32(xs + static_cast<size_t>(m) * K + h * 32);
          uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 xv = lds128(xa + q * 16);       // same address in every lane: one broadcast wavefront
            a0 = hfma2(o[4 * q + 0], xv.x, a0);
            a1 = hfma2(o[4 * q + 1], xv.y, a1);
            a0 = hfma2(o[4 * q + 2], xv.z, a0);
            a1 = hfma2(o[4 * q + 3], xv.w, a1);
          }
          const float2 f0 = unpack_half2(a0), f1 = unpack_half2(a1);
          acc[m] += (f0.x + f0.y) + (f1.x + f1.y);
        }
      }
    }
    issue(grp + 2);              // refill the slots just read (own slots only: no barrier needed)
  }

#pragma unroll
  for (int m = 0; m < MROWS; ++m) red[(warp * MROWS + m) * kGemvCh + lane] = acc[m];
  __syncthreads();
  if (warp < MROWS && warp < args.M) {
    const int m = warp;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kGemvWarps; ++w) t += red[(w * MROWS + m) * kGemvCh + lane];
    const int n = ch0 + lane;
    if (args.bias != nullptr) t += __half2float(args.bias[n]);
    __half hv = __float2half_rn(t);
    const size_t off = static_cast<size_t>(m) * args.ldc + args.col0 + n;
    if (args.residual != nullptr) hv = __hadd(args.residual[off], hv);
    args.C[off] = hv;
  }
  if (independent) pdl_wait_prior_grid();      // completion stays transitive (see w4a16_umma_kernel)
}

}  // namespace qb200
