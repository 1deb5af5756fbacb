// quick_b200 — C-ABI implementation (see include/quick_b200.h for the contract and the reference
// interfaces each entry point replaces).  sm_100a only; there is no CPU or non-tcgen05 fallback on
// the hot path: if the device is not compute capability 10.x the GEMM entry points return QB200_ECUDA.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/quick_b200.h"
#include "w4a16_umma.cuh"

namespace {

thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};
long long* g_trace = nullptr;   // debug trace buffer (qb200_debug_set_trace)

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define QB_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return fail(QB200_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------------
// Layout kernels
// ------------------------------------------------------------------------------------------------

// (k, n) -> flat QUICK qweight word index and nibble position.
// Inverse of the reference kernel's fragment addressing (gemm_cuda_quick.cu:1354,:1372 + mma.m16n8k16
// B-fragment ownership); closed form in SURVEY.md Appendix A-1.
__device__ __forceinline__ void quick_locate(int k, int n, int N, size_t& word, int& nib) {
  const int kt = k >> 5, kk = k & 31, ks = kk >> 4, k16 = kk & 15;
  const int hi = k16 >> 3;                 // 0: rows k0,k0+1   1: rows k0+8,k0+9
  const int l4 = (k16 & 7) >> 1;           // lane % 4
  const int odd = k16 & 1;
  const int bx = n >> 7, nn = n & 127, ty = nn >> 6, n64 = nn & 63, chk = n64 >> 4, n16 = n64 & 15;
  const int dc = n16 >> 3;                 // 0: c0   1: c0 + 8
  const int lane = ((n16 & 7) << 2) | l4;
  word = static_cast<size_t>(kt) * 4 * N + static_cast<size_t>(2 * ty + (lane >> 4)) * N + bx * 128 + (lane & 15) * 8 +
         ks * 4 + chk;
  nib = odd * 4 + dc * 2 + hi;
}
// column n -> scale/zero slot x of a packed row (inverse of quick.py:125-128)
__device__ __forceinline__ int quick_slot(int n, int N) {
  const int bx = n >> 7, ty = (n & 127) >> 6, r = n & 63;
  const int m = ((r >> 4) << 1) | ((r & 15) >> 3);
  const int lh = (r & 7) >> 2, j4 = r & 3;
  return ty * (N >> 1) + lh * (N >> 2) + bx * 32 + j4 * 8 + m;
}

// One thread per B200 word: gathers 8 nibbles (4 QUICK words x 2 nibbles).
__global__ void relayout_wq_kernel(const uint32_t* __restrict__ qweight, uint32_t* __restrict__ wq, int K, int N) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(K) * N / 8;
  if (idx >= total) return;
  const int KB = K / 64;
  const int j = idx & 3;
  const int c = (idx >> 2) & 127;
  const int h = (idx >> 9) & 1;
  const size_t blk = idx >> 10;
  const int kb = static_cast<int>(blk % KB);
  const int nt = static_cast<int>(blk / KB);
  const int n = nt * 128 + c;
  const int kbase = kb * 64 + h * 32 + j * 8;
  uint32_t out = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int i = (p < 4) ? 2 * p : 2 * (p - 4) + 1;   // nibble order k0,k2,k4,k6,k1,k3,k5,k7
    size_t w;
    int nib;
    quick_locate(kbase + i, n, N, w, nib);
    out |= ((__ldg(qweight + w) >> (4 * nib)) & 0xFu) << (4 * p);
  }
  wq[idx] = out;
}

__global__ void relayout_sz_kernel(const uint32_t* __restrict__ qzeros, const __half* __restrict__ scales,
                                   uint32_t* __restrict__ sz, int NG, int N) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(NG) * N) return;
  const int c = idx & 127;
  const size_t blk = idx >> 7;
  const int g = static_cast<int>(blk % NG);
  const int nt = static_cast<int>(blk / NG);
  const int n = nt * 128 + c;
  const int x = quick_slot(n, N);
  const uint32_t zw = __ldg(qzeros + static_cast<size_t>(g) * (N / 4) + (x >> 2));
  const uint32_t z = (zw >> (4 * (x & 3))) & 0xFu;
  const uint32_t s = __half_as_ushort(__ldg(scales + static_cast<size_t>(g) * 2 * N + 2 * x));
  sz[idx] = s | ((0x6400u + z) << 16);
}

// logical -> QUICK layout (GPU packer).  One thread per qweight word / per scale-zero slot.
__global__ void pack_qweight_kernel(const uint8_t* __restrict__ q, uint32_t* __restrict__ qweight, int K, int N) {
  const size_t f = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (f >= static_cast<size_t>(K) * N / 8) return;
  const int kt = static_cast<int>(f / (4 * static_cast<size_t>(N)));
  const int r = static_cast<int>(f % (4 * static_cast<size_t>(N)));
  const int r4 = r / N, c = r % N, bx = c >> 7, w = c & 127;
  const int lane = 16 * (r4 & 1) + (w >> 3), ty = r4 >> 1, ks = (w & 7) >> 2, chk = w & 3;
  const int k0 = 32 * kt + 16 * ks + 2 * (lane & 3);
  const int c0 = 128 * bx + 64 * ty + 16 * chk + (lane >> 2);
  uint32_t out = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int dk = ((p >> 2) & 1) + 8 * (p & 1);
    const int dc = 8 * ((p >> 1) & 1);
    out |= (static_cast<uint32_t>(q[static_cast<size_t>(k0 + dk) * N + c0 + dc]) & 0xFu) << (4 * p);
  }
  qweight[f] = out;
}
__global__ void pack_sz_kernel(const uint8_t* __restrict__ z, const __half* __restrict__ s, uint32_t* __restrict__ qzeros,
                               __half* __restrict__ scales, int NG, int N) {
  // one thread per zero word (4 slots): writes 1 int32 of zeros and 8 halves of (duplicated) scales
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(NG) * (N / 4)) return;
  const int g = static_cast<int>(idx / (N / 4));
  const int xw = static_cast<int>(idx % (N / 4));
  const int nb = N >> 7;
  uint32_t zw = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = 4 * xw + i;
    const int ty = x / (N >> 1), lh = (x / (N >> 2)) & 1, bx = (x >> 5) % nb, j4 = (x & 31) >> 3, m = x & 7;
    const int n = 128 * bx + 64 * ty + 16 * (m >> 1) + 8 * (m & 1) + 4 * lh + j4;
    zw |= (static_cast<uint32_t>(z[static_cast<size_t>(g) * N + n]) & 0xFu) << (4 * i);
    const __half sv = s[static_cast<size_t>(g) * N + n];
    scales[static_cast<size_t>(g) * 2 * N + 2 * x] = sv;
    scales[static_cast<size_t>(g) * 2 * N + 2 * x + 1] = sv;
  }
  qzeros[idx] = zw | (zw << 16);
}

// ---- AWQ "GEMM" checkpoint layout (the format public AWQ checkpoints ship in) ----
// qweight int32 [K][N/8], qzeros int32 [K/G][N/8]: nibble i of word (row, c) is column 8c + ORDER[i],
// ORDER = {0,2,4,6,1,3,5,7} (reference packer quick/awq/modules/linear/gemm.py:108-143, inverse in
// quick/awq/utils/packing_utils.py:4-39); scales fp16 [K/G][N].
__device__ __forceinline__ uint32_t awq_gemm_nibble(const uint32_t* __restrict__ packed, int row, int n, int N) {
  const int inv = ((n & 1) << 2) | ((n & 7) >> 1);   // position of column (n % 8) inside the word: inverse of ORDER
  return (__ldg(packed + static_cast<size_t>(row) * (N >> 3) + (n >> 3)) >> (4 * inv)) & 0xFu;
}

// AWQ-GEMM -> QUICK layout: one thread per QUICK qweight word / per zero word (same closed form as pack_*_kernel).
__global__ void gemm_to_quick_qweight_kernel(const uint32_t* __restrict__ gq, uint32_t* __restrict__ qweight, int K, int N) {
  const size_t f = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (f >= static_cast<size_t>(K) * N / 8) return;
  const int kt = static_cast<int>(f / (4 * static_cast<size_t>(N)));
  const int r = static_cast<int>(f % (4 * static_cast<size_t>(N)));
  const int r4 = r / N, c = r % N, bx = c >> 7, w = c & 127;
  const int lane = 16 * (r4 & 1) + (w >> 3), ty = r4 >> 1, ks = (w & 7) >> 2, chk = w & 3;
  const int k0 = 32 * kt + 16 * ks + 2 * (lane & 3);
  const int c0 = 128 * bx + 64 * ty + 16 * chk + (lane >> 2);
  uint32_t out = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int dk = ((p >> 2) & 1) + 8 * (p & 1);
    const int dc = 8 * ((p >> 1) & 1);
    out |= awq_gemm_nibble(gq, k0 + dk, c0 + dc, N) << (4 * p);
  }
  qweight[f] = out;
}
__global__ void gemm_to_quick_sz_kernel(const uint32_t* __restrict__ gz, const __half* __restrict__ gs,
                                        uint32_t* __restrict__ qzeros, __half* __restrict__ scales, int NG, int N) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(NG) * (N / 4)) return;
  const int g = static_cast<int>(idx / (N / 4));
  const int xw = static_cast<int>(idx % (N / 4));
  const int nb = N >> 7;
  uint32_t zw = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = 4 * xw + i;
    const int ty = x / (N >> 1), lh = (x / (N >> 2)) & 1, bx = (x >> 5) % nb, j4 = (x & 31) >> 3, m = x & 7;
    const int n = 128 * bx + 64 * ty + 16 * (m >> 1) + 8 * (m & 1) + 4 * lh + j4;
    zw |= awq_gemm_nibble(gz, g, n, N) << (4 * i);
    const __half sv = gs[static_cast<size_t>(g) * N + n];
    scales[static_cast<size_t>(g) * 2 * N + 2 * x] = sv;
    scales[static_cast<size_t>(g) * 2 * N + 2 * x + 1] = sv;
  }
  qzeros[idx] = zw | (zw << 16);
}

// AWQ-GEMM -> B200 layout directly (what the kernel streams): one thread per B200 word / per (group, channel).
__global__ void gemm_to_b200_wq_kernel(const uint32_t* __restrict__ gq, uint32_t* __restrict__ wq, int K, int N) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(K) * N / 8) return;
  const int KB = K / 64;
  const int j = idx & 3, c = (idx >> 2) & 127, h = (idx >> 9) & 1;
  const size_t blk = idx >> 10;
  const int kb = static_cast<int>(blk % KB), nt = static_cast<int>(blk / KB);
  const int n = nt * 128 + c;
  const int kbase = kb * 64 + h * 32 + j * 8;
  uint32_t out = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int i = (p < 4) ? 2 * p : 2 * (p - 4) + 1;   // nibble order k0,k2,k4,k6,k1,k3,k5,k7
    out |= awq_gemm_nibble(gq, kbase + i, n, N) << (4 * p);
  }
  wq[idx] = out;
}
__global__ void gemm_to_b200_sz_kernel(const uint32_t* __restrict__ gz, const __half* __restrict__ gs,
                                       uint32_t* __restrict__ sz, int NG, int N) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(NG) * N) return;
  const int c = idx & 127;
  const size_t blk = idx >> 7;
  const int g = static_cast<int>(blk % NG), nt = static_cast<int>(blk / NG);
  const int n = nt * 128 + c;
  const uint32_t z = awq_gemm_nibble(gz, g, n, N);
  const uint32_t sbits = __half_as_ushort(__ldg(gs + static_cast<size_t>(g) * N + n));
  sz[idx] = sbits | ((0x6400u + z) << 16);
}

// B200 layout -> W16[K][N]
__global__ void dequant_kernel(const uint32_t* __restrict__ wq, const uint32_t* __restrict__ sz, __half* __restrict__ W,
                               int K, int N, int G) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(K) * N / 8) return;
  const int KB = K / 64, NG = K / G;
  const int j = idx & 3, c = (idx >> 2) & 127, h = (idx >> 9) & 1;
  const size_t blk = idx >> 10;
  const int kb = static_cast<int>(blk % KB), nt = static_cast<int>(blk / KB);
  const int n = nt * 128 + c;
  const int kbase = kb * 64 + h * 32 + j * 8;
  const uint32_t szw = sz[(static_cast<size_t>(nt) * NG + kbase / G) * 128 + c];
  const qb200::GroupConsts g = qb200::make_group_consts(szw);
  uint32_t o[4];
  qb200::dequant_word(wq[idx], g, o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    W[static_cast<size_t>(kbase + 2 * i) * N + n] = __ushort_as_half(static_cast<unsigned short>(o[i] & 0xFFFFu));
    W[static_cast<size_t>(kbase + 2 * i + 1) * N + n] = __ushort_as_half(static_cast<unsigned short>(o[i] >> 16));
  }
}

// CUDA-core cross-check: one CTA per (n-tile, row m); thread = channel; fp32 accumulation in k order.
__global__ void gemm_simt_kernel(const __half* __restrict__ A, const uint32_t* __restrict__ wq,
                                 const uint32_t* __restrict__ sz, __half* __restrict__ C, int M, int K, int N, int G) {
  extern __shared__ __half a_row[];
  const int nt = blockIdx.x, m = blockIdx.y, c = threadIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) a_row[k] = A[static_cast<size_t>(m) * K + k];
  __syncthreads();
  const int KB = K / 64, NG = K / G;
  float acc = 0.f;
  for (int kb = 0; kb < KB; ++kb) {
    for (int h = 0; h < 2; ++h) {
      const int k32 = kb * 64 + h * 32;
      const uint32_t szw = sz[(static_cast<size_t>(nt) * NG + k32 / G) * 128 + c];
      const qb200::GroupConsts g = qb200::make_group_consts(szw);
      const uint4 w = *reinterpret_cast<const uint4*>(wq + ((static_cast<size_t>(nt) * KB + kb) * 2 + h) * 512 + c * 4);
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t o[4];
        qb200::dequant_word(ws[j], g, o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = k32 + j * 8 + 2 * i;
          acc += __half2float(a_row[k]) * __half2float(__ushort_as_half(static_cast<unsigned short>(o[i] & 0xFFFFu)));
          acc += __half2float(a_row[k + 1]) * __half2float(__ushort_as_half(static_cast<unsigned short>(o[i] >> 16)));
        }
      }
    }
  }
  C[static_cast<size_t>(m) * N + nt * 128 + c] = __float2half_rn(acc);
}

// ------------------------------------------------------------------------------------------------
// Decoder-layer glue kernels: what sits between the GEMMs of a Llama-like layer.  The reference's
// fused modules do the same jobs with awq_ext kernels that are not part of its tree (modules/fused/norm.py:18
// layernorm_forward_cuda, attn.py:100-245 RoPE + cache update, mlp.py:52-76 silu * up).  Arithmetic follows the
// torch expressions of quick_b200/awq/models/llama_like.py step by step (same roundings), so they are drop-ins.
// ------------------------------------------------------------------------------------------------

// y = fp16( fp16( x * rsqrt(mean(x^2) + eps) ) * w ), statistics in fp32.  One CTA per row.
__global__ void rmsnorm_kernel(const __half* __restrict__ x, const __half* __restrict__ w, __half* __restrict__ y, int H, float eps,
                               const qb200::PeerWait wait) {
  // programmatic dependent launch: let the next kernel (usually a GEMM: barrier init, TMEM, weight prefetch) start
  // now; our own input comes from the previous kernel, so wait for it before the first load
  qb200::pdl_launch_dependents();
  qb200::pdl_wait_prior_grid();
  if (wait.epoch != nullptr) {          // tensor parallel: x is a gathered buffer, meet the ranks that fill it
    if (threadIdx.x < 32) qb200::peer_wait_warp(wait, blockIdx.x == 0);
    __syncthreads();
  }
  const __half* xr = x + static_cast<size_t>(blockIdx.x) * H;
  __half* yr = y + static_cast<size_t>(blockIdx.x) * H;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); ss += f.x * f.x + f.y * f.y; }
  }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) red[0] = rsqrtf(t / static_cast<float>(H) + eps);
  }
  __syncthreads();
  const float r = red[0];
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint4 wv = *reinterpret_cast<const uint4*>(w + i);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      oh[j] = __hmul2(__floats2half2_rn(f.x * r, f.y * r), wh[j]);
    }
    *reinterpret_cast<uint4*>(yr + i) = o;
  }
}

// qkv [B][T][(nh + 2 nkv) * hd] -> q_rot [B][T][nh][hd] (token-major, like qkv: attention over the transposed view then
// returns token-major rows, which o_proj reads without a transpose copy); k_rot, v written into the static caches
// [B][nkv][S][hd] at pos[t].
// rope(t) = fp16(fp16(t * cos) + fp16(rot(t) * sin)), rot = (-t2, t1) over the two halves of the head (same as _rope()).
// One work item = (row b·T + t, head, 8 dims of the lower half + the matching 8 dims of the upper half): 16-byte loads
// and stores, a grid-strided loop (a prefill of 8192 rows is 6 M items; the first version launched one 128-thread block
// per (head, row): 786 k blocks).
__device__ __forceinline__ uint4 rope8(uint4 v, uint4 other, uint4 c, uint4 s, bool negate_other) {
  uint4 r;
  const __half2* pv = reinterpret_cast<const __half2*>(&v);
  const __half2* po = reinterpret_cast<const __half2*>(&other);
  const __half2* pc = reinterpret_cast<const __half2*>(&c);
  const __half2* ps = reinterpret_cast<const __half2*>(&s);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 o = negate_other ? __hneg2(po[i]) : po[i];
    pr[i] = __hadd2_rn(__hmul2_rn(pv[i], pc[i]), __hmul2_rn(o, ps[i]));   // _rn: no contraction into fma, torch rounds each step
  }
  return r;
}
__global__ void rope_kv_kernel(const __half* __restrict__ qkv, const __half* __restrict__ cosb, const __half* __restrict__ sinb,
                               const long long* __restrict__ pos, __half* __restrict__ q_out, __half* __restrict__ cache_k,
                               __half* __restrict__ cache_v, int B, int T, int nh, int nkv, int hd, int S) {
  qb200::pdl_launch_dependents();
  qb200::pdl_wait_prior_grid();
  const int heads = nh + 2 * nkv, half_hd = hd >> 1, per_head = hd >> 4;     // items per head (hd is a multiple of 16)
  const long long total = static_cast<long long>(B) * T * heads * per_head;
  for (long long it = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; it < total;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(it % per_head) * 8;
    long long r = it / per_head;
    const int head = static_cast<int>(r % heads);
    r /= heads;
    const int t = static_cast<int>(r % T), b = static_cast<int>(r / T);
    const __half* src = qkv + (static_cast<size_t>(b) * T + t) * heads * hd + static_cast<size_t>(head) * hd;
    const uint4 lo = *reinterpret_cast<const uint4*>(src + c8), hi = *reinterpret_cast<const uint4*>(src + half_hd + c8);
    const long long p = pos[t];
    if (head >= nh + nkv) {               // value head: plain copy into the cache
      __half* dst = cache_v + ((static_cast<size_t>(b) * nkv + (head - nh - nkv)) * S + p) * hd;
      *reinterpret_cast<uint4*>(dst + c8) = lo;
      *reinterpret_cast<uint4*>(dst + half_hd + c8) = hi;
      continue;
    }
    const __half* cp = cosb + static_cast<size_t>(p) * hd;
    const __half* sp = sinb + static_cast<size_t>(p) * hd;
    const uint4 out_lo = rope8(lo, hi, *reinterpret_cast<const uint4*>(cp + c8), *reinterpret_cast<const uint4*>(sp + c8), true);
    const uint4 out_hi = rope8(hi, lo, *reinterpret_cast<const uint4*>(cp + half_hd + c8), *reinterpret_cast<const uint4*>(sp + half_hd + c8), false);
    __half* dst = head < nh ? q_out + ((static_cast<size_t>(b) * T + t) * nh + head) * hd
                            : cache_k + ((static_cast<size_t>(b) * nkv + (head - nh)) * S + p) * hd;
    *reinterpret_cast<uint4*>(dst + c8) = out_lo;
    *reinterpret_cast<uint4*>(dst + half_hd + c8) = out_hi;
  }
}

// Decode-step attention (ONE new token per sequence) fused with the rotary embedding and the KV-cache update — the job of
// the reference's QuantAttentionFused decode branch (modules/fused/attn.py:187-245: awq_ext.single_query_attention over
// its rolling cache).  A CLUSTER of `nsplit` CTAs per (kv head, sequence) — flash-decoding inside a thread-block cluster:
// CTA r of the cluster takes a contiguous range of the cached positions, a warp streams its positions' K and V rows in
// ONE pass (online softmax, kBatch K rows + kBatch V rows in flight per lane, a lane owns hd/32 consecutive dims), the
// eight warps merge in shared memory and the cluster's partial (max, sum, P·V) triples are merged by CTA 0 through
// distributed shared memory.  No score buffer (any cache length), no mask tensor, no q / k / v round trip through HBM.
// The g = nh / nkv query heads that share a kv head are processed together (each K / V row is read once).
// CTA 0 also rotates k, writes k / v into the static caches [B][nkv][S][hd] at position p = pos[0] (same arithmetic as
// rope_kv_kernel) and adds the new position from shared memory.
//   out [B][1][nh * hd] fp16 (the layout o_proj consumes), or (tensor parallel) column col0 of every rank's buffer.
// Before griddepcontrol.wait the CTAs prefetch their K / V rows into L2 (prefetch.global.L2: no data is consumed, so a
// stale position can only cost bandwidth) — the HBM latency of the cache overlaps the tail of the q|k|v GEMM.
constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = kAttnThreads / 32;
constexpr int kAttnMaxSplit = 8;
__host__ __device__ constexpr size_t attn_decode_smem(int g, int hd) {
  // q_rot g·hd, k, v (fp16) | per-warp (max, sum) | per-warp P·V | per-CTA (max, sum) x cluster | per-CTA P·V x cluster | mbarrier
  return static_cast<size_t>(g + 2) * hd * 2 + static_cast<size_t>(kAttnWarps) * g * 8 + static_cast<size_t>(kAttnWarps) * g * hd * 4 +
         static_cast<size_t>(kAttnMaxSplit) * g * 8 + static_cast<size_t>(kAttnMaxSplit) * g * hd * 4 + 16;
}

template <int VEC> struct AttnRow;      // VEC consecutive fp16 of a cache row, as loaded
template <> struct AttnRow<2> { uint32_t v; };
template <> struct AttnRow<4> { uint2 v; };
template <> struct AttnRow<8> { uint4 v; };
template <int VEC>
__device__ __forceinline__ void attn_unpack(const AttnRow<VEC>& r, float (&f)[VEC]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r.v);
#pragma unroll
  for (int u = 0; u < VEC / 2; ++u) {
    const float2 t = __half22float2(h[u]);
    f[2 * u] = t.x; f[2 * u + 1] = t.y;
  }
}

// Destination of a column slab that every rank needs (tensor parallel): the [rows][ld] buffers of all ranks at column
// col0 — ONE multimem.st per 16-byte chunk through the NVSwitch multicast mapping when there is one, else a loop of
// peer stores.
struct PeerDst {
  __half* peer[8];
  __half* mc;
  int n, ld, col0;
};
__device__ __forceinline__ void store16_all(const PeerDst& d, size_t off, uint4 v) {
  if (d.mc != nullptr) qb200::multimem_st_v4(d.mc + off, v);
  else for (int p = 0; p < d.n; ++p) *reinterpret_cast<uint4*>(d.peer[p] + off) = v;
}
template <int HD, int G>
__global__ void __launch_bounds__(kAttnThreads)
attn_decode_kernel(const __half* __restrict__ qkv, const __half* __restrict__ cosb, const __half* __restrict__ sinb,
                   const long long* __restrict__ pos, __half* __restrict__ out, __half* __restrict__ cache_k,
                   __half* __restrict__ cache_v, int nh, int nkv, int S, float scale, int nsplit, const PeerDst dst,
                   const qb200::PeerSignal sig) {
  constexpr int hd = HD, g = G;
  constexpr int VEC = HD / 32;                           // halves of a row owned by one lane (2, 4 or 8)
  constexpr int kBatch = (G * VEC >= 32) ? 2 : 4;        // cache positions per batch (K and V rows each); two batches in flight
  const int kvh = blockIdx.x, b = blockIdx.y;
  const int rank = blockIdx.z;                           // == %cluster_ctarank: the cluster spans the z dimension
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t crow = (static_cast<size_t>(b) * nkv + kvh) * S;   // first cache row of this (sequence, kv head)
  const __half* kbase = cache_k + crow * hd + lane * VEC;
  const __half* vbase = cache_v + crow * hd + lane * VEC;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __half* qs = reinterpret_cast<__half*>(attn_smem);     // [g][hd]
  __half* ks = qs + g * hd;                              // [hd]   (CTA 0)
  __half* vs = ks + hd;                                  // [hd]   (CTA 0)
  float* wm = reinterpret_cast<float*>(vs + hd);         // [warps][g] running max
  float* wl = wm + kAttnWarps * g;                       // [warps][g] running sum
  float* wacc = wl + kAttnWarps * g;                     // [warps][g][hd]
  float* cm = wacc + kAttnWarps * g * hd;                // [cluster][g]      per-CTA triples, gathered in CTA 0 of the cluster
  float* cl = cm + kAttnMaxSplit * g;                    // [cluster][g]
  float* cacc = cl + kAttnMaxSplit * g;                  // [cluster][g][hd]
  const uint32_t bar = qb200::smem_u32(cacc + kAttnMaxSplit * g * hd);   // CTA 0: the other CTAs' triples have landed
  qb200::pdl_launch_dependents();
  if (nsplit > 1) {
    if (rank == 0 && tid == 0) {
      qb200::mbar_init(bar, 1);
      qb200::fence_barrier_init();
    }
    qb200::cluster_arrive_relaxed();                      // matched by a wait just before the first remote store
  }
  {   // L2 prefetch of this warp's cache rows; pos may still be the previous step's value here (harmless, see above)
    const int pp = min(max(static_cast<int>(pos[0]), 0), S - 1);
    const int per = (pp + nsplit - 1) / nsplit;
    const int lo = rank * per, hi = min(pp, lo + per);
    for (int j = lo + warp; j < hi; j += kAttnWarps) {
      if (lane * VEC * 2 % 128 == 0) {                   // one prefetch per 128-byte line of the row
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kbase + static_cast<size_t>(j) * hd));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(vbase + static_cast<size_t>(j) * hd));
      }
    }
  }
  qb200::pdl_wait_prior_grid();
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) qb200::peer_begin_fill(sig);   // tensor parallel: out is a gathered buffer
  const int p = static_cast<int>(pos[0]);                // the new position; positions 0 .. p are visible
  const int per = (p + nsplit - 1) / nsplit;             // cached positions 0 .. p-1 split over the cluster
  const int lo = rank * per, hi = min(p, lo + per);

  // first batch of this warp's K / V rows: requested before the rotary phase (they do not depend on this step's q)
  AttnRow<VEC> kr0[kBatch], vr0[kBatch], kr1[kBatch], vr1[kBatch];   // two batches (named, so they stay in registers)
  auto load_batch = [&](int j0, AttnRow<VEC> (&kb)[kBatch], AttnRow<VEC> (&vb)[kBatch]) {   // warp-uniform bounds; returns the valid rows
    int nvalid = 0;
#pragma unroll
    for (int t = 0; t < kBatch; ++t) {
      const int j = j0 + t * kAttnWarps;
      if (j < hi) {
        kb[t] = *reinterpret_cast<const AttnRow<VEC>*>(kbase + static_cast<size_t>(j) * hd);
        vb[t] = *reinterpret_cast<const AttnRow<VEC>*>(vbase + static_cast<size_t>(j) * hd);
        nvalid = t + 1;
      } else {
        kb[t] = AttnRow<VEC>{}; vb[t] = AttnRow<VEC>{};
      }
    }
    return nvalid;
  };
  int j0 = lo + warp;
  int nv = j0 < hi ? load_batch(j0, kr0, vr0) : 0;

  // rotary embedding of the g query heads (every CTA) and of k (CTA 0, which also updates the cache): one item = 8 dims of
  // the lower half of a head + the matching 8 of the upper half, 16-byte accesses (same arithmetic as rope_kv_kernel)
  const __half* src = qkv + static_cast<size_t>(b) * (nh + 2 * nkv) * hd;
  constexpr int half_hd = hd >> 1, per_head = hd >> 4;
  for (int idx = tid; idx < (rank == 0 ? g + 2 : g) * per_head; idx += kAttnThreads) {
    const int hl = idx / per_head, c8 = (idx - hl * per_head) * 8;
    const __half* hsrc = src + static_cast<size_t>(hl < g ? kvh * g + hl : hl == g ? nh + kvh : nh + nkv + kvh) * hd;
    const uint4 lo = *reinterpret_cast<const uint4*>(hsrc + c8), hi = *reinterpret_cast<const uint4*>(hsrc + half_hd + c8);
    if (hl == g + 1) {                                   // value: plain copy
      *reinterpret_cast<uint4*>(vs + c8) = lo;
      *reinterpret_cast<uint4*>(vs + half_hd + c8) = hi;
      __half* dstv = cache_v + (crow + p) * hd;
      *reinterpret_cast<uint4*>(dstv + c8) = lo;
      *reinterpret_cast<uint4*>(dstv + half_hd + c8) = hi;
      continue;
    }
    const __half* cp = cosb + static_cast<size_t>(p) * hd;
    const __half* sp = sinb + static_cast<size_t>(p) * hd;
    const uint4 out_lo = rope8(lo, hi, *reinterpret_cast<const uint4*>(cp + c8), *reinterpret_cast<const uint4*>(sp + c8), true);
    const uint4 out_hi = rope8(hi, lo, *reinterpret_cast<const uint4*>(cp + half_hd + c8), *reinterpret_cast<const uint4*>(sp + half_hd + c8), false);
    __half* dsts = hl < g ? qs + hl * hd : ks;
    *reinterpret_cast<uint4*>(dsts + c8) = out_lo;
    *reinterpret_cast<uint4*>(dsts + half_hd + c8) = out_hi;
    if (hl == g) {
      __half* dstk = cache_k + (crow + p) * hd;
      *reinterpret_cast<uint4*>(dstk + c8) = out_lo;
      *reinterpret_cast<uint4*>(dstk + half_hd + c8) = out_hi;
    }
  }
  __syncthreads();

  // one pass over this warp's cached positions: online softmax, fp32
  float qf[G][VEC], acc[G][VEC], mx[G], sum[G];
#pragma unroll
  for (int h = 0; h < G; ++h) {
    mx[h] = -INFINITY; sum[h] = 0.f;
#pragma unroll
    for (int u = 0; u < VEC; ++u) { qf[h][u] = __half2float(qs[h * hd + lane * VEC + u]) * scale; acc[h][u] = 0.f; }
  }
  auto fold = [&](const AttnRow<VEC> (&kb)[kBatch], const AttnRow<VEC> (&vb)[kBatch], int nvalid) {
    float s[G][kBatch];
#pragma unroll
    for (int t = 0; t < kBatch; ++t) {
      float kf[VEC];
      attn_unpack<VEC>(kb[t], kf);
#pragma unroll
      for (int h = 0; h < G; ++h) {
        float a = 0.f;
#pragma unroll
        for (int u = 0; u < VEC; ++u) a = fmaf(qf[h][u], kf[u], a);
        s[h][t] = a;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int h = 0; h < G; ++h)
#pragma unroll
        for (int t = 0; t < kBatch; ++t) s[h][t] += __shfl_xor_sync(0xffffffffu, s[h][t], o);
#pragma unroll
    for (int h = 0; h < G; ++h) {
      float m_new = mx[h];
#pragma unroll
      for (int t = 0; t < kBatch; ++t) if (t < nvalid) m_new = fmaxf(m_new, s[h][t]);
      const float corr = __expf(mx[h] - m_new);          // first batch: exp(-inf) = 0
      mx[h] = m_new;
      sum[h] *= corr;
#pragma unroll
      for (int u = 0; u < VEC; ++u) acc[h][u] *= corr;
#pragma unroll
      for (int t = 0; t < kBatch; ++t) {
        s[h][t] = t < nvalid ? __expf(s[h][t] - m_new) : 0.f;
        sum[h] += s[h][t];
      }
    }
#pragma unroll
    for (int t = 0; t < kBatch; ++t) {
      float vf[VEC];
      attn_unpack<VEC>(vb[t], vf);
#pragma unroll
      for (int h = 0; h < G; ++h)
#pragma unroll
        for (int u = 0; u < VEC; ++u) acc[h][u] = fmaf(s[h][t], vf[u], acc[h][u]);
    }
  };
  {
    constexpr int kStep = kAttnWarps * kBatch;
    while (nv > 0) {                                       // warp-uniform; the next batch is in flight while this one is folded
      j0 += kStep;
      const int n1 = j0 < hi ? load_batch(j0, kr1, vr1) : 0;
      fold(kr0, vr0, nv);
      if (n1 == 0) break;
      j0 += kStep;
      nv = j0 < hi ? load_batch(j0, kr0, vr0) : 0;
      fold(kr1, vr1, n1);
    }
    if (rank == 0 && warp == 0) {                          // the new position, from shared memory
#pragma unroll
      for (int t = 0; t < kBatch; ++t) { kr0[t] = AttnRow<VEC>{}; vr0[t] = AttnRow<VEC>{}; }
      kr0[0] = *reinterpret_cast<const AttnRow<VEC>*>(ks + lane * VEC);
      vr0[0] = *reinterpret_cast<const AttnRow<VEC>*>(vs + lane * VEC);
      fold(kr0, vr0, 1);
    }
  }
#pragma unroll
  for (int h = 0; h < G; ++h) {
    if (lane == 0) { wm[warp * g + h] = mx[h]; wl[warp * g + h] = sum[h]; }
#pragma unroll
    for (int u = 0; u < VEC; ++u) wacc[(warp * g + h) * hd + lane * VEC + u] = acc[h][u];
  }
  __syncthreads();

  auto store_out = [&](int h, int d, float val) {
    const __half v = __float2half_rn(val);
    if (dst.n == 0) {
      out[static_cast<size_t>(b) * nh * hd + static_cast<size_t>(kvh * g + h) * hd + d] = v;
    } else {   // tensor parallel: this rank's heads go to column col0 of every rank's [B][ld] attention buffer
      const size_t off = static_cast<size_t>(b) * dst.ld + dst.col0 + static_cast<size_t>(kvh * g + h) * hd + d;
      for (int pr = 0; pr < dst.n; ++pr) dst.peer[pr][off] = v;
    }
  };
  // merge the eight warps: (max, sum, P·V) of this CTA's positions -> slot `rank` of CTA 0's gather buffers
  // (st.async through distributed shared memory, completion counted on CTA 0's mbarrier; CTA 0 writes its own slot)
  uint32_t r_cm = 0, r_cl = 0, r_cacc = 0, r_bar = 0;
  if (nsplit > 1 && rank != 0) {
    qb200::cluster_wait();                                 // CTA 0 has initialised its mbarrier
    r_cm = qb200::mapa_shared(qb200::smem_u32(cm + rank * g), 0);
    r_cl = qb200::mapa_shared(qb200::smem_u32(cl + rank * g), 0);
    r_cacc = qb200::mapa_shared(qb200::smem_u32(cacc + rank * g * hd), 0);
    r_bar = qb200::mapa_shared(bar, 0);
  }
  for (int idx = tid; idx < g * hd; idx += kAttnThreads) {
    const int h = idx / hd, d = idx - h * hd;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, wm[w * g + h]);
    float L = 0.f, A = 0.f;
    if (M != -INFINITY) {
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) {
        const float mw = wm[w * g + h];
        const float e = mw == -INFINITY ? 0.f : __expf(mw - M);
        L = fmaf(wl[w * g + h], e, L);
        A = fmaf(wacc[(w * g + h) * hd + d], e, A);
      }
    }
    if (nsplit == 1) {
      store_out(h, d, A / L);                              // the new position is always present: L > 0
    } else if (rank == 0) {
      cacc[h * hd + d] = A;
      if (d == 0) { cm[h] = M; cl[h] = L; }
    } else {
      uint32_t w = __float_as_uint(A);
      qb200::st_async<1>(r_cacc + (h * hd + d) * 4, &w, r_bar);
      if (d == 0) {
        w = __float_as_uint(M); qb200::st_async<1>(r_cm + h * 4, &w, r_bar);
        w = __float_as_uint(L); qb200::st_async<1>(r_cl + h * 4, &w, r_bar);
      }
    }
  }
  if (nsplit == 1 || rank != 0) return;                    // outbound st.async data is in flight from registers

  // CTA 0: wait for the other CTAs' triples, merge, store
  qb200::cluster_wait();                                   // (pairs with the arrive at the top; every thread waits once)
  if (tid == 0) qb200::mbar_arrive_expect_tx(bar, static_cast<uint32_t>((nsplit - 1) * (g * hd * 4 + g * 8)));
  __syncthreads();                                         // own slot written
  qb200::mbar_wait_cluster(bar, 0);
  for (int idx = tid; idx < g * hd; idx += kAttnThreads) {
    const int h = idx / hd, d = idx - h * hd;
    float M = -INFINITY;
#pragma unroll
    for (int r = 0; r < kAttnMaxSplit; ++r) if (r < nsplit) M = fmaxf(M, cm[r * g + h]);
    float L = 0.f, A = 0.f;
#pragma unroll
    for (int r = 0; r < kAttnMaxSplit; ++r) {
      if (r < nsplit) {
        const float mr = cm[r * g + h];
        const float e = mr == -INFINITY ? 0.f : __expf(mr - M);
        L = fmaf(cl[r * g + h], e, L);
        A = fmaf(cacc[(r * g + h) * hd + d], e, A);
      }
    }
    store_out(h, d, A / L);
  }
}

// act[M][I] = fp16( fp16(silu(g)) * u ),  gu = [g | u] per row ([M][2I]); silu in fp32 like torch's half kernel.
__global__ void silu_mul_kernel(const __half* __restrict__ gu, __half* __restrict__ act, size_t M, int I) {
  qb200::pdl_launch_dependents();
  qb200::pdl_wait_prior_grid();
  const size_t idx = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (idx >= M * static_cast<size_t>(I)) return;
  const size_t m = idx / I;
  const int i = static_cast<int>(idx % I);
  const uint4 g = *reinterpret_cast<const uint4*>(gu + m * 2 * I + i);
  const uint4 u = *reinterpret_cast<const uint4*>(gu + m * 2 * I + I + i);
  const __half2* gh = reinterpret_cast<const __half2*>(&g);
  const __half2* uh = reinterpret_cast<const __half2*>(&u);
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(gh[j]);
    oh[j] = __hmul2(__floats2half2_rn(f.x / (1.f + expf(-f.x)), f.y / (1.f + expf(-f.y))), uh[j]);
  }
  *reinterpret_cast<uint4*>(act + idx) = o;
}

// Same for a gate|up GEMM output whose columns are interleaved (g_0, u_0, g_1, u_1, ...): what a QB200_GEMM_SILU_MUL weight
// produces when it is run WITHOUT the fused epilogue (large token tiles, where the epilogue math would idle the tensor pipe).
__global__ void silu_mul_pairs_kernel(const __half* __restrict__ gu, __half* __restrict__ act, size_t total) {
  qb200::pdl_launch_dependents();
  qb200::pdl_wait_prior_grid();
  const size_t idx = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;   // 8 outputs = 16 inputs
  if (idx >= total) return;
  const uint4 a = *reinterpret_cast<const uint4*>(gu + 2 * idx);
  const uint4 b = *reinterpret_cast<const uint4*>(gu + 2 * idx + 8);
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
  const __half2* bh = reinterpret_cast<const __half2*>(&b);
  uint4 o;
  __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __half2 p = j < 4 ? ah[j] : bh[j - 4];
    const float g = __low2float(p);
    oh[j] = __hmul(__float2half_rn(g / (1.f + expf(-g))), __high2half(p));
  }
  *reinterpret_cast<uint4*>(act + idx) = o;
}

// act[rows][I] = silu(g) * u of this rank's gate|up slab [rows][2 I], written to every rank's [rows][ld] activation
// buffer at column col0 (the input of the column-parallel down projection), then published.
__global__ void silu_mul_scatter_kernel(const __half* __restrict__ gu, size_t M, int I, const PeerDst dst, const qb200::PeerSignal sig) {
  qb200::pdl_launch_dependents();
  qb200::pdl_wait_prior_grid();
  if (blockIdx.x == 0 && threadIdx.x == 0) qb200::peer_begin_fill(sig);
  const size_t total = M * static_cast<size_t>(I);
  for (size_t idx = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x * 8) {
    const size_t m = idx / I;
    const int i = static_cast<int>(idx % I);
    const uint4 g = *reinterpret_cast<const uint4*>(gu + m * 2 * I + i);
    const uint4 u = *reinterpret_cast<const uint4*>(gu + m * 2 * I + I + i);
    const __half2* gh = reinterpret_cast<const __half2*>(&g);
    const __half2* uh = reinterpret_cast<const __half2*>(&u);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(gh[j]);
      oh[j] = __hmul2(__floats2half2_rn(f.x / (1.f + expf(-f.x)), f.y / (1.f + expf(-f.y))), uh[j]);
    }
    store16_all(dst, m * dst.ld + dst.col0 + i, o);
  }
}
// src[rows][n_local] -> every rank's [rows][ld] buffer at column col0, then published (the attention output of this
// rank's heads, input of the column-parallel output projection).
__global__ void scatter_cols_kernel(const __half* __restrict__ src, size_t M, int n_local, const PeerDst dst, const qb200::PeerSignal sig) {
  qb200::pdl_launch_dependents();
  qb200::pdl_wait_prior_grid();
  if (blockIdx.x == 0 && threadIdx.x == 0) qb200::peer_begin_fill(sig);
  const size_t total = M * static_cast<size_t>(n_local);
  for (size_t idx = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x * 8) {
    const size_t m = idx / n_local;
    const int i = static_cast<int>(idx % n_local);
    store16_all(dst, m * dst.ld + dst.col0 + i, *reinterpret_cast<const uint4*>(src + idx));
  }
}

// Cross-GPU hand-over for the fused all-gather: every rank bumps its own epoch counter, publishes the epoch in its slot
// of every peer's flag array (peer-mapped memory) and waits until all peers have published the same epoch in ITS array.
// Launched right behind the GEMM on the same stream: the GEMM's peer stores are complete (kernel boundary) before the
// flags go out, and whatever runs next on this stream sees every peer's slab.  The epoch lives on the device, so a
// CUDA-graph replay hands out fresh epochs.
struct PeerFlagPtrs { unsigned* p[8]; };
// timeout_ns: 0 = wait for ever (like NCCL); otherwise a rank that has not seen all peers after that long (%globaltimer,
// independent of the SM clock) traps so that a lost peer does not hang the GPU.  Host-side skew between ranks (rank-0-only
// tokenizing, GC pauses, lazy relayout, graph capture) is legitimate, so the default is long (QB200_PEER_TIMEOUT_S, 120 s).
__global__ void peer_barrier_kernel(unsigned* epoch_counter, PeerFlagPtrs flags, int rank, int n_peers, unsigned long long timeout_ns) {
  __shared__ unsigned e_sh;
  if (threadIdx.x == 0) e_sh = atomicAdd(epoch_counter, 1u) + 1u;
  __syncthreads();
  const unsigned e = e_sh;
  if (threadIdx.x < n_peers) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[threadIdx.x] + rank), "r"(e) : "memory");
    const unsigned* mine = flags.p[rank] + threadIdx.x;
    unsigned long long t0 = 0;
    unsigned v, polls = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (timeout_ns != 0 && (++polls & 1023u) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) __trap();
      }
    } while (static_cast<int>(v - e) < 0);
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-map encode (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_x_tensor_map(CUtensorMap* map, const void* A, int M, int K, int tok) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(QB200_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(M)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(tok)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(A), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(QB200_ECUDA, "cuTensorMapEncodeTiled failed (%d): A=%p M=%d K=%d tok=%d", (int)r, A, M, K, tok);
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------
// GEMM launch
// ------------------------------------------------------------------------------------------------
// Everything CUDA keeps per device is cached per device here (a process may drive several GPUs: device_map,
// the quantizer's work device, tests): SM count, architecture check, dynamic shared-memory opt-ins.
constexpr int kMaxDevices = 64;
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}
int device_sm_count() {
  static std::atomic<int> sms[kMaxDevices];   // zero-initialised: 0 = not queried yet
  const int dev = current_device();
  if (dev < 0) return 148;
  int v = sms[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// One bit per device: has cudaFuncSetAttribute(MaxDynamicSharedMemorySize) been applied to this kernel there?
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  template <typename F>
  cudaError_t ensure(F&& set) {
    const int dev = current_device();
    if (dev < 0) return set();
    const unsigned long long bit = 1ull << dev;
    if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = set();
    if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

bool use_pdl() {   // QB200_NO_PDL=1 disables programmatic dependent launch (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("QB200_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// Launch with the programmatic-stream-serialization attribute (the kernel must call griddepcontrol.wait before it
// touches its inputs); plain launch when PDL is disabled.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

unsigned long long peer_timeout_ns() {   // QB200_PEER_TIMEOUT_S (default 120 s, 0 = never): a lost rank traps instead of hanging
  static const unsigned long long v = [] {
    const char* e = getenv("QB200_PEER_TIMEOUT_S");
    const double sec = e ? atof(e) : 120.0;
    return sec <= 0 ? 0ull : static_cast<unsigned long long>(sec * 1e9);
  }();
  return v;
}
qb200::PeerWait make_wait(const qb200_peer_wait* w) {
  qb200::PeerWait r{};
  if (w != nullptr && w->epoch != nullptr) {
    r.epoch = w->epoch; r.flags = w->flag_arrays[w->rank]; r.rank = w->rank; r.n = w->n_peers; r.timeout_ns = peer_timeout_ns();
    // 2 (default): device-scope acquire fence after the last poll — the rows being read are in THIS device's memory (L2 is
    // where the peers' stores land, and they were released at system scope by the announcing ranks); 0: system-scope fence,
    // the conservative form, 2.3 us (M = 1) to 7 us (64 reader CTAs) slower per hand-over on 2 GPUs (profiles/README.md)
    static const int mode = [] { const char* e = getenv("QB200_TP_WAITMODE"); return e ? atoi(e) : 2; }();
    r.mode = mode;
    for (int p = 0; p < 8; ++p) r.peer_flags[p] = p < w->n_peers ? w->flag_arrays[p] : nullptr;
  }
  return r;
}
qb200::PeerSignal make_signal(const qb200_peer_signal* s) {
  qb200::PeerSignal r{};
  if (s != nullptr && s->epoch != nullptr) r.epoch = s->epoch;
  return r;
}
int check_peer_args(const qb200_peer_wait* w, const qb200_peer_signal* s) {
  (void)s;
  if (w != nullptr && w->epoch != nullptr) {
    if (w->flag_arrays == nullptr || w->n_peers < 1 || w->n_peers > 8 || w->rank < 0 || w->rank >= w->n_peers)
      return fail(QB200_EINVAL, "peer wait: flag arrays, rank and n_peers (1..8) required");
    for (int p = 0; p < w->n_peers; ++p)
      if (w->flag_arrays[p] == nullptr) return fail(QB200_EINVAL, "peer wait: null flag array %d", p);
  }
  return QB200_OK;
}

std::atomic<int> g_variant{-1};        // tile-configuration variant forced by qb200_debug_set_variant (QB200_VARIANTS builds); -1 = planner
// Streams on which a layout kernel has just written wq / sz: the next GEMM on THAT stream must not prefetch its
// weights ahead of griddepcontrol.wait, so it launches without the programmatic attribute.  Tracked per stream
// (a process-wide flag could be consumed by an unrelated GEMM on another stream or thread).
std::mutex g_relayout_mu;
cudaStream_t g_relayout_streams[16];
int g_relayout_n = 0;
std::atomic<int> g_relayout_overflow{0};
void note_relayout(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_relayout_mu);
  for (int i = 0; i < g_relayout_n; ++i)
    if (g_relayout_streams[i] == st) return;
  if (g_relayout_n == 16) {   // table full (16 streams with a pending relayout): be conservative, the next GEMM
    g_relayout_overflow.store(1, std::memory_order_relaxed);   // launched anywhere serialises fully
    return;
  }
  g_relayout_streams[g_relayout_n++] = st;
}
bool take_relayout(cudaStream_t st) {   // true: a layout kernel precedes this launch on the stream
  std::lock_guard<std::mutex> lk(g_relayout_mu);
  for (int i = 0; i < g_relayout_n; ++i) {
    if (g_relayout_streams[i] == st) {
      g_relayout_streams[i] = g_relayout_streams[--g_relayout_n];
      return true;
    }
  }
  return g_relayout_overflow.exchange(0, std::memory_order_relaxed) != 0;
}

template <int TOK, int SPLIT, int VAR, bool SILU>
int launch_umma_impl(const CUtensorMap& map, const qb200::GemmArgs& args, int m_tiles, cudaStream_t stream) {
  using Cfg = qb200::TileCfg<TOK, VAR>;
  auto kfn = qb200::w4a16_umma_kernel<TOK, SPLIT, VAR, SILU>;
  static PerDeviceOnce attr_once;   // per instantiation and device
  QB_CUDA(attr_once.ensure([&] { return cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes(SPLIT)); }));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(args.N / qb200::kChan, m_tiles, SPLIT);
  cfg.blockDim = dim3(Cfg::kNumThreads);
  cfg.dynamicSmemBytes = Cfg::smem_bytes(SPLIT);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = SPLIT;
  // PDL: this GEMM's CTAs start (barrier init, TMEM allocation, weight TMA, first dequant) while the previous
  // kernel of the stream is still running; the weights must therefore not be an output of that kernel.
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const bool pdl = !take_relayout(stream) && use_pdl();
  cfg.numAttrs = pdl ? 2 : 1;
  QB_CUDA(cudaLaunchKernelEx(&cfg, kfn, map, args));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QB200_OK;
}

template <int TOK, int SPLIT, int VAR>
int launch_umma(const CUtensorMap& map, const qb200::GemmArgs& args, int m_tiles, cudaStream_t stream) {
  if (args.flags & qb200::kFlagSiluMul) {
    if constexpr (VAR <= 1) return launch_umma_impl<TOK, SPLIT, VAR, true>(map, args, m_tiles, stream);   // experiment slots: no fused SwiGLU
    else return fail(QB200_EINVAL, "QB200_GEMM_SILU_MUL is not instantiated for experiment variants");
  }
  return launch_umma_impl<TOK, SPLIT, VAR, false>(map, args, m_tiles, stream);
}

template <int TOK, int VAR>
int dispatch_split(int split, const CUtensorMap& map, const qb200::GemmArgs& args, int m_tiles, cudaStream_t st) {
  switch (split) {
    case 1: return launch_umma<TOK, 1, VAR>(map, args, m_tiles, st);
    case 2: return launch_umma<TOK, 2, VAR>(map, args, m_tiles, st);
    case 4: return launch_umma<TOK, 4, VAR>(map, args, m_tiles, st);
    case 8:
      if constexpr (TOK <= 32) return launch_umma<TOK, 8, VAR>(map, args, m_tiles, st);
      break;
  }
  return fail(QB200_EINVAL, "unsupported split %d for tok %d", split, TOK);
}

// Which ring configuration a launch gets (Variant<TOK, VAR> in w4a16_umma.cuh); measured, tools/tune.py.
int pick_variant(int tok, int split, int ctas, int K) {
  const int stages = ((K / 64 + split - 1) / split + 1) / 2;   // 128-k stages per CTA
  if (tok <= 32) return stages <= 4 ? 1 : 0;
  if (tok == 64) return ctas <= device_sm_count() ? 1 : 0;
  if (tok == 128) return stages >= 16 ? 1 : 0;
  return 0;
}

template <int TOK>
int dispatch_variant(int split, const CUtensorMap& map, const qb200::GemmArgs& args, int m_tiles, cudaStream_t st) {
  int var = pick_variant(TOK, split, (args.N / qb200::kChan) * m_tiles * split, args.K);
#ifdef QB200_VARIANTS
  const int forced = g_variant.load(std::memory_order_relaxed);   // -1 = planner's choice
  if (forced >= 0) var = forced;
  switch (var) {
    case 2: return dispatch_split<TOK, 2>(split, map, args, m_tiles, st);
    case 3: return dispatch_split<TOK, 3>(split, map, args, m_tiles, st);
  }
#endif
  if (var == 1) return dispatch_split<TOK, 1>(split, map, args, m_tiles, st);
  return dispatch_split<TOK, 0>(split, map, args, m_tiles, st);
}

int check_device() {
  static std::atomic<signed char> ok[kMaxDevices];   // 0 = unknown, 1 = sm_100, -1 = something else
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(QB200_ECUDA, "no CUDA device");
  signed char v = (dev >= 0 && dev < kMaxDevices) ? ok[dev].load(std::memory_order_relaxed) : 0;
  if (v == 0) {
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    v = (major == 10) ? 1 : -1;
    if (dev >= 0 && dev < kMaxDevices) ok[dev].store(v, std::memory_order_relaxed);
  }
  if (v != 1) return fail(QB200_ECUDA, "quick_b200 requires an sm_100a (B200) device; no fallback path exists");
  return QB200_OK;
}

// Two plans (measured on B200, tools/tune.py, K = N = 4096):
//   ordered (default)  — the GEMM must finish as early as possible on its own: smallest tile that still covers M
//                        without duplicating too much dequant work, then grow the split-K cluster until the grid
//                        fills the 148 SMs.
//   independent        — QB200_GEMM_INDEPENDENT launches overlap each other, so throughput wins: no split-K
//                        (no exchange, one CTA per 128-channel tile streams the whole K) and the largest token
//                        tile, which leaves the other SMs to the neighbouring GEMMs.
void plan(int M, int K, int N, int split_hint, unsigned flags, int* tok_out, int* split_out) {
  (void)split_hint;   // the reference's split_k_iters is accepted but only a hint
  if (flags & QB200_GEMM_INDEPENDENT) {
    *tok_out = M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256;
    *split_out = 1;
    return;
  }
  const int tok = M <= 16 ? 16 : M <= 64 ? 64 : M <= 256 ? 128 : 256;   // 32-token tiles never won a measurement
  const int tiles = (N / 128) * ((M + tok - 1) / tok);
  const int KB = K / 64;
  const int sms = device_sm_count();
  const int max_split = tok <= 32 ? 8 : 4;
  int split = 1;
  if (tok <= 64) {
    // Co-resident tiles (two CTAs per SM): pick the split with the smallest modelled time
    //   waves(tiles*split over 2*SMs slots) x (stages per CTA x ~600 cycles per stage + fixed cost),
    // fixed = setup + epilogue (~2500 cycles) + ~300 per extra cluster rank for the split-K exchange.  Constants
    // from the in-kernel traces (profiles/README.md); the model reproduces the measured optimum for N = 4096
    // (split 8 for 16-token tiles, 4 for 64) and for the wide / deep projections of a 7B layer (qkv, gate|up: 2;
    // down: 8).
    const int slots = 2 * sms;
    long best = -1;
    for (int s = 1; s <= max_split; s *= 2) {
      if (s > 1 && KB / s < 4) break;                      // every rank keeps >= 2 stages
      const int stages = ((KB + s - 1) / s + 1) / 2;
      // a partly filled last wave costs less than a full one but more than its share: midpoint of both (x2)
      const long ctas = static_cast<long>(tiles) * s;
      const long waves2 = (ctas + slots - 1) / slots * slots + (ctas > slots ? ctas : slots);   // (ceil + continuous) * slots
      const long cost = waves2 * (stages * 600L + 2500L + 300L * (s - 1));
      if (best < 0 || cost < best) { best = cost; split = s; }
    }
    // More tiles than CTA slots (Llama-2-70B gate|up on one GPU: 448 tiles, K = 8192): the model over-rates the fixed
    // cost of a CTA once the SMs are throughput-bound (a starting CTA's prologue overlaps its neighbour's streaming);
    // measured 77.9 us unsplit vs 70.1 with split 2 (tools/tune_shapes.py, SHAPES=70b).
    if (split == 1 && tiles > slots && KB >= 64) split = 2;
  } else {
    // one CTA per SM: grow the cluster while the grid still under-fills the machine and every rank keeps >= 4 k-blocks
    while (split * 2 <= max_split && tiles * split * 2 <= sms + sms / 4 && KB / (split * 2) >= 4) split *= 2;
  }
  *tok_out = tok;
  *split_out = split;
}

}  // namespace

namespace {
int make_peer_dst(PeerDst* d, void* const* peers, void* mc, int n_peers, int ld, int col0, int width) {
  if (n_peers < 1 || n_peers > 8 || peers == nullptr) return fail(QB200_EINVAL, "1..8 peer buffers required");
  if (ld < col0 + width || col0 < 0 || (ld % 8) != 0 || (col0 % 8) != 0 || (width % 8) != 0) return fail(QB200_EINVAL, "bad ld / col0 / width");
  for (int p = 0; p < 8; ++p) {
    d->peer[p] = p < n_peers ? reinterpret_cast<__half*>(peers[p]) : nullptr;
    if (p < n_peers && (peers[p] == nullptr || (reinterpret_cast<uintptr_t>(peers[p]) & 15))) return fail(QB200_EINVAL, "peer buffer %d is null or unaligned", p);
  }
  if (mc != nullptr && (reinterpret_cast<uintptr_t>(mc) & 15)) return fail(QB200_EINVAL, "multicast pointer unaligned");
  d->mc = reinterpret_cast<__half*>(mc);
  d->n = n_peers; d->ld = ld; d->col0 = col0;
  return QB200_OK;
}
}  // namespace

namespace {
int gemm_impl(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, const void* residual, void* C,
              void* const* C_peers, int n_peers, int ld_c, int col0, int M, int K, int N, int G, int tok, int split,
              unsigned flags, void* stream, const qb200_peer_wait* wait = nullptr, const qb200_peer_signal* signal = nullptr,
              const qb200_norm_fusion* norm = nullptr);
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* qb200_version(void) { return "quick_b200 0.1 (sm_100a tcgen05/TMEM/TMA W4A16)"; }
const char* qb200_last_error(void) { return g_last_error.c_str(); }
unsigned long long qb200_launch_count(void) { return g_launches.load(); }
void qb200_debug_set_variant(int variant) { g_variant.store(variant); }   // -1 = planner's choice
void qb200_debug_set_trace(void* device_buffer) {
  g_trace = reinterpret_cast<long long*>(device_buffer);
  // the same buffer (host-mapped if it should survive a trap) receives the timed-out-wait report
  unsigned long long* p = reinterpret_cast<unsigned long long*>(device_buffer);
  cudaMemcpyToSymbol(qb200::g_qb_timeout_report, &p, sizeof(p));
}

size_t qb200_wq_bytes(int K, int N) { return static_cast<size_t>(K) * N / 2; }
size_t qb200_sz_bytes(int K, int N, int G) { return static_cast<size_t>(K / G) * N * 4; }

int qb200_check_shape(int M, int K, int N, int G) {
  if (M < 0 || K <= 0 || N <= 0 || G <= 0) return fail(QB200_EINVAL, "non-positive dimension");
  if (N % 128 != 0) return fail(QB200_EINVAL, "OC is not multiple of cta_N = 128");
  if (N % 8 != 0) return fail(QB200_EINVAL, "OC is not multiple of pack_num = 8");
  if (G % 32 != 0) return fail(QB200_EINVAL, "Group size should be a multiple of 32");
  if (K % G != 0) return fail(QB200_EINVAL, "IC is not a multiple of the group size");
  if (K % 64 != 0) return fail(QB200_EINVAL, "IC is not a multiple of 64");
  return QB200_OK;
}

int qb200_relayout_from_quick(const int32_t* qweight, const int32_t* qzeros, const void* scales, int K, int N, int G,
                              uint32_t* wq, uint32_t* sz, void* stream) {
  int rc = qb200_check_shape(1, K, N, G);
  if (rc) return rc;
  const size_t words = static_cast<size_t>(K) * N / 8;
  relayout_wq_kernel<<<static_cast<unsigned>((words + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint32_t*>(qweight), wq, K, N);
  const size_t nsz = static_cast<size_t>(K / G) * N;
  relayout_sz_kernel<<<static_cast<unsigned>((nsz + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint32_t*>(qzeros), reinterpret_cast<const __half*>(scales), sz, K / G, N);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  note_relayout(as_stream(stream));   // wq/sz are being written: the next GEMM on this stream launches fully serialised
  QB_CUDA(cudaGetLastError());
  return QB200_OK;
}

int qb200_pack_quick(const uint8_t* q, const uint8_t* z, const void* s, int K, int N, int G, int32_t* qweight,
                     int32_t* qzeros, void* scales, void* stream) {
  int rc = qb200_check_shape(1, K, N, G);
  if (rc) return rc;
  const size_t words = static_cast<size_t>(K) * N / 8;
  pack_qweight_kernel<<<static_cast<unsigned>((words + 255) / 256), 256, 0, as_stream(stream)>>>(
      q, reinterpret_cast<uint32_t*>(qweight), K, N);
  const size_t nz = static_cast<size_t>(K / G) * (N / 4);
  pack_sz_kernel<<<static_cast<unsigned>((nz + 255) / 256), 256, 0, as_stream(stream)>>>(
      z, reinterpret_cast<const __half*>(s), reinterpret_cast<uint32_t*>(qzeros), reinterpret_cast<__half*>(scales), K / G, N);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  QB_CUDA(cudaGetLastError());
  return QB200_OK;
}

int qb200_awq_gemm_to_quick(const int32_t* gemm_qweight, const int32_t* gemm_qzeros, const void* gemm_scales, int K, int N,
                            int G, int32_t* qweight, int32_t* qzeros, void* scales, void* stream) {
  int rc = qb200_check_shape(1, K, N, G);
  if (rc) return rc;
  const size_t words = static_cast<size_t>(K) * N / 8;
  gemm_to_quick_qweight_kernel<<<static_cast<unsigned>((words + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint32_t*>(gemm_qweight), reinterpret_cast<uint32_t*>(qweight), K, N);
  const size_t nz = static_cast<size_t>(K / G) * (N / 4);
  gemm_to_quick_sz_kernel<<<static_cast<unsigned>((nz + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint32_t*>(gemm_qzeros), reinterpret_cast<const __half*>(gemm_scales),
      reinterpret_cast<uint32_t*>(qzeros), reinterpret_cast<__half*>(scales), K / G, N);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  QB_CUDA(cudaGetLastError());
  return QB200_OK;
}

int qb200_relayout_from_awq_gemm(const int32_t* gemm_qweight, const int32_t* gemm_qzeros, const void* gemm_scales, int K,
                                 int N, int G, uint32_t* wq, uint32_t* sz, void* stream) {
  int rc = qb200_check_shape(1, K, N, G);
  if (rc) return rc;
  const size_t words = static_cast<size_t>(K) * N / 8;
  gemm_to_b200_wq_kernel<<<static_cast<unsigned>((words + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint32_t*>(gemm_qweight), wq, K, N);
  const size_t nsz = static_cast<size_t>(K / G) * N;
  gemm_to_b200_sz_kernel<<<static_cast<unsigned>((nsz + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint32_t*>(gemm_qzeros), reinterpret_cast<const __half*>(gemm_scales), sz, K / G, N);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  note_relayout(as_stream(stream));   // wq/sz are being written: the next GEMM on this stream launches fully serialised
  QB_CUDA(cudaGetLastError());
  return QB200_OK;
}

int qb200_dequantize(const uint32_t* wq, const uint32_t* sz, int K, int N, int G, void* w16, void* stream) {
  int rc = qb200_check_shape(1, K, N, G);
  if (rc) return rc;
  const size_t words = static_cast<size_t>(K) * N / 8;
  dequant_kernel<<<static_cast<unsigned>((words + 255) / 256), 256, 0, as_stream(stream)>>>(
      wq, sz, reinterpret_cast<__half*>(w16), K, N, G);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QB_CUDA(cudaGetLastError());
  return QB200_OK;
}

int qb200_gemm_plan(int M, int K, int N, int G, int split_k_hint, int* tok, int* split, int* ctas) {
  return qb200_gemm_plan_ex(M, K, N, G, split_k_hint, 0u, tok, split, ctas);
}

int qb200_gemm_plan_ex(int M, int K, int N, int G, int split_k_hint, unsigned flags, int* tok, int* split, int* ctas) {
  int rc = qb200_check_shape(M, K, N, G);
  if (rc) return rc;
  int t, s;
  plan(M, K, N, split_k_hint, flags, &t, &s);
  if (tok) *tok = t;
  if (split) *split = s;
  if (ctas) *ctas = (N / 128) * ((M + t - 1) / t) * s;
  return QB200_OK;
}

int qb200_gemm_w4a16_cfg(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, void* C, int M, int K,
                         int N, int G, int tok, int split, void* stream) {
  return qb200_gemm_w4a16_ex(A, wq, sz, bias, C, M, K, N, G, tok, split, 0u, stream);
}

int qb200_gemm_w4a16_ex(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, void* C, int M, int K,
                        int N, int G, int tok, int split, unsigned flags, void* stream) {
  return qb200_gemm_w4a16_fused(A, wq, sz, bias, nullptr, C, M, K, N, G, tok, split, flags, stream);
}

int qb200_gemm_w4a16_fused(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, const void* residual,
                           void* C, int M, int K, int N, int G, int tok, int split, unsigned flags, void* stream) {
  return gemm_impl(A, wq, sz, bias, residual, C, nullptr, 0, N, 0, M, K, N, G, tok, split, flags, stream);
}

int qb200_gemm_w4a16_norm(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, const void* residual, void* C,
                          int M, int K, int N, int G, int tok, int split, unsigned flags, const qb200_norm_fusion* norm,
                          void* stream) {
  return gemm_impl(A, wq, sz, bias, residual, C, nullptr, 0, N, 0, M, K, N, G, tok, split, flags, stream, nullptr, nullptr, norm);
}

int qb200_gemm_w4a16_allgather(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, const void* residual,
                               void* const* C_peers, void* C_multicast, int n_peers, int ld_c, int col0, int M, int K, int N,
                               int G, int tok, int split, unsigned flags, void* stream) {
  if (n_peers < 1 || n_peers > 8 || C_peers == nullptr) return fail(QB200_EINVAL, "allgather: 1..8 peer buffers required");
  if (ld_c < col0 + N || col0 < 0 || (ld_c % 8) != 0 || (col0 % 8) != 0) return fail(QB200_EINVAL, "allgather: bad ld_c / col0");
  for (int p = 0; p < n_peers; ++p)
    if (C_peers[p] == nullptr || (reinterpret_cast<uintptr_t>(C_peers[p]) & 15)) return fail(QB200_EINVAL, "allgather: peer buffer %d is null or unaligned", p);
  if (C_multicast != nullptr && (reinterpret_cast<uintptr_t>(C_multicast) & 15)) return fail(QB200_EINVAL, "allgather: multicast pointer unaligned");
  return gemm_impl(A, wq, sz, bias, residual, C_multicast, C_peers, n_peers, ld_c, col0, M, K, N, G, tok, split, flags, stream);
}

int qb200_gemm_w4a16_tp(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, const void* residual,
                        void* C_local, void* const* C_peers, void* C_multicast, int n_peers, int ld_c, int col0, int M, int K,
                        int N, int G, int tok, int split, unsigned flags, const qb200_peer_wait* wait,
                        const qb200_peer_signal* signal, void* stream) {
  if (n_peers == 0) {   // local output (row stride N), possibly reading a gathered buffer
    if (C_local == nullptr) return fail(QB200_EINVAL, "tp gemm: C_local required when n_peers = 0");
    return gemm_impl(A, wq, sz, bias, residual, C_local, nullptr, 0, N, 0, M, K, N, G, tok, split, flags, stream, wait, signal);
  }
  if (n_peers < 1 || n_peers > 8 || C_peers == nullptr) return fail(QB200_EINVAL, "tp gemm: 1..8 peer buffers required");
  const int n_out = (flags & QB200_GEMM_SILU_MUL) ? N / 2 : N;
  if (ld_c < col0 + n_out || col0 < 0 || (ld_c % 8) != 0 || (col0 % 8) != 0) return fail(QB200_EINVAL, "tp gemm: bad ld_c / col0");
  for (int p = 0; p < n_peers; ++p)
    if (C_peers[p] == nullptr || (reinterpret_cast<uintptr_t>(C_peers[p]) & 15)) return fail(QB200_EINVAL, "tp gemm: peer buffer %d is null or unaligned", p);
  if (C_multicast != nullptr && (reinterpret_cast<uintptr_t>(C_multicast) & 15)) return fail(QB200_EINVAL, "tp gemm: multicast pointer unaligned");
  return gemm_impl(A, wq, sz, bias, residual, C_multicast, C_peers, n_peers, ld_c, col0, M, K, N, G, tok, split, flags, stream, wait, signal);
}

}  // extern "C" (reopened below)

namespace {
int gemm_impl(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, const void* residual, void* C,
              void* const* C_peers, int n_peers, int ld_c, int col0, int M, int K, int N, int G, int tok, int split,
              unsigned flags, void* stream, const qb200_peer_wait* wait, const qb200_peer_signal* signal,
              const qb200_norm_fusion* norm) {
  int rc = qb200_check_shape(M, K, N, G);
  if (rc) return rc;
  if (M == 0) return (wait || signal) ? fail(QB200_EINVAL, "tensor-parallel GEMM with M = 0") : QB200_OK;
  rc = check_peer_args(wait, signal);
  if (rc) return rc;
  if ((flags & QB200_GEMM_INDEPENDENT) && ((wait != nullptr && wait->epoch != nullptr) || (signal != nullptr && signal->epoch != nullptr)))
    return fail(QB200_EINVAL, "an independent launch cannot take part in a peer hand-over (it does not wait for its own predecessor)");
  rc = check_device();
  if (rc) return rc;
  if (flags & ~(QB200_GEMM_INDEPENDENT | QB200_GEMM_SILU_MUL)) return fail(QB200_EINVAL, "unknown flags 0x%x", flags);
  if (flags & QB200_GEMM_SILU_MUL) {
    if (residual != nullptr) return fail(QB200_EINVAL, "QB200_GEMM_SILU_MUL takes no residual");
    if (n_peers == 0) ld_c = N / 2;          // local output [M][N/2]
    if ((reinterpret_cast<uintptr_t>(C) & 15) && n_peers == 0) return fail(QB200_EINVAL, "C must be 16-byte aligned");
  }
  if (tok == 0 || split == 0) {
    int t, s;
    plan(M, K, N, 0, flags, &t, &s);
    if (tok == 0) tok = t;
    if (split == 0) split = s;
  }
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(wq) & 15))
    return fail(QB200_EINVAL, "A and wq must be 16-byte aligned");
  if (residual != nullptr && ((reinterpret_cast<uintptr_t>(residual) & 15) || (reinterpret_cast<uintptr_t>(C) & 15)))
    return fail(QB200_EINVAL, "residual and C must be 16-byte aligned");
  if ((ld_c % 8) != 0 && (residual != nullptr || n_peers > 0)) return fail(QB200_EINVAL, "row stride must be a multiple of 8");
  const int KB = K / 64;
  if (split < 1 || split > KB) return fail(QB200_EINVAL, "split %d out of range for K=%d", split, K);
  const int kbps = (KB + split - 1) / split;
  if ((split - 1) * kbps >= KB) return fail(QB200_EINVAL, "split %d leaves an empty k-range for K=%d", split, K);
  const int m_tiles = (M + tok - 1) / tok;
  if (m_tiles > 65535) return fail(QB200_EINVAL, "M too large for one launch");

  CUtensorMap map;
  rc = make_x_tensor_map(&map, A, M, K, tok);
  if (rc) return rc;
  qb200::GemmArgs args;
  args.wq = wq;
  args.sz = sz;
  args.bias = reinterpret_cast<const __half*>(bias);
  args.residual = reinterpret_cast<const __half*>(residual);
  args.C = n_peers > 0 ? nullptr : reinterpret_cast<__half*>(C);
  args.mcC = n_peers > 0 ? reinterpret_cast<__half*>(C) : nullptr;   // gemm_impl's C doubles as the multicast pointer in peer mode
  args.n_peers = n_peers;
  args.ldc = ld_c;
  args.col0 = col0;
  for (int p = 0; p < 8; ++p) args.peerC[p] = p < n_peers ? reinterpret_cast<__half*>(C_peers[p]) : nullptr;
  args.M = M;
  args.K = K;
  args.N = N;
  args.G = G;
  args.kb_per_split = kbps;
  args.flags = flags;
  args.wait = make_wait(wait);
  args.signal = make_signal(signal);
  args.norm_gamma = nullptr; args.norm_out = nullptr; args.ssq_out = nullptr; args.ssq_in = nullptr; args.ssq_parts = 0; args.rms_eps = 0.f;
  if (norm != nullptr && norm->gamma_fp16 != nullptr) {
    if ((flags & (QB200_GEMM_SILU_MUL | QB200_GEMM_INDEPENDENT)) || n_peers > 0)
      return fail(QB200_EINVAL, "fused RMSNorm (producer side) needs a plain local output: no SiLU, no gathered buffer, ordered launch");
    if (norm->normed_out_fp16 == nullptr || norm->ssq_out == nullptr ||
        ((reinterpret_cast<uintptr_t>(norm->gamma_fp16) | reinterpret_cast<uintptr_t>(norm->normed_out_fp16) | reinterpret_cast<uintptr_t>(C)) & 15))
      return fail(QB200_EINVAL, "fused RMSNorm: gamma, normed_out (16-byte aligned) and ssq_out required");
    args.norm_gamma = reinterpret_cast<const __half*>(norm->gamma_fp16);
    args.norm_out = reinterpret_cast<__half*>(norm->normed_out_fp16);
    args.ssq_out = norm->ssq_out;
  }
  if (norm != nullptr && norm->ssq_in != nullptr) {
    if ((flags & QB200_GEMM_INDEPENDENT) || (wait != nullptr && wait->epoch != nullptr))
      return fail(QB200_EINVAL, "fused RMSNorm (consumer side) needs an ordered, local launch");
    if (norm->ssq_parts != K / 128 || K % 128 != 0) return fail(QB200_EINVAL, "fused RMSNorm: ssq_parts must be K / 128");
    args.ssq_in = norm->ssq_in;
    args.ssq_parts = norm->ssq_parts;
    args.rms_eps = norm->eps;
  }
  args.launch_id = static_cast<unsigned>(g_launches.load(std::memory_order_relaxed));
  args.trace = g_trace;
  cudaStream_t st = as_stream(stream);
  switch (tok) {
    case 16: return dispatch_variant<16>(split, map, args, m_tiles, st);
    case 32: return dispatch_variant<32>(split, map, args, m_tiles, st);
    case 64: return dispatch_variant<64>(split, map, args, m_tiles, st);
    case 128: return dispatch_variant<128>(split, map, args, m_tiles, st);
    case 256: return dispatch_variant<256>(split, map, args, m_tiles, st);
  }
  return fail(QB200_EINVAL, "unsupported token tile %d", tok);
}
}  // namespace

extern "C" {

int qb200_gemm_w4a16(const void* A, const uint32_t* wq, const uint32_t* sz, const void* bias, void* C, int M, int K,
                     int N, int G, int split_k_hint, void* stream) {
  int rc = qb200_check_shape(M, K, N, G);
  if (rc) return rc;
  int tok, split;
  plan(M, K, N, split_k_hint, 0u, &tok, &split);
  return qb200_gemm_w4a16_cfg(A, wq, sz, bias, C, M, K, N, G, tok, split, stream);
}

int qb200_gemm_forward_quick(const void* A, const int32_t* qweight, const void* scales, const int32_t* qzeros, void* C,
                             int M, int K, int N, int G, int split_k_iters, void* workspace, size_t workspace_bytes,
                             void* stream) {
  int rc = qb200_check_shape(M, K, N, G);
  if (rc) return rc;
  const size_t wqb = qb200_wq_bytes(K, N), szb = qb200_sz_bytes(K, N, G);
  if (workspace == nullptr || workspace_bytes < wqb + szb)
    return fail(QB200_ENOSPC, "workspace needs %zu bytes", wqb + szb);
  uint32_t* wq = reinterpret_cast<uint32_t*>(workspace);
  uint32_t* sz = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + wqb);
  rc = qb200_relayout_from_quick(qweight, qzeros, scales, K, N, G, wq, sz, stream);
  if (rc) return rc;
  return qb200_gemm_w4a16(A, wq, sz, nullptr, C, M, K, N, G, split_k_iters, stream);
}

int qb200_gemm_w4a16_simt(const void* A, const uint32_t* wq, const uint32_t* sz, void* C, int M, int K, int N, int G,
                          void* stream) {
  int rc = qb200_check_shape(M, K, N, G);
  if (rc) return rc;
  if (M == 0) return QB200_OK;
  if (static_cast<size_t>(K) * 2 > 48 * 1024) {
    QB_CUDA(cudaFuncSetAttribute(gemm_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K * 2));
  }
  gemm_simt_kernel<<<dim3(N / 128, M), 128, static_cast<size_t>(K) * 2, as_stream(stream)>>>(
      reinterpret_cast<const __half*>(A), wq, sz, reinterpret_cast<__half*>(C), M, K, N, G);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QB_CUDA(cudaGetLastError());
  return QB200_OK;
}

int qb200_peer_barrier(unsigned* epoch_counter, unsigned* const* flag_arrays, int rank, int n_peers, void* stream) {
  if (n_peers < 1 || n_peers > 8 || rank < 0 || rank >= n_peers || epoch_counter == nullptr || flag_arrays == nullptr)
    return fail(QB200_EINVAL, "peer_barrier: bad arguments");
  PeerFlagPtrs f{};
  for (int p = 0; p < n_peers; ++p) {
    if (flag_arrays[p] == nullptr) return fail(QB200_EINVAL, "peer_barrier: null flag array %d", p);
    f.p[p] = flag_arrays[p];
  }
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("QB200_PEER_TIMEOUT_S");
    const double sec = e ? atof(e) : 120.0;
    return sec <= 0 ? 0ull : static_cast<unsigned long long>(sec * 1e9);
  }();
  peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(epoch_counter, f, rank, n_peers, timeout_ns);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QB_CUDA(cudaGetLastError());
  return QB200_OK;
}

int qb200_rmsnorm_tp(const void* x, const void* weight, void* y, int rows, int H, float eps, const qb200_peer_wait* wait,
                     void* stream) {
  { const int prc = check_peer_args(wait, nullptr); if (prc) return prc; }
  if (rows < 0 || H <= 0 || H % 8 != 0) return fail(QB200_EINVAL, "rmsnorm: H must be a positive multiple of 8");
  if (rows == 0) return QB200_OK;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(y)) & 15)
    return fail(QB200_EINVAL, "rmsnorm: pointers must be 16-byte aligned");
  const int threads = H >= 4096 ? 512 : H >= 1024 ? 128 : 64;
  QB_CUDA(launch_pdl(rmsnorm_kernel, dim3(rows), dim3(threads), as_stream(stream), reinterpret_cast<const __half*>(x),
                     reinterpret_cast<const __half*>(weight), reinterpret_cast<__half*>(y), H, eps, make_wait(wait)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QB200_OK;
}

int qb200_rmsnorm(const void* x, const void* weight, void* y, int rows, int H, float eps, void* stream) {
  return qb200_rmsnorm_tp(x, weight, y, rows, H, eps, nullptr, stream);
}

int qb200_rope_kv_update(const void* qkv, const void* cos_table, const void* sin_table, const long long* pos, void* q_out,
                         void* cache_k, void* cache_v, int B, int T, int nh, int nkv, int hd, int S, void* stream) {
  if (B <= 0 || T <= 0 || nh <= 0 || nkv <= 0 || hd <= 0 || hd % 16 != 0 || S <= 0)
    return fail(QB200_EINVAL, "rope_kv_update: bad dimensions (the head dimension must be a multiple of 16)");
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(cos_table) | reinterpret_cast<uintptr_t>(sin_table) |
       reinterpret_cast<uintptr_t>(q_out) | reinterpret_cast<uintptr_t>(cache_k) | reinterpret_cast<uintptr_t>(cache_v)) & 15)
    return fail(QB200_EINVAL, "rope_kv_update: pointers must be 16-byte aligned");
  const long long items = static_cast<long long>(B) * T * (nh + 2 * nkv) * (hd / 16);
  const unsigned blocks = static_cast<unsigned>(std::min<long long>((items + 255) / 256, 8LL * device_sm_count()));
  QB_CUDA(launch_pdl(rope_kv_kernel, dim3(blocks), dim3(256), as_stream(stream),
                     reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(cos_table),
                     reinterpret_cast<const __half*>(sin_table), pos, reinterpret_cast<__half*>(q_out),
                     reinterpret_cast<__half*>(cache_k), reinterpret_cast<__half*>(cache_v), B, T, nh, nkv, hd, S));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QB200_OK;
}

int qb200_attn_decode_smem_bytes(int nh, int nkv, int hd, int S) {
  if (nh <= 0 || nkv <= 0 || nh % nkv != 0 || (hd != 64 && hd != 128 && hd != 256) || S <= 0) return -1;
  const int grp = nh / nkv;
  if (grp != 1 && grp != 2 && grp != 4 && grp != 8) return -1;     // instantiated query-group sizes
  return static_cast<int>(attn_decode_smem(grp, hd));               // independent of the cache length
}

static int attn_decode_impl(const void* qkv, const void* cos_table, const void* sin_table, const long long* pos, void* out,
                            void* cache_k, void* cache_v, int B, int nh, int nkv, int hd, int S, float scale, const PeerDst& dst,
                            const qb200_peer_signal* signal, void* stream);

int qb200_attn_decode(const void* qkv, const void* cos_table, const void* sin_table, const long long* pos, void* out,
                      void* cache_k, void* cache_v, int B, int nh, int nkv, int hd, int S, float scale, void* stream) {
  PeerDst none{};
  return attn_decode_impl(qkv, cos_table, sin_table, pos, out, cache_k, cache_v, B, nh, nkv, hd, S, scale, none, nullptr, stream);
}

int qb200_attn_decode_tp(const void* qkv, const void* cos_table, const void* sin_table, const long long* pos, void* cache_k,
                         void* cache_v, int B, int nh, int nkv, int hd, int S, float scale, void* const* out_peers, int n_peers,
                         int ld, int col0, const qb200_peer_signal* signal, void* stream) {
  PeerDst d;
  int rc = make_peer_dst(&d, out_peers, nullptr, n_peers, ld, col0, nh * hd);
  if (rc) return rc;
  return attn_decode_impl(qkv, cos_table, sin_table, pos, nullptr, cache_k, cache_v, B, nh, nkv, hd, S, scale, d, signal, stream);
}

static int attn_decode_impl(const void* qkv, const void* cos_table, const void* sin_table, const long long* pos, void* out,
                            void* cache_k, void* cache_v, int B, int nh, int nkv, int hd, int S, float scale, const PeerDst& dst,
                            const qb200_peer_signal* signal, void* stream) {
  const int smem = qb200_attn_decode_smem_bytes(nh, nkv, hd, S);
  if (B <= 0 || B > 65535 || smem < 0)
    return fail(QB200_EINVAL, "attn_decode: needs nh / nkv in {1, 2, 4, 8} and hd in {64, 128, 256}");
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(cache_k) | reinterpret_cast<uintptr_t>(cache_v)) & 15)
    return fail(QB200_EINVAL, "attn_decode: pointers must be 16-byte aligned");
  using AttnFn = void (*)(const __half*, const __half*, const __half*, const long long*, __half*, __half*, __half*, int, int, int, float,
                          int, const PeerDst, const qb200::PeerSignal);
  static const AttnFn table[3][4] = {
      {attn_decode_kernel<64, 1>, attn_decode_kernel<64, 2>, attn_decode_kernel<64, 4>, attn_decode_kernel<64, 8>},
      {attn_decode_kernel<128, 1>, attn_decode_kernel<128, 2>, attn_decode_kernel<128, 4>, attn_decode_kernel<128, 8>},
      {attn_decode_kernel<256, 1>, attn_decode_kernel<256, 2>, attn_decode_kernel<256, 4>, attn_decode_kernel<256, 8>}};
  static PerDeviceOnce attr_once[12];
  const int grp = nh / nkv;
  const int hi = hd == 64 ? 0 : hd == 128 ? 1 : 2, gi = grp == 1 ? 0 : grp == 2 ? 1 : grp == 4 ? 2 : 3;
  const AttnFn kfn = table[hi][gi];
  const int ki = hi * 4 + gi;
  QB_CUDA(attr_once[ki].ensure([&] { return cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); }));
  // Positions of one (kv head, sequence) split over a cluster of nsplit CTAs: the largest cluster for which every CTA of
  // the launch is still resident at once (SMs x occupancy of this instantiation) — the kernel is latency-bound, a
  // second wave costs more than shorter per-CTA ranges gain (tools/tune_attn.py) — and at least 16 cache positions per
  // CTA.  The cache LENGTH decides, not the current position (a device value under CUDA-graph replay).
  int nsplit = 1;
  {
    static int occ[12][64];                                          // CTAs per SM of the instantiation, per device
    const int dev = current_device();
    int* o = &occ[ki][dev >= 0 && dev < 64 ? dev : 0];
    if (*o == 0) {
      int n = 0;
      QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kfn, kAttnThreads, static_cast<size_t>(smem)));
      *o = n > 0 ? n : 1;
    }
    const long long resident = static_cast<long long>(device_sm_count()) * *o;
    const long long base = static_cast<long long>(nkv) * B;
    while (nsplit < kAttnMaxSplit && base * nsplit * 2 <= resident && S / (nsplit * 2) >= 16) nsplit *= 2;
    const char* env = std::getenv("QB200_ATTN_SPLIT");             // tests / tuning: force the cluster size
    const int forced = env ? std::atoi(env) : 0;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) nsplit = forced;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nkv, B, nsplit);
  cfg.blockDim = dim3(kAttnThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (nsplit > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = static_cast<unsigned>(nsplit);
    ++na;
  }
  if (use_pdl()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  QB_CUDA(cudaLaunchKernelEx(&cfg, kfn, reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(cos_table),
                             reinterpret_cast<const __half*>(sin_table), pos, reinterpret_cast<__half*>(out),
                             reinterpret_cast<__half*>(cache_k), reinterpret_cast<__half*>(cache_v), nh, nkv, S, scale, nsplit, dst,
                             make_signal(signal)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QB200_OK;
}

int qb200_silu_mul(const void* gate_up, void* act, long long rows, int I, void* stream) {
  if (rows < 0 || I <= 0 || I % 8 != 0) return fail(QB200_EINVAL, "silu_mul: I must be a positive multiple of 8");
  if (rows == 0) return QB200_OK;
  if ((reinterpret_cast<uintptr_t>(gate_up) | reinterpret_cast<uintptr_t>(act)) & 15) return fail(QB200_EINVAL, "silu_mul: pointers must be 16-byte aligned");
  const size_t vecs = static_cast<size_t>(rows) * I / 8;
  QB_CUDA(launch_pdl(silu_mul_kernel, dim3(static_cast<unsigned>((vecs + 255) / 256)), dim3(256), as_stream(stream),
                     reinterpret_cast<const __half*>(gate_up), reinterpret_cast<__half*>(act), static_cast<size_t>(rows), I));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QB200_OK;
}


int qb200_silu_mul_interleaved(const void* gate_up, void* act, long long rows, int I, void* stream) {
  if (rows < 0 || I <= 0 || I % 8 != 0) return fail(QB200_EINVAL, "silu_mul_interleaved: I must be a positive multiple of 8");
  if (rows == 0) return QB200_OK;
  if ((reinterpret_cast<uintptr_t>(gate_up) | reinterpret_cast<uintptr_t>(act)) & 15) return fail(QB200_EINVAL, "silu_mul_interleaved: pointers must be 16-byte aligned");
  const size_t total = static_cast<size_t>(rows) * I;
  QB_CUDA(launch_pdl(silu_mul_pairs_kernel, dim3(static_cast<unsigned>((total / 8 + 255) / 256)), dim3(256), as_stream(stream),
                     reinterpret_cast<const __half*>(gate_up), reinterpret_cast<__half*>(act), total));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QB200_OK;
}



int qb200_silu_mul_tp(const void* gate_up, long long rows, int I, void* const* act_peers, void* act_multicast, int n_peers, int ld,
                      int col0, const qb200_peer_signal* signal, void* stream) {
  if (rows <= 0 || I <= 0 || I % 8 != 0) return fail(QB200_EINVAL, "silu_mul_tp: rows > 0 and I a positive multiple of 8 required");
  if (reinterpret_cast<uintptr_t>(gate_up) & 15) return fail(QB200_EINVAL, "silu_mul_tp: pointers must be 16-byte aligned");
  int rc = check_peer_args(nullptr, signal);
  if (rc) return rc;
  PeerDst d;
  rc = make_peer_dst(&d, act_peers, act_multicast, n_peers, ld, col0, I);
  if (rc) return rc;
  const size_t vecs = static_cast<size_t>(rows) * I / 8;
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>((vecs + 255) / 256, 2 * static_cast<size_t>(device_sm_count())));
  QB_CUDA(launch_pdl(silu_mul_scatter_kernel, dim3(blocks), dim3(256), as_stream(stream), reinterpret_cast<const __half*>(gate_up),
                     static_cast<size_t>(rows), I, d, make_signal(signal)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QB200_OK;
}

int qb200_scatter_cols(const void* src, long long rows, int n_local, void* const* dst_peers, void* dst_multicast, int n_peers, int ld,
                       int col0, const qb200_peer_signal* signal, void* stream) {
  if (rows <= 0 || n_local <= 0) return fail(QB200_EINVAL, "scatter_cols: rows and n_local must be positive");
  if (reinterpret_cast<uintptr_t>(src) & 15) return fail(QB200_EINVAL, "scatter_cols: pointers must be 16-byte aligned");
  int rc = check_peer_args(nullptr, signal);
  if (rc) return rc;
  PeerDst d;
  rc = make_peer_dst(&d, dst_peers, dst_multicast, n_peers, ld, col0, n_local);
  if (rc) return rc;
  const size_t vecs = static_cast<size_t>(rows) * n_local / 8;
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>((vecs + 255) / 256, 2 * static_cast<size_t>(device_sm_count())));
  QB_CUDA(launch_pdl(scatter_cols_kernel, dim3(blocks), dim3(256), as_stream(stream), reinterpret_cast<const __half*>(src),
                     static_cast<size_t>(rows), n_local, d, make_signal(signal)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QB200_OK;
}

// ---- host-buffer handle ----
struct qb200_linear {
  int K, N, G, max_m, device;
  uint32_t* wq;
  uint32_t* sz;
  __half* bias;
  __half* x;
  __half* y;
  cudaStream_t stream;
};

int qb200_linear_create(qb200_linear** out, const int32_t* qweight_host, const int32_t* qzeros_host,
                        const void* scales_host, const void* bias_host, int K, int N, int G, int max_m, int device) {
  if (!out) return fail(QB200_EINVAL, "null out");
  int rc = qb200_check_shape(max_m, K, N, G);
  if (rc) return rc;
  QB_CUDA(cudaSetDevice(device));
  rc = check_device();
  if (rc) return rc;
  qb200_linear* h = new qb200_linear();
  std::memset(h, 0, sizeof(*h));
  h->K = K; h->N = N; h->G = G; h->max_m = max_m; h->device = device;
  const size_t qw_b = static_cast<size_t>(K) * N / 2, qz_b = static_cast<size_t>(K / G) * N, sc_b = static_cast<size_t>(K / G) * N * 4;
  void *dqw = nullptr, *dqz = nullptr, *dsc = nullptr;
  QB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  QB_CUDA(cudaMalloc(&dqw, qw_b));
  QB_CUDA(cudaMalloc(&dqz, qz_b));
  QB_CUDA(cudaMalloc(&dsc, sc_b));
  QB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->wq), qb200_wq_bytes(K, N)));
  QB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->sz), qb200_sz_bytes(K, N, G)));
  QB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->x), static_cast<size_t>(max_m) * K * 2));
  QB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->y), static_cast<size_t>(max_m) * N * 2));
  QB_CUDA(cudaMemcpyAsync(dqw, qweight_host, qw_b, cudaMemcpyHostToDevice, h->stream));
  QB_CUDA(cudaMemcpyAsync(dqz, qzeros_host, qz_b, cudaMemcpyHostToDevice, h->stream));
  QB_CUDA(cudaMemcpyAsync(dsc, scales_host, sc_b, cudaMemcpyHostToDevice, h->stream));
  if (bias_host) {
    QB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->bias), static_cast<size_t>(N) * 2));
    QB_CUDA(cudaMemcpyAsync(h->bias, bias_host, static_cast<size_t>(N) * 2, cudaMemcpyHostToDevice, h->stream));
  }
  rc = qb200_relayout_from_quick(reinterpret_cast<int32_t*>(dqw), reinterpret_cast<int32_t*>(dqz), dsc, K, N, G, h->wq,
                                 h->sz, h->stream);
  if (rc) return rc;
  QB_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(dqw); cudaFree(dqz); cudaFree(dsc);
  *out = h;
  return QB200_OK;
}

int qb200_linear_forward(qb200_linear* h, const void* x_dev, void* y_dev, int M, void* stream) {
  if (!h) return fail(QB200_EINVAL, "null handle");
  return qb200_gemm_w4a16(x_dev, h->wq, h->sz, h->bias, y_dev, M, h->K, h->N, h->G, 0, stream);
}

int qb200_linear_forward_host(qb200_linear* h, const void* x_host, void* y_host, int M) {
  if (!h) return fail(QB200_EINVAL, "null handle");
  if (M > h->max_m) return fail(QB200_EINVAL, "M=%d exceeds the handle's max_m=%d", M, h->max_m);
  if (M == 0) return QB200_OK;
  QB_CUDA(cudaMemcpyAsync(h->x, x_host, static_cast<size_t>(M) * h->K * 2, cudaMemcpyHostToDevice, h->stream));
  int rc = qb200_gemm_w4a16(h->x, h->wq, h->sz, h->bias, h->y, M, h->K, h->N, h->G, 0, h->stream);
  if (rc) return rc;
  QB_CUDA(cudaMemcpyAsync(y_host, h->y, static_cast<size_t>(M) * h->N * 2, cudaMemcpyDeviceToHost, h->stream));
  QB_CUDA(cudaStreamSynchronize(h->stream));
  return QB200_OK;
}

int qb200_linear_forward_host_async(qb200_linear* h, const void* x_host, void* y_host, int M) {
  if (!h) return fail(QB200_EINVAL, "null handle");
  if (M > h->max_m) return fail(QB200_EINVAL, "M=%d exceeds the handle's max_m=%d", M, h->max_m);
  if (M == 0) return QB200_OK;
  // Everything is ordered on the handle's own stream, so the staging buffers are reused safely by the next
  // call on the same handle, while calls on OTHER handles (other streams) overlap their copies with this GEMM.
  QB_CUDA(cudaMemcpyAsync(h->x, x_host, static_cast<size_t>(M) * h->K * 2, cudaMemcpyHostToDevice, h->stream));
  int rc = qb200_gemm_w4a16(h->x, h->wq, h->sz, h->bias, h->y, M, h->K, h->N, h->G, 0, h->stream);
  if (rc) return rc;
  QB_CUDA(cudaMemcpyAsync(y_host, h->y, static_cast<size_t>(M) * h->N * 2, cudaMemcpyDeviceToHost, h->stream));
  return QB200_OK;
}

int qb200_linear_synchronize(qb200_linear* h) {
  if (!h) return fail(QB200_EINVAL, "null handle");
  QB_CUDA(cudaStreamSynchronize(h->stream));
  return QB200_OK;
}

void qb200_linear_destroy(qb200_linear* h) {
  if (!h) return;
  cudaFree(h->wq); cudaFree(h->sz); cudaFree(h->bias); cudaFree(h->x); cudaFree(h->y);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

}  // extern "C"
