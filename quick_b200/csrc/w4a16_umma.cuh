// quick_b200 — tcgen05 / TMEM / TMA W4A16 grouped GEMM for sm_100a.
//
// Computes  C[M][N] = A[M][K] (fp16) · W16[K][N],  W16 = fp16(q - z) * s  (one rounding,
// bit-identical to what the reference materialises in registers: csrc/gemm_cuda_quick.cu:52-60),
// fp32 accumulation in TMEM, one final fp16 rounding.
//
// Blackwell mapping (not a port of the reference's mma.sync fragment scheme):
//   * swap A/B:  D^T[128 channels][TOK tokens] += W^T[128][16] · X^T[16][TOK]
//       - UMMA M = 128 output channels = 128 TMEM lanes, UMMA N = TOK tokens, K = 16 / instruction
//       - the WEIGHTS are the A operand and are sourced from TMEM (tcgen05.mma [d],[a_tmem],b_desc):
//         dequantised fragments go registers -> tcgen05.st -> tensor core, never through shared memory
//       - the ACTIVATIONS are the B operand in shared memory (K-major, 128B swizzle) loaded by TMA
//   * packed weights in the "B200 layout" (see include/quick_b200.h): one thread = one TMEM lane =
//     one output channel, its 16-byte shared-memory read = 32 consecutive k = 16 TMEM columns
//   * warp roles: 4*NWG dequant warps (warpgroups take 128-k stages round-robin; the first two also run the
//     epilogue), one TMA producer warp, one MMA issuer warp (+ TMEM owner)
//   * three decoupled rings: W stages (shared memory, filled before griddepcontrol.wait), X stages (shared
//     memory), A slots (tensor memory)
//   * split-K lives inside a thread-block cluster (1,1,SPLIT): partial tiles are exchanged through
//     distributed shared memory (st.async register fragments for small tiles, TMA bulk copies for large
//     ones), each CTA reduces and stores TOK/SPLIT token columns; no HBM temp, no second kernel
//     (reference: (split_k,M,N) temp + at::sum, gemm_cuda_quick.cu:1468,1515)
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace qb200 {

constexpr int kChan = 128;          // channels per tile = UMMA M
constexpr int kBK = 64;             // k per pipeline stage
constexpr int kWStageBytes = kChan * kBK / 2;   // 4096
constexpr int kEpilogueWarps = 8;    // warps 0..7 run the epilogue
// The single-thread issuers get the HIGHEST warp ids: the SM's warp arbiter favours higher warp ids, and a
// starved TMA/MMA issuer stalls the whole pipeline.

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes)
               : "memory");
}
#ifndef QB200_WAIT_TIMEOUT_CYCLES
#define QB200_WAIT_TIMEOUT_CYCLES 2000000000ll  // ~1 s: a lost barrier traps instead of hanging the GPU
#endif
// Optional host-mapped buffer (qb200_debug_set_trace): a timed-out wait records who was waiting on what
// before trapping, so a protocol bug is diagnosable after the context is gone.
__device__ unsigned long long* g_qb_timeout_report = nullptr;
// Polls in a tight PTX loop (try_wait itself suspends the warp in hardware for a bounded time, so idle warps
// do not steal issue slots); the clock is only read on the slow path, once per 4096 polls.
__device__ __forceinline__ uint32_t mbar_poll(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 cnt;\n"
      "mov.u32 cnt, 0;\n"
      "QB_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "@p bra QB_WAIT_DONE;\n"
      "add.u32 cnt, cnt, 1;\n"
      "setp.lt.u32 p, cnt, 4096;\n"
      "@p bra QB_WAIT_LOOP;\n"
      "mov.u32 %0, 0;\n"
      "bra QB_WAIT_EXIT;\n"
      "QB_WAIT_DONE:\n"
      "mov.u32 %0, 1;\n"
      "QB_WAIT_EXIT:\n"
      "}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
__device__ __noinline__ void mbar_timeout_report(uint32_t bar, uint32_t parity, int tag, int iter) {
  unsigned long long* rep = g_qb_timeout_report;
  if (rep != nullptr && (threadIdx.x & 31) == 0) {
    // first block to time out claims the report; each of its warps fills its own row, then lingers so
    // the other warps of the block (stuck on the same lost event) can report before the trap
    const unsigned long long me = 1ull + ((static_cast<unsigned long long>(blockIdx.x) << 20) | (blockIdx.y << 10) | blockIdx.z);
    const unsigned long long prev = atomicCAS(rep, 0ull, me);
    if (prev == 0ull || prev == me) {
      unsigned long long st;
      asm volatile("ld.shared.u64 %0, [%1];" : "=l"(st) : "r"(bar) : "memory");
      unsigned long long* row = rep + 8 + (threadIdx.x >> 5) * 8;
      row[0] = tag; row[1] = iter; row[2] = bar; row[3] = parity; row[4] = st; row[5] = 1;
      __threadfence_system();
    }
    const long long t1 = clock64();
    while (clock64() - t1 < QB200_WAIT_TIMEOUT_CYCLES / 4) { }
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0, int iter = -1) {
  if (mbar_poll(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_poll(bar, parity)) {
    if (clock64() - t0 > QB200_WAIT_TIMEOUT_CYCLES) mbar_timeout_report(bar, parity, tag, iter);
  }
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 2-D tiled TMA load (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// TMEM management (one warp)
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[tmem] · B[smem desc]   (SASS: UTCHMMA)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// registers -> TMEM, 16 consecutive 32-bit columns of this thread's lane (SASS: STTM)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM -> registers, NCOL consecutive 32-bit columns of this thread's lane (SASS: LDTM)
template <int NCOL>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r) {
  if constexpr (NCOL == 1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
  } else if constexpr (NCOL == 2) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
  } else if constexpr (NCOL == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
  } else if constexpr (NCOL == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
  } else {
    static_assert(NCOL == 16, "tmem_ld: 1,2,4,8,16 columns");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
  }
}

// Accumulator columns -> registers (and wait).  With two issuers the tile's accumulator is D0 + D1 (D1 = D0 + d1_off
// columns; has_d1 is false when the second issuer had no stage, its columns are then uninitialised).
template <int NCOL, int NI>
__device__ __forceinline__ void tmem_ld_acc(uint32_t taddr, uint32_t d1_off, bool has_d1, uint32_t* v) {
  tmem_ld<NCOL>(taddr, v);
  if constexpr (NI == 2) {
    if (has_d1) {
      uint32_t u[NCOL];
      tmem_ld<NCOL>(taddr + d1_off, u);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < NCOL; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
      return;
    }
  }
  tmem_wait_ld();
}

// cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// 4-bit unpack.  One B200-layout word = 8 consecutive k of one channel, nibble order
// k0,k2,k4,k6,k1,k3,k5,k7, so the lop3 extraction yields (k0,k1) (k2,k3) (k4,k5) (k6,k7) as half2.
//   bottom nibbles: (w & 0x000f000f) | 0x6400_6400 = 1024 + q          -> sub (1024 + z)        = q - z
//   top    nibbles: (w & 0x00f000f0) | 0x6400_6400 = 1024 + 16 q       -> fma(·, 1/16, -(64+z)) = q - z
// both exact in fp16; then one mul.rn by the scale == the reference's sub.f16x2 + mul.rn.f16x2
// (gemm_cuda_quick.cu:53-54) on the reference's 1024+q / 1024+z operands (dequantize_quick.cuh:35-60).
// ------------------------------------------------------------------------------------------------
struct GroupConsts {
  uint32_t zb;   // (1024 + z) x2
  uint32_t zt;   // -(64 + z)  x2
  uint32_t sc;   // scale      x2
};
__device__ __forceinline__ GroupConsts make_group_consts(uint32_t szw) {
  GroupConsts g;
  asm("prmt.b32 %0, %1, %1, 0x1010;" : "=r"(g.sc) : "r"(szw));   // low half duplicated
  asm("prmt.b32 %0, %1, %1, 0x3232;" : "=r"(g.zb) : "r"(szw));   // high half duplicated
  const uint32_t k960 = 0x63806380u;                              // 960 x2
  asm("sub.f16x2 %0, %1, %2;" : "=r"(g.zt) : "r"(k960), "r"(g.zb));   // 960 - (1024+z) = -(64+z), exact
  return g;
}
__device__ __forceinline__ void dequant_word(uint32_t w, const GroupConsts& g, uint32_t* out) {
  constexpr uint32_t kLut = (0xf0 & 0xcc) | 0xaa;   // (a & b) | c
  const uint32_t top = w >> 8;
  uint32_t h0, h1, h2, h3;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h0) : "r"(w), "n"(0x000f000f), "n"(0x64006400), "n"(kLut));
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h1) : "r"(w), "n"(0x00f000f0), "n"(0x64006400), "n"(kLut));
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h2) : "r"(top), "n"(0x000f000f), "n"(0x64006400), "n"(kLut));
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h3) : "r"(top), "n"(0x00f000f0), "n"(0x64006400), "n"(kLut));
  const uint32_t k16th = 0x2c002c00u;   // 1/16 x2
  asm("sub.f16x2 %0, %1, %2;" : "=r"(h0) : "r"(h0), "r"(g.zb));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h1) : "r"(h1), "r"(k16th), "r"(g.zt));
  asm("sub.f16x2 %0, %1, %2;" : "=r"(h2) : "r"(h2), "r"(g.zb));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h3) : "r"(h3), "r"(k16th), "r"(g.zt));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(out[0]) : "r"(h0), "r"(g.sc));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(out[1]) : "r"(h1), "r"(g.sc));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(out[2]) : "r"(h2), "r"(g.sc));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(out[3]) : "r"(h3), "r"(g.sc));
}

// ------------------------------------------------------------------------------------------------
// Descriptors
// ------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, rows of 64 fp16 (128 B), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);   // start address
  d |= static_cast<uint64_t>(1) << 16;                     // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;             // stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                     // SWIZZLE_128B
  return d;
}
// Instruction descriptor: kind::f16, A = B = fp16, D = fp32, both K-major, M = 128, N = n.
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// Pipeline geometry.  A pipeline stage is 128 k (two k64 blocks): W nibbles 8 KB, X tile 2 x [TOK][128 B]
// swizzled panels, TMEM A slot 64 columns.  Two rings:
//   W ring (DS stages, shared memory)   — hides HBM latency; filled BEFORE griddepcontrol.wait, i.e. while the
//                                         previous kernel of the stream is still running (weights are constants)
//   operand ring (D2 slots)             — slot = X stage in shared memory (hides the L2 latency of the
//                                         activations, which come from the previous kernel) + A slot in tensor
//                                         memory (dequant -> MMA hand-off; the first D2 stages are dequantised
//                                         before the activations arrive)
// Per operand slot ONE "ready" barrier (4 dequant-warp arrivals + the producer's expect_tx arrival + the X
// bytes) and ONE "free" barrier (a single tcgen05.commit): the MMA issuer — the serial spine of the kernel —
// does one wait and one commit per stage.
// Small tiles (TOK <= 64) are sized so that two CTAs are co-resident on an SM (<= 72 registers x 448 threads,
// <= 256 TMEM columns, <= 113 KB shared memory): under programmatic dependent launch the next GEMM's CTAs sit
// next to the running ones with their weights already in shared memory and TMEM.
// ------------------------------------------------------------------------------------------------
// SUB = k64 blocks per pipeline stage (2 -> 128 k, 3 -> 192 k, 4 -> 256 k).  The MMA issuer pays a fixed ≈650 cycles
// per stage (mbarrier wait, commit, loop) whatever the stage holds, so tiles whose tensor work per 128 k is below
// that (TOK <= 128) want longer stages; the price is shared memory (X stage = SUB x TOK x 128 B) and TMEM
// (A slot = 32 x SUB columns).
// NI = number of MMA issuer warps.  One issuer iteration (mbarrier wait, eight tcgen05.mma, tcgen05.commit, warp
// re-convergence) costs ≈600-750 cycles whatever the tile holds (round-2 traces, profiles/README.md), more than the
// tensor work of a 128-k stage for every tile below 256 tokens.  Two issuers take alternate stages and accumulate
// into their own TMEM accumulator (D0 / D1, summed by the epilogue): the per-stage issue cost halves without
// ordering two threads' tcgen05.mma on one accumulator.
template <int NWG_, int DS_, int D2_, int SUB_ = 2, int NI_ = 1>
struct Rings {
  static constexpr int NWG = NWG_, DS = DS_, D2 = D2_, SUB = SUB_, NI = NI_;
};
// Two configurations per small tile, chosen by the launch planner (quick_b200.cu, pick_variant):
//   VAR 0  three dequant warpgroups, one issuer  — many stages per CTA / co-resident CTAs that both stream
//   VAR 1  two dequant warpgroups, two issuers   — short tiles (<= 4 stages) and lone 64-token CTAs
//   128-token tiles: VAR 0 = (2 warpgroups, 4 operand slots, 2 issuers) for <= 8 stages, VAR 1 = (3, 3, 2) for longer ones
// (measured with tools/tune.py on K = N = 4096 and the 7B layer shapes, profiles/r2_tune_*.json).  VAR 2..3 exist in
// QB200_VARIANTS builds only (A/B slots for tools/tune.py, qb200_debug_set_variant).
template <int TOK, int VAR> struct Variant;
template <> struct Variant<16, 0> : Rings<3, 6, 3> {};
template <> struct Variant<16, 1> : Rings<2, 6, 3, 2, 2> {};
template <> struct Variant<32, 0> : Rings<3, 6, 3> {};
template <> struct Variant<32, 1> : Rings<2, 6, 3, 2, 2> {};
template <> struct Variant<64, 0> : Rings<3, 6, 3> {};
template <> struct Variant<64, 1> : Rings<2, 6, 3, 2, 2> {};
template <> struct Variant<128, 0> : Rings<2, 6, 4, 2, 2> {};
template <> struct Variant<128, 1> : Rings<3, 6, 3, 2, 2> {};
template <> struct Variant<256, 0> : Rings<2, 4, 3> {};
template <> struct Variant<256, 1> : Rings<2, 4, 2> {};
#ifdef QB200_VARIANTS   // experiment slots (tools/tune.py)
template <> struct Variant<16, 2> : Rings<3, 9, 3> {};
template <> struct Variant<16, 3> : Rings<3, 6, 3, 2, 2> {};
template <> struct Variant<32, 2> : Rings<3, 9, 3> {};
template <> struct Variant<32, 3> : Rings<3, 6, 3, 2, 2> {};
template <> struct Variant<64, 2> : Rings<3, 6, 3, 2, 2> {};
template <> struct Variant<64, 3> : Rings<4, 8, 4, 2, 2> {};
template <> struct Variant<128, 2> : Rings<4, 8, 4, 2, 2> {};
template <> struct Variant<128, 3> : Rings<3, 6, 3> {};
template <> struct Variant<256, 2> : Rings<2, 4, 3> {};
template <> struct Variant<256, 3> : Rings<2, 4, 3> {};
#endif

template <int TOK, int VAR = 0>
struct TileCfg {
  using R = Variant<TOK, VAR>;
  static constexpr int kNumWG = R::NWG, kDS = R::DS, kD2 = R::D2, kSub = R::SUB, kNI = R::NI;
  static_assert(kNI == 1 || (kNI == 2 && TOK <= 128 && kD2 >= 2), "one or two MMA issuers (two accumulators must fit tensor memory)");
  static constexpr int kWStage = kSub * kWStageBytes;         // packed nibbles of one stage (SUB x 4 KB)
  static constexpr int kASlotCols = 32 * kSub;                // TMEM columns of one A slot (fp16 pairs)
  // A W slot is always served by the same warpgroup (DS % NWG == 0), so its warps observe every phase of
  // "their" TMA barriers in order (mbarrier parity waits are only valid one phase ahead, and TMA completions
  // arrive out of order).  The "free" barriers complete in MMA order, so D2 >= NWG suffices for them.
  static_assert(kDS % kNumWG == 0 && kD2 >= kNumWG, "ring depths vs warpgroup count");
  static_assert(kDS >= kD2, "the W ring is at least as deep as the operand ring");
  static constexpr int kProducerWarp = 4 * kNumWG;
  static constexpr int kMmaWarp = 4 * kNumWG + 1;
  static constexpr int kNumThreads = (4 * kNumWG + 1 + kNI) * 32;     // dequant warps, producer, issuer(s)
  static constexpr int kXPanelBytes = TOK * 128;                      // one k64 panel
  static constexpr int kXStageBytes = kSub * kXPanelBytes;
  static constexpr int kACol0 = kNI * TOK < 32 ? 32 : kNI * TOK;       // accumulator(s) first, then the A ring
  static constexpr int kColsNeeded = kACol0 + kASlotCols * kD2;
  static constexpr int kTmemCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128
                                   : kColsNeeded <= 256 ? 256 : 512;
  static_assert(kColsNeeded <= 512, "TMEM budget");
  // Barrier ring length.  An mbarrier parity wait is only valid if the waiter observes EVERY phase of the barrier in
  // order (waiting for use u while use u-1 has not completed returns true: aliasing).  With two issuers taking
  // alternate stages and an odd number of operand slots, a slot alternates between the issuers, so the ready / free
  // barriers form a ring of 2 x D2 (stage it uses barrier it % kNB and slot it % D2): barrier b then always belongs
  // to issuer b % 2, and to one dequant warpgroup.
  static constexpr int kNB = (kNI == 2 && kD2 % 2 == 1) ? 2 * kD2 : kD2;
  static_assert(kNI == 1 || kNB % kNumWG == 0, "two issuers: every free barrier must be observed by one dequant warpgroup");
  static constexpr int kNumBars = kDS + 2 * kNB + 2;
  static constexpr int kBarBytes = kNumBars * 8 + 16;
  static constexpr int kRstdBytes = TOK * 4 + 128 + 16;        // fused RMSNorm (GemmArgs): per-token 1/rms + 128 B of sum-of-squares scratch
  static constexpr int kPipeBytes = kD2 * kXStageBytes + kDS * kWStage;
  // split-K exchange: TOK <= 64 sends register fragments with st.async straight into the owner's receive
  // buffer; larger tiles stage packed fp16 slices and move them with one TMA bulk DSMEM copy per owner.
  static constexpr bool kAsyncExchange = TOK <= 64;
  // dedicated receive buffer (not aliasing the pipeline stages): senders need no "owner finished its main
  // loop" barrier
  static constexpr bool kDedicatedRecv = TOK <= 64 || (TOK == 128 && kD2 >= 3 && kPipeBytes <= 184 * 1024);
  __host__ __device__ static constexpr int slice(int split) { return TOK / split; }
  // bytes per exchanged element: fp32 when a thread owns <= 4 columns of a slice, packed fp16 otherwise
  __host__ __device__ static constexpr int elem_bytes(int split) { return (kAsyncExchange && slice(split) / 2 <= 4) ? 4 : 2; }
  __host__ __device__ static constexpr int recv_bytes(int split) {
    return split > 1 ? (split - 1) * kChan * slice(split) * elem_bytes(split) : 0;
  }
  __host__ __device__ static constexpr int smem_bytes(int split) {
    return kPipeBytes + kBarBytes + kRstdBytes + (kDedicatedRecv ? recv_bytes(split) + 16 : 0) + 1024;   // + 1024-B alignment slack
  }
  // two CTAs per SM: 233472 B per SM, 1 KB reserved per CTA
  static_assert(smem_bytes(4) <= 232448, "shared-memory budget (227 KB per CTA)");
  static constexpr bool kCoResident = kTmemCols <= 256 && 2 * (smem_bytes(4) + 1024) <= 233472;
  static constexpr int kMinBlocks = kCoResident ? 2 : 1;
};

constexpr unsigned kFlagIndependent = 1u;   // == QB200_GEMM_INDEPENDENT
constexpr unsigned kFlagSiluMul = 2u;       // == QB200_GEMM_SILU_MUL

// SwiGLU fused into the epilogue (reference modules/fused/mlp.py:52-76: gate and up projections, silu(gate) * up).
// The weight's output channels are interleaved gate_0, up_0, gate_1, up_1, ... (prepared once at load), so in the
// swap-A/B accumulator — one lane per channel — a gate channel and its up channel are NEIGHBOURING LANES of one warp:
// one shuffle, no second pass over HBM.  Rounding is that of the unfused pair of kernels (fp16 GEMM outputs, SiLU in
// fp32 rounded to fp16, fp16 product), so results are bit-identical to GEMM + qb200_silu_mul.  Valid in even lanes.
__device__ __forceinline__ __half silu_mul_pair(float acc, int lane) {
  const __half h = __float2half_rn(acc);
  const __half other = __ushort_as_half(static_cast<unsigned short>(__shfl_xor_sync(0xffffffffu, static_cast<unsigned>(__half_as_ushort(h)), 1)));
  const float g = __half2float(h);
  (void)lane;
  return __hmul(__float2half_rn(g / (1.f + expf(-g))), other);
}

// ------------------------------------------------------------------------------------------------
// Tensor-parallel hand-over without a barrier kernel (include/quick_b200.h: qb200_peer_wait / qb200_peer_signal).
// A gathered buffer (every rank's kernel stores its column slab into ALL ranks' copies over NVLink) has a local epoch
// counter and, on every rank, one flag word per producing rank.
//   producer kernel : plain peer / multicast stores, no fence; ONE thread bumps the local epoch after
//                     griddepcontrol.wait (every reader of the previous fill has completed by then).
//   consumer kernel : after ITS griddepcontrol.wait the local producer grid has completed, i.e. this rank's slab has
//                     been delivered everywhere; one thread announces that (epoch -> slot [rank] of every rank's flag
//                     array, st.release.sys) and every CTA polls its own flag array until all ranks have announced.
// The producer needs no system-scope fence and no CTA counter (an earlier version fenced in every CTA and published
// from the last one: 7-12 us per hand-over on 2 GPUs against 2-3 us for this form), the epoch lives on the device
// (CUDA-graph replays hand out fresh epochs) and the wait rides in the consumer's prologue — a GEMM keeps prefetching
// its weights meanwhile — instead of a barrier kernel per gather.
// ------------------------------------------------------------------------------------------------
struct PeerWait {
  const unsigned* epoch;   // local epoch counter of the buffer about to be read (nullptr: nothing to wait for)
  const unsigned* flags;   // this rank's flag array: flags[p] = last fill rank p has announced as delivered
  unsigned* peer_flags[8]; // every rank's flag array (peer-mapped); slot [rank] is ours to write
  int rank, n;
  int mode;                // QB200_TP_WAITMODE: 2 = device-scope acquire fence after the last poll (default), 0 = system scope
  unsigned long long timeout_ns;   // 0 = wait for ever
};
struct PeerSignal {
  unsigned* epoch;         // local epoch counter of the buffer this kernel fills (nullptr: not a gathered buffer)
};
// Producer side: called by ONE thread of the launch, after griddepcontrol.wait.
__device__ __forceinline__ void peer_begin_fill(const PeerSignal& s) {
  if (s.epoch != nullptr) *reinterpret_cast<volatile unsigned*>(s.epoch) = *reinterpret_cast<volatile unsigned*>(s.epoch) + 1u;
}
// Consumer side: called by one converged warp per CTA after griddepcontrol.wait; `announce` is true in exactly one
// warp of the launch.
__device__ __forceinline__ void peer_wait_warp(const PeerWait& w, bool announce) {
  if (w.epoch == nullptr) return;
  const int lane = threadIdx.x & 31;
  const unsigned e = *reinterpret_cast<const volatile unsigned*>(w.epoch);
  if (lane < w.n) {
    if (announce) {
      if (w.mode & 4) __threadfence_system();     // experiment: explicit system-scope fence before the announcement
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(w.peer_flags[lane] + w.rank), "r"(e) : "memory");
    }
    unsigned long long t0 = 0;
    unsigned v, polls = 0;
    do {   // relaxed polls (an acquire per poll would put a system-scope fence into the loop); one fence after the last
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w.flags + lane) : "memory");
      if (w.timeout_ns != 0 && (++polls & 1023u) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > w.timeout_ns) __trap();
      }
    } while (static_cast<int>(v - e) < 0);
    if ((w.mode & 3) == 0) asm volatile("fence.acq_rel.sys;" ::: "memory");
    else if ((w.mode & 3) == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncwarp();
  asm volatile("fence.proxy.async;" ::: "memory");   // the gathered rows may be read by TMA (async proxy) next
}

struct GemmArgs {
  const uint32_t* wq;
  const uint32_t* sz;
  const __half* bias;
  const __half* residual;   // optional [M][N] fp16: C = residual + fp16(A·W + bias), the add in fp16 like torch's x + linear(x)
  __half* C;
  // Fused all-gather (column-parallel linears): the output slab is stored into the buffers of ALL n_peers ranks
  // (peer-mapped pointers, this rank's included) at column offset col0 of rows ldc wide; plain GEMM: n_peers = 0,
  // ldc = N, col0 = 0 and C is the only destination.
  __half* peerC[8];
  __half* mcC;        // optional NVSwitch multicast mapping of the same buffers: ONE multimem.st reaches every rank
  int n_peers, ldc, col0;
  int M, K, N, G;
  int kb_per_split;   // k64 blocks per cluster rank
  unsigned flags;     // QB200_GEMM_* (include/quick_b200.h)
  PeerWait wait;      // tensor parallel: A lives in a gathered buffer — meet its producers before the first load
  PeerSignal signal;  // tensor parallel: C is a gathered buffer — publish once every CTA has stored its slab
  // RMSNorm folded around the GEMM (reference modules/fused/block.py:61-74, norm.py:16-19: norm -> linear).
  //   producer side (C is the residual stream that an RMSNorm with weight gamma reads next): the epilogue also writes
  //     norm_out = fp16(C * gamma) and, per 128-channel tile, the sum of squares of its C values per row
  //     (ssq_out [N/128][M] fp32, fixed summation order: bit-reproducible);
  //   consumer side (A is such a norm_out): every output row is scaled by rsqrt(sum of the row's ssq parts / K + eps)
  //     before the bias — x·rstd·gamma·W == rstd · ((x·gamma)·W), so the RMSNorm kernel between the two GEMMs disappears.
  const __half* norm_gamma;
  __half* norm_out;
  float* ssq_out;
  const float* ssq_in;
  int ssq_parts;
  float rms_eps;
  unsigned launch_id; // host-side launch counter (QB200_TRACE builds: row of the cross-launch timeline)
  long long* trace;   // debug (QB200_TRACE builds only): clock64 stamps of CTA (0,0,0)
};
#ifndef QB_TQ
#define QB_TQ 2   // TMEM lane quadrant whose dequant warps are traced in detail
#endif
#ifdef QB200_TRACE
#define QB_TRACE(slot, it, k)                                                                  \
  do {                                                                                         \
    if (args.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)         \
      args.trace[((slot) * 256 + (it)) * 4 + (k)] = clock64();                                 \
  } while (0)
// Cross-launch timeline (tools/timeline.py): CTA (0,0,0) of every launch takes a sequence number and stamps
// %globaltimer at its start / after griddepcontrol.wait / accumulator complete / exit.
#define QB_TL_BASE (6 * 256 * 4)
__device__ __forceinline__ long long qb_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define QB_TL(k)                                                                                        \
  do {                                                                                                  \
    if (args.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)                  \
      args.trace[QB_TL_BASE + 8 + (qb_tl_seq & 255) * 4 + (k)] = qb_globaltimer();                       \
  } while (0)
// every CTA of the launch: min/max over the grid of (0) CTA start, (1) griddepcontrol.wait returned, (2) CTA exit
#define QB_TLALL_BASE (QB_TL_BASE + 8 + 256 * 4)
#define QB_TLALL(k)                                                                                     \
  do {                                                                                                  \
    if (args.trace != nullptr) {                                                                        \
      const unsigned long long t_ = static_cast<unsigned long long>(qb_globaltimer());                  \
      unsigned long long* row_ = reinterpret_cast<unsigned long long*>(args.trace) + QB_TLALL_BASE + (args.launch_id & 255) * 8; \
      atomicMin(row_ + 2 * (k), t_);                                                                    \
      atomicMax(row_ + 2 * (k) + 1, t_);                                                                \
    }                                                                                                   \
  } while (0)
#else
#define QB_TRACE(slot, it, k) do { } while (0)
#define QB_TL(k) do { } while (0)
#define QB_TLALL(k) do { } while (0)
#endif

// Programmatic dependent launch (PDL): the next kernel in the stream starts while this one is running; it
// waits for this kernel only where its results (the activations) are first consumed.  No-ops when launched
// without the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
// shared::cta -> (remote) shared::cluster bulk copy by the TMA engine, completing on the destination CTA's mbarrier
__device__ __forceinline__ void bulk_s2dsmem(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remote_dst),
               "r"(local_src), "r"(bytes), "r"(remote_bar)
               : "memory");
}
// register -> (remote) shared::cluster store that completes bytes on the destination CTA's mbarrier (SASS: STAS)
template <int NW>
__device__ __forceinline__ void st_async(uint32_t remote_dst, const uint32_t* v, uint32_t remote_bar) {
  if constexpr (NW == 1) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_dst), "r"(v[0]),
                 "r"(remote_bar)
                 : "memory");
  } else if constexpr (NW == 2) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(remote_dst),
                 "r"(v[0]), "r"(v[1]), "r"(remote_bar)
                 : "memory");
  } else {
    static_assert(NW == 4, "st_async: 1, 2 or 4 words");
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_dst),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(remote_bar)
                 : "memory");
  }
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  bool timed = false;
  while (true) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (!timed) { t0 = clock64(); timed = true; }
    else if (clock64() - t0 > QB200_WAIT_TIMEOUT_CYCLES) { __trap(); }
  }
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}
// NVSwitch multicast stores (multimem.st on a multicast mapping: the switch replicates the write to every rank's copy)
__device__ __forceinline__ void multimem_st_v4(void* p, uint4 v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
               "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}
__device__ __forceinline__ void multimem_st_b32(void* p, uint32_t v) {
  asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// a + b on eight packed fp16 values (residual + GEMM result, one rounding each like torch's fp16 add)
__device__ __forceinline__ uint4 hadd2x4(uint4 a, uint4 b) {
  uint4 r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r.x) : "r"(a.x), "r"(b.x));
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r.y) : "r"(a.y), "r"(b.y));
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r.z) : "r"(a.z), "r"(b.z));
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r.w) : "r"(a.w), "r"(b.w));
  return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// Epilogue staging: this CTA's [rows tokens][128 channels] fp16 tile in shared memory (or [rows][64] products when
// SiLU·up is fused), then 16-byte coalesced stores to C / every peer / the multicast mapping.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_out(uint32_t smem_out, int row, int ch, int lane, float acc, bool silu_mul) {
  if (silu_mul) {
    const __half h = silu_mul_pair(acc, lane);
    if ((lane & 1) == 0) sts_u16(smem_out + static_cast<uint32_t>((row * (kChan / 2) + (ch >> 1)) * 2), __half_as_ushort(h));
  } else {
    sts_u16(smem_out + static_cast<uint32_t>((row * kChan + ch) * 2), __half_as_ushort(__float2half_rn(acc)));
  }
}
__device__ __forceinline__ uint4 hmul2x4(uint4 a, uint4 b) {
  uint4 r;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r.x) : "r"(a.x), "r"(b.x));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r.y) : "r"(a.y), "r"(b.y));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r.z) : "r"(a.z), "r"(b.z));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r.w) : "r"(a.w), "r"(b.w));
  return r;
}
__device__ __forceinline__ float sumsq8(uint4 v) {
  const float2 a = unpack_half2(v.x), b = unpack_half2(v.y), c = unpack_half2(v.z), d = unpack_half2(v.w);
  return ((a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y)) + ((c.x * c.x + c.y * c.y) + (d.x * d.x + d.y * d.y));
}
// called by the 256 epilogue threads after the tile has been staged (and a barrier); rows is even, so the two rows a warp
// covers per iteration (16 lanes each) exist or not together and the shuffles below see full warps
__device__ __forceinline__ void store_tile(const GemmArgs& args, uint32_t smem_out, int rows, int m_base, int n0, int nt, bool silu_mul) {
  const int tid = threadIdx.x;             // 0..255
  const int cpr = silu_mul ? 8 : 16;       // 16-byte chunks per staged row (128 B of products, or 256 B)
  const int chunk = tid % cpr;
  const int nbase = silu_mul ? (n0 >> 1) : n0;
  const bool normed = args.norm_gamma != nullptr && !silu_mul;
  uint4 gam = make_uint4(0, 0, 0, 0);
  if (normed) gam = *reinterpret_cast<const uint4*>(args.norm_gamma + nbase + chunk * 8);
#pragma unroll 4
  for (int row = tid / cpr; row < rows; row += (kEpilogueWarps * 32) / cpr) {
    const int m = m_base + row;
    float sq = 0.f;
    if (m < args.M) {
      uint4 v = lds128(smem_out + static_cast<uint32_t>((row * cpr + chunk) * 16));
      const size_t off = static_cast<size_t>(m) * args.ldc + args.col0 + nbase + chunk * 8;
      if (args.residual != nullptr) v = hadd2x4(*reinterpret_cast<const uint4*>(args.residual + off), v);
      if (normed) {   // fused RMSNorm, producer side (GemmArgs): gamma-scaled copy + sum of squares of the row's 128 channels
        *reinterpret_cast<uint4*>(args.norm_out + off) = hmul2x4(v, gam);
        sq = sumsq8(v);
      }
      if (args.n_peers == 0) *reinterpret_cast<uint4*>(args.C + off) = v;
      else if (args.mcC != nullptr) multimem_st_v4(args.mcC + off, v);
      else for (int p = 0; p < args.n_peers; ++p) *reinterpret_cast<uint4*>(args.peerC[p] + off) = v;
    }
    if (normed) {     // 16 lanes = one row: fixed-order butterfly
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (chunk == 0 && m < args.M) args.ssq_out[static_cast<size_t>(nt) * args.M + m] = sq;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
template <int TOK, int SPLIT, int VAR = 0, bool SILU = false>
__global__ void __launch_bounds__(TileCfg<TOK, VAR>::kNumThreads, TileCfg<TOK, VAR>::kMinBlocks)
w4a16_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmArgs args) {
  using Cfg = TileCfg<TOK, VAR>;
  constexpr int NWG = Cfg::kNumWG, DS = Cfg::kDS, D2 = Cfg::kD2;
  constexpr int kSubPerStage = Cfg::kSub, kWStage = Cfg::kWStage, kASlotCols = Cfg::kASlotCols;
  constexpr int kProducerWarp = Cfg::kProducerWarp;
  constexpr int kMmaWarp = Cfg::kMmaWarp;
  constexpr int kNI = Cfg::kNI;
  constexpr int NB = Cfg::kNB;                // ready / free barrier ring length (stage it -> barrier it % NB)
  constexpr int SLICE = TOK / SPLIT;          // token columns owned by one cluster rank
  constexpr int CH = SLICE / 2;               // columns per (owner, epilogue warpgroup)
  static_assert(CH >= 1, "TOK / SPLIT must be >= 2");
  constexpr bool kAsync = Cfg::kAsyncExchange;
  constexpr bool kF32X = kAsync && CH <= 4;   // exchange fp32 fragments (else packed fp16)
  constexpr bool kDirectStore = kAsync && CH <= 4;   // C rows written straight from registers
  constexpr int PIECE = CH < 16 ? CH : 16;    // columns per tcgen05.ld
  constexpr int UNIT = CH < 8 ? CH : 8;       // columns per exchanged vector
  constexpr int ES = Cfg::elem_bytes(SPLIT);
  static_assert(ES == (kF32X ? 4 : 2), "exchange element size");
  // Epilogue staging (shared memory).
  //   recv : partial slices from the other SPLIT-1 ranks.
  //            st.async path : [src][channel][SLICE] elements (fp32 or fp16), dedicated region
  //            bulk path     : packed fp16 [src][unit][channel][UNIT] (32 lanes of a warp write one contiguous
  //                            run); dedicated for TOK = 128, else aliases the (dead) pipeline stages behind a
  //                            cluster barrier
  //   stage: (bulk path) the partial slices this CTA sends, same layout, LOCAL shared memory (dead pipeline stages)
  //   out  : this CTA's [SLICE tokens][128 channels] fp16 tile (dead pipeline stages), stored with 16-byte writes
  constexpr int kRecvBytes = Cfg::recv_bytes(SPLIT);
  constexpr int kOutBytes = SLICE * kChan * 2;
  static_assert(kRecvBytes % 16 == 0, "recv alignment");
  static_assert((kAsync ? 0 : (Cfg::kDedicatedRecv ? 0 : kRecvBytes) + kRecvBytes) + kOutBytes <= Cfg::kPipeBytes,
                "epilogue staging must fit the (dead) pipeline stages");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;                                   // D2 x 2 x [TOK rows][128 B] swizzled
  const uint32_t smem_w = smem_base + D2 * Cfg::kXStageBytes;          // DS x 2 x [2][128][16 B]
  const uint32_t bar_wfull = smem_w + DS * kWStage;                    // W stage landed (TMA bytes)
  const uint32_t bar_ready = bar_wfull + 8 * DS;                       // operand slot ready: A in TMEM (4 warps) + X landed
  const uint32_t bar_free = bar_ready + 8 * NB;                        // operand slot read by its MMAs (one tcgen05.commit)
  const uint32_t bar_accum = bar_free + 8 * NB;                        // all MMAs of the tile done
  const uint32_t bar_recv = bar_accum + 8;                             // split-K partials from the other ranks landed
  const uint32_t tmem_ptr_smem = bar_recv + 8;
  const uint32_t smem_rstd = (tmem_ptr_smem + 16 + 15) & ~15u;         // [TOK] fp32
  const uint32_t smem_ssq = smem_rstd + TOK * 4;                       // [8 warps][<= 4 rows] fp32 (producer side, direct stores)
  const uint32_t smem_recv = Cfg::kDedicatedRecv ? ((smem_ssq + 128 + 15) & ~15u) : smem_x;
  const uint32_t smem_stage = Cfg::kDedicatedRecv ? smem_x : smem_x + kRecvBytes;   // bulk path only
  const uint32_t smem_out = kAsync ? smem_x : smem_stage + kRecvBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Independent launch (QB200_GEMM_INDEPENDENT): the caller guarantees that A, C and the weights are not
  // touched by any kernel this launch may overlap under programmatic dependent launch, so nothing waits for
  // the previous grid until the very end (one griddepcontrol.wait before exit keeps completion transitive:
  // "this kernel finished" still implies "everything launched before it finished").
  const bool independent = (args.flags & kFlagIndependent) != 0;
  // SwiGLU fused into the epilogue (gate / up channels interleaved): a separate instantiation — as a run-time branch in
  // the epilogue's inner loops it cost every GEMM 3 % (16-token tiles) to 10 % (256-token tiles)
  constexpr bool silu_mul = SILU;
  const int nt = blockIdx.x;
  const int mt = blockIdx.y;
  const int rank = SPLIT > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int KB = args.K / kBK;
  const int kb0 = rank * args.kb_per_split;
  const int nkb = min(args.kb_per_split, KB - kb0);            // k64 blocks of this CTA
  const int nst = (nkb + kSubPerStage - 1) / kSubPerStage;     // pipeline stages (last may hold one block)
  const uint32_t* wsrc = args.wq + (static_cast<size_t>(nt) * KB + kb0) * (kWStageBytes / 4);
  auto issue_w = [&](int it) {   // one elected producer lane
    const int s = it % DS;
    const int nsub = min(kSubPerStage, nkb - it * kSubPerStage);
    const uint32_t bar = bar_wfull + 8 * s;
    mbar_arrive_expect_tx(bar, nsub * kWStageBytes);
    bulk_g2s(smem_w + s * kWStage, wsrc + static_cast<size_t>(it) * (kWStage / 4), nsub * kWStageBytes, bar);
  };

  // ---------------- setup: three warps work in parallel, everybody meets once ----------------
  pdl_launch_dependents();
#ifdef QB200_TRACE
  __shared__ int qb_tl_seq_s;
  if (threadIdx.x == 0 && args.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    qb_tl_seq_s = static_cast<int>(atomicAdd(reinterpret_cast<unsigned long long*>(args.trace + QB_TL_BASE), 1ull));
    const int qb_tl_seq = qb_tl_seq_s;
    QB_TL(0);
  }
#endif
  if (threadIdx.x == 0) QB_TRACE(3, 0, 0);
  if (threadIdx.x == 32) QB_TLALL(0);
  if (warp == kProducerWarp) {
    // One elected producer lane owns the whole setup: tensor-map prefetch first (the descriptor fetch is a cold
    // miss whose latency should overlap everything else), all barriers, then the whole W ring — requested before
    // the TMEM allocation and before griddepcontrol.wait, so under PDL the HBM stream of this GEMM overlaps the
    // previous kernel.
    if (elect_one()) {
      prefetch_tmap(&tmap_x);
      for (int i = 0; i < DS; ++i) mbar_init(bar_wfull + 8 * i, 1);
      for (int i = 0; i < NB; ++i) {
        mbar_init(bar_ready + 8 * i, 5);   // 4 dequant warps + the producer's expect_tx arrival
        mbar_init(bar_free + 8 * i, 1);
      }
      mbar_init(bar_accum, Cfg::kNI);   // one tcgen05.commit (or plain arrival) per issuer
      mbar_init(bar_recv, 1);   // one expect_tx arrive by the owner; the senders' copies complete the bytes
      fence_barrier_init();
      fence_proxy_async();
      const int npre = min(nst, DS);
      for (int it = 0; it < npre; ++it) issue_w(it);
      QB_TRACE(0, 0, 2);
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);   // blocks while a co-resident CTA of the previous kernel holds the columns
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds_u32(tmem_ptr_smem);
#ifdef QB200_TRACE
  const int qb_tl_seq = qb_tl_seq_s;
#endif
  if constexpr (SPLIT > 1) cluster_arrive_relaxed();   // matched by a wait just before the first remote access
  if (threadIdx.x == 0) QB_TRACE(3, 0, 1);

  if (warp == kProducerWarp) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    // One in-order loop over the X stages: X(j) reuses the operand slot of stage j-D2 and waits for that stage's
    // MMAs ("free").  The W ring is not refilled here: a W slot is dead as soon as its dequant warpgroup has loaded
    // it into registers, long before the stage's MMAs complete, so the warpgroup itself requests the slot's next
    // stage (round-2 traces: riding on "free", weights were requested only DS - D2 stages ahead and 128-token tiles
    // waited ~400 cycles per stage for them; polling a separate "loaded" barrier from this warp cost more than it
    // gained because it delayed the X loads).
    if (!independent) pdl_wait_prior_grid();   // the activations come from the previous kernel
    // tensor parallel: the activations may live in a gathered buffer (meet the ranks that fill it), and C may be one
    const bool first_cta = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    peer_wait_warp(args.wait, first_cta);
    if (first_cta && lane == 0) peer_begin_fill(args.signal);
    if (lane == 0) QB_TL(1);
    if (lane == 0) QB_TLALL(1);
    for (int j = 0; j < nst; ++j) {
      const int x = j % D2;                                       // operand slot
      if (j >= D2) mbar_wait(bar_free + 8 * ((j - D2) % NB), static_cast<uint32_t>((j - D2) / NB) & 1u, 1, j);
      if (elect_one()) {
        QB_TRACE(0, j, 0);
        const int nsub = min(kSubPerStage, nkb - j * kSubPerStage);
        const uint32_t bar = bar_ready + 8 * (j % NB);
        mbar_arrive_expect_tx(bar, nsub * Cfg::kXPanelBytes);
#pragma unroll
        for (int p = 0; p < kSubPerStage; ++p)
          if (p < nsub)
            tma_load_2d(smem_x + x * Cfg::kXStageBytes + p * Cfg::kXPanelBytes, &tmap_x, bar, (kb0 + j * kSubPerStage + p) * kBK, mt * TOK);
        QB_TRACE(0, j, 1);
      }
      __syncwarp();
    }
  } else if (warp >= kMmaWarp) {
    // ===================== MMA issuer(s): one wait, 8 MMAs, (at most) one commit per stage =====================
    // Issuer iw takes stages iw, iw + NI, ... and accumulates into its own accumulator D_iw.
    constexpr uint32_t idesc = make_idesc_f16(TOK);
    const int iw = warp - kMmaWarp;
    const uint32_t d_acc = tmem_base + static_cast<uint32_t>(iw * TOK);
    for (int it = iw; it < nst; it += kNI) {
      const int t = it % D2;                                      // operand slot
      const int b = it % NB;                                      // its barrier pair
      if (lane == 0 && iw == 0) QB_TRACE(5, it, 0);
      mbar_wait(bar_ready + 8 * b, static_cast<uint32_t>(it / NB) & 1u, 2, it);
      tc_fence_after();
      if (elect_one()) {
        if (iw == 0) QB_TRACE(1, it, 0);
        const int nsub = min(kSubPerStage, nkb - it * kSubPerStage);
        const uint64_t bdesc = make_smem_desc_sw128(smem_x + t * Cfg::kXStageBytes);
        const uint32_t a_tmem = tmem_base + Cfg::kACol0 + t * kASlotCols;
#pragma unroll
        for (int p = 0; p < kSubPerStage; ++p) {
          if (p < nsub) {
            const uint64_t bdesc_p = bdesc + static_cast<uint64_t>(p * (Cfg::kXPanelBytes >> 4));
#pragma unroll
            for (int j = 0; j < kBK / 16; ++j) {
              // +32 B (= 2 in 16-B units) of start address per k16 step inside the 128-B swizzle row
              umma_f16_ts(d_acc, a_tmem + p * 32 + j * 8, bdesc_p + 2 * j, idesc, (it != iw || (p | j) != 0) ? 1u : 0u);
            }
          }
        }
        if (iw == 0) QB_TRACE(1, it, 1);
        // X stage, W stage and TMEM A slot are free once these MMAs have completed.  Committed only when a later
        // stage will reuse the slot: its dequant warps and the producer wait for exactly this phase, so every
        // commit arrival is observed before the accumulator barrier completes (no arrival can outlive the CTA
        // and hit the barrier words of the next CTA scheduled on this SM), and a tile whose stages all fit the
        // ring pays for a single tcgen05.commit per issuer.
        if (it + D2 < nst) umma_commit(bar_free + 8 * b);
        if (it + kNI >= nst) umma_commit(bar_accum);   // this issuer's last stage
        if (iw == 0) QB_TRACE(1, it, 2);
      }
      __syncwarp();
    }
    if (iw >= nst && lane == 0) mbar_arrive(bar_accum);   // an issuer without a stage still completes the barrier
  } else {
    // ===================== dequant warps: smem nibbles -> registers -> TMEM A operand =====================
    // Stage order follows issue priority: the SM sub-partition arbiter serves the highest warp id first, so
    // the LAST warpgroup takes stages 0, NWG, ... — the stage the MMA issuer needs next is also the one
    // whose dequant warps win the FMA pipe.
    const int wg = NWG - 1 - (warp >> 2);           // logical warpgroup w takes stages w, w + NWG, ...
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int ch = quad * 32 + lane;                // output channel within the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const int NG = args.K / args.G;
    const int g32 = args.G >> 5;                    // k32 blocks per group
    const uint32_t* szp = args.sz + static_cast<size_t>(nt) * NG * kChan + ch;
    constexpr int kQ = 2 * kSubPerStage;            // k32 blocks per stage
    const int last_nsub = nkb - (nst - 1) * kSubPerStage;   // only the last stage can be partial
    // One scale / zero word per stage when a whole stage lies inside one quantisation group (G a multiple of the stage
    // length and the CTA's k range starting on a stage boundary: the common case, G = 128 with 128-k stages); otherwise
    // one word per k32 block.  The loop is instantiated for both (the choice is uniform over the CTA): the common case
    // saves three loads and nine constant-preparation instructions per stage and thread.
    const bool one_group = (g32 % kQ == 0) && ((kb0 * 2) % kQ == 0);
    auto run = [&](auto one_tag) {
      constexpr bool kOne = decltype(one_tag)::value;
      constexpr int kNZ = kOne ? 1 : kQ;            // scale / zero words per stage
      // group index of the k32 blocks of the next stage, advanced incrementally (no divisions in the loop)
      int kq = (kb0 + wg * kSubPerStage) * 2;
      int grp = kq / g32, rem = kq % g32;
      uint32_t szw[kNZ];
      auto load_sz = [&]() {
        int g = grp, r = rem;
#pragma unroll
        for (int q = 0; q < kNZ; ++q) {
          szw[q] = __ldg(szp + static_cast<size_t>(min(g, NG - 1)) * kChan);
          if (++r >= g32) { r = 0; ++g; }
        }
      };
      if (wg < nst) load_sz();
      int s = wg % DS;                              // W ring slot (same warpgroup every time: DS % NWG == 0)
      uint32_t sph = 0;
      int b = wg % NB;                              // ready / free barrier of the stage (it % NB), kept incrementally
      uint32_t bph = 0;                             // (it / NB) & 1
      for (int it = wg; it < nst; it += NWG) {
        const int nsub = it == nst - 1 ? last_nsub : kSubPerStage;
        mbar_wait(bar_wfull + 8 * s, sph, 3, it);
        if (lane == 0 && quad == QB_TQ) QB_TRACE(2, it, 0);
        const uint32_t wbase = smem_w + s * kWStage + ch * 16;
        uint4 w[kQ];
#pragma unroll
        for (int p = 0; p < kSubPerStage; ++p) {
          if (p < nsub) {
            w[2 * p] = lds128(wbase + p * 4096);
            w[2 * p + 1] = lds128(wbase + p * 4096 + 2048);
          }
        }
        GroupConsts gc[kNZ];
#pragma unroll
        for (int q = 0; q < kNZ; ++q) gc[q] = make_group_consts(szw[q]);
        // advance NWG stages and prefetch the next scale / zero words
        rem += kQ * NWG;
        while (rem >= g32) { rem -= g32; ++grp; }
        if (it + NWG < nst) load_sz();
        if (lane == 0 && quad == QB_TQ) QB_TRACE(4, it, 0);
        // operand slot t = it % D2 is free once the MMAs of stage it - D2 have completed (barrier (it - D2) % NB)
        int t, fb;
        uint32_t fph;
        if constexpr (NB == D2) { t = b; fb = b; fph = bph ^ 1u; }
        else { t = b >= D2 ? b - D2 : b; fb = b >= D2 ? b - D2 : b + D2; fph = b >= D2 ? bph : bph ^ 1u; }
        if (it >= D2) {
          mbar_wait(bar_free + 8 * fb, fph, 4, it);
          tc_fence_after();
        }
        if (lane == 0 && quad == QB_TQ) QB_TRACE(4, it, 1);
        const uint32_t a_tmem = tmem_base + lane_addr + Cfg::kACol0 + t * kASlotCols;
#pragma unroll
        for (int sub = 0; sub < kSubPerStage; ++sub) {
          if (sub < nsub) {
            const GroupConsts& g0 = gc[kOne ? 0 : 2 * sub];
            const GroupConsts& g1 = gc[kOne ? 0 : 2 * sub + 1];
            uint32_t r[32];
            dequant_word(w[2 * sub].x, g0, r + 0);
            dequant_word(w[2 * sub].y, g0, r + 4);
            dequant_word(w[2 * sub].z, g0, r + 8);
            dequant_word(w[2 * sub].w, g0, r + 12);
            dequant_word(w[2 * sub + 1].x, g1, r + 16);
            dequant_word(w[2 * sub + 1].y, g1, r + 20);
            dequant_word(w[2 * sub + 1].z, g1, r + 24);
            dequant_word(w[2 * sub + 1].w, g1, r + 28);
            tmem_st16(a_tmem + sub * 32, r);
            tmem_st16(a_tmem + sub * 32 + 16, r + 16);
            if (sub == 0 && lane == 0 && quad == QB_TQ) QB_TRACE(4, it, 2);
          }
        }
        if (lane == 0 && quad == QB_TQ) QB_TRACE(2, it, 1);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && quad == QB_TQ) QB_TRACE(4, it, 3);
        if (lane == 0) mbar_arrive(bar_ready + 8 * b);
        if (lane == 0 && quad == QB_TQ) QB_TRACE(2, it, 2);
        if (lane == 0 && quad != QB_TQ) QB_TRACE(5, it, 1 + (quad < QB_TQ ? quad : quad - 1));   // hand-off of the other quadrants
        // Refill this W slot with the stage DS ahead: all four warps of the warpgroup have loaded it (named barrier
        // per warpgroup), one of them — rotating, so no quadrant is always the late one — issues the bulk copy.
        if (it + DS < nst) {
          named_bar_sync(3 + (warp >> 2), 128);
          if (quad == (it & 3) && elect_one()) issue_w(it + DS);
          __syncwarp();
        }
        s += NWG;
        if (s >= DS) { s -= DS; sph ^= 1; }
        b += NWG;
        if (b >= NB) { b -= NB; bph ^= 1u; }
      }
    };
    if (one_group) run(std::true_type{}); else run(std::false_type{});
  }

  // ===================== epilogue =====================
  const int quad = warp & 3;
  const int wg = warp >> 2;
  const int ch = quad * 32 + lane;
  const bool is_epi = warp < kEpilogueWarps;   // the first two dequant warpgroups run the epilogue
  const bool has_d1 = kNI == 2 && nst > 1;     // the second issuer accumulated at least one stage into D1
  const uint32_t d_tmem = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const int n0 = nt * kChan;
  const int valid = min(TOK, args.M - mt * TOK);    // token columns of this tile that exist (the rest are zero rows)
  const bool i_own = rank * SLICE < valid;          // this rank's slice holds at least one real row
  const float bias_v = (args.bias != nullptr && is_epi) ? __half2float(args.bias[n0 + ch]) : 0.f;
  const int m_base = mt * TOK + rank * SLICE;

  // Fused RMSNorm, consumer side: 1/rms of this tile's token rows from the producer's per-tile sums of squares (they are
  // final once the previous grid has completed), computed while the last MMAs drain.  The 256 epilogue threads split
  // into TOK tokens x P part-slices (P adjacent lanes per token): all of a thread's loads are independent (one round of
  // L2 latency), the slices are added by a butterfly over the P lanes — a fixed order, bit-reproducible.
  const bool has_rstd = args.ssq_in != nullptr;
  if (has_rstd && is_epi) {
    pdl_wait_prior_grid();
    constexpr int P = (kEpilogueWarps * 32) / TOK >= 32 ? 16 : (kEpilogueWarps * 32) / TOK;   // 16, 8 (TOK 32), 4, 2, 1
    constexpr int kTokPerPass = (kEpilogueWarps * 32) / P;
    const int slice = threadIdx.x % P;
#pragma unroll 1
    for (int tk = threadIdx.x / P; tk < TOK; tk += kTokPerPass) {      // one pass unless TOK = 16 (P capped at 16)
      float sq = 0.f;
      if (tk < valid) {
        const float* src = args.ssq_in + (mt * TOK + tk);
#pragma unroll 8
        for (int part = slice; part < args.ssq_parts; part += P) sq += __ldcg(src + static_cast<size_t>(part) * args.M);
      }
#pragma unroll
      for (int o = P / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (slice == 0 && tk < valid) sts_u32(smem_rstd + tk * 4, __float_as_uint(rsqrtf(sq / static_cast<float>(args.K) + args.rms_eps)));
    }
    named_bar_sync(1, kEpilogueWarps * 32);
  }
  if (is_epi) {
    mbar_wait(bar_accum, 0, 5, nst);   // every TMA write landed and every MMA read of this CTA's smem is complete
    tc_fence_after();
    if (threadIdx.x == 0) QB_TRACE(3, 0, 2);
    if (threadIdx.x == 0) QB_TL(2);
  }

  if constexpr (kAsync) {
    // ---------- small tiles: register fragments go straight to the owner with st.async ----------
    if constexpr (SPLIT > 1) {
      // rank o owns token columns [o*SLICE, (o+1)*SLICE); owners whose slice has no real row neither receive
      // nor store.  The own partial stays fp32 in TMEM.
      cluster_wait();              // every CTA of the cluster has initialised its barriers
      if (threadIdx.x == 0) QB_TRACE(3, 2, 0);
      if (threadIdx.x == 0 && i_own) mbar_arrive_expect_tx(bar_recv, static_cast<uint32_t>(kRecvBytes));
      if (is_epi) {
#pragma unroll 1
        for (int oo = 1; oo < SPLIT; ++oo) {
          const int o = (rank + oo) % SPLIT;
          if (o * SLICE >= valid) continue;
          const int src_slot = rank < o ? rank : rank - 1;    // my slot among the owner's SPLIT-1 sources
          const uint32_t dst = mapa_shared(
              smem_recv + static_cast<uint32_t>((src_slot * kChan + ch) * (SLICE * ES) + wg * CH * ES), static_cast<uint32_t>(o));
          const uint32_t rbar = mapa_shared(bar_recv, static_cast<uint32_t>(o));
          if constexpr (kF32X) {
            uint32_t v[CH];
            tmem_ld_acc<CH, kNI>(d_tmem + o * SLICE + wg * CH, TOK, has_d1, v);
            st_async<CH>(dst, v, rbar);
          } else {
#pragma unroll 1
            for (int p = 0; p < CH / 8; ++p) {
              uint32_t v[8], h[4];
              tmem_ld_acc<8, kNI>(d_tmem + o * SLICE + wg * CH + p * 8, TOK, has_d1, v);
#pragma unroll
              for (int i = 0; i < 4; ++i) h[i] = pack_half2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
              st_async<4>(dst + p * 16, h, rbar);
            }
          }
        }
        if (threadIdx.x == 0) QB_TRACE(3, 2, 1);
        if (i_own) mbar_wait_cluster(bar_recv, 0);   // all partial slices for my columns have landed
        if (threadIdx.x == 0) QB_TRACE(3, 2, 2);
      }
    }
    if (is_epi && i_own) {
      // No griddepcontrol.wait here: in ordered mode everything the epilogue does is causally after the producer
      // lane's wait (its X loads fed the MMAs whose completion barrier these warps have observed), i.e. after the
      // previous grid has completed and flushed; in independent mode the caller has declared C unrelated.
#pragma unroll 1
      for (int p = 0; p < CH / PIECE; ++p) {
        const int j0 = wg * CH + p * PIECE;
        uint32_t v[PIECE];
        tmem_ld_acc<PIECE, kNI>(d_tmem + rank * SLICE + j0, TOK, has_d1, v);
        float acc[PIECE];
#pragma unroll
        for (int i = 0; i < PIECE; ++i) acc[i] = __uint_as_float(v[i]) + (has_rstd ? 0.f : bias_v);
        if constexpr (SPLIT > 1) {
#pragma unroll
          for (int r = 0; r < SPLIT - 1; ++r) {
            const uint32_t addr = smem_recv + static_cast<uint32_t>((r * kChan + ch) * (SLICE * ES) + j0 * ES);
            if constexpr (kF32X) {
              static_assert(!kF32X || PIECE == CH, "fp32 exchange: one piece per thread");
              if constexpr (CH == 4) {
                const uint4 q = lds128(addr);
                acc[0] += __uint_as_float(q.x); acc[1] += __uint_as_float(q.y);
                acc[2] += __uint_as_float(q.z); acc[3] += __uint_as_float(q.w);
              } else if constexpr (CH == 2) {
                const uint2 q = lds64(addr);
                acc[0] += __uint_as_float(q.x); acc[1] += __uint_as_float(q.y);
              } else {
                acc[0] += __uint_as_float(lds_u32(addr));
              }
            } else {
#pragma unroll
              for (int i = 0; i < PIECE; i += 8) {
                const uint4 q = lds128(addr + i * 2);
                const float2 a = unpack_half2(q.x), b = unpack_half2(q.y), c = unpack_half2(q.z), d = unpack_half2(q.w);
                acc[i] += a.x; acc[i + 1] += a.y; acc[i + 2] += b.x; acc[i + 3] += b.y;
                acc[i + 4] += c.x; acc[i + 5] += c.y; acc[i + 6] += d.x; acc[i + 7] += d.y;
              }
            }
          }
        }
        if (has_rstd) {   // fused RMSNorm: scale the row by its 1/rms, then the bias
#pragma unroll
          for (int i = 0; i < PIECE; ++i) acc[i] = fmaf(acc[i], lds_f32(smem_rstd + (rank * SLICE + j0 + i) * 4), bias_v);
        }
        if constexpr (kDirectStore) {
          // 32 lanes = 32 consecutive channels of one token row: 64-byte runs, no staging round trip
#pragma unroll
          for (int i = 0; i < PIECE; ++i) {
            const int m = m_base + j0 + i;
            if (m < args.M) {
              if (silu_mul) {
                // lanes (2i, 2i+1) = (gate_i, up_i): even lanes hold the product, output column (n0 + ch) / 2
                const __half h = silu_mul_pair(acc[i], lane);
                const size_t off = static_cast<size_t>(m) * args.ldc + args.col0 + ((n0 + ch) >> 1);
                if (args.n_peers == 0) {
                  if ((lane & 1) == 0) args.C[off] = h;
                } else if (args.mcC != nullptr) {
                  const uint32_t mine = __half_as_ushort(h);
                  const uint32_t next = __shfl_down_sync(0xffffffffu, mine, 2);
                  if ((lane & 3) == 0) multimem_st_b32(args.mcC + off, mine | (next << 16));
                } else if ((lane & 1) == 0) {
                  for (int p = 0; p < args.n_peers; ++p) args.peerC[p][off] = h;
                }
                continue;
              }
              const size_t off = static_cast<size_t>(m) * args.ldc + args.col0 + n0 + ch;
              __half h = __float2half_rn(acc[i]);
              if (args.residual != nullptr) h = __hadd(args.residual[off], h);
              if (args.norm_gamma != nullptr) {
                // fused RMSNorm, producer side: gamma-scaled copy + this warp's (32 channels) sum of squares of the row;
                // the four warps of the warpgroup (= the tile's 128 channels) are added in a fixed order below
                args.norm_out[off] = __hmul(h, args.norm_gamma[n0 + ch]);
                const float hf = __half2float(h);
                float sq = hf * hf;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                if (lane == 0) sts_u32(smem_ssq + ((wg * 4 + quad) * PIECE + i) * 4, __float_as_uint(sq));
              }
              if (args.n_peers == 0) {
                args.C[off] = h;
              } else if (args.mcC != nullptr) {
                // multimem.st moves at least 32 bits: even lanes store their channel and the next one
                const uint32_t mine = __half_as_ushort(h);
                const uint32_t next = __shfl_down_sync(0xffffffffu, mine, 1);
                if ((lane & 1) == 0) multimem_st_b32(args.mcC + off, mine | (next << 16));
              } else {
                for (int p = 0; p < args.n_peers; ++p) args.peerC[p][off] = h;
              }
            }
          }
          if (args.norm_gamma != nullptr) {
            static_assert(!kDirectStore || PIECE == CH, "direct store: one piece per thread");
            named_bar_sync(3 + wg, 128);          // the warpgroup's four warps (ids 3 + wg: the dequant loop is over)
            if (quad == 0 && lane < PIECE && m_base + j0 + lane < args.M) {
              float sq = 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) sq += lds_f32(smem_ssq + ((wg * 4 + q) * PIECE + lane) * 4);
              args.ssq_out[static_cast<size_t>(nt) * args.M + m_base + j0 + lane] = sq;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < PIECE; ++i)
            stage_out(smem_out, j0 + i, ch, lane, acc[i], silu_mul);
        }
      }
      if constexpr (!kDirectStore) {
        named_bar_sync(1, kEpilogueWarps * 32);
        store_tile(args, smem_out, SLICE, m_base, n0, nt, silu_mul);
      }
      if (threadIdx.x == 0) QB_TRACE(3, 0, 3);
    }
    // No exit barrier on this path: an owner leaves only after all its inbound bytes have landed (bar_recv),
    // and nobody writes into a CTA that owns nothing; outbound st.async data is in flight from registers.
  } else {
    // ---------- large tiles: packed fp16 slices staged locally, one TMA bulk DSMEM copy per owner ----------
    if constexpr (SPLIT > 1) {
      cluster_wait();              // every CTA of the cluster has initialised its barriers
      if constexpr (!Cfg::kDedicatedRecv) {
        cluster_arrive();
        cluster_wait();            // every CTA is past its main loop: the aliased pipeline stages are dead
      }
      if (threadIdx.x == 0) QB_TRACE(3, 2, 0);
      constexpr uint32_t kSliceBytes = kChan * SLICE * 2;
      if (threadIdx.x == 0) mbar_arrive_expect_tx(bar_recv, (SPLIT - 1) * kSliceBytes);
      if (is_epi) {
#pragma unroll 1
        for (int oo = 1; oo < SPLIT; ++oo) {
          const int o = (rank + oo) % SPLIT;
          const uint32_t dst = smem_stage + static_cast<uint32_t>((oo - 1) * kSliceBytes);
#pragma unroll 1
          for (int p = 0; p < CH / PIECE; ++p) {
            uint32_t v[PIECE];
            tmem_ld_acc<PIECE, kNI>(d_tmem + o * SLICE + wg * CH + p * PIECE, TOK, has_d1, v);
#pragma unroll
            for (int i = 0; i < PIECE; i += UNIT) {
              static_assert(kAsync || UNIT == 8, "bulk path exchanges 8-column units");
              const int unit = (wg * CH + p * PIECE + i) / UNIT;       // unit index inside the slice
              const uint32_t addr = dst + static_cast<uint32_t>((unit * kChan + ch) * (UNIT * 2));
              sts_v4(addr, pack_half2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                     pack_half2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                     pack_half2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                     pack_half2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
            }
          }
        }
        fence_proxy_async();                                  // generic-proxy smem writes -> visible to the TMA engine
        named_bar_sync(1, kEpilogueWarps * 32);
        if (threadIdx.x < SPLIT - 1) {                        // one thread per owner issues that owner's slice
          const int oo = threadIdx.x + 1;
          const int o = (rank + oo) % SPLIT;
          const int src_slot = rank < o ? rank : rank - 1;    // my slot among the owner's SPLIT-1 sources
          bulk_s2dsmem(mapa_shared(smem_recv + static_cast<uint32_t>(src_slot * kSliceBytes), static_cast<uint32_t>(o)),
                       smem_stage + static_cast<uint32_t>((oo - 1) * kSliceBytes), kSliceBytes,
                       mapa_shared(bar_recv, static_cast<uint32_t>(o)));
        }
        if (threadIdx.x == 0) QB_TRACE(3, 2, 1);
        mbar_wait_cluster(bar_recv, 0);   // all partial slices for my columns have landed
        if (threadIdx.x == 0) QB_TRACE(3, 2, 2);
      }
      // Every bulk copy is some CTA's inbound slice: once every CTA of the cluster has seen its bar_recv
      // complete, no TMA engine is still reading anybody's staging buffer.  Arrive now, wait just before exit,
      // so no CTA frees its shared memory under an in-flight copy.
      cluster_arrive_relaxed();
    }
    if (is_epi) {
      // this thread's CH columns of the owned slice (+ the other ranks' partials) -> fp16 -> staging tile [token][channel]
#pragma unroll 1
      for (int p = 0; p < CH / PIECE; ++p) {
        const int j0 = wg * CH + p * PIECE;
        uint32_t v[PIECE];
        tmem_ld_acc<PIECE, kNI>(d_tmem + rank * SLICE + j0, TOK, has_d1, v);
        float acc[PIECE];
#pragma unroll
        for (int i = 0; i < PIECE; ++i) acc[i] = __uint_as_float(v[i]) + (has_rstd ? 0.f : bias_v);
        if constexpr (SPLIT > 1) {
#pragma unroll
          for (int r = 0; r < SPLIT - 1; ++r) {
#pragma unroll
            for (int i = 0; i < PIECE; i += UNIT) {
              const int unit = (j0 + i) / UNIT;
              const uint32_t addr = smem_recv + static_cast<uint32_t>(r * (kChan * SLICE * 2) + (unit * kChan + ch) * (UNIT * 2));
              const uint4 q = lds128(addr);
              const float2 a = unpack_half2(q.x), b = unpack_half2(q.y), c = unpack_half2(q.z), d = unpack_half2(q.w);
              acc[i] += a.x; acc[i + 1] += a.y; acc[i + 2] += b.x; acc[i + 3] += b.y;
              acc[i + 4] += c.x; acc[i + 5] += c.y; acc[i + 6] += d.x; acc[i + 7] += d.y;
            }
          }
        }
        if (has_rstd) {
#pragma unroll
          for (int i = 0; i < PIECE; ++i) acc[i] = fmaf(acc[i], lds_f32(smem_rstd + (rank * SLICE + j0 + i) * 4), bias_v);
        }
#pragma unroll
        for (int i = 0; i < PIECE; ++i)
          stage_out(smem_out, j0 + i, ch, lane, acc[i], silu_mul);
      }
      if (threadIdx.x == 0) QB_TRACE(3, 2, 3);
      named_bar_sync(1, kEpilogueWarps * 32);
      // No griddepcontrol.wait here: in ordered mode everything the epilogue does is causally after the producer
      // lane's wait (its X loads fed the MMAs whose completion barrier these warps have observed), i.e. after the
      // previous grid has completed and flushed; in independent mode the caller has declared C unrelated.
      // coalesced 16-byte stores: 16 threads cover one 256-byte token row of the tile
      store_tile(args, smem_out, SLICE, m_base, n0, nt, silu_mul);
      if (threadIdx.x == 0) QB_TRACE(3, 0, 3);
    }
    if constexpr (SPLIT > 1) cluster_wait();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  if (independent) pdl_wait_prior_grid();
  if (threadIdx.x == kMmaWarp * 32) QB_TRACE(3, 1, 0);
  if (threadIdx.x == kMmaWarp * 32) QB_TL(3);
  if (threadIdx.x == kMmaWarp * 32) QB_TLALL(2);
}

}  // namespace qb200
