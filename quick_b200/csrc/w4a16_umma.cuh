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
//   * warp roles: warps 0..7 = dequant, warp 8 = TMA producer, warp 9 = MMA issuer (+TMEM alloc)
//     (two warpgroups alternating k64 stages) and epilogue
//   * split-K lives inside a thread-block cluster (1,1,SPLIT): partial tiles are exchanged through
//     distributed shared memory, each CTA reduces and stores TOK/SPLIT token columns; no HBM temp,
//     no second kernel (reference: (split_k,M,N) temp + at::sum, gemm_cuda_quick.cu:1468,1515)
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qb200 {

constexpr int kChan = 128;          // channels per tile = UMMA M
constexpr int kBK = 64;             // k per pipeline stage
constexpr int kWStageBytes = kChan * kBK / 2;   // 4096
constexpr int kEpilogueWarps = 8;    // warps 0..7 run the epilogue
// Warp roles.  The single-thread issuers get the HIGHEST warp ids: the SM's warp arbiter favours
// higher warp ids, and a starved TMA/MMA issuer stalls the whole pipeline (measured: with the
// issuers on warps 0/1 every already-complete mbarrier wait cost ~300 cycles).
// NWG dequant warpgroups (4 warps each, warp % 4 = TMEM lane quadrant) take pipeline stages round-robin.
template <int TOK>
constexpr int default_nwg() { return TOK <= 64 ? 3 : 2; }   // measured best of (2,6) (3,3) (3,6) (4,4) (5,5) (6,6)

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

// Pipeline stage = 128 k (two k64 blocks): W nibbles 8 KB, X tile 2 x [TOK][128 B] swizzled panels, TMEM A
// slot 64 columns.  The shared-memory stages and the TMEM A slots form ONE ring of depth D, so a single
// "consumed" barrier per slot (4 dequant-warp arrivals + 1 tcgen05.commit) releases both.  Depths are
// chosen so that TOK <= 64 tiles fit 2 CTAs / SM (TMEM <= 256 columns, smem <= ~110 KB).
constexpr int kSubPerStage = 2;                          // k64 blocks per stage
constexpr int kWStageBytesV3 = kSubPerStage * kWStageBytes;   // 8192
template <int TOK>
constexpr int default_depth() { return TOK <= 64 ? 6 : TOK == 128 ? 4 : 3; }

template <int TOK, int D = default_depth<TOK>(), int NWG = default_nwg<TOK>()>
struct TileCfg {
  static constexpr int kDepth = D;
  static constexpr int kNumWG = NWG;
  static constexpr int kProducerWarp = 4 * NWG;
  static constexpr int kMmaWarp = 4 * NWG + 1;
  static constexpr int kNumThreads = (4 * NWG + 2) * 32;
  // when D is a multiple of NWG a slot is always consumed by the same warpgroup, so its warps see every
  // phase of "their" full barriers in order; otherwise each warp must also observe the skipped stages
  static constexpr bool kInOrderWaits = (D % NWG) != 0;
  static constexpr int kXPanelBytes = TOK * 128;                      // one k64 panel
  static constexpr int kXStageBytes = kSubPerStage * kXPanelBytes;
  static constexpr int kStageBytes = kXStageBytes + kWStageBytesV3;
  static constexpr int kACol0 = TOK < 32 ? 32 : TOK;
  static constexpr int kASlotCols = 32 * kSubPerStage;
  static constexpr int kColsNeeded = kACol0 + kASlotCols * D;
  static constexpr int kTmemCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128
                                   : kColsNeeded <= 256 ? 256 : 512;
  static_assert(kColsNeeded <= 512, "TMEM budget");
  static constexpr int kBarBytes = (3 * D + 2) * 8 + 16;
  static constexpr int kPipeBytes = D * kStageBytes;
  // split-K receive buffer: (SPLIT-1) fp16 partial slices of 128 x (TOK/SPLIT); dedicated (not aliasing the
  // pipeline stages) for TOK <= 128 so that senders need no "owner finished its main loop" barrier
  static constexpr bool kDedicatedRecv = TOK <= 128;
  __host__ __device__ static constexpr int recv_bytes(int split) { return split > 1 ? (split - 1) * kChan * (TOK / split) * 2 : 0; }
  __host__ __device__ static constexpr int smem_bytes(int split) {
    return kPipeBytes + kBarBytes + (kDedicatedRecv ? recv_bytes(split) + 16 : 0) + 1024;   // + 1024-B alignment slack
  }
};

struct GemmArgs {
  const uint32_t* wq;
  const uint32_t* sz;
  const __half* bias;
  __half* C;
  int M, K, N, G;
  int kb_per_split;   // k64 blocks per cluster rank
  long long* trace;   // debug (QB200_TRACE builds only): clock64 stamps of CTA (0,0,0)
};
#ifdef QB200_TRACE
#define QB_TRACE(slot, it, k)                                                                  \
  do {                                                                                         \
    if (args.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)         \
      args.trace[((slot) * 256 + (it)) * 4 + (k)] = clock64();                                 \
  } while (0)
#else
#define QB_TRACE(slot, it, k) do { } while (0)
#endif

// Programmatic dependent launch (PDL): let the next kernel in the stream start its prologue (barrier init,
// TMEM allocation, weight TMA) while this one is still running, and wait for the previous kernel only where
// its results (the activations) are first consumed.  No-ops when launched without the PDL attribute.
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
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
// shared::cta -> (remote) shared::cluster bulk copy by the TMA engine, completing on the destination CTA's mbarrier
__device__ __forceinline__ void bulk_s2dsmem(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remote_dst),
               "r"(local_src), "r"(bytes), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
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
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
template <int TOK, int SPLIT, int D = default_depth<TOK>(), int NWG = default_nwg<TOK>()>
__global__ void __launch_bounds__(TileCfg<TOK, D, NWG>::kNumThreads, 1)
w4a16_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmArgs args) {
  using Cfg = TileCfg<TOK, D, NWG>;
  constexpr int kProducerWarp = Cfg::kProducerWarp;
  constexpr int kMmaWarp = Cfg::kMmaWarp;
  constexpr int SLICE = TOK / SPLIT;          // token columns owned by one cluster rank
  constexpr int CH = SLICE / 2;               // columns per (owner, warpgroup)
  static_assert(CH >= 1, "TOK / SPLIT must be >= 2");
  constexpr int PIECE = CH < 16 ? CH : 16;    // columns per tcgen05.ld
  constexpr int UNIT = CH < 8 ? CH : 8;       // columns per exchanged vector (UNIT halves = 2*UNIT bytes per lane)
  // Epilogue staging.
  //   recv : partial slices from the other SPLIT-1 ranks as packed fp16, laid out [src][unit][channel][UNIT]
  //          so that the 32 lanes of a warp (consecutive channels) write one contiguous run — DSMEM, like
  //          global memory, wants coalesced warps.  Dedicated region for TOK <= 128, else aliases the
  //          (dead) pipeline stages behind a cluster barrier.
  //   out  : this CTA's [SLICE tokens][128 channels] fp16 tile (aliases the dead pipeline stages), stored
  //          with 16-byte coalesced writes.
  constexpr int kRecvBytes = Cfg::recv_bytes(SPLIT);
  constexpr int kOutBytes = SLICE * kChan * 2;
  static_assert(kRecvBytes % 16 == 0, "recv alignment");
  //   stage: the partial slices this CTA sends, same layout, in LOCAL shared memory (aliases the dead pipeline
  //          stages); one TMA bulk copy per owner moves a slice into the owner's recv slot and completes on the
  //          owner's mbarrier (cp.async.bulk.shared::cluster.shared::cta) — far faster than per-thread
  //          st.shared::cluster, and no release/acquire round trip per owner.
  static_assert((Cfg::kDedicatedRecv ? 0 : kRecvBytes) + kRecvBytes + kOutBytes <= Cfg::kPipeBytes,
                "epilogue staging must fit the (dead) pipeline stages: X stages then W stages are contiguous");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;                                   // D x 2 x [TOK rows][128 B] swizzled
  const uint32_t smem_w = smem_base + D * Cfg::kXStageBytes;           // D x 2 x [2][128][16 B]
  const uint32_t bar_full = smem_w + D * kWStageBytesV3;               // TMA landed (W + X)
  const uint32_t bar_tfull = bar_full + 8 * D;                         // A operand written to TMEM slot
  const uint32_t bar_cons = bar_tfull + 8 * D;                         // slot consumed: 4 dequant warps + MMA commit
  const uint32_t bar_accum = bar_cons + 8 * D;                         // all MMAs of the tile done
  const uint32_t bar_recv = bar_accum + 8;                             // split-K partials from the other ranks landed
  const uint32_t tmem_ptr_smem = bar_recv + 8;
  const uint32_t smem_recv = Cfg::kDedicatedRecv ? ((tmem_ptr_smem + 16 + 15) & ~15u) : smem_x;
  const uint32_t smem_stage = Cfg::kDedicatedRecv ? smem_x : smem_x + kRecvBytes;
  const uint32_t smem_out = smem_stage + kRecvBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = blockIdx.x;
  const int mt = blockIdx.y;
  const int rank = SPLIT > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int KB = args.K / kBK;
  const int kb0 = rank * args.kb_per_split;
  const int nkb = min(args.kb_per_split, KB - kb0);            // k64 blocks of this CTA
  const int nst = (nkb + kSubPerStage - 1) / kSubPerStage;     // pipeline stages (last may hold one block)

  pdl_launch_dependents();
  if (threadIdx.x == 0) QB_TRACE(3, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < D; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_tfull + 8 * i, 4);
      mbar_init(bar_cons + 8 * i, 5);
    }
    mbar_init(bar_accum, 1);
    if constexpr (SPLIT > 1) mbar_init(bar_recv, 1);   // one expect_tx arrive; the senders' bulk copies complete the bytes
    fence_barrier_init();
    fence_proxy_async();
    prefetch_tmap(&tmap_x);
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem) : "memory");
  if (threadIdx.x == 0) QB_TRACE(3, 0, 1);
  if constexpr (SPLIT > 1) cluster_arrive();   // matched by a wait just before the first remote access

  if (warp == kProducerWarp) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t* wsrc = args.wq + (static_cast<size_t>(nt) * KB + kb0) * (kWStageBytes / 4);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nst; ++it) {
      if (it >= D) mbar_wait(bar_cons + 8 * s, ph ^ 1, 1, it);   // the first D slots are free by construction
      if (elect_one()) {
        QB_TRACE(0, it, 0);
        const int nsub = min(kSubPerStage, nkb - it * kSubPerStage);
        const uint32_t bar = bar_full + 8 * s;
        mbar_arrive_expect_tx(bar, nsub * (kWStageBytes + Cfg::kXPanelBytes));
        bulk_g2s(smem_w + s * kWStageBytesV3, wsrc + static_cast<size_t>(it) * (kWStageBytesV3 / 4), nsub * kWStageBytes, bar);
        if (it == 0) pdl_wait_prior_grid();   // weights are constants; the activations come from the previous kernel
        tma_load_2d(smem_x + s * Cfg::kXStageBytes, &tmap_x, bar, (kb0 + it * kSubPerStage) * kBK, mt * TOK);
        if (nsub > 1)
          tma_load_2d(smem_x + s * Cfg::kXStageBytes + Cfg::kXPanelBytes, &tmap_x, bar, (kb0 + it * kSubPerStage + 1) * kBK, mt * TOK);
        QB_TRACE(0, it, 1);
      }
      __syncwarp();
      if (++s == D) { s = 0; ph ^= 1; }
    }
    // Tail: observe the final "consumed" phase of every slot that was used.  The tcgen05.commit arrives are
    // asynchronous; if the CTA exited before they landed they would hit the barrier words of the NEXT CTA
    // scheduled on this SM (same shared-memory layout) and corrupt its phase accounting (seen as rare hangs
    // under back-to-back launches).
    for (int i = 0; i < min(D, nst); ++i) {
      const int it = nst - 1 - i;
      mbar_wait(bar_cons + 8 * (it % D), (it / D) & 1, 6, it);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // Waits only on "A operand in TMEM": the dequant warps observed the TMA barrier of the same slot before
    // writing it, so the X tile of the slot is complete (and ordered) by then.
    constexpr uint32_t idesc = make_idesc_f16(TOK);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nst; ++it) {
      mbar_wait(bar_tfull + 8 * s, ph, 2, it);
      tc_fence_after();
      if (elect_one()) {
        QB_TRACE(1, it, 0);
        const int nsub = min(kSubPerStage, nkb - it * kSubPerStage);
        const uint64_t bdesc = make_smem_desc_sw128(smem_x + s * Cfg::kXStageBytes);
        const uint32_t a_tmem = tmem_base + Cfg::kACol0 + s * Cfg::kASlotCols;
#pragma unroll
        for (int j = 0; j < kBK / 16; ++j) {
          // +32 B (= 2 in 16-B units) of start address per k16 step inside the 128-B swizzle row
          umma_f16_ts(tmem_base, a_tmem + j * 8, bdesc + 2 * j, idesc, (it | j) != 0 ? 1u : 0u);
        }
        if (nsub > 1) {
          const uint64_t bdesc1 = bdesc + (Cfg::kXPanelBytes >> 4);
#pragma unroll
          for (int j = 0; j < kBK / 16; ++j) umma_f16_ts(tmem_base, a_tmem + 32 + j * 8, bdesc1 + 2 * j, idesc, 1u);
        }
        QB_TRACE(1, it, 1);
        umma_commit(bar_cons + 8 * s);     // smem slot + TMEM slot free once these MMAs have completed
        if (it == nst - 1) umma_commit(bar_accum);
        QB_TRACE(1, it, 2);
      }
      __syncwarp();
      if (++s == D) { s = 0; ph ^= 1; }
    }
  } else {
    // ===================== dequant warps: smem nibbles -> registers -> TMEM A operand =====================
    const int wg = warp >> 2;                       // warpgroup w takes stages w, w + NWG, ...
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int ch = quad * 32 + lane;                // output channel within the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const int NG = args.K / args.G;
    const int g32 = args.G >> 5;                    // k32 blocks per group
    const uint32_t* szp = args.sz + static_cast<size_t>(nt) * NG * kChan + ch;
    // group index of each of the 4 k32 blocks of the next stage, advanced incrementally (no divisions in the loop)
    int kq = (kb0 + wg * kSubPerStage) * 2;
    int grp = kq / g32, rem = kq % g32;
    uint32_t szw[4];
    auto load_sz = [&]() {
      int g = grp, r = rem;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        szw[q] = __ldg(szp + static_cast<size_t>(min(g, NG - 1)) * kChan);
        if (++r >= g32) { r = 0; ++g; }
      }
    };
    if (wg < nst) load_sz();
    int s = wg % D;
    uint32_t ph = 0;
    for (int it = wg; it < nst; it += NWG) {
      const int nsub = min(kSubPerStage, nkb - it * kSubPerStage);
      // mbarrier parity waits are only valid one phase ahead.  The other warpgroup consumes stage it-1, and
      // TMA completions arrive out of order, so this warp must first observe stage it-1's barrier itself:
      // otherwise, with slot reuse distance D odd, it could test slot s for stage `it` while the slot is
      // still in the phase of stage it-D and the parity test would alias and pass (seen as rare hangs).
      if constexpr (Cfg::kInOrderWaits) {
        static_assert(!Cfg::kInOrderWaits || NWG == 2, "in-order observation is implemented for two warpgroups");
        if (it > 0) {
          const int sp = s == 0 ? D - 1 : s - 1;
          const uint32_t php = s == 0 ? ph ^ 1 : ph;
          mbar_wait(bar_full + 8 * sp, php, 7, it - 1);
        }
      }
      mbar_wait(bar_full + 8 * s, ph, 3, it);
      if (lane == 0 && quad == 2) QB_TRACE(2, it, 0);
      const uint32_t wbase = smem_w + s * kWStageBytesV3 + ch * 16;
      uint4 w[4];
      w[0] = lds128(wbase);
      w[1] = lds128(wbase + 2048);
      if (nsub > 1) {
        w[2] = lds128(wbase + 4096);
        w[3] = lds128(wbase + 6144);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_cons + 8 * s);     // W nibbles are in registers
      GroupConsts gc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) gc[q] = make_group_consts(szw[q]);
      // advance NWG stages (4 k32 blocks each) and prefetch the next scale/zero words
      rem += 2 * kSubPerStage * NWG;
      while (rem >= g32) { rem -= g32; ++grp; }
      if (it + NWG < nst) load_sz();
      // TMEM slot s is free once the MMAs of stage it - D have completed (the previous phase of bar_cons);
      // the producer already observed that phase before refilling the slot, this wait is the acquire.
      mbar_wait(bar_cons + 8 * s, ph ^ 1, 4, it);
      tc_fence_after();
      const uint32_t a_tmem = tmem_base + lane_addr + Cfg::kACol0 + s * Cfg::kASlotCols;
#pragma unroll
      for (int sub = 0; sub < kSubPerStage; ++sub) {
        if (sub < nsub) {
          uint32_t r[32];
          dequant_word(w[2 * sub].x, gc[2 * sub], r + 0);
          dequant_word(w[2 * sub].y, gc[2 * sub], r + 4);
          dequant_word(w[2 * sub].z, gc[2 * sub], r + 8);
          dequant_word(w[2 * sub].w, gc[2 * sub], r + 12);
          dequant_word(w[2 * sub + 1].x, gc[2 * sub + 1], r + 16);
          dequant_word(w[2 * sub + 1].y, gc[2 * sub + 1], r + 20);
          dequant_word(w[2 * sub + 1].z, gc[2 * sub + 1], r + 24);
          dequant_word(w[2 * sub + 1].w, gc[2 * sub + 1], r + 28);
          tmem_st16(a_tmem + sub * 32, r);
          tmem_st16(a_tmem + sub * 32 + 16, r + 16);
        }
      }
      if (lane == 0 && quad == 2) QB_TRACE(2, it, 1);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tfull + 8 * s);
      if (lane == 0 && quad == 2) QB_TRACE(2, it, 2);
      s += NWG;
      while (s >= D) { s -= D; ph ^= 1; }
    }
  }

  // ===================== epilogue =====================
  const int quad = warp & 3;
  const int wg = warp >> 2;
  const int ch = quad * 32 + lane;
  const bool is_dq = warp < kEpilogueWarps;   // the first two warpgroups run the epilogue
  const uint32_t d_tmem = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const int n0 = nt * kChan;
  const float bias_v = (args.bias != nullptr && is_dq) ? __half2float(args.bias[n0 + ch]) : 0.f;

  if (is_dq) {
    mbar_wait(bar_accum, 0, 5, nst);   // every TMA write landed and every MMA read of this CTA's smem is complete
    tc_fence_after();
    if (threadIdx.x == 0) QB_TRACE(3, 0, 2);
  }
  if constexpr (SPLIT > 1) {
    // Exchange partial tiles through distributed shared memory: rank o owns token columns
    // [o*SLICE, (o+1)*SLICE) and receives the other ranks' partials for them as packed fp16; its own partial
    // stays in TMEM as fp32.  Point-to-point: each sending warp arrives (release.cluster) on the owner's
    // bar_recv after its stores; no cluster-wide barrier on the critical path for TOK <= 128.
    cluster_wait();              // every CTA of the cluster has initialised its barriers
    if constexpr (!Cfg::kDedicatedRecv) {
      cluster_arrive();
      cluster_wait();            // every CTA is past its main loop: the aliased pipeline stages are dead
    }
    if (threadIdx.x == 0) QB_TRACE(3, 2, 0);
    constexpr uint32_t kSliceBytes = kChan * SLICE * 2;
    if (threadIdx.x == 0) mbar_arrive_expect_tx(bar_recv, (SPLIT - 1) * kSliceBytes);
    if (is_dq) {
#pragma unroll 1
      for (int oo = 1; oo < SPLIT; ++oo) {
        const int o = (rank + oo) % SPLIT;
        const uint32_t dst = smem_stage + static_cast<uint32_t>((oo - 1) * kSliceBytes);
#pragma unroll 1
        for (int p = 0; p < CH / PIECE; ++p) {
          uint32_t v[PIECE];
          tmem_ld<PIECE>(d_tmem + o * SLICE + wg * CH + p * PIECE, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < PIECE; i += UNIT) {
            const int unit = (wg * CH + p * PIECE + i) / UNIT;       // unit index inside the slice
            const uint32_t addr = dst + static_cast<uint32_t>((unit * kChan + ch) * (UNIT * 2));
            if constexpr (UNIT == 8) {
              sts_v4(addr, pack_half2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                     pack_half2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                     pack_half2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                     pack_half2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
            } else if constexpr (UNIT == 4) {
              sts_u32(addr, pack_half2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
              sts_u32(addr + 4, pack_half2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
            } else if constexpr (UNIT == 2) {
              sts_u32(addr, pack_half2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            } else {
              sts_u16(addr, __half_as_ushort(__float2half_rn(__uint_as_float(v[i]))));
            }
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
    cluster_arrive();
  }
  if (is_dq) {
    // this thread's CH columns of the owned slice (+ the other ranks' partials) -> fp16 -> staging tile [token][channel]
#pragma unroll 1
    for (int p = 0; p < CH / PIECE; ++p) {
      const int j0 = wg * CH + p * PIECE;
      uint32_t v[PIECE];
      tmem_ld<PIECE>(d_tmem + rank * SLICE + j0, v);
      tmem_wait_ld();
      float acc[PIECE];
#pragma unroll
      for (int i = 0; i < PIECE; ++i) acc[i] = __uint_as_float(v[i]) + bias_v;
      if constexpr (SPLIT > 1) {
#pragma unroll
        for (int r = 0; r < SPLIT - 1; ++r) {
#pragma unroll
          for (int i = 0; i < PIECE; i += UNIT) {
            const int unit = (j0 + i) / UNIT;
            const uint32_t addr = smem_recv + static_cast<uint32_t>(r * (kChan * SLICE * 2) + (unit * kChan + ch) * (UNIT * 2));
            if constexpr (UNIT == 8) {
              const uint4 q = lds128(addr);
              const float2 a = unpack_half2(q.x), b = unpack_half2(q.y), c = unpack_half2(q.z), d = unpack_half2(q.w);
              acc[i] += a.x; acc[i + 1] += a.y; acc[i + 2] += b.x; acc[i + 3] += b.y;
              acc[i + 4] += c.x; acc[i + 5] += c.y; acc[i + 6] += d.x; acc[i + 7] += d.y;
            } else if constexpr (UNIT == 4) {
              const float2 a = unpack_half2(lds_u32(addr)), b = unpack_half2(lds_u32(addr + 4));
              acc[i] += a.x; acc[i + 1] += a.y; acc[i + 2] += b.x; acc[i + 3] += b.y;
            } else if constexpr (UNIT == 2) {
              const float2 a = unpack_half2(lds_u32(addr));
              acc[i] += a.x; acc[i + 1] += a.y;
            } else {
              unsigned short hv;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(addr) : "memory");
              acc[i] += __half2float(__ushort_as_half(hv));
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < PIECE; ++i)
        sts_u16(smem_out + static_cast<uint32_t>(((j0 + i) * kChan + ch) * 2), __half_as_ushort(__float2half_rn(acc[i])));
    }
    if (threadIdx.x == 0) QB_TRACE(3, 2, 3);
    named_bar_sync(1, kEpilogueWarps * 32);
    pdl_wait_prior_grid();   // C may alias a buffer the previous kernel still reads/writes
    // coalesced 16-byte stores: 16 threads cover one 256-byte token row of the tile
    const int tid = threadIdx.x;             // 0..255
    const int chunk = tid & 15;
    const int m_base = mt * TOK + rank * SLICE;
#pragma unroll 1
    for (int row = tid >> 4; row < SLICE; row += (kEpilogueWarps * 32) / 16) {
      const int m = m_base + row;
      if (m < args.M) {
        const uint4 v = lds128(smem_out + static_cast<uint32_t>(row * kChan * 2 + chunk * 16));
        *reinterpret_cast<uint4*>(args.C + static_cast<size_t>(m) * args.N + n0 + chunk * 8) = v;
      }
    }
    if (threadIdx.x == 0) QB_TRACE(3, 0, 3);
  }

  if constexpr (SPLIT > 1) cluster_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  if (threadIdx.x == kMmaWarp * 32) QB_TRACE(3, 1, 0);
}

}  // namespace qb200
