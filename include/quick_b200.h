/*
 * quick_b200 — C-ABI of the B200-native W4A16 grouped GEMM (drop-in for the one
 * native entry point of SqueezeBits/QUICK).
 *
 * Reference interface replaced (all paths under /root/reference):
 *   csrc/gemm_cuda_quick.h:3-8      torch::Tensor gemm_forward_cuda_quick(in_feats, kernel,
 *                                   scaling_factors, zeros, split_k_iters)
 *   csrc/pybind.cpp:5-8             module `quick_kernels`, symbol `gemm_forward_cuda_quick`
 *   csrc/gemm_cuda_quick.cu:1456-1517  host dispatcher (argument checks :1479-1484)
 *   quick/awq/modules/linear/quick.py:52-54   packed operand shapes
 *   quick/awq/modules/linear/quick.py:88-150  offline interleave/packer (qb200_pack_quick)
 *
 * Plain pointers and sizes only; no torch types.  Unless a function says "host",
 * every pointer is a DEVICE pointer and work is enqueued on `stream` (a
 * cudaStream_t passed as void*; NULL = legacy default stream) without
 * synchronising.  All functions return 0 on success or a negative QB200_E* code;
 * qb200_last_error() gives the message of the calling thread's last failure.
 *
 * Operand formats
 *   QUICK layout (what the reference's checkpoints and WQLinear_QUICK hold):
 *     qweight int32 [K/4][N/2], qzeros int32 [K/G][N/4], scales fp16 [K/G][2N]
 *   B200 layout (what the tcgen05 kernel streams; produced once per weight):
 *     wq  uint32 [N/128][K/64][2][128][4]  word = 8 consecutive k of one output channel,
 *                                          nibble order k0,k2,k4,k6,k1,k3,k5,k7
 *     sz  uint32 [N/128][K/G][128]         low half = fp16 scale, high half = fp16(1024 + zero)
 */
#ifndef QUICK_B200_H
#define QUICK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB200_OK 0
#define QB200_EINVAL (-1)   /* bad shape/argument: the reference's std::invalid_argument cases and the checks it lacks */
#define QB200_ECUDA (-2)    /* CUDA runtime / driver error */
#define QB200_ENOSPC (-3)   /* workspace too small */

const char* qb200_version(void);
const char* qb200_last_error(void);

/* Bytes of the B200-layout buffers for a (K, N, G) weight. */
size_t qb200_wq_bytes(int K, int N);
size_t qb200_sz_bytes(int K, int N, int G);

/* Argument validation shared by every entry point: N % 128, G % 32 (reference :1479-1484),
 * plus K % 64, K % G, G <= K which the reference leaves unchecked. */
int qb200_check_shape(int M, int K, int N, int G);

/* QUICK layout -> B200 layout (bit-exact nibble permutation + de-duplication of scales/zeros).
 * Replaces nothing in the reference: it is the load-time transform that lets the unchanged
 * checkpoint format feed the tcgen05 kernel. */
int qb200_relayout_from_quick(const int32_t* qweight, const int32_t* qzeros, const void* scales_fp16,
                              int K, int N, int G, uint32_t* wq, uint32_t* sz, void* stream);

/* Logical (q[K][N] uint8 0..15, z[K/G][N] uint8, s[K/G][N] fp16) -> QUICK layout.
 * GPU replacement of the python packer quick.py:88-150 (no N==128 / N%256 restriction). */
int qb200_pack_quick(const uint8_t* q, const uint8_t* z, const void* s_fp16, int K, int N, int G,
                     int32_t* qweight, int32_t* qzeros, void* scales_fp16, void* stream);

/* AWQ "GEMM" checkpoint layout (what public AWQ checkpoints ship: qweight int32 [K][N/8], qzeros int32 [K/G][N/8],
 * nibble i of a word = column 8c + {0,2,4,6,1,3,5,7}[i]; scales fp16 [K/G][N]) -> QUICK-layout tensors, bit-exact.
 * The reference has no such converter: its GEMM packer is quick/awq/modules/linear/gemm.py:108-143 and the inverse
 * nibble order is in quick/awq/utils/packing_utils.py:4-39; QUICK models had to be re-quantized from fp16. */
int qb200_awq_gemm_to_quick(const int32_t* gemm_qweight, const int32_t* gemm_qzeros, const void* gemm_scales_fp16,
                            int K, int N, int G, int32_t* qweight, int32_t* qzeros, void* scales_fp16, void* stream);
/* Same source format straight into the B200 layout the kernel streams (skips the QUICK intermediate). */
int qb200_relayout_from_awq_gemm(const int32_t* gemm_qweight, const int32_t* gemm_qzeros, const void* gemm_scales_fp16,
                                 int K, int N, int G, uint32_t* wq, uint32_t* sz, void* stream);

/* B200 layout -> W16[K][N] fp16 = fp16((q - z)) * s, one rounding (gemm_cuda_quick.cu:52-60). */
int qb200_dequantize(const uint32_t* wq, const uint32_t* sz, int K, int N, int G, void* w16_fp16, void* stream);

/* C[M][N] fp16 = A[M][K] fp16 · dequant(W) — the hot path (tcgen05/TMEM/TMA, sm_100a).
 * split_k_hint: the reference's split_k_iters; treated as a hint (0 = auto).
 * bias (fp16 [N]) is added in the epilogue when non-NULL (reference: quick.py:165, a separate torch add).
 * No workspace: split-K partials are reduced inside a thread-block cluster through distributed shared memory. */
int qb200_gemm_w4a16(const void* A_fp16, const uint32_t* wq, const uint32_t* sz, const void* bias_fp16_or_null,
                     void* C_fp16, int M, int K, int N, int G, int split_k_hint, void* stream);

/* Same, forcing the tile configuration (tuning / tests): tok in {16,32,64,128,256}, split in {1,2,4,8}. */
int qb200_gemm_w4a16_cfg(const void* A_fp16, const uint32_t* wq, const uint32_t* sz, const void* bias_fp16_or_null,
                         void* C_fp16, int M, int K, int N, int G, int tok, int split, void* stream);

/* Same with launch flags; tok = 0 / split = 0 let the planner choose.
 * QB200_GEMM_INDEPENDENT: the caller guarantees that A, C and the weights of this launch are neither written
 * nor (for C) read by any earlier kernel of the stream that may still be running — e.g. sibling projections
 * of the same activations, or a batch of unrelated GEMMs.  The launch then overlaps its predecessor completely
 * under programmatic dependent launch instead of waiting for it before the first activation load; stream order
 * of COMPLETION is preserved (a later ordinary launch still sees all results).  The reference has no such
 * mode: its launches serialise on the legacy default stream (gemm_cuda_quick.cu:1491-1513). */
#define QB200_GEMM_INDEPENDENT 1u
/* SwiGLU fused into the epilogue (reference modules/fused/mlp.py:52-76: silu(gate_proj(x)) * up_proj(x)): the weight is
 * the gate|up pair with its OUTPUT CHANNELS INTERLEAVED (channel 2i = gate_i, 2i+1 = up_i; quick_b200.ops.interleave_pairs
 * prepares it once from the [gate | up] concatenation), C is [M][N/2] (row stride N/2, or ld_c for gathered buffers, col0 in
 * units of output columns) and C[m][i] = fp16(silu(fp16(gate_i)) ) * fp16(up_i) — the rounding of the unfused GEMM +
 * qb200_silu_mul pair, so the results are bit-identical to it.  No residual. */
#define QB200_GEMM_SILU_MUL 2u
int qb200_gemm_w4a16_ex(const void* A_fp16, const uint32_t* wq, const uint32_t* sz, const void* bias_fp16_or_null,
                        void* C_fp16, int M, int K, int N, int G, int tok, int split, unsigned flags, void* stream);

/* Same with a fused residual: C = residual + fp16(A·W + bias), residual fp16 [M][N] or NULL (the `x + proj(...)` of a
 * decoder layer, reference modules/fused/block.py:61-74, without a separate add kernel).  C may alias residual. */
int qb200_gemm_w4a16_fused(const void* A_fp16, const uint32_t* wq, const uint32_t* sz, const void* bias_fp16_or_null,
                           const void* residual_fp16_or_null, void* C_fp16, int M, int K, int N, int G, int tok, int split,
                           unsigned flags, void* stream);

/* RMSNorm folded around the GEMMs of a decoder layer (reference modules/fused/block.py:61-74 and
 * norm.py:16-19 run norm -> linear as separate kernels).  With x·W' where x = rmsnorm(h) = h · rstd(h) ⊙ gamma:
 *     rmsnorm(h) · W  ==  rstd(h) · ((h ⊙ gamma) · W)            (rstd is a per-row scalar),
 * so the GEMM that PRODUCES h (o_proj / down_proj with the fused residual) also emits h ⊙ gamma and the row's sum of
 * squares, and the GEMM that CONSUMES the normed rows scales its output rows by rstd — no RMSNorm kernel in between.
 *   producer side:  gamma_fp16 [N] != NULL -> the epilogue writes C as usual, normed_out[M][N] = fp16(C ⊙ gamma) and
 *                   ssq_out[N/128][M] (fp32): per 128-column tile, the sum of squares of the fp16 values stored to C
 *                   (fixed summation order: bit-reproducible).  Not with QB200_GEMM_SILU_MUL / gathered outputs.
 *   consumer side:  ssq_in [ssq_parts][M] != NULL (A is a producer's normed_out, ssq_parts = K / 128): every output row m
 *                   becomes  rsqrt(sum_p ssq_in[p][m] / K + eps) · (A·W)[m] + bias  (then SiLU·up / residual as usual).
 * Rounding differs from rmsnorm-then-GEMM only in where the fp16 roundings sit (one of h·gamma instead of one of
 * h·rstd and one of ·gamma); tests bound it against the unfused pair.  Either side may be used alone. */
typedef struct qb200_norm_fusion {
  const void* gamma_fp16;   /* producer: weight of the RMSNorm that reads C next, or NULL */
  void* normed_out_fp16;    /* producer: [M][N] */
  float* ssq_out;           /* producer: [N/128][M] */
  const float* ssq_in;      /* consumer: [ssq_parts][M], or NULL */
  int ssq_parts;            /* consumer: K / 128 */
  float eps;                /* consumer: the RMSNorm's epsilon */
} qb200_norm_fusion;
int qb200_gemm_w4a16_norm(const void* A_fp16, const uint32_t* wq, const uint32_t* sz, const void* bias_fp16_or_null,
                          const void* residual_fp16_or_null, void* C_fp16, int M, int K, int N, int G, int tok, int split,
                          unsigned flags, const qb200_norm_fusion* norm, void* stream);

/* Fused GEMM + all-gather for column-parallel (tensor-parallel) linears: this rank computes its N output columns and the
 * epilogue stores the slab straight into the full-width buffers of ALL ranks — C_peers[r] is rank r's [rows][ld_c] buffer
 * mapped into this process (CUDA IPC / torch symmetric memory; C_peers[rank] is the local one), the slab lands at column
 * col0 (= rank * N) — so the transfer over NVLink rides on the GEMM's own stores instead of a separate NCCL all-gather and
 * re-layout copy.  residual (optional) is this rank's local full-width [rows][ld_c] tensor, read at the same columns.
 * C_multicast (optional): the NVSwitch multicast mapping of the same buffers (torch symmetric memory `multicast_ptr`);
 * when given, every store is ONE multimem.st that the switch replicates to all ranks instead of a loop over C_peers.
 * Follow it with qb200_peer_barrier on the same stream before anything reads the gathered rows.  The reference has no
 * multi-GPU path for this operator; the NCCL baseline is quick_b200/parallel.py. */
int qb200_gemm_w4a16_allgather(const void* A_fp16, const uint32_t* wq, const uint32_t* sz, const void* bias_fp16_or_null,
                               const void* residual_fp16_or_null, void* const* C_peers, void* C_multicast_or_null,
                               int n_peers, int ld_c, int col0, int M, int K, int N, int G, int tok, int split,
                               unsigned flags, void* stream);
/* All ranks meet: rank publishes a fresh epoch (device counter *epoch_counter, bumped by the kernel) into slot [rank] of
 * every peer's flag array and waits for all n_peers slots of its own array.  flag_arrays[r] = rank r's array of >= n_peers
 * zero-initialised uint32 in peer-mapped memory.  Traps after ~2 s instead of hanging if a peer never arrives. */
int qb200_peer_barrier(unsigned* epoch_counter, unsigned* const* flag_arrays, int rank, int n_peers, void* stream);

/* ---- Tensor parallelism without barrier kernels: the hand-over rides in the producing and the consuming kernels. ----
 * A "gathered buffer" is a [rows][ld] fp16 buffer that exists on every rank (peer-mapped, e.g. torch symmetric memory);
 * every rank's producer stores its column slab into ALL copies.  Per gathered buffer and rank there is a local epoch
 * counter (zero-initialised uint32 in ordinary device memory) and a flag array of n_peers zero-initialised uint32 in
 * peer-mapped memory.
 *   qb200_peer_signal  given to the kernel that FILLS the buffer: one of its threads bumps *epoch (after the kernel's own
 *                      stream predecessor has completed).  The fill itself is plain peer / multicast stores.
 *   qb200_peer_wait    given to the FIRST kernel that reads the buffer after a fill, on every rank: once its own stream
 *                      predecessor — hence this rank's producer grid, hence the delivery of this rank's slab — has
 *                      completed, it writes *epoch into slot [rank] of every rank's flag array (st.release.sys) and polls its
 *                      own array until every rank has announced that epoch (traps after QB200_PEER_TIMEOUT_S, default 120 s,
 *                      0 = never).  Later readers on the same stream need no wait.
 * Rules: on every rank the same sequence of fills per buffer (SPMD); every fill is followed by a waiting reader; the readers
 * of fill i on a rank precede that rank's producer of fill i+1 in stream order; and between two fills of one buffer every
 * rank runs the producer of some other gathered buffer that it launches after its readers of the first fill (a decoder
 * layer alternates four buffers), so no rank's slab can overwrite rows a slower rank is still reading.  The epoch lives on
 * the device: CUDA-graph replays work.  The reference has no multi-GPU path for this operator. */
typedef struct { const unsigned* epoch; unsigned* const* flag_arrays; int rank; int n_peers; } qb200_peer_wait;
typedef struct { unsigned* epoch; } qb200_peer_signal;

/* The GEMM of a tensor-parallel layer.  n_peers = 0: local output C_local [M][N] (e.g. head-sharded q|k|v, gate|up);
 * n_peers >= 1: column-parallel with the gather fused into the epilogue exactly like qb200_gemm_w4a16_allgather (C_peers,
 * C_multicast_or_null, ld_c, col0; C_local ignored).  wait (or NULL): A lives in a gathered buffer; signal (or NULL): C is a
 * gathered buffer.  Weights keep streaming while the wait polls: the hand-over costs no kernel of its own. */
int qb200_gemm_w4a16_tp(const void* A_fp16, const uint32_t* wq, const uint32_t* sz, const void* bias_fp16_or_null,
                        const void* residual_fp16_or_null, void* C_local_fp16, void* const* C_peers, void* C_multicast_or_null,
                        int n_peers, int ld_c, int col0, int M, int K, int N, int G, int tok, int split, unsigned flags,
                        const qb200_peer_wait* wait, const qb200_peer_signal* signal, void* stream);
/* qb200_rmsnorm on rows of a gathered buffer (wait may be NULL). */
int qb200_rmsnorm_tp(const void* x_fp16, const void* weight_fp16, void* y_fp16, int rows, int H, float eps,
                     const qb200_peer_wait* wait, void* stream);
/* silu(g) * u of this rank's gate|up slab [rows][2 I], written into every rank's [rows][ld] activation buffer at column
 * col0 (the gathered input of the column-parallel down projection) and published. */
int qb200_silu_mul_tp(const void* gate_up_fp16, long long rows, int I, void* const* act_peers, void* act_multicast_or_null,
                      int n_peers, int ld, int col0, const qb200_peer_signal* signal, void* stream);
/* src [rows][n_local] -> every rank's [rows][ld] buffer at column col0, published (this rank's attention heads -> the
 * gathered input of the column-parallel output projection). */
int qb200_scatter_cols(const void* src_fp16, long long rows, int n_local, void* const* dst_peers, void* dst_multicast_or_null,
                       int n_peers, int ld, int col0, const qb200_peer_signal* signal, void* stream);

/* Reports the configuration qb200_gemm_w4a16 would pick. */
int qb200_gemm_plan(int M, int K, int N, int G, int split_k_hint, int* tok, int* split, int* ctas);
/* Same for a given set of launch flags (the independent plan never splits K). */
int qb200_gemm_plan_ex(int M, int K, int N, int G, int split_k_hint, unsigned flags, int* tok, int* split, int* ctas);

/* Stateless drop-in for gemm_forward_cuda_quick on QUICK-layout operands: relayout into the caller's
 * workspace (>= qb200_wq_bytes + qb200_sz_bytes, 256-B aligned) then GEMM.  The torch binding
 * caches the relayout per weight instead (quick_kernels_ext.cpp). */
int qb200_gemm_forward_quick(const void* A_fp16, const int32_t* qweight, const void* scales_fp16,
                             const int32_t* qzeros, void* C_fp16, int M, int K, int N, int G,
                             int split_k_iters, void* workspace, size_t workspace_bytes, void* stream);

/* CUDA-core cross-check of the same contraction (tests only; never dispatched to by the hot path). */
int qb200_gemm_w4a16_simt(const void* A_fp16, const uint32_t* wq, const uint32_t* sz, void* C_fp16,
                          int M, int K, int N, int G, void* stream);

/* ---- Decoder-layer glue: the kernels between the GEMMs of a Llama-like layer.  The reference's fused
 * modules call awq_ext for these (modules/fused/norm.py:18, attn.py:100-245, mlp.py:52-76); awq_ext is not in its tree. ---- */
/* y[rows][H] = fp16(fp16(x * rsqrt(mean(x^2) + eps)) * weight), statistics in fp32. */
int qb200_rmsnorm(const void* x_fp16, const void* weight_fp16, void* y_fp16, int rows, int H, float eps, void* stream);
/* qkv [B][T][(nh + 2 nkv) hd] -> rotary-embedded q_out [B][T][nh][hd] (token-major); rotary k and plain v are written into the static
 * caches [B][nkv][S][hd] at positions pos[t] (int64 device array); cos/sin tables are [S][hd] fp16.  hd is a multiple of
 * 16 and every pointer 16-byte aligned (16-byte vector accesses). */
int qb200_rope_kv_update(const void* qkv_fp16, const void* cos_table_fp16, const void* sin_table_fp16, const long long* pos,
                         void* q_out_fp16, void* cache_k_fp16, void* cache_v_fp16, int B, int T, int nh, int nkv, int hd, int S,
                         void* stream);
/* Decode-step attention fused with rotary embedding and KV-cache update: qkv [B][1][(nh + 2 nkv) hd] -> out [B][1][nh hd];
 * rotary k and plain v are written into the static caches [B][nkv][S][hd] at position pos[0] and the pos[0] + 1 cached
 * positions are attended (softmax in fp32, scale = 1/sqrt(hd) usually).  The reference does this in
 * QuantAttentionFused.forward's single-token branch (quick/awq/modules/fused/attn.py:187-245) through awq_ext kernels
 * that are not part of its tree.  qb200_attn_decode_smem_bytes returns -1 when the configuration is not supported
 * (nh / nkv not in {1, 2, 4, 8}, hd not in {64, 128, 256}, or a cache too long for shared memory). */
int qb200_attn_decode_smem_bytes(int nh, int nkv, int hd, int S);
int qb200_attn_decode(const void* qkv_fp16, const void* cos_table_fp16, const void* sin_table_fp16, const long long* pos,
                      void* out_fp16, void* cache_k_fp16, void* cache_v_fp16, int B, int nh, int nkv, int hd, int S, float scale,
                      void* stream);
/* qb200_attn_decode of this rank's heads (tensor parallel: head-sharded attention) whose output goes straight into every
 * rank's [B][ld] attention buffer at column col0 (a gathered buffer: the input of the column-parallel output projection);
 * plain peer stores, `signal` as for the other producers.  nh / nkv are the LOCAL head counts. */
int qb200_attn_decode_tp(const void* qkv_fp16, const void* cos_table_fp16, const void* sin_table_fp16, const long long* pos,
                         void* cache_k_fp16, void* cache_v_fp16, int B, int nh, int nkv, int hd, int S, float scale,
                         void* const* out_peers, int n_peers, int ld, int col0, const qb200_peer_signal* signal, void* stream);
/* act[rows][I] = silu(g) * u for gate_up rows [g | u] of width 2I. */
int qb200_silu_mul(const void* gate_up_fp16, void* act_fp16, long long rows, int I, void* stream);
/* Same for rows with interleaved columns (g_0, u_0, g_1, u_1, ...): the output of a QB200_GEMM_SILU_MUL weight run without
 * the flag (large token tiles, where the fused epilogue's math would idle the tensor pipe). */
int qb200_silu_mul_interleaved(const void* gate_up_fp16, void* act_fp16, long long rows, int I, void* stream);

/* ---- HOST-buffer handle API (the end-to-end path: H2D, GEMM, D2H inside the call) ---- */
typedef struct qb200_linear qb200_linear;

/* Uploads QUICK-layout HOST tensors, converts them to the B200 layout on the device, keeps both
 * staging buffers for activations/outputs of up to max_m rows.  bias_fp16 may be NULL. */
int qb200_linear_create(qb200_linear** out, const int32_t* qweight_host, const int32_t* qzeros_host,
                        const void* scales_fp16_host, const void* bias_fp16_host,
                        int K, int N, int G, int max_m, int device);
/* y[M][N] = x[M][K] · W (+ bias): copies x from host, runs the kernel, copies y back, synchronises. */
int qb200_linear_forward_host(qb200_linear* h, const void* x_fp16_host, void* y_fp16_host, int M);
/* Asynchronous form: enqueues H2D(x) -> GEMM -> D2H(y) on the handle's private stream and returns; x_host and
 * y_host must be pinned and stay valid until qb200_linear_synchronize(h).  Calls on different handles overlap
 * (PCIe in both directions and the GEMMs run concurrently); calls on one handle run in order. */
int qb200_linear_forward_host_async(qb200_linear* h, const void* x_fp16_host, void* y_fp16_host, int M);
int qb200_linear_synchronize(qb200_linear* h);
/* Device-pointer forward on the handle's weights (async on `stream`). */
int qb200_linear_forward(qb200_linear* h, const void* x_fp16_dev, void* y_fp16_dev, int M, void* stream);
void qb200_linear_destroy(qb200_linear* h);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches claim). */
unsigned long long qb200_launch_count(void);

/* Debug only: device buffer of >= 6*256*4 int64 that CTA (0,0,0) of every later GEMM launch fills with
 * clock64() stamps per warp role and k-stage (NULL disables; disabled by default). */
void qb200_debug_set_trace(void* device_buffer);
/* Debug only: forces one of the tile configurations (ring depths / warpgroup and issuer counts) compiled into
 * QB200_VARIANTS builds for A/B measurements; -1 = the launch planner's choice (default).  No effect in the
 * shipped library, where the planner always chooses. */
void qb200_debug_set_variant(int variant);

#ifdef __cplusplus
}
#endif
#endif /* QUICK_B200_H */
