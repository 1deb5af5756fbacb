// Python module `quick_kernels` — the drop-in for the reference's only native module
// (/root/reference/csrc/pybind.cpp:5-8, csrc/gemm_cuda_quick.h:3-8).  Same module name, same symbol,
// same positional signature, same return-shape quirk, same ValueErrors; underneath it calls the
// quick_b200 C-ABI (include/quick_b200.h).  torch is plumbing here: allocation, current stream,
// and the lifetime tracking that lets the one-time QUICK -> B200 relayout be cached per weight.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>
#include <cmath>

#include <mutex>
#include <stdexcept>
#include <unordered_map>

#include "../../include/quick_b200.h"

namespace {

void check(int rc) {
  if (rc == QB200_OK) return;
  // the reference throws std::invalid_argument for shape errors (gemm_cuda_quick.cu:1479-1484) -> ValueError
  if (rc == QB200_EINVAL) throw std::invalid_argument(qb200_last_error());
  throw std::runtime_error(qb200_last_error());
}

struct Prepacked {
  c10::weak_intrusive_ptr<c10::StorageImpl> w_store, z_store, s_store;
  const void *w_ptr, *z_ptr, *s_ptr;
  uint32_t w_ver, z_ver, s_ver;
  int K, N, G;
  torch::Tensor wq, sz;
};

std::mutex g_mu;
std::unordered_map<const void*, Prepacked> g_cache;   // keyed by qweight StorageImpl*
size_t g_hits = 0, g_misses = 0;

uint32_t version_of(const torch::Tensor& t) { return t.is_inference() ? 0u : static_cast<uint32_t>(t._version()); }

struct Shapes { int M, K, N, G; };

Shapes derive_shapes(const torch::Tensor& in_feats, const torch::Tensor& kernel, const torch::Tensor& scales,
                     const torch::Tensor& zeros) {
  TORCH_CHECK(in_feats.dim() == 2, "in_feats must be 2-D (M, K)");
  TORCH_CHECK(kernel.dim() == 2 && scales.dim() == 2 && zeros.dim() == 2, "packed operands must be 2-D");
  TORCH_CHECK(in_feats.is_cuda() && kernel.is_cuda() && scales.is_cuda() && zeros.is_cuda(),
              "quick_kernels: all operands must be CUDA tensors (there is no CPU path)");
  Shapes s;
  s.M = static_cast<int>(in_feats.size(0));
  s.K = static_cast<int>(in_feats.size(1));
  s.N = static_cast<int>(kernel.size(1) / 4 * 8);                        // reference: gemm_cuda_quick.cu:1468
  TORCH_CHECK(scales.size(0) > 0, "scales has no rows");
  s.G = static_cast<int>(s.K / scales.size(0));                         // reference: :1477
  check(qb200_check_shape(s.M, s.K, s.N, s.G));
  TORCH_CHECK(kernel.size(0) * 4 == s.K, "qweight rows (", kernel.size(0), ") != K/4 for K=", s.K);
  TORCH_CHECK(scales.size(1) == 2 * s.N, "scales must be (K/G, 2N)");
  TORCH_CHECK(zeros.size(0) == scales.size(0) && zeros.size(1) * 4 == s.N, "qzeros must be (K/G, N/4)");
  return s;
}

std::pair<torch::Tensor, torch::Tensor> relayout(const torch::Tensor& kernel, const torch::Tensor& scales,
                                                 const torch::Tensor& zeros, int K, int N, int G) {
  auto opts = torch::TensorOptions().dtype(torch::kInt32).device(kernel.device());
  torch::Tensor wq = torch::empty({static_cast<int64_t>(qb200_wq_bytes(K, N) / 4)}, opts);
  torch::Tensor sz = torch::empty({static_cast<int64_t>(qb200_sz_bytes(K, N, G) / 4)}, opts);
  auto stream = at::cuda::getCurrentCUDAStream();
  check(qb200_relayout_from_quick(kernel.data_ptr<int>(), zeros.data_ptr<int>(), scales.data_ptr<at::Half>(), K, N, G,
                                  reinterpret_cast<uint32_t*>(wq.data_ptr<int>()),
                                  reinterpret_cast<uint32_t*>(sz.data_ptr<int>()), stream.stream()));
  return {wq, sz};
}

// Returns the cached B200-layout copy of (kernel, scales, zeros), building it on first use.  An entry
// is valid only while all three storages are alive (weak refs keep the StorageImpl addresses from
// being recycled) and unmodified (version counters), so a freed-and-reallocated or in-place-updated
// weight can never hit a stale entry.
std::pair<torch::Tensor, torch::Tensor> get_prepacked(const torch::Tensor& kernel, const torch::Tensor& scales,
                                                      const torch::Tensor& zeros, int K, int N, int G) {
  const void* key = kernel.storage().unsafeGetStorageImpl();
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    Prepacked& e = it->second;
    const bool ok = !e.w_store.expired() && !e.z_store.expired() && !e.s_store.expired() &&
                    e.w_ptr == kernel.data_ptr() && e.z_ptr == zeros.data_ptr() && e.s_ptr == scales.data_ptr() &&
                    e.w_ver == version_of(kernel) && e.z_ver == version_of(zeros) && e.s_ver == version_of(scales) &&
                    e.K == K && e.N == N && e.G == G;
    if (ok) {
      ++g_hits;
      return {e.wq, e.sz};
    }
    g_cache.erase(it);
  }
  ++g_misses;
  if (g_cache.size() >= 64 && (g_misses % 64) == 0) {   // sweep entries whose weights were freed
    for (auto i = g_cache.begin(); i != g_cache.end();) i = i->second.w_store.expired() ? g_cache.erase(i) : std::next(i);
  }
  auto packed = relayout(kernel, scales, zeros, K, N, G);
  Prepacked e{kernel.storage().getWeakStorageImpl(), zeros.storage().getWeakStorageImpl(),
              scales.storage().getWeakStorageImpl(), kernel.data_ptr(), zeros.data_ptr(), scales.data_ptr(),
              version_of(kernel), version_of(zeros), version_of(scales), K, N, G, packed.first, packed.second};
  g_cache.emplace(key, std::move(e));
  return packed;
}

}  // namespace

// Reference signature: csrc/gemm_cuda_quick.h:3-8.  Positional call order from Python is
// (x2d, qweight, scales, qzeros, split_k)  (quick/awq/modules/linear/quick.py:162,164).
torch::Tensor gemm_forward_cuda_quick(torch::Tensor _in_feats, torch::Tensor _kernel, torch::Tensor _scaling_factors,
                                      torch::Tensor _zeros, int split_k_iters) {
  TORCH_CHECK(_in_feats.is_cuda(), "quick_kernels: all operands must be CUDA tensors (there is no CPU path)");
  const at::cuda::OptionalCUDAGuard device_guard(device_of(_in_feats));
  const Shapes s = derive_shapes(_in_feats, _kernel, _scaling_factors, _zeros);
  if (split_k_iters < 1) throw std::invalid_argument("split_k_iters must be >= 1");
  torch::Tensor in_feats = _in_feats.contiguous();
  torch::Tensor kernel = _kernel.contiguous(), scales = _scaling_factors.contiguous(), zeros = _zeros.contiguous();
  (void)in_feats.data_ptr<at::Half>();   // dtype errors surface exactly like the reference's data_ptr<T>() calls
  auto packed = get_prepacked(kernel, scales, zeros, s.K, s.N, s.G);
  auto options = torch::TensorOptions().dtype(in_feats.dtype()).device(in_feats.device());
  torch::Tensor out = torch::empty({s.M, s.N}, options);
  auto stream = at::cuda::getCurrentCUDAStream();
  check(qb200_gemm_w4a16(in_feats.data_ptr<at::Half>(), reinterpret_cast<const uint32_t*>(packed.first.data_ptr<int>()),
                         reinterpret_cast<const uint32_t*>(packed.second.data_ptr<int>()), nullptr,
                         out.data_ptr<at::Half>(), s.M, s.K, s.N, s.G, split_k_iters, stream.stream()));
  // the reference returns (1, M, N) when split_k_iters == 1 and (M, N) otherwise (…cu:1515-1516)
  if (split_k_iters == 1) return out.view({1, s.M, s.N});
  return out;
}

// Explicit one-time conversion for callers that hold weights for a long time (WQLinear_QUICK).
std::vector<torch::Tensor> prepack_quick(torch::Tensor kernel, torch::Tensor scales, torch::Tensor zeros, int64_t K) {
  TORCH_CHECK(kernel.is_cuda() && scales.is_cuda() && zeros.is_cuda(), "prepack_quick: CUDA tensors required (there is no CPU path)");
  const at::cuda::OptionalCUDAGuard device_guard(device_of(kernel));
  const int N = static_cast<int>(kernel.size(1) / 4 * 8);
  const int G = static_cast<int>(K / scales.size(0));
  check(qb200_check_shape(1, static_cast<int>(K), N, G));
  auto p = relayout(kernel.contiguous(), scales.contiguous(), zeros.contiguous(), static_cast<int>(K), N, G);
  return {p.first, p.second};
}

// GEMM on already-converted weights, bias fused into the epilogue.
// independent = true passes QB200_GEMM_INDEPENDENT (include/quick_b200.h): the caller guarantees that the operands
// are not produced by a kernel that may still be running (e.g. sibling projections of one activation tensor).
torch::Tensor gemm_forward_b200(torch::Tensor in_feats, torch::Tensor wq, torch::Tensor sz,
                                c10::optional<torch::Tensor> bias, int64_t N, int64_t G, bool independent,
                                c10::optional<torch::Tensor> residual, bool silu_mul) {
  TORCH_CHECK(in_feats.dim() == 2 && in_feats.is_cuda(), "in_feats must be a 2-D CUDA tensor (there is no CPU path)");
  const at::cuda::OptionalCUDAGuard device_guard(device_of(in_feats));
  torch::Tensor x = in_feats.contiguous();
  const int M = static_cast<int>(x.size(0)), K = static_cast<int>(x.size(1));
  check(qb200_check_shape(M, K, static_cast<int>(N), static_cast<int>(G)));
  TORCH_CHECK(static_cast<size_t>(wq.numel()) * 4 == qb200_wq_bytes(K, N), "wq size mismatch");
  TORCH_CHECK(static_cast<size_t>(sz.numel()) * 4 == qb200_sz_bytes(K, N, G), "sz size mismatch");
  const void* bias_ptr = nullptr;
  torch::Tensor b;
  if (bias.has_value() && bias->defined()) {
    b = bias->contiguous();
    TORCH_CHECK(b.numel() == N && b.scalar_type() == torch::kHalf, "bias must be fp16 [N]");
    bias_ptr = b.data_ptr<at::Half>();
  }
  // residual (fp16 [M, N]): out = residual + fp16(x·W + bias), fused into the epilogue (qb200_gemm_w4a16_fused)
  const void* res_ptr = nullptr;
  torch::Tensor r;
  if (residual.has_value() && residual->defined()) {
    r = residual->contiguous();
    TORCH_CHECK(r.numel() == static_cast<int64_t>(M) * N && r.scalar_type() == torch::kHalf && r.is_cuda(), "residual must be CUDA fp16 [M, N]");
    res_ptr = r.data_ptr<at::Half>();
  }
  // silu_mul: the weight is a gate|up pair with interleaved output channels; out = silu(gate) * up, [M, N/2]
  TORCH_CHECK(!(silu_mul && res_ptr != nullptr), "silu_mul takes no residual");
  torch::Tensor out = torch::empty({M, silu_mul ? N / 2 : N}, x.options());
  auto stream = at::cuda::getCurrentCUDAStream();
  check(qb200_gemm_w4a16_fused(x.data_ptr<at::Half>(), reinterpret_cast<const uint32_t*>(wq.data_ptr<int>()),
                               reinterpret_cast<const uint32_t*>(sz.data_ptr<int>()), bias_ptr, res_ptr, out.data_ptr<at::Half>(), M, K,
                               static_cast<int>(N), static_cast<int>(G), /*tok*/ 0, /*split*/ 0,
                               (independent ? QB200_GEMM_INDEPENDENT : 0u) | (silu_mul ? QB200_GEMM_SILU_MUL : 0u), stream.stream()));
  return out;
}

// The per-call work of WQLinear_QUICK.forward in one object: holds the B200-layout copy and the bias, flattens the leading
// dimensions, launches, reshapes — the Python side of a forward is then one method call (the eager module surface is
// host-bound below M = 64: 14 us per call through the Python-level wrapper).
struct B200Linear {
  torch::Tensor wq, sz, bias;
  int64_t K, N, G;
  const void* bias_ptr = nullptr;
  B200Linear(torch::Tensor wq_, torch::Tensor sz_, c10::optional<torch::Tensor> bias_, int64_t K_, int64_t N_, int64_t G_)
      : wq(std::move(wq_)), sz(std::move(sz_)), K(K_), N(N_), G(G_) {
    TORCH_CHECK(wq.is_cuda() && sz.is_cuda(), "B200Linear: CUDA tensors required (there is no CPU path)");
    check(qb200_check_shape(1, static_cast<int>(K), static_cast<int>(N), static_cast<int>(G)));
    TORCH_CHECK(static_cast<size_t>(wq.numel()) * 4 == qb200_wq_bytes(K, N) && static_cast<size_t>(sz.numel()) * 4 == qb200_sz_bytes(K, N, G),
                "B200Linear: wq / sz size mismatch");
    if (bias_.has_value() && bias_->defined()) {
      bias = bias_->contiguous();
      TORCH_CHECK(bias.numel() == N && bias.scalar_type() == torch::kHalf && bias.is_cuda(), "bias must be CUDA fp16 [N]");
      bias_ptr = bias.data_ptr<at::Half>();
    }
  }
  torch::Tensor forward(const torch::Tensor& x_in, const c10::optional<torch::Tensor>& residual) {
    TORCH_CHECK(x_in.is_cuda() && x_in.scalar_type() == torch::kHalf && x_in.dim() >= 1 && x_in.size(-1) == K,
                "input must be a CUDA fp16 tensor [..., in_features] (there is no CPU path)");
    const at::cuda::OptionalCUDAGuard device_guard(device_of(x_in));
    torch::Tensor x = x_in.contiguous();
    const int64_t M = x.numel() / K;
    std::vector<int64_t> shape(x.sizes().begin(), x.sizes().end());
    shape.back() = N;
    torch::Tensor out = torch::empty(shape, x.options());
    const void* res_ptr = nullptr;
    torch::Tensor r;
    if (residual.has_value() && residual->defined()) {
      r = residual->contiguous();
      TORCH_CHECK(r.numel() == M * N && r.scalar_type() == torch::kHalf && r.is_cuda(), "residual must be CUDA fp16 [..., out_features]");
      res_ptr = r.data_ptr<at::Half>();
    }
    if (M == 0) return out;
    check(qb200_gemm_w4a16_fused(x.data_ptr<at::Half>(), reinterpret_cast<const uint32_t*>(wq.data_ptr<int>()),
                                 reinterpret_cast<const uint32_t*>(sz.data_ptr<int>()), bias_ptr, res_ptr, out.data_ptr<at::Half>(),
                                 static_cast<int>(M), static_cast<int>(K), static_cast<int>(N), static_cast<int>(G), 0, 0, 0u,
                                 at::cuda::getCurrentCUDAStream().stream()));
    return out;
  }
};

// GEMM with an RMSNorm folded around it (C-ABI qb200_gemm_w4a16_norm, include/quick_b200.h):
//   norm_gamma given -> producer side: returns {out, out ⊙ gamma (fp16 [M, N]), per-tile sums of squares (fp32 [N/128, M])}
//   ssq_in given     -> consumer side: in_feats is a producer's gamma-scaled copy, rows are scaled by 1/rms before the bias
std::vector<torch::Tensor> gemm_forward_b200_norm(torch::Tensor in_feats, torch::Tensor wq, torch::Tensor sz,
                                                  c10::optional<torch::Tensor> bias, int64_t N, int64_t G,
                                                  c10::optional<torch::Tensor> residual, bool silu_mul,
                                                  c10::optional<torch::Tensor> norm_gamma, c10::optional<torch::Tensor> ssq_in,
                                                  double eps) {
  TORCH_CHECK(in_feats.dim() == 2 && in_feats.is_cuda(), "in_feats must be a 2-D CUDA tensor (there is no CPU path)");
  const at::cuda::OptionalCUDAGuard device_guard(device_of(in_feats));
  torch::Tensor x = in_feats.contiguous();
  const int M = static_cast<int>(x.size(0)), K = static_cast<int>(x.size(1));
  check(qb200_check_shape(M, K, static_cast<int>(N), static_cast<int>(G)));
  TORCH_CHECK(static_cast<size_t>(wq.numel()) * 4 == qb200_wq_bytes(K, N), "wq size mismatch");
  TORCH_CHECK(static_cast<size_t>(sz.numel()) * 4 == qb200_sz_bytes(K, N, G), "sz size mismatch");
  const void* bias_ptr = nullptr;
  torch::Tensor b, r, gam, sq;
  if (bias.has_value() && bias->defined()) {
    b = bias->contiguous();
    TORCH_CHECK(b.numel() == N && b.scalar_type() == torch::kHalf, "bias must be fp16 [N]");
    bias_ptr = b.data_ptr<at::Half>();
  }
  const void* res_ptr = nullptr;
  if (residual.has_value() && residual->defined()) {
    r = residual->contiguous();
    TORCH_CHECK(r.numel() == static_cast<int64_t>(M) * N && r.scalar_type() == torch::kHalf && r.is_cuda(), "residual must be CUDA fp16 [M, N]");
    res_ptr = r.data_ptr<at::Half>();
  }
  TORCH_CHECK(!(silu_mul && res_ptr != nullptr), "silu_mul takes no residual");
  qb200_norm_fusion nf{};
  torch::Tensor out = torch::empty({M, silu_mul ? N / 2 : N}, x.options());
  std::vector<torch::Tensor> ret{out};
  if (norm_gamma.has_value() && norm_gamma->defined()) {
    gam = norm_gamma->contiguous();
    TORCH_CHECK(gam.numel() == N && gam.scalar_type() == torch::kHalf && gam.is_cuda() && !silu_mul, "norm_gamma must be CUDA fp16 [N] (no silu_mul)");
    torch::Tensor normed = torch::empty({M, N}, x.options());
    torch::Tensor parts = torch::empty({N / 128, M}, x.options().dtype(torch::kFloat));
    nf.gamma_fp16 = gam.data_ptr<at::Half>();
    nf.normed_out_fp16 = normed.data_ptr<at::Half>();
    nf.ssq_out = parts.data_ptr<float>();
    ret.push_back(normed);
    ret.push_back(parts);
  }
  if (ssq_in.has_value() && ssq_in->defined()) {
    sq = ssq_in->contiguous();
    TORCH_CHECK(sq.scalar_type() == torch::kFloat && sq.is_cuda() && sq.dim() == 2 && sq.size(0) == K / 128 && sq.size(1) == M,
                "ssq_in must be CUDA fp32 [K/128, M]");
    nf.ssq_in = sq.data_ptr<float>();
    nf.ssq_parts = static_cast<int>(sq.size(0));
    nf.eps = static_cast<float>(eps);
  }
  auto stream = at::cuda::getCurrentCUDAStream();
  check(qb200_gemm_w4a16_norm(x.data_ptr<at::Half>(), reinterpret_cast<const uint32_t*>(wq.data_ptr<int>()),
                              reinterpret_cast<const uint32_t*>(sz.data_ptr<int>()), bias_ptr, res_ptr, out.data_ptr<at::Half>(), M, K,
                              static_cast<int>(N), static_cast<int>(G), /*tok*/ 0, /*split*/ 0, silu_mul ? QB200_GEMM_SILU_MUL : 0u, &nf,
                              stream.stream()));
  return ret;
}

// ---- decoder-layer glue (C-ABI qb200_rmsnorm / qb200_rope_kv_update / qb200_silu_mul) ----
torch::Tensor rmsnorm(torch::Tensor x, torch::Tensor weight, double eps) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kHalf && weight.scalar_type() == torch::kHalf, "rmsnorm: CUDA fp16 tensors required");
  const at::cuda::OptionalCUDAGuard device_guard(device_of(x));
  torch::Tensor xc = x.contiguous(), w = weight.contiguous();
  const int H = static_cast<int>(xc.size(-1));
  TORCH_CHECK(w.numel() == H, "rmsnorm: weight must have H elements");
  torch::Tensor y = torch::empty_like(xc);
  check(qb200_rmsnorm(xc.data_ptr<at::Half>(), w.data_ptr<at::Half>(), y.data_ptr<at::Half>(), static_cast<int>(xc.numel() / H), H,
                      static_cast<float>(eps), at::cuda::getCurrentCUDAStream().stream()));
  return y;
}

torch::Tensor rope_kv_update(torch::Tensor qkv, torch::Tensor cos_table, torch::Tensor sin_table, torch::Tensor pos,
                             torch::Tensor cache_k, torch::Tensor cache_v, int64_t nh, int64_t nkv) {
  TORCH_CHECK(qkv.is_cuda() && qkv.dim() == 3 && qkv.scalar_type() == torch::kHalf, "rope_kv_update: qkv must be CUDA fp16 [B, T, (nh + 2 nkv) hd]");
  TORCH_CHECK(cache_k.is_contiguous() && cache_v.is_contiguous() && cache_k.dim() == 4, "rope_kv_update: caches must be contiguous [B, nkv, S, hd]");
  TORCH_CHECK(pos.scalar_type() == torch::kLong && pos.is_cuda(), "rope_kv_update: pos must be a CUDA int64 tensor");
  const at::cuda::OptionalCUDAGuard device_guard(device_of(qkv));
  torch::Tensor x = qkv.contiguous(), c = cos_table.contiguous(), s = sin_table.contiguous(), p = pos.contiguous();
  const int B = static_cast<int>(x.size(0)), T = static_cast<int>(x.size(1));
  const int hd = static_cast<int>(x.size(2) / (nh + 2 * nkv)), S = static_cast<int>(cache_k.size(2));
  TORCH_CHECK(x.size(2) == (nh + 2 * nkv) * hd && cache_k.size(3) == hd && cache_k.size(1) == nkv && cache_k.size(0) >= B, "rope_kv_update: shape mismatch");
  TORCH_CHECK(c.size(-1) == hd && c.size(0) >= S && p.numel() == T, "rope_kv_update: table / pos shape mismatch");
  torch::Tensor q = torch::empty({B, T, nh, hd}, x.options());     // token-major memory, returned as the [B, nh, T, hd] view
  check(qb200_rope_kv_update(x.data_ptr<at::Half>(), c.data_ptr<at::Half>(), s.data_ptr<at::Half>(),
                             reinterpret_cast<const long long*>(p.data_ptr<int64_t>()), q.data_ptr<at::Half>(),
                             cache_k.data_ptr<at::Half>(), cache_v.data_ptr<at::Half>(), B, T, static_cast<int>(nh),
                             static_cast<int>(nkv), hd, S, at::cuda::getCurrentCUDAStream().stream()));
  return q.transpose(1, 2);
}

// Decode-step attention fused with rotary embedding + KV-cache update (C-ABI qb200_attn_decode): qkv [B, 1, (nh + 2 nkv) hd]
// -> [B, 1, nh hd]; the caches are updated in place at position pos[0].
torch::Tensor attn_decode(torch::Tensor qkv, torch::Tensor cos_table, torch::Tensor sin_table, torch::Tensor pos,
                          torch::Tensor cache_k, torch::Tensor cache_v, int64_t nh, int64_t nkv) {
  TORCH_CHECK(qkv.is_cuda() && qkv.dim() == 3 && qkv.size(1) == 1 && qkv.scalar_type() == torch::kHalf,
              "attn_decode: qkv must be CUDA fp16 [B, 1, (nh + 2 nkv) hd]");
  TORCH_CHECK(cache_k.is_contiguous() && cache_v.is_contiguous() && cache_k.dim() == 4 && cache_k.scalar_type() == torch::kHalf,
              "attn_decode: caches must be contiguous fp16 [B, nkv, S, hd]");
  TORCH_CHECK(pos.scalar_type() == torch::kLong && pos.is_cuda() && pos.numel() == 1, "attn_decode: pos must be a CUDA int64 tensor with one element");
  const at::cuda::OptionalCUDAGuard device_guard(device_of(qkv));
  torch::Tensor x = qkv.contiguous(), c = cos_table.contiguous(), s = sin_table.contiguous(), p = pos.contiguous();
  const int B = static_cast<int>(x.size(0));
  const int hd = static_cast<int>(x.size(2) / (nh + 2 * nkv)), S = static_cast<int>(cache_k.size(2));
  TORCH_CHECK(x.size(2) == (nh + 2 * nkv) * hd && cache_k.size(3) == hd && cache_k.size(1) == nkv && cache_k.size(0) == B &&
              cache_v.sizes() == cache_k.sizes(), "attn_decode: shape mismatch");
  TORCH_CHECK(c.size(-1) == hd && c.size(0) >= S && s.sizes() == c.sizes(), "attn_decode: rotary table shape mismatch");
  torch::Tensor out = torch::empty({B, 1, nh * hd}, x.options());
  check(qb200_attn_decode(x.data_ptr<at::Half>(), c.data_ptr<at::Half>(), s.data_ptr<at::Half>(),
                          reinterpret_cast<const long long*>(p.data_ptr<int64_t>()), out.data_ptr<at::Half>(),
                          cache_k.data_ptr<at::Half>(), cache_v.data_ptr<at::Half>(), B, static_cast<int>(nh), static_cast<int>(nkv),
                          hd, S, 1.0f / std::sqrt(static_cast<float>(hd)), at::cuda::getCurrentCUDAStream().stream()));
  return out;
}

torch::Tensor silu_mul(torch::Tensor gate_up) {
  TORCH_CHECK(gate_up.is_cuda() && gate_up.scalar_type() == torch::kHalf, "silu_mul: CUDA fp16 tensor required");
  const at::cuda::OptionalCUDAGuard device_guard(device_of(gate_up));
  torch::Tensor gu = gate_up.contiguous();
  const int64_t I = gu.size(-1) / 2;
  auto shape = gu.sizes().vec();
  shape.back() = I;
  torch::Tensor act = torch::empty(shape, gu.options());
  check(qb200_silu_mul(gu.data_ptr<at::Half>(), act.data_ptr<at::Half>(), static_cast<long long>(gu.numel() / (2 * I)), static_cast<int>(I),
                       at::cuda::getCurrentCUDAStream().stream()));
  return act;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  py::class_<B200Linear>(m, "B200Linear", "linear(x) (+ bias, + residual) on B200-layout weights: WQLinear_QUICK.forward's per-call work in one call")
      .def(py::init<torch::Tensor, torch::Tensor, c10::optional<torch::Tensor>, int64_t, int64_t, int64_t>(), py::arg("wq"), py::arg("sz"),
           py::arg("bias"), py::arg("K"), py::arg("N"), py::arg("G"))
      .def("forward", &B200Linear::forward, py::arg("x"), py::arg("residual") = py::none());
  m.def("rmsnorm", &rmsnorm, "RMSNorm (fp16 in/out, fp32 statistics)");
  m.def("rope_kv_update", &rope_kv_update, "rotary embedding of q/k + static KV-cache update; returns q [B, nh, T, hd]");
  m.def("silu_mul", &silu_mul, "silu(gate) * up for rows [gate | up]");
  m.def("gemm_forward_b200_norm", &gemm_forward_b200_norm,
        "GEMM with an RMSNorm folded around it: producer side (norm_gamma) returns [out, out*gamma, ssq parts], consumer side (ssq_in) scales rows by 1/rms",
        py::arg("in_feats"), py::arg("wq"), py::arg("sz"), py::arg("bias"), py::arg("N"), py::arg("G"), py::arg("residual") = py::none(),
        py::arg("silu_mul") = false, py::arg("norm_gamma") = py::none(), py::arg("ssq_in") = py::none(), py::arg("eps") = 1e-6);
  m.def("attn_decode", &attn_decode, "one-token attention fused with rotary embedding + static KV-cache update; returns [B, 1, nh*hd]");
  m.def("attn_decode_supported", [](int64_t nh, int64_t nkv, int64_t hd, int64_t S) {
    return qb200_attn_decode_smem_bytes(static_cast<int>(nh), static_cast<int>(nkv), static_cast<int>(hd), static_cast<int>(S)) >= 0;
  });
  m.def("gemm_forward_cuda_quick", &gemm_forward_cuda_quick, "QUICK AWQ GEMM kernel.");
  m.def("prepack_quick", &prepack_quick, "QUICK layout -> B200 layout (wq, sz)");
  m.def("gemm_forward_b200", &gemm_forward_b200, "W4A16 GEMM on B200-layout weights (bias fused)", pybind11::arg("in_feats"),
        pybind11::arg("wq"), pybind11::arg("sz"), pybind11::arg("bias"), pybind11::arg("N"), pybind11::arg("G"),
        pybind11::arg("independent") = false, pybind11::arg("residual") = pybind11::none(), pybind11::arg("silu_mul") = false);
  m.def("cache_stats", [] {
    std::lock_guard<std::mutex> lock(g_mu);
    return std::vector<int64_t>{static_cast<int64_t>(g_cache.size()), static_cast<int64_t>(g_hits), static_cast<int64_t>(g_misses)};
  });
  m.def("clear_cache", [] {
    std::lock_guard<std::mutex> lock(g_mu);
    g_cache.clear();
  });
  m.def("launch_count", [] { return static_cast<int64_t>(qb200_launch_count()); });
  m.def("version", [] { return std::string(qb200_version()); });
}
