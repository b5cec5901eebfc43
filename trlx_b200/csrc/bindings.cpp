// Python bindings (torch tensors in, raw pointers out) + the host-side C++ runtime pieces:
//   * PagedKVAllocator — free-list page allocator behind the paged KV cache used by the rollout engine
// Kernels live in the .cu files of this directory and export a plain C ABI so that they compile in seconds
// without torch headers.
#include <algorithm>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <deque>
#include <stdexcept>
#include <unordered_map>
#include <vector>

using torch::Tensor;
typedef c10::optional<Tensor> OptTensor;

extern "C" {
int b200_gemm_bf16(const void*, const void*, void*, int, int, int, long long, long long, long long, const void*, const void*,
                   long long, const float*, float, int, int, int, cudaStream_t);
int b200_lmhead_tiles(int);
int b200_lmhead_bf16(const void*, const void*, int, int, int, long long, long long, const void*, const long long*, float*,
                     float*, float*, int, float, unsigned long long, const long long*, const int*, int, int, long long*, float*,
                     cudaStream_t);
int b200_norm_bf16(const void*, const void*, const void*, void*, int, int, long long, long long, float, int, cudaStream_t);
int b200_embed_bf16(const long long*, const int*, const void*, const void*, int, void*, int, int, float*, cudaStream_t);
int b200_decode_attention_bf16(const void*, void*, void*, const int*, const int*, const int*, void*, int, int, int, int, int,
                               int, float, int, float, int, const float*, int, cudaStream_t);
int b200_rowdot_bf16(const void*, const void*, const void*, float*, int, int, long long, cudaStream_t);
int b200_ilql_sample(const float*, const float*, const float*, const float*, long long, int, int, float, int, float,
                     unsigned long long, const long long*, const int*, const unsigned char*, int, int, const long long*, long long*,
                     cudaStream_t);
int b200_sample_filtered(const float*, long long, int, int, int, float, float, unsigned long long, const long long*, const int*,
                         int, int, long long*, float*, cudaStream_t);
int b200_decode_step(const long long*, const float*, const float*, const float*, int*, int, int, long long, long long,
                     long long*, float*, float*, float*, int*, int*, int*, int*, long long*, int*, cudaStream_t);
int b200_paged_kv_write(const void*, const void*, void*, void*, const int*, const int*, const int*, int, int, int, int, int,
                        int, long long, long long, cudaStream_t);
int b200_logprob_from_logits(const void*, const long long*, float*, float*, long long, int, long long, int, cudaStream_t);
int b200_gemm_fp8(const void*, const void*, void*, int, int, int, long long, long long, long long, const float*, const float*,
                  const void*, const void*, long long, int, cudaStream_t);
int b200_quant_rows_fp8(const void*, void*, float*, int, int, long long, long long, cudaStream_t);
int b200_norm_quant_fp8(const void*, const void*, const void*, void*, float*, int, int, long long, long long, float, int,
                        cudaStream_t);
int b200_gemm_bf16_ln(const void*, const void*, void*, int, int, int, long long, long long, long long, const void*, const void*,
                      long long, int, const float*, const float*, float, int, float*, cudaStream_t);
int b200_gemm_bf16_ex(const void*, const void*, void*, int, int, int, long long, long long, long long, int, int, int, int, float*,
                      int, cudaStream_t);
int b200_gemm_splitk_plan(int, int, int);
int b200_lmhead_dlogits_bf16(const void*, const void*, void*, int, int, int, long long, long long, long long, const void*,
                             const long long*, const float*, const float*, cudaStream_t);
int b200_logprob_backward_inplace(void*, const long long*, const float*, const float*, long long, int, long long, int,
                                  cudaStream_t);
int b200_gae(const float*, const float*, float*, float*, int, int, int, const int*, long long, float, float, double*,
             cudaStream_t);
int b200_whiten(float*, int, int, const int*, long long, const double*, int, cudaStream_t);
int b200_ppo_loss_num_outputs();
int b200_ppo_loss_workspace_floats(int);
int b200_ppo_loss(const float*, const float*, const float*, const float*, const float*, const float*, const float*, int, float,
                  float, float, float*, float*, float*, int, float*, int, const int*, cudaStream_t);
int b200_rollout_rewards(const float*, const float*, const float*, const long long*, const float*, int, int, int, float, float*,
                         float*, float*, int*, double*, cudaStream_t);
int b200_adam8bit(void*, const void*, int, void*, float*, void*, float*, long long, float, float, float, float, int, float, float,
                  float, cudaStream_t);
int b200_adamw_flat(void*, float*, const void*, int, float*, float*, long long, float, float, float, float, int, const float*,
                    cudaStream_t);
int b200_sqnorm(const void*, int, long long, double*, cudaStream_t);
int b200_clip_coef(const double*, float, float*, float*, cudaStream_t);
int b200_signal_barrier(void* const*, int, int, unsigned int*, cudaStream_t);
int b200_rs_adamw_ag(void* const*, void* const*, int, long long, long long, float*, float*, float*, float*, int, float, float,
                     float, float, int, const float*, double*, cudaStream_t);
int b200_lerp_bf16(void*, const void*, long long, float, cudaStream_t);
int b200_rs_adamw_ag_bucket(void* const*, void* const*, int, int, long long, long long, float*, float*, float*, float*, int,
                            float, float, float, float, int, const float*, double*, void* const*, long long, unsigned int*,
                            unsigned int*, int, const void*, void*, float, cudaStream_t);
int b200_clip_exchange(void* const*, void* const*, long long, int, int, const double*, unsigned int*, float, float*, float*,
                       cudaStream_t);
int b200_attn_short_ok(int, int, int, int);
int b200_attn_tc_ok(int, int, int);
int b200_attn_tc_fwd(const void*, const void*, const void*, const float*, void*, float*, int, int, int, int, int,
                     const long long*, const long long*, const long long*, const long long*, float, int, cudaStream_t);
int b200_attn_short_fwd(const void*, const void*, const void*, const float*, void*, float*, int, int, int, int, int,
                        const long long*, const long long*, const long long*, const long long*, float, int, cudaStream_t);
int b200_attn_short_bwd(const void*, const void*, const void*, const float*, const void*, const void*, const long long*,
                        const float*, void*, void*, void*, int, int, int, int, int, const long long*, const long long*,
                        const long long*, const long long*, float, int, cudaStream_t);
int b200_ln_train_ok(int);
int b200_ln_bwd_ctas(int);
int b200_ln_fwd(const void*, const void*, const void*, void*, float*, int, int, long long, float, int, cudaStream_t);
int b200_ln_bwd(const void*, const void*, const float*, const void*, void*, float*, void*, void*, int, int, long long, long long,
                int, cudaStream_t);
int b200_colsum_bf16(const void*, void*, int, int, long long, float*, cudaStream_t);
int b200_colsum_rows(int, int);
void b200_set_pdl(int);
void b200_set_static_weights(int);
int b200_get_static_weights();
int b200_get_pdl();
int b200_gemm_allgather_bf16(void* const*, int, const void*, void*, int, int, int, long long, long long, long long, const void*,
                             int, cudaStream_t);
int b200_gemm_reduce_scatter_bf16(const void*, const void*, float* const*, int, int, int, int, long long, long long, long long,
                                  const void*, cudaStream_t);
int b200_gemm_flagged_bf16(const void*, const void*, void*, int, int, int, long long, long long, long long, const void*, int,
                           const uint32_t*, uint32_t, int, int, cudaStream_t);
int b200_stage_reduce(const void*, int, const void*, const void*, void*, long long, int, long long, long long, long long,
                      cudaStream_t);
int b200_mc_reduce_rows(const void*, const void*, const void*, void*, long long, int, int, long long, long long, long long, int,
                        cudaStream_t);
int b200_gemm_stage_scatter_bf16(const void*, const void*, void* const*, int, int, int, int, int, long long, long long, long long,
                                 const void*, cudaStream_t);
int b200_rs_finalize(const float*, const void*, const void*, void*, long long, int, long long, long long, long long,
                     cudaStream_t);
int b200_umma_probe(int, int, int, int, long long*, cudaStream_t);
int b200_decode_mega_make_map(void*, const void*, long long, long long, long long);
int b200_decode_mega_layer_bytes();
int b200_decode_mega_stages(int, int);
int b200_decode_mega_max_clusters(int, int, int);
int b200_decode_mega_cluster_size(int, int, int);
int b200_decode_mega(int, int, int, int, int, int, int, float, float, int, int, const int*, const int*, void*, void*, void*,
                     const void*, const void*, void*, long long, const long long*, int, const float*, long long*, int,
                     cudaStream_t);
}

namespace {

inline cudaStream_t stream() { return at::cuda::getCurrentCUDAStream().stream(); }

inline void check(int code, const char* what) {
  if (code == 0) return;
  if (code > 0) throw std::runtime_error(std::string(what) + ": CUDA error " + cudaGetErrorString((cudaError_t)code));
  throw std::runtime_error(std::string(what) + ": invalid arguments / driver entry point unavailable (code " +
                           std::to_string(code) + ")");
}

inline const void* optptr(const OptTensor& t) { return t.has_value() ? t->data_ptr() : nullptr; }

#define CHECK_BF16(x) TORCH_CHECK((x).is_cuda() && (x).scalar_type() == at::kBFloat16, #x " must be a CUDA bf16 tensor")
#define CHECK_F32(x) TORCH_CHECK((x).is_cuda() && (x).scalar_type() == at::kFloat, #x " must be a CUDA fp32 tensor")

int act_code(const std::string& a) {
  if (a.empty() || a == "none") return 0;
  if (a == "gelu_new" || a == "gelu_tanh" || a == "gelu_pytorch_tanh" || a == "gelu_fast") return 1;
  if (a == "gelu") return 2;
  if (a == "relu") return 3;
  if (a == "silu" || a == "swish") return 4;
  throw std::invalid_argument("unknown activation " + a);
}

// y[M,N] = act(alpha * x[M,K] @ w[N,K]^T * col_scale + bias) + residual
Tensor gemm(const Tensor& x, const Tensor& w, const OptTensor& bias, const OptTensor& residual, const std::string& act,
            bool out_f32, const OptTensor& out_, const OptTensor& col_scale, double alpha, int64_t force_bn) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1), "gemm: shape mismatch");
  TORCH_CHECK(x.stride(1) == 1 && w.stride(1) == 1, "gemm: operands must be K-major");
  const int64_t M = x.size(0), N = w.size(0), K = x.size(1);
  TORCH_CHECK(K % 8 == 0 && x.stride(0) % 8 == 0 && w.stride(0) % 8 == 0, "gemm: K and row pitches must be multiples of 8");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(x.data_ptr()) % 16 == 0 && reinterpret_cast<uintptr_t>(w.data_ptr()) % 16 == 0,
              "gemm: operands must be 16-byte aligned");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = out_.has_value() ? *out_ : torch::empty({M, N}, x.options().dtype(out_f32 ? at::kFloat : at::kBFloat16));
  TORCH_CHECK(out.size(0) == M && out.size(1) == N && out.stride(1) == 1, "gemm: bad out");
  if (bias.has_value()) { CHECK_BF16(*bias); TORCH_CHECK(bias->numel() == N); }
  long long ldr = 0;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->size(0) == M && residual->size(1) == N && residual->stride(1) == 1); ldr = residual->stride(0); }
  if (col_scale.has_value()) { CHECK_F32(*col_scale); TORCH_CHECK(col_scale->numel() == N); }
  check(b200_gemm_bf16(x.data_ptr(), w.data_ptr(), out.data_ptr(), (int)M, (int)N, (int)K, x.stride(0), w.stride(0),
                       out.stride(0), optptr(bias), optptr(residual), ldr, (const float*)optptr(col_scale), (float)alpha,
                       act_code(act), out.scalar_type() == at::kFloat, (int)force_bn, stream()),
        "gemm");
  return out;
}

// y = act(LN(x) @ w^T + b) + residual with the normalisation FOLDED into the GEMM: `x` is the raw residual stream, `w` already
// carries gamma (w * gamma[None, :]), `bias` carries w @ beta + b, `ln_c1[n] = sum_k w[n, k]`, and `ln_stats` = [M, 2] row
// (sum, sum of squares) of x that the producer of x accumulated through its own `stats_out`.
Tensor gemm_ln(const Tensor& x, const Tensor& w, const OptTensor& bias, const OptTensor& residual, const std::string& act,
               const OptTensor& ln_stats, const OptTensor& ln_c1, double ln_eps, bool ln_rms, const OptTensor& stats_out) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && x.stride(1) == 1 && w.stride(1) == 1);
  const int64_t M = x.size(0), N = w.size(0), K = x.size(1);
  TORCH_CHECK(K % 8 == 0 && N % 16 == 0 && x.stride(0) % 8 == 0 && w.stride(0) % 8 == 0, "gemm_ln: K % 8, N % 16");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = torch::empty({M, N}, x.options());
  long long ldr = 0;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->size(0) == M && residual->size(1) == N && residual->stride(1) == 1); ldr = residual->stride(0); }
  if (bias.has_value()) { CHECK_BF16(*bias); TORCH_CHECK(bias->numel() == N); }
  const float* st = nullptr; const float* c1 = nullptr; float* so = nullptr;
  if (ln_stats.has_value()) {
    CHECK_F32(*ln_stats); TORCH_CHECK(ln_stats->numel() == 2 * M && ln_stats->is_contiguous() && ln_c1.has_value());
    CHECK_F32(*ln_c1); TORCH_CHECK(ln_c1->numel() == N && ln_c1->is_contiguous());
    st = ln_stats->data_ptr<float>(); c1 = ln_c1->data_ptr<float>();
  }
  if (stats_out.has_value()) { CHECK_F32(*stats_out); TORCH_CHECK(stats_out->numel() == 2 * M && stats_out->is_contiguous()); so = stats_out->data_ptr<float>(); }
  check(b200_gemm_bf16_ln(x.data_ptr(), w.data_ptr(), out.data_ptr(), (int)M, (int)N, (int)K, x.stride(0), w.stride(0), out.stride(0),
                          optptr(bias), optptr(residual), ldr, act_code(act), st, c1, (float)ln_eps, ln_rms ? 1 : 0, so, stream()),
        "gemm_ln");
  return out;
}

// e4m3 x e4m3 GEMM with per-row (A) and per-column (B) dequantisation scales; bf16 output.
Tensor gemm_fp8(const Tensor& a8, const Tensor& b8, const Tensor& row_scale, const Tensor& col_scale, const OptTensor& bias,
                const OptTensor& residual, const std::string& act) {
  TORCH_CHECK(a8.is_cuda() && b8.is_cuda() && a8.element_size() == 1 && b8.element_size() == 1, "gemm_fp8: 1-byte operands");
  TORCH_CHECK(a8.dim() == 2 && b8.dim() == 2 && a8.size(1) == b8.size(1) && a8.stride(1) == 1 && b8.stride(1) == 1);
  const int64_t M = a8.size(0), N = b8.size(0), K = a8.size(1);
  TORCH_CHECK(K % 16 == 0 && a8.stride(0) % 16 == 0 && b8.stride(0) % 16 == 0, "gemm_fp8: K and pitches must be multiples of 16");
  CHECK_F32(row_scale); CHECK_F32(col_scale);
  TORCH_CHECK(row_scale.numel() == M && col_scale.numel() == N && row_scale.is_contiguous() && col_scale.is_contiguous());
  c10::cuda::CUDAGuard guard(a8.device());
  Tensor out = torch::empty({M, N}, a8.options().dtype(at::kBFloat16));
  long long ldr = 0;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->size(0) == M && residual->size(1) == N && residual->stride(1) == 1); ldr = residual->stride(0); }
  if (bias.has_value()) { CHECK_BF16(*bias); TORCH_CHECK(bias->numel() == N); }
  check(b200_gemm_fp8(a8.data_ptr(), b8.data_ptr(), out.data_ptr(), (int)M, (int)N, (int)K, a8.stride(0), b8.stride(0), out.stride(0),
                      row_scale.data_ptr<float>(), col_scale.data_ptr<float>(), optptr(bias), optptr(residual), ldr, act_code(act),
                      stream()),
        "gemm_fp8");
  return out;
}

// (y8 [rows, H] e4m3 bytes, scale [rows]) = quantise(norm(x))
std::vector<Tensor> norm_quant(const Tensor& x, const Tensor& w, const OptTensor& b, double eps, bool rms) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.size(1) % 16 == 0 && x.stride(0) % 8 == 0);
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t rows = x.size(0), H = x.size(1);
  Tensor y8 = torch::empty({rows, H}, x.options().dtype(at::kByte));
  Tensor scale = torch::empty({rows}, x.options().dtype(at::kFloat));
  check(b200_norm_quant_fp8(x.data_ptr(), w.data_ptr(), optptr(b), y8.data_ptr(), scale.data_ptr<float>(), (int)rows, (int)H,
                            x.stride(0), y8.stride(0), (float)eps, rms ? 1 : 0, stream()),
        "norm_quant");
  return {y8, scale};
}

// (y8 [rows, H] e4m3 bytes, scale [rows]) = per-row quantisation of a bf16 matrix
std::vector<Tensor> quant_rows(const Tensor& x) {
  CHECK_BF16(x);
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.size(1) % 16 == 0 && x.stride(0) % 8 == 0 &&
              reinterpret_cast<uintptr_t>(x.data_ptr()) % 16 == 0, "quant_rows: [rows, H] bf16 with H % 16 == 0");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t rows = x.size(0), H = x.size(1);
  Tensor y8 = torch::empty({rows, H}, x.options().dtype(at::kByte));
  Tensor scale = torch::empty({rows}, x.options().dtype(at::kFloat));
  check(b200_quant_rows_fp8(x.data_ptr(), y8.data_ptr(), scale.data_ptr<float>(), (int)rows, (int)H, x.stride(0), y8.stride(0),
                            stream()),
        "quant_rows");
  return {y8, scale};
}

// out[M, N] = sum_k A(m, k) B(n, k) with either operand optionally MN-major (stored transposed: A as [K, M], B as [K, N]).
// These are the layouts of the backward GEMMs (dX = dY·W, dW = dYᵀ·X) — no transposed copies are made.
Tensor gemm_ex(const Tensor& a, const Tensor& b, bool a_mn, bool b_mn, bool out_f32, int64_t split_k, const OptTensor& out_,
               bool accumulate) {
  CHECK_BF16(a); CHECK_BF16(b);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1, "gemm_ex: operands must be row-major 2-D");
  const int64_t M = a_mn ? a.size(1) : a.size(0), Ka = a_mn ? a.size(0) : a.size(1);
  const int64_t N = b_mn ? b.size(1) : b.size(0), Kb = b_mn ? b.size(0) : b.size(1);
  TORCH_CHECK(Ka == Kb, "gemm_ex: contraction sizes differ");
  TORCH_CHECK(a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0, "gemm_ex: row pitches must be multiples of 8 elements");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(a.data_ptr()) % 16 == 0 && reinterpret_cast<uintptr_t>(b.data_ptr()) % 16 == 0);
  c10::cuda::CUDAGuard guard(a.device());
  const int splits = split_k < 0 ? b200_gemm_splitk_plan((int)M, (int)N, (int)Ka) : (int)std::max<int64_t>(split_k, 1);
  Tensor ws;
  if (splits > 1) ws = torch::zeros({M, N}, a.options().dtype(at::kFloat));  // partial products are red.add'ed into it
  Tensor out;
  if (out_.has_value()) {  // caller-provided destination (accumulate: out += a.b — weight gradients added straight into .grad)
    out = *out_;
    CHECK_BF16(out);
    TORCH_CHECK(!out_f32 && out.dim() == 2 && out.size(0) == M && out.size(1) == N && out.stride(1) == 1 && out.stride(0) % 8 == 0 &&
                reinterpret_cast<uintptr_t>(out.data_ptr()) % 16 == 0, "gemm_ex: bad out");
  } else {
    TORCH_CHECK(!accumulate, "gemm_ex: accumulate needs out");
    out = (splits > 1 && out_f32) ? ws : torch::empty({M, N}, a.options().dtype(out_f32 ? at::kFloat : at::kBFloat16));
  }
  check(b200_gemm_bf16_ex(a.data_ptr(), b.data_ptr(), out.data_ptr(), (int)M, (int)N, (int)Ka, a.stride(0), b.stride(0),
                          out.stride(0), a_mn ? 1 : 0, b_mn ? 1 : 0, out_f32 ? 1 : 0, splits,
                          splits > 1 ? ws.data_ptr<float>() : nullptr, accumulate ? 1 : 0, stream()),
        "gemm_ex");
  return out;
}

// ---- short-sequence attention (attention.cu): q/k/v are [B, H, T, d] views with a contiguous head dimension
namespace {
struct AttnArgs {
  long long qs[3], ks[3], vs[3], bs[3];
  const float* bias = nullptr;
};
AttnArgs attn_args(const Tensor& q, const Tensor& k, const Tensor& v, const OptTensor& bias) {
  CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v);
  TORCH_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4, "attn_short: q/k/v must be [B, H, T, d]");
  TORCH_CHECK(q.stride(3) == 1 && k.stride(3) == 1 && v.stride(3) == 1, "attn_short: head dimension must be contiguous");
  TORCH_CHECK(k.sizes() == v.sizes() && q.size(0) == k.size(0) && q.size(1) == k.size(1) && q.size(3) == k.size(3));
  for (const Tensor* t : {&q, &k, &v}) {
    TORCH_CHECK(reinterpret_cast<uintptr_t>(t->data_ptr()) % 16 == 0 && t->stride(0) % 8 == 0 && t->stride(1) % 8 == 0 &&
                t->stride(2) % 8 == 0, "attn_short: 16-byte aligned rows required");
  }
  AttnArgs a;
  for (int i = 0; i < 3; ++i) { a.qs[i] = q.stride(i); a.ks[i] = k.stride(i); a.vs[i] = v.stride(i); a.bs[i] = 0; }
  if (bias.has_value()) {
    const Tensor& bt = *bias;
    CHECK_F32(bt);
    TORCH_CHECK(bt.dim() == 4 && bt.size(3) == k.size(2) && (bt.size(3) == 1 || bt.stride(3) == 1), "attn_short: bad bias");
    TORCH_CHECK((bt.size(0) == 1 || bt.size(0) == q.size(0)) && (bt.size(1) == 1 || bt.size(1) == q.size(1)) &&
                (bt.size(2) == 1 || bt.size(2) == q.size(2)), "attn_short: bias must broadcast to [B, H, Tq, Tk]");
    a.bs[0] = bt.size(0) == 1 ? 0 : bt.stride(0);
    a.bs[1] = bt.size(1) == 1 ? 0 : bt.stride(1);
    a.bs[2] = bt.size(2) == 1 ? 0 : bt.stride(2);
    a.bias = bt.data_ptr<float>();
  }
  return a;
}
}  // namespace

// returns (o [B, Tq, H, d] bf16 contiguous, stats [B, H, Tq, 2] fp32 = row max and 1 / row sum)
std::vector<Tensor> attn_short_fwd(const Tensor& q, const Tensor& k, const Tensor& v, const OptTensor& bias, bool causal,
                                   double scale) {
  AttnArgs a = attn_args(q, k, v, bias);
  const int B = (int)q.size(0), H = (int)q.size(1), Tq = (int)q.size(2), Tk = (int)k.size(2), d = (int)q.size(3);
  c10::cuda::CUDAGuard guard(q.device());
  Tensor o = torch::empty({B, Tq, H, d}, q.options());
  Tensor stats = torch::empty({B, H, Tq, 2}, q.options().dtype(at::kFloat));
  check(b200_attn_short_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), a.bias, o.data_ptr(), stats.data_ptr<float>(), B, H, Tq,
                            Tk, d, a.qs, a.ks, a.vs, a.bs, (float)scale, causal ? 1 : 0, stream()),
        "attn_short_fwd");
  return {o, stats};
}

// same contract on the tcgen05 kernel (Tq, Tk <= 128, d in {64, 128})
std::vector<Tensor> attn_tc_fwd(const Tensor& q, const Tensor& k, const Tensor& v, const OptTensor& bias, bool causal,
                                double scale) {
  AttnArgs a = attn_args(q, k, v, bias);
  const int B = (int)q.size(0), H = (int)q.size(1), Tq = (int)q.size(2), Tk = (int)k.size(2), d = (int)q.size(3);
  TORCH_CHECK(b200_attn_tc_ok(Tq, Tk, d), "attn_tc_fwd: unsupported shape");
  c10::cuda::CUDAGuard guard(q.device());
  Tensor o = torch::empty({B, Tq, H, d}, q.options());
  Tensor stats = torch::empty({B, H, Tq, 2}, q.options().dtype(at::kFloat));
  check(b200_attn_tc_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), a.bias, o.data_ptr(), stats.data_ptr<float>(), B, H, Tq, Tk, d,
                         a.qs, a.ks, a.vs, a.bs, (float)scale, causal ? 1 : 0, stream()),
        "attn_tc_fwd");
  return {o, stats};
}

// d_o: [B, H, Tq, d] view (any strides, contiguous head dim); returns dq, dk, dv as [B, T, H, d] contiguous
std::vector<Tensor> attn_short_bwd(const Tensor& q, const Tensor& k, const Tensor& v, const OptTensor& bias, const Tensor& o,
                                   const Tensor& d_o, const Tensor& stats, bool causal, double scale) {
  AttnArgs a = attn_args(q, k, v, bias);
  const int B = (int)q.size(0), H = (int)q.size(1), Tq = (int)q.size(2), Tk = (int)k.size(2), d = (int)q.size(3);
  CHECK_BF16(o); CHECK_BF16(d_o); CHECK_F32(stats);
  TORCH_CHECK(o.is_contiguous() && o.size(0) == B && o.size(1) == Tq && o.size(2) == H && o.size(3) == d);
  TORCH_CHECK(d_o.dim() == 4 && d_o.size(0) == B && d_o.size(1) == H && d_o.size(2) == Tq && d_o.size(3) == d &&
              d_o.stride(3) == 1 && d_o.stride(0) % 8 == 0 && d_o.stride(1) % 8 == 0 && d_o.stride(2) % 8 == 0 &&
              reinterpret_cast<uintptr_t>(d_o.data_ptr()) % 16 == 0, "attn_short_bwd: bad d_o");
  TORCH_CHECK(stats.is_contiguous() && stats.numel() == (int64_t)B * H * Tq * 2);
  c10::cuda::CUDAGuard guard(q.device());
  Tensor dq = torch::empty({B, Tq, H, d}, q.options());
  Tensor dk = torch::empty({B, Tk, H, d}, q.options());
  Tensor dv = torch::empty({B, Tk, H, d}, q.options());
  const long long dos[3] = {d_o.stride(0), d_o.stride(1), d_o.stride(2)};
  check(b200_attn_short_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), a.bias, o.data_ptr(), d_o.data_ptr(), dos,
                            stats.data_ptr<float>(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, H, Tq, Tk, d, a.qs, a.ks,
                            a.vs, a.bs, (float)scale, causal ? 1 : 0, stream()),
        "attn_short_bwd");
  return {dq, dk, dv};
}

// ---- training-side norms and the bias-gradient column sum (norm_train.cu)
std::vector<Tensor> ln_fwd(const Tensor& x, const Tensor& w, const OptTensor& b, double eps, bool rms) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.stride(0) % 8 == 0 && w.is_contiguous() && w.numel() == x.size(1));
  TORCH_CHECK(b200_ln_train_ok((int)x.size(1)), "ln_fwd: unsupported hidden size");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(x.data_ptr()) % 16 == 0 && reinterpret_cast<uintptr_t>(w.data_ptr()) % 16 == 0);
  if (b.has_value()) { CHECK_BF16(*b); TORCH_CHECK(b->is_contiguous() && reinterpret_cast<uintptr_t>(b->data_ptr()) % 16 == 0); }
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = torch::empty({x.size(0), x.size(1)}, x.options());
  Tensor stats = torch::empty({x.size(0), 2}, x.options().dtype(at::kFloat));
  check(b200_ln_fwd(x.data_ptr(), w.data_ptr(), rms ? nullptr : optptr(b), y.data_ptr(), stats.data_ptr<float>(), (int)x.size(0),
                    (int)x.size(1), x.stride(0), (float)eps, rms ? 1 : 0, stream()),
        "ln_fwd");
  return {y, stats};
}

// returns (dx [M, H], d-gamma [H], d-beta [H] or empty)
std::vector<Tensor> ln_bwd(const Tensor& x, const Tensor& w, const Tensor& stats, const Tensor& dy, bool rms, bool has_bias) {
  CHECK_BF16(x); CHECK_BF16(w); CHECK_BF16(dy); CHECK_F32(stats);
  TORCH_CHECK(x.dim() == 2 && dy.dim() == 2 && x.sizes() == dy.sizes() && x.stride(1) == 1 && dy.stride(1) == 1 &&
              x.stride(0) % 8 == 0 && dy.stride(0) % 8 == 0 && stats.is_contiguous() && stats.numel() == 2 * x.size(0));
  TORCH_CHECK(reinterpret_cast<uintptr_t>(x.data_ptr()) % 16 == 0 && reinterpret_cast<uintptr_t>(dy.data_ptr()) % 16 == 0);
  const int rows = (int)x.size(0), H = (int)x.size(1);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor dx = torch::empty({rows, H}, x.options());
  Tensor partial = torch::empty({(int64_t)b200_ln_bwd_ctas(rows), 2, H}, x.options().dtype(at::kFloat));
  Tensor dgamma = torch::empty({H}, x.options());
  Tensor dbeta = (has_bias && !rms) ? torch::empty({H}, x.options()) : Tensor();
  check(b200_ln_bwd(x.data_ptr(), w.data_ptr(), stats.data_ptr<float>(), dy.data_ptr(), dx.data_ptr(), partial.data_ptr<float>(),
                    dgamma.data_ptr(), dbeta.defined() ? dbeta.data_ptr() : nullptr, rows, H, x.stride(0), dy.stride(0),
                    rms ? 1 : 0, stream()),
        "ln_bwd");
  return {dx, dgamma, dbeta};
}

Tensor colsum(const Tensor& x) {
  CHECK_BF16(x);
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.size(1) % 2 == 0 && x.stride(0) % 2 == 0 &&
              reinterpret_cast<uintptr_t>(x.data_ptr()) % 4 == 0, "colsum: bf16 [M, N] with even N and pitch");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = torch::empty({x.size(1)}, x.options());
  const int64_t N = x.size(1);
  // one zeroed workspace per (device, width), reused by every call: the kernel restores it to zero before it finishes, and calls
  // on one stream are ordered (the bias gradients of a step are reduced one after the other)
  static std::unordered_map<int64_t, Tensor> ws_cache;
  Tensor ws;
  if (b200_colsum_rows((int)x.size(0), (int)N) > 1) {
    const int64_t key = (int64_t)x.device().index() * (1ll << 32) + N;
    auto it = ws_cache.find(key);
    if (it == ws_cache.end()) it = ws_cache.emplace(key, torch::zeros({N + (N + 63) / 64 + 1}, x.options().dtype(at::kFloat))).first;
    ws = it->second;
  }
  check(b200_colsum_bf16(x.data_ptr(), out.data_ptr(), (int)x.size(0), (int)N, x.stride(0),
                         ws.defined() ? ws.data_ptr<float>() : nullptr, stream()),
        "colsum");
  return out;
}

// LM-head backward: d-logits[M, N] (bf16, row pitch padded to a multiple of 64 so every consumer is vectorised) from
// h[M,K], w[N,K], the saved logsumexp and the incoming gradient of log p(label).
Tensor lmhead_dlogits(const Tensor& h, const Tensor& w, const OptTensor& bias, const Tensor& labels, const Tensor& lse,
                      const Tensor& grad) {
  CHECK_BF16(h); CHECK_BF16(w); CHECK_F32(lse); CHECK_F32(grad);
  TORCH_CHECK(h.dim() == 2 && w.dim() == 2 && h.size(1) == w.size(1) && h.stride(1) == 1 && w.stride(1) == 1);
  const int64_t M = h.size(0), N = w.size(0), K = h.size(1);
  TORCH_CHECK(K % 8 == 0 && h.stride(0) % 8 == 0 && w.stride(0) % 8 == 0);
  TORCH_CHECK(labels.scalar_type() == at::kLong && labels.numel() == M && labels.is_contiguous());
  TORCH_CHECK(lse.numel() == M && grad.numel() == M && lse.is_contiguous() && grad.is_contiguous());
  c10::cuda::CUDAGuard guard(h.device());
  const int64_t ld = (N + 63) / 64 * 64;
  Tensor buf = torch::empty({M, ld}, h.options());
  if (ld != N) buf.slice(1, N, ld).zero_();
  check(b200_lmhead_dlogits_bf16(h.data_ptr(), w.data_ptr(), buf.data_ptr(), (int)M, (int)N, (int)K, h.stride(0), w.stride(0),
                                 ld, optptr(bias), (const long long*)labels.data_ptr<int64_t>(), lse.data_ptr<float>(),
                                 grad.data_ptr<float>(), stream()),
        "lmhead_dlogits");
  return buf.slice(1, 0, N);
}

// Fused LM head: returns (lse[M], logprob[M], token[M], token_logprob[M]); unused outputs are empty tensors.
std::vector<Tensor> lmhead(const Tensor& h, const Tensor& w, const OptTensor& bias, const OptTensor& labels, bool sample,
                           double temperature, int64_t seed, const OptTensor& step, int64_t suppress_col,
                           int64_t suppress_until, const OptTensor& workspace_, const OptTensor& seed_tensor) {
  CHECK_BF16(h); CHECK_BF16(w);
  TORCH_CHECK(h.dim() == 2 && w.dim() == 2 && h.size(1) == w.size(1) && h.stride(1) == 1 && w.stride(1) == 1);
  const int64_t M = h.size(0), N = w.size(0), K = h.size(1);
  TORCH_CHECK(K % 8 == 0 && h.stride(0) % 8 == 0 && w.stride(0) % 8 == 0);
  c10::cuda::CUDAGuard guard(h.device());
  const int64_t nt = b200_lmhead_tiles((int)N);
  auto f32 = h.options().dtype(at::kFloat);
  Tensor ws = workspace_.has_value() ? *workspace_ : torch::empty({5 * M * nt + M}, f32);
  TORCH_CHECK(ws.numel() >= 5 * M * nt + M);
  Tensor lse = torch::empty({M}, f32), lp = torch::empty({M}, f32);
  Tensor tok = sample ? torch::empty({M}, h.options().dtype(at::kLong)) : Tensor();
  Tensor tlp = sample ? torch::empty({M}, f32) : Tensor();
  const long long* lab = nullptr;
  if (labels.has_value()) { TORCH_CHECK(labels->scalar_type() == at::kLong && labels->numel() == M && labels->is_contiguous()); lab = labels->data_ptr<int64_t>() ? (const long long*)labels->data_ptr<int64_t>() : nullptr; }
  const long long* seedp = nullptr;
  if (seed_tensor.has_value()) { TORCH_CHECK(seed_tensor->scalar_type() == at::kLong && seed_tensor->is_cuda()); seedp = (const long long*)seed_tensor->data_ptr<int64_t>(); }
  const int* sp = nullptr;
  if (step.has_value()) { TORCH_CHECK(step->scalar_type() == at::kInt); sp = step->data_ptr<int>(); }
  check(b200_lmhead_bf16(h.data_ptr(), w.data_ptr(), (int)M, (int)N, (int)K, h.stride(0), w.stride(0), optptr(bias), lab,
                         ws.data_ptr<float>(), lse.data_ptr<float>(), lp.data_ptr<float>(), sample ? 1 : 0,
                         (float)temperature, (unsigned long long)seed, seedp, sp, (int)suppress_col, (int)suppress_until,
                         sample ? (long long*)tok.data_ptr<int64_t>() : nullptr, sample ? tlp.data_ptr<float>() : nullptr,
                         stream()),
        "lmhead");
  return {lse, lp, tok, tlp};
}

// temperature / top-k / top-p sampling from fp32 logits [B, >= V] (row pitch = stride(0)) -> (token, raw log-prob of it)
std::vector<Tensor> sample_filtered(const Tensor& logits, int64_t V, int64_t top_k, double top_p, double temperature,
                                    int64_t seed, const OptTensor& step, int64_t suppress_col, int64_t suppress_until,
                                    const OptTensor& seed_tensor) {
  CHECK_F32(logits);
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1 && logits.size(1) >= V);
  c10::cuda::CUDAGuard guard(logits.device());
  const int64_t B = logits.size(0);
  Tensor tok = torch::empty({B}, logits.options().dtype(at::kLong)), lp = torch::empty({B}, logits.options());
  const long long* seedp = nullptr;
  if (seed_tensor.has_value()) { TORCH_CHECK(seed_tensor->scalar_type() == at::kLong && seed_tensor->is_cuda()); seedp = (const long long*)seed_tensor->data_ptr<int64_t>(); }
  const int* sp = nullptr;
  if (step.has_value()) { TORCH_CHECK(step->scalar_type() == at::kInt); sp = step->data_ptr<int>(); }
  check(b200_sample_filtered(logits.data_ptr<float>(), logits.stride(0), (int)B, (int)V, (int)top_k, (float)top_p,
                             (float)temperature, (unsigned long long)seed, seedp, sp, (int)suppress_col, (int)suppress_until,
                             (long long*)tok.data_ptr<int64_t>(), lp.data_ptr<float>(), stream()),
        "sample_filtered");
  return {tok, lp};
}

// ILQL advantage-shifted sampling step: logits / q1 / q2 fp32 [B, >= V] sharing one row pitch, vs fp32 [B] -> tokens [B]
Tensor ilql_sample(const Tensor& logits, const Tensor& q1, const OptTensor& q2, const Tensor& vs, int64_t V, double beta,
                   int64_t top_k, double temperature, int64_t seed, const OptTensor& step, const OptTensor& seed_tensor,
                   const OptTensor& logit_mask, const OptTensor& last_tokens) {
  CHECK_F32(logits); CHECK_F32(q1); CHECK_F32(vs);
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1 && logits.size(1) >= V && q1.stride(0) == logits.stride(0) &&
              q1.stride(1) == 1 && vs.is_contiguous() && vs.numel() == logits.size(0));
  const float* q2p = nullptr;
  if (q2.has_value()) { CHECK_F32(*q2); TORCH_CHECK(q2->stride(0) == logits.stride(0) && q2->stride(1) == 1); q2p = q2->data_ptr<float>(); }
  c10::cuda::CUDAGuard guard(logits.device());
  const int64_t B = logits.size(0);
  Tensor tok = torch::empty({B}, logits.options().dtype(at::kLong));
  const long long* seedp = nullptr;
  if (seed_tensor.has_value()) { TORCH_CHECK(seed_tensor->scalar_type() == at::kLong && seed_tensor->is_cuda()); seedp = (const long long*)seed_tensor->data_ptr<int64_t>(); }
  const int* sp = nullptr;
  if (step.has_value()) { TORCH_CHECK(step->scalar_type() == at::kInt); sp = step->data_ptr<int>(); }
  const unsigned char* mp = nullptr;
  int mr = 0, mc = 0;
  const long long* lastp = nullptr;
  if (logit_mask.has_value()) {
    TORCH_CHECK(logit_mask->is_cuda() && logit_mask->dim() == 2 && logit_mask->is_contiguous() &&
                (logit_mask->scalar_type() == at::kBool || logit_mask->scalar_type() == at::kByte));
    TORCH_CHECK(last_tokens.has_value() && last_tokens->scalar_type() == at::kLong && last_tokens->numel() == B);
    mp = reinterpret_cast<const unsigned char*>(logit_mask->data_ptr());
    mr = (int)logit_mask->size(0); mc = (int)logit_mask->size(1);
    lastp = (const long long*)last_tokens->data_ptr<int64_t>();
  }
  check(b200_ilql_sample(logits.data_ptr<float>(), q1.data_ptr<float>(), q2p, vs.data_ptr<float>(), logits.stride(0), (int)B, (int)V,
                         (float)beta, (int)top_k, (float)temperature, (unsigned long long)seed, seedp, sp, mp, mr, mc, lastp,
                         (long long*)tok.data_ptr<int64_t>(), stream()),
        "ilql_sample");
  return tok;
}

Tensor norm(const Tensor& x, const Tensor& w, const OptTensor& b, double eps, bool rms, const OptTensor& out_) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.stride(0) % 8 == 0);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = out_.has_value() ? *out_ : torch::empty_like(x, at::MemoryFormat::Contiguous);
  check(b200_norm_bf16(x.data_ptr(), w.data_ptr(), optptr(b), y.data_ptr(), (int)x.size(0), (int)x.size(1), x.stride(0),
                       y.stride(0), (float)eps, rms ? 1 : 0, stream()),
        "norm");
  return y;
}

Tensor embed(const Tensor& tokens, const Tensor& positions, const Tensor& wte, const OptTensor& wpe, int64_t pos_offset,
             const OptTensor& out_, const OptTensor& stats_out) {
  CHECK_BF16(wte);
  TORCH_CHECK(tokens.scalar_type() == at::kLong && positions.scalar_type() == at::kInt);
  c10::cuda::CUDAGuard guard(wte.device());
  const int64_t B = tokens.numel(), H = wte.size(1);
  TORCH_CHECK(H % 8 == 0);
  Tensor x = out_.has_value() ? *out_ : torch::empty({B, H}, wte.options());
  check(b200_embed_bf16((const long long*)tokens.data_ptr<int64_t>(), positions.data_ptr<int>(), wte.data_ptr(), optptr(wpe),
                        (int)pos_offset, x.data_ptr(), (int)B, (int)H,
                        stats_out.has_value() ? stats_out->data_ptr<float>() : nullptr, stream()),
        "embed");
  return x;
}

Tensor decode_attention(const Tensor& qkv, Tensor& kcache, Tensor& vcache, const Tensor& block_table, const Tensor& seq_lens,
                        const Tensor& positions, int64_t nq, int64_t nkv, int64_t d, double scale, int64_t rot_dim,
                        double rot_base, bool rot_interleaved, const OptTensor& alibi, int64_t window, const OptTensor& out_) {
  CHECK_BF16(qkv); CHECK_BF16(kcache); CHECK_BF16(vcache);
  TORCH_CHECK(qkv.is_contiguous() && kcache.is_contiguous() && vcache.is_contiguous() && block_table.is_contiguous());
  TORCH_CHECK(block_table.scalar_type() == at::kInt && seq_lens.scalar_type() == at::kInt && positions.scalar_type() == at::kInt);
  c10::cuda::CUDAGuard guard(qkv.device());
  const int64_t B = qkv.size(0);
  Tensor out = out_.has_value() ? *out_ : torch::empty({B, nq * d}, qkv.options());
  check(b200_decode_attention_bf16(qkv.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), block_table.data_ptr<int>(),
                                   seq_lens.data_ptr<int>(), positions.data_ptr<int>(), out.data_ptr(), (int)B, (int)nq,
                                   (int)nkv, (int)d, (int)kcache.size(1), (int)block_table.size(1), (float)scale,
                                   (int)rot_dim, (float)rot_base, rot_interleaved ? 1 : 0, (const float*)optptr(alibi),
                                   (int)window, stream()),
        "decode_attention");
  return out;
}

Tensor rowdot(const Tensor& x, const Tensor& w, const OptTensor& bias, const OptTensor& out_) {
  CHECK_BF16(x); CHECK_BF16(w);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = out_.has_value() ? *out_ : torch::empty({x.size(0)}, x.options().dtype(at::kFloat));
  check(b200_rowdot_bf16(x.data_ptr(), w.data_ptr(), optptr(bias), out.data_ptr<float>(), (int)x.size(0), (int)x.size(1),
                         x.stride(0), stream()),
        "rowdot");
  return out;
}

void decode_step(const Tensor& sampled, const Tensor& lp, const OptTensor& ref_lp, const OptTensor& value, Tensor& step,
                 int64_t max_new, int64_t eos_id, int64_t pad_id, Tensor& tokens_out, Tensor& logprobs_out,
                 const OptTensor& ref_logprobs_out, const OptTensor& values_out, Tensor& finished, Tensor& resp_lens,
                 Tensor& seq_lens, Tensor& positions, Tensor& next_tokens, Tensor& n_running) {
  c10::cuda::CUDAGuard guard(sampled.device());
  const int B = (int)sampled.numel();
  check(b200_decode_step((const long long*)sampled.data_ptr<int64_t>(), lp.data_ptr<float>(), (const float*)optptr(ref_lp),
                         (const float*)optptr(value), step.data_ptr<int>(), (int)max_new, B, eos_id, pad_id,
                         (long long*)tokens_out.data_ptr<int64_t>(), logprobs_out.data_ptr<float>(),
                         (float*)optptr(ref_logprobs_out), (float*)optptr(values_out), finished.data_ptr<int>(),
                         resp_lens.data_ptr<int>(), seq_lens.data_ptr<int>(), positions.data_ptr<int>(),
                         (long long*)next_tokens.data_ptr<int64_t>(), n_running.data_ptr<int>(), stream()),
        "decode_step");
}

// k, v: [B, T, nkv*d] (any batch/time strides, unit inner stride)
void paged_kv_write(const Tensor& k, const Tensor& v, Tensor& kcache, Tensor& vcache, const Tensor& block_table,
                    const Tensor& first, const Tensor& lens, int64_t nkv, int64_t d) {
  CHECK_BF16(k); CHECK_BF16(v);
  TORCH_CHECK(k.dim() == 3 && k.stride(2) == 1 && v.stride(2) == 1 && k.stride(0) == v.stride(0) && k.stride(1) == v.stride(1));
  TORCH_CHECK((nkv * d) % 8 == 0 && k.stride(0) % 8 == 0 && k.stride(1) % 8 == 0);
  c10::cuda::CUDAGuard guard(k.device());
  check(b200_paged_kv_write(k.data_ptr(), v.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), block_table.data_ptr<int>(),
                            first.data_ptr<int>(), lens.data_ptr<int>(), (int)k.size(0), (int)k.size(1), (int)nkv, (int)d,
                            (int)kcache.size(1), (int)block_table.size(1), k.stride(0), k.stride(1), stream()),
        "paged_kv_write");
}

int dtype_code(const Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return 0;
    case at::kBFloat16: return 1;
    case at::kHalf: return 2;
    default: throw std::invalid_argument("logits must be fp32 / bf16 / fp16");
  }
}

std::vector<Tensor> logprob_from_logits(const Tensor& logits, const Tensor& labels) {
  TORCH_CHECK(logits.is_cuda() && logits.stride(-1) == 1 && labels.scalar_type() == at::kLong);
  c10::cuda::CUDAGuard guard(logits.device());
  const int64_t V = logits.size(-1);
  Tensor l2 = logits.reshape({-1, V});
  if (l2.stride(1) != 1) l2 = l2.contiguous();
  Tensor lab = labels.reshape({-1}).contiguous();
  TORCH_CHECK(lab.numel() == l2.size(0));
  auto f32 = logits.options().dtype(at::kFloat);
  Tensor out = torch::empty({l2.size(0)}, f32), lse = torch::empty({l2.size(0)}, f32);
  check(b200_logprob_from_logits(l2.data_ptr(), (const long long*)lab.data_ptr<int64_t>(), out.data_ptr<float>(),
                                 lse.data_ptr<float>(), l2.size(0), (int)V, l2.stride(0), dtype_code(l2), stream()),
        "logprob_from_logits");
  return {out.reshape(labels.sizes()), lse.reshape(labels.sizes())};
}

void logprob_backward_inplace(Tensor& logits, const Tensor& labels, const Tensor& lse, const Tensor& grad) {
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1);
  CHECK_F32(lse); CHECK_F32(grad);
  c10::cuda::CUDAGuard guard(logits.device());
  check(b200_logprob_backward_inplace(logits.data_ptr(), (const long long*)labels.data_ptr<int64_t>(), lse.data_ptr<float>(),
                                      grad.data_ptr<float>(), logits.size(0), (int)logits.size(1), logits.stride(0),
                                      dtype_code(logits), stream()),
        "logprob_backward_inplace");
}

// returns (advantages, returns, stats[3] double = (count, sum, sumsq)); whitening applied unless `stats_only`
std::vector<Tensor> gae(const Tensor& values, const Tensor& rewards, int64_t width, double gamma, double lam, bool do_whiten,
                        bool unbiased, const OptTensor& width_tensor) {
  const int* wp = nullptr;
  if (width_tensor.has_value()) { TORCH_CHECK(width_tensor->scalar_type() == at::kInt && width_tensor->is_cuda()); wp = width_tensor->data_ptr<int>(); }
  CHECK_F32(values); CHECK_F32(rewards);
  TORCH_CHECK(values.dim() == 2 && values.is_contiguous() && rewards.is_contiguous() && values.sizes() == rewards.sizes());
  c10::cuda::CUDAGuard guard(values.device());
  const int B = (int)values.size(0), R = (int)values.size(1);
  Tensor adv = torch::empty_like(values), ret = torch::empty_like(values);
  Tensor stats = torch::zeros({3}, values.options().dtype(at::kDouble));
  check(b200_gae(values.data_ptr<float>(), rewards.data_ptr<float>(), adv.data_ptr<float>(), ret.data_ptr<float>(), B, R,
                 (int)width, wp, R, (float)gamma, (float)lam, stats.data_ptr<double>(), stream()),
        "gae");
  if (do_whiten)
    check(b200_whiten(adv.data_ptr<float>(), B, (int)width, wp, R, stats.data_ptr<double>(), unbiased ? 1 : 0, stream()), "whiten");
  return {adv, ret, stats};
}

void whiten_(Tensor& adv, int64_t width, const Tensor& stats, bool unbiased, const OptTensor& width_tensor) {
  CHECK_F32(adv);
  TORCH_CHECK(stats.scalar_type() == at::kDouble);
  c10::cuda::CUDAGuard guard(adv.device());
  const int* wp = width_tensor.has_value() ? width_tensor->data_ptr<int>() : nullptr;
  check(b200_whiten(adv.data_ptr<float>(), (int)adv.size(0), (int)width, wp, adv.size(1), stats.data_ptr<double>(),
                    unbiased ? 1 : 0, stream()),
        "whiten");
}

// returns (out[O_COUNT], dlogprobs, dvalues)
std::vector<Tensor> ppo_loss(const Tensor& logprobs, const Tensor& values, const Tensor& old_logprobs, const Tensor& old_values,
                             const Tensor& adv, const Tensor& ret, const Tensor& mask, double clip, double clip_v,
                             double vf_coef, const OptTensor& width_tensor) {
  for (const Tensor* t : {&logprobs, &values, &old_logprobs, &old_values, &adv, &ret, &mask}) {
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kFloat && t->is_contiguous() && t->numel() == logprobs.numel(),
                "ppo_loss: all inputs must be contiguous fp32 CUDA tensors of equal size");
  }
  c10::cuda::CUDAGuard guard(logprobs.device());
  const int total = (int)logprobs.numel();
  int nblocks = (total + 255) / 256;
  if (nblocks > 64) nblocks = 64;
  if (nblocks < 1) nblocks = 1;
  auto f32 = logprobs.options();
  Tensor ws = torch::empty({b200_ppo_loss_workspace_floats(nblocks)}, f32);
  Tensor out = torch::empty({b200_ppo_loss_num_outputs()}, f32);
  Tensor dlp = torch::empty_like(logprobs), dv = torch::empty_like(values);
  check(b200_ppo_loss(logprobs.data_ptr<float>(), values.data_ptr<float>(), old_logprobs.data_ptr<float>(),
                      old_values.data_ptr<float>(), adv.data_ptr<float>(), ret.data_ptr<float>(), mask.data_ptr<float>(),
                      total, (float)clip, (float)clip_v, (float)vf_coef, dlp.data_ptr<float>(), dv.data_ptr<float>(),
                      ws.data_ptr<float>(), nblocks, out.data_ptr<float>(), (int)logprobs.size(0),
                      width_tensor.has_value() ? width_tensor->data_ptr<int>() : nullptr, stream()),
        "ppo_loss");
  return {out, dlp, dv};
}

// returns (rewards[B,R], kl_stats double[2] = (sum of per-row KL, rows))

// make_experience's per-chunk post-processing in one launch → (rewards, logprobs, values on the response window, slice_len, Σk3)
std::vector<Tensor> rollout_rewards(const Tensor& lp, const Tensor& ref_lp, const Tensor& values, const Tensor& mask,
                                    const Tensor& scores, int64_t start, double kl_coef) {
  CHECK_F32(lp); CHECK_F32(ref_lp); CHECK_F32(values); CHECK_F32(scores);
  TORCH_CHECK(lp.is_contiguous() && ref_lp.is_contiguous() && values.is_contiguous() && mask.is_contiguous() &&
              scores.is_contiguous() && mask.scalar_type() == at::kLong);
  const int64_t B = lp.size(0), Tm1 = lp.size(1), R = Tm1 - start;
  TORCH_CHECK(mask.size(0) == B && mask.size(1) == Tm1 + 1 && scores.numel() == B && R > 0);
  c10::cuda::CUDAGuard guard(lp.device());
  Tensor rewards = torch::empty({B, R}, lp.options()), lp_out = torch::empty({B, R}, lp.options()),
         v_out = torch::empty({B, R}, lp.options());
  Tensor slice_len = torch::empty({B}, lp.options().dtype(at::kInt));
  Tensor kl = torch::zeros({1}, lp.options().dtype(at::kDouble));
  check(b200_rollout_rewards(lp.data_ptr<float>(), ref_lp.data_ptr<float>(), values.data_ptr<float>(),
                             reinterpret_cast<const long long*>(mask.data_ptr<int64_t>()), scores.data_ptr<float>(), (int)B,
                             (int)Tm1, (int)start, (float)kl_coef, rewards.data_ptr<float>(), lp_out.data_ptr<float>(),
                             v_out.data_ptr<float>(), slice_len.data_ptr<int>(), kl.data_ptr<double>(), stream()),
        "rollout_rewards");
  return {rewards, lp_out, v_out, slice_len, kl};
}

// 8-bit-state AdamW step on one parameter tensor (in place): mq int8 / vq uint8 codes padded to 256-element blocks + fp32 block scales
void adam8bit(Tensor& param, const Tensor& grad, Tensor& mq, Tensor& mscale, Tensor& vq, Tensor& vscale, double lr, double beta1,
              double beta2, double eps, double weight_decay, bool decoupled, int64_t step) {
  TORCH_CHECK(param.is_cuda() && param.is_contiguous() && grad.is_contiguous() && grad.scalar_type() == param.scalar_type());
  TORCH_CHECK(param.scalar_type() == at::kFloat || param.scalar_type() == at::kBFloat16, "adam8bit: fp32 or bf16 parameters");
  const int64_t n = param.numel(), blocks = (n + 255) / 256;
  TORCH_CHECK(mq.scalar_type() == at::kChar && vq.scalar_type() == at::kByte && mq.numel() == blocks * 256 && vq.numel() == blocks * 256);
  CHECK_F32(mscale); CHECK_F32(vscale);
  TORCH_CHECK(mscale.numel() == blocks && vscale.numel() == blocks && mq.is_contiguous() && vq.is_contiguous());
  c10::cuda::CUDAGuard guard(param.device());
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  check(b200_adam8bit(param.data_ptr(), grad.data_ptr(), param.scalar_type() == at::kFloat, mq.data_ptr(), mscale.data_ptr<float>(),
                      vq.data_ptr(), vscale.data_ptr<float>(), n, (float)beta1, (float)beta2, (float)eps, (float)weight_decay,
                      decoupled ? 1 : 0, (float)lr, (float)bc1, (float)bc2, stream()),
        "adam8bit");
}

void adamw_flat(Tensor& param, Tensor& master, const Tensor& grad, Tensor& exp_avg, Tensor& exp_avg_sq, double beta1,
                double beta2, double eps, double weight_decay, bool decoupled, const Tensor& hyper) {
  CHECK_BF16(param); CHECK_F32(master); CHECK_F32(exp_avg); CHECK_F32(exp_avg_sq); CHECK_F32(hyper);
  TORCH_CHECK(grad.is_cuda() && (grad.scalar_type() == at::kFloat || grad.scalar_type() == at::kBFloat16));
  c10::cuda::CUDAGuard guard(param.device());
  check(b200_adamw_flat(param.data_ptr(), master.data_ptr<float>(), grad.data_ptr(), grad.scalar_type() == at::kFloat,
                        exp_avg.data_ptr<float>(), exp_avg_sq.data_ptr<float>(), param.numel(), (float)beta1, (float)beta2,
                        (float)eps, (float)weight_decay, decoupled ? 1 : 0, hyper.data_ptr<float>(), stream()),
        "adamw_flat");
}

void sqnorm_(const Tensor& x, Tensor& out) {
  TORCH_CHECK(out.scalar_type() == at::kDouble);
  c10::cuda::CUDAGuard guard(x.device());
  check(b200_sqnorm(x.data_ptr(), x.scalar_type() == at::kFloat, x.numel(), out.data_ptr<double>(), stream()), "sqnorm");
}

void clip_coef_(const Tensor& sqsum, double max_norm, Tensor& hyper, const OptTensor& norm_out) {
  c10::cuda::CUDAGuard guard(hyper.device());
  check(b200_clip_coef(sqsum.data_ptr<double>(), (float)max_norm, hyper.data_ptr<float>(), (float*)optptr(norm_out), stream()),
        "clip_coef");
}

// `epoch` is a 1-element int32 CUDA tensor owned by the caller (zero-initialised); every call increments it on the device
void signal_barrier(const std::vector<int64_t>& pads, int64_t rank, Tensor& epoch) {
  TORCH_CHECK(epoch.is_cuda() && epoch.scalar_type() == at::kInt && epoch.numel() == 1);
  c10::cuda::CUDAGuard guard(epoch.device());
  std::vector<void*> p(pads.size());
  for (size_t i = 0; i < pads.size(); ++i) p[i] = reinterpret_cast<void*>(pads[i]);
  check(b200_signal_barrier(p.data(), (int)rank, (int)pads.size(), reinterpret_cast<unsigned int*>(epoch.data_ptr<int>()), stream()),
        "signal_barrier");
}

void rs_adamw_ag(const std::vector<int64_t>& grads, const std::vector<int64_t>& params, int64_t lo, int64_t n, Tensor& master,
                 Tensor& exp_avg, Tensor& exp_avg_sq, const OptTensor& gshard, int64_t mode, double beta1, double beta2,
                 double eps, double weight_decay, bool decoupled, const Tensor& hyper, const OptTensor& sq_out) {
  TORCH_CHECK(grads.size() == params.size());
  std::vector<void*> g(grads.size()), p(params.size());
  for (size_t i = 0; i < grads.size(); ++i) { g[i] = reinterpret_cast<void*>(grads[i]); p[i] = reinterpret_cast<void*>(params[i]); }
  c10::cuda::CUDAGuard guard(master.device());
  check(b200_rs_adamw_ag(g.data(), p.data(), (int)grads.size(), lo, n, master.data_ptr<float>(), exp_avg.data_ptr<float>(),
                         exp_avg_sq.data_ptr<float>(), (float*)optptr(gshard), (int)mode, (float)beta1, (float)beta2,
                         (float)eps, (float)weight_decay, decoupled ? 1 : 0, hyper.data_ptr<float>(),
                         (double*)optptr(sq_out), stream()),
        "rs_adamw_ag");
}

// One gradient bucket of the overlapped distributed optimizer; master / exp_avg / exp_avg_sq / gshard are the views of this
// bucket's shard, `epoch` / `done` 1-element int32 views of the per-bucket device counters.
void rs_adamw_ag_bucket(const std::vector<int64_t>& grads, const std::vector<int64_t>& params, int64_t rank, int64_t lo,
                        int64_t n, Tensor& master, Tensor& exp_avg, Tensor& exp_avg_sq, const OptTensor& gshard, int64_t mode,
                        double beta1, double beta2, double eps, double weight_decay, bool decoupled, const Tensor& hyper,
                        const OptTensor& sq_out, const std::vector<int64_t>& flags, int64_t flag_offset, Tensor& epoch,
                        Tensor& done, int64_t max_blocks, int64_t mc_grad, int64_t mc_param, double grad_scale) {
  TORCH_CHECK(grads.size() == params.size() && grads.size() == flags.size());
  TORCH_CHECK(epoch.scalar_type() == at::kInt && done.scalar_type() == at::kInt);
  std::vector<void*> g(grads.size()), p(params.size()), f(flags.size());
  for (size_t i = 0; i < grads.size(); ++i) {
    g[i] = reinterpret_cast<void*>(grads[i]);
    p[i] = reinterpret_cast<void*>(params[i]);
    f[i] = reinterpret_cast<void*>(flags[i]);
  }
  c10::cuda::CUDAGuard guard(master.device());
  check(b200_rs_adamw_ag_bucket(g.data(), p.data(), (int)grads.size(), (int)rank, lo, n, master.data_ptr<float>(),
                                exp_avg.data_ptr<float>(), exp_avg_sq.data_ptr<float>(), (float*)optptr(gshard), (int)mode,
                                (float)beta1, (float)beta2, (float)eps, (float)weight_decay, decoupled ? 1 : 0,
                                hyper.data_ptr<float>(), (double*)optptr(sq_out), f.data(), flag_offset,
                                reinterpret_cast<unsigned int*>(epoch.data_ptr<int>()),
                                reinterpret_cast<unsigned int*>(done.data_ptr<int>()), (int)max_blocks,
                                reinterpret_cast<const void*>(mc_grad), reinterpret_cast<void*>(mc_param), (float)grad_scale,
                                stream()),
        "rs_adamw_ag_bucket");
}

void clip_exchange(const std::vector<int64_t>& sqbufs, const std::vector<int64_t>& flags, int64_t flag_offset, int64_t rank,
                   const Tensor& sq_local, Tensor& epoch, double max_norm, Tensor& hyper, const OptTensor& norm_out) {
  TORCH_CHECK(sqbufs.size() == flags.size() && sq_local.scalar_type() == at::kDouble && epoch.scalar_type() == at::kInt);
  std::vector<void*> s(sqbufs.size()), f(flags.size());
  for (size_t i = 0; i < sqbufs.size(); ++i) { s[i] = reinterpret_cast<void*>(sqbufs[i]); f[i] = reinterpret_cast<void*>(flags[i]); }
  c10::cuda::CUDAGuard guard(hyper.device());
  check(b200_clip_exchange(s.data(), f.data(), flag_offset, (int)rank, (int)sqbufs.size(), sq_local.data_ptr<double>(),
                           reinterpret_cast<unsigned int*>(epoch.data_ptr<int>()), (float)max_norm, hyper.data_ptr<float>(),
                           (float*)optptr(norm_out), stream()),
        "clip_exchange");
}

void lerp_(Tensor& tgt, const Tensor& src, double alpha) {
  CHECK_BF16(tgt); CHECK_BF16(src);
  TORCH_CHECK(tgt.is_contiguous() && src.is_contiguous() && tgt.numel() == src.numel());
  c10::cuda::CUDAGuard guard(tgt.device());
  check(b200_lerp_bf16(tgt.data_ptr(), src.data_ptr(), tgt.numel(), (float)alpha, stream()), "lerp");
}

// out[M, N] = act(concat_r A_r . w^T + bias) where A_r ([M/world, K], row pitch lda) lives at device address peers[r]
Tensor gemm_allgather(const std::vector<int64_t>& peers, int64_t rows_per_rank, int64_t K, int64_t lda, const Tensor& w,
                      const OptTensor& bias, const std::string& act, const OptTensor& out_) {
  CHECK_BF16(w);
  TORCH_CHECK(w.dim() == 2 && w.size(1) == K && w.stride(1) == 1 && K % 8 == 0 && lda % 8 == 0);
  c10::cuda::CUDAGuard guard(w.device());
  const int64_t world = (int64_t)peers.size(), M = rows_per_rank * world, N = w.size(0);
  std::vector<void*> p(world);
  for (int64_t i = 0; i < world; ++i) p[i] = reinterpret_cast<void*>(peers[i]);
  Tensor out = out_.has_value() ? *out_ : torch::empty({M, N}, w.options());
  check(b200_gemm_allgather_bf16(p.data(), (int)world, w.data_ptr(), out.data_ptr(), (int)M, (int)N, (int)K, lda, w.stride(0),
                                 out.stride(0), optptr(bias), act_code(act), stream()),
        "gemm_allgather");
  return out;
}

// GEMM whose A rows become readable chunk by chunk (flags[c] == epoch), starting with chunk `first_chunk` (see AReady)
Tensor gemm_flagged(const Tensor& a, const Tensor& w, const OptTensor& bias, const std::string& act, const Tensor& flags,
                    int64_t epoch, int64_t rows_per_flag, int64_t first_chunk) {
  CHECK_BF16(a); CHECK_BF16(w);
  TORCH_CHECK(a.dim() == 2 && w.dim() == 2 && a.size(1) == w.size(1) && a.stride(1) == 1 && w.stride(1) == 1);
  TORCH_CHECK(flags.scalar_type() == at::kInt && flags.is_cuda() && flags.is_contiguous());
  TORCH_CHECK(a.size(0) % rows_per_flag == 0 && flags.numel() >= a.size(0) / rows_per_flag);
  c10::cuda::CUDAGuard guard(a.device());
  Tensor out = torch::empty({a.size(0), w.size(0)}, a.options());
  check(b200_gemm_flagged_bf16(a.data_ptr(), w.data_ptr(), out.data_ptr(), (int)a.size(0), (int)w.size(0), (int)a.size(1),
                               a.stride(0), w.stride(0), out.stride(0), optptr(bias), act_code(act),
                               reinterpret_cast<const uint32_t*>(flags.data_ptr<int>()), (uint32_t)epoch, (int)rows_per_flag,
                               (int)first_chunk, stream()),
        "gemm_flagged");
  return out;
}

// partial product x[M, K_local] . w[N, K_local]^T stored (bf16) into slot `rank` of every owner's staging buffer
void gemm_stage_scatter(const Tensor& x, const Tensor& w, const std::vector<int64_t>& stage_peers, int64_t rank, int64_t ldstage,
                        const OptTensor& bias) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && x.stride(1) == 1 && w.stride(1) == 1);
  c10::cuda::CUDAGuard guard(x.device());
  std::vector<void*> p(stage_peers.size());
  for (size_t i = 0; i < stage_peers.size(); ++i) p[i] = reinterpret_cast<void*>(stage_peers[i]);
  check(b200_gemm_stage_scatter_bf16(x.data_ptr(), w.data_ptr(), p.data(), (int)p.size(), (int)rank, (int)x.size(0), (int)w.size(0),
                                     (int)x.size(1), x.stride(0), w.stride(0), ldstage, optptr(bias), stream()),
        "gemm_stage_scatter");
}

// out[rows, N] = sum over the `world` slots of stage [world, rows, N] (+ bias + residual)
Tensor stage_reduce(const Tensor& stage, const OptTensor& bias, const OptTensor& residual) {
  CHECK_BF16(stage);
  TORCH_CHECK(stage.dim() == 3 && stage.is_contiguous());
  c10::cuda::CUDAGuard guard(stage.device());
  const int64_t world = stage.size(0), rows = stage.size(1), N = stage.size(2);
  Tensor out = torch::empty({rows, N}, stage.options());
  long long ldr = 0;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->size(0) == rows && residual->size(1) == N && residual->stride(1) == 1); ldr = residual->stride(0); }
  check(b200_stage_reduce(stage.data_ptr(), (int)world, optptr(bias), optptr(residual), out.data_ptr(), rows, (int)N, N, ldr,
                          out.stride(0), stream()),
        "stage_reduce");
  return out;
}

// out[rows, col0 : col0 + ncols] = NVLS sum over all ranks of their partial [.., ldp] rows behind the multicast address `mc_ptr`
// (already offset to this rank's first row) (+ bias + residual)
void mc_reduce_rows(int64_t mc_ptr, Tensor& out, const OptTensor& bias, const OptTensor& residual, int64_t col0, int64_t ncols,
                    int64_t ldp, int64_t blocks) {
  CHECK_BF16(out);
  TORCH_CHECK(out.dim() == 2 && out.stride(1) == 1 && mc_ptr != 0);
  c10::cuda::CUDAGuard guard(out.device());
  long long ldr = 0;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->size(0) == out.size(0) && residual->stride(1) == 1); ldr = residual->stride(0); }
  if (bias.has_value()) { CHECK_BF16(*bias); }
  check(b200_mc_reduce_rows(reinterpret_cast<const void*>(mc_ptr), optptr(bias), optptr(residual), out.data_ptr(), out.size(0),
                            (int)ncols, (int)col0, ldp, ldr, out.stride(0), (int)blocks, stream()),
        "mc_reduce_rows");
}

// adds this rank's partial product x[M, K_local] . w[N, K_local]^T into the peers' fp32 accumulators (row-blocks by owner)
void gemm_reduce_scatter(const Tensor& x, const Tensor& w, const std::vector<int64_t>& acc_peers, int64_t ldacc,
                         const OptTensor& bias) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && x.stride(1) == 1 && w.stride(1) == 1);
  TORCH_CHECK(x.size(1) % 8 == 0 && x.stride(0) % 8 == 0 && w.stride(0) % 8 == 0);
  c10::cuda::CUDAGuard guard(x.device());
  std::vector<float*> p(acc_peers.size());
  for (size_t i = 0; i < acc_peers.size(); ++i) p[i] = reinterpret_cast<float*>(acc_peers[i]);
  check(b200_gemm_reduce_scatter_bf16(x.data_ptr(), w.data_ptr(), p.data(), (int)p.size(), (int)x.size(0), (int)w.size(0),
                                      (int)x.size(1), x.stride(0), w.stride(0), ldacc, optptr(bias), stream()),
        "gemm_reduce_scatter");
}

Tensor rs_finalize(const Tensor& acc, const OptTensor& bias, const OptTensor& residual, const OptTensor& out_) {
  CHECK_F32(acc);
  TORCH_CHECK(acc.dim() == 2 && acc.stride(1) == 1);
  c10::cuda::CUDAGuard guard(acc.device());
  Tensor out = out_.has_value() ? *out_ : torch::empty({acc.size(0), acc.size(1)}, acc.options().dtype(at::kBFloat16));
  long long ldr = 0;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->stride(1) == 1); ldr = residual->stride(0); }
  check(b200_rs_finalize(acc.data_ptr<float>(), optptr(bias), optptr(residual), out.data_ptr(), acc.size(0), (int)acc.size(1),
                         acc.stride(0), ldr, out.stride(0), stream()),
        "rs_finalize");
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// Page allocator for the paged KV cache.  Pages are fixed-size token slots shared by every layer (each layer has its
// own [num_pages, page_size, nkv, d] tensor but the same page ids).  Sequences are identified by an integer handle.
class PagedKVAllocator {
 public:
  PagedKVAllocator(int64_t num_pages, int64_t page_size) : page_size_(page_size), num_pages_(num_pages) {
    for (int64_t i = 0; i < num_pages; ++i) free_.push_back((int32_t)i);
  }
  // ensure `seq` can hold `n_tokens`; returns the page ids of the sequence
  std::vector<int32_t> reserve(int64_t seq, int64_t n_tokens) {
    auto& pages = seqs_[seq];
    const int64_t need = (n_tokens + page_size_ - 1) / page_size_;
    while ((int64_t)pages.size() < need) {
      if (free_.empty()) throw std::runtime_error("PagedKVAllocator: out of pages");
      pages.push_back(free_.front());
      free_.pop_front();
    }
    return pages;
  }
  void release(int64_t seq) {
    auto it = seqs_.find(seq);
    if (it == seqs_.end()) return;
    for (int32_t p : it->second) free_.push_back(p);
    seqs_.erase(it);
  }
  void reset() {
    seqs_.clear();
    free_.clear();
    for (int64_t i = 0; i < num_pages_; ++i) free_.push_back((int32_t)i);
  }
  // block table for a batch of sequence handles, -1 padded: [len(seqs), max_pages] int32 (CPU tensor)
  Tensor block_table(const std::vector<int64_t>& seqs, int64_t max_pages) {
    Tensor t = torch::full({(int64_t)seqs.size(), max_pages}, 0, torch::dtype(torch::kInt32));
    auto a = t.accessor<int32_t, 2>();
    for (size_t i = 0; i < seqs.size(); ++i) {
      auto it = seqs_.find(seqs[i]);
      if (it == seqs_.end()) continue;
      TORCH_CHECK((int64_t)it->second.size() <= max_pages, "block_table: max_pages too small");
      for (size_t j = 0; j < it->second.size(); ++j) a[i][j] = it->second[j];
    }
    return t;
  }
  int64_t free_pages() const { return (int64_t)free_.size(); }
  int64_t page_size() const { return page_size_; }
  int64_t num_pages() const { return num_pages_; }

 private:
  int64_t page_size_, num_pages_;
  std::deque<int32_t> free_;
  std::unordered_map<int64_t, std::vector<int32_t>> seqs_;
};


// ---- decode megakernel (csrc/decode_mega.cu): per-layer pointer table + weight tensor maps, built once per engine ----
// `layers`: list (one per block) of lists [ln1_w, ln1_b, ln2_w, ln2_b, qkv_w, qkv_b, out_w, out_b, fc_w, fc_b, fc2_w, fc2_b,
// kcache, vcache]; biases / ln biases may be None.  Returns (pointer table, tensor maps) as uint8 CUDA tensors.
std::vector<Tensor> decode_mega_build(const std::vector<std::vector<OptTensor>>& layers) {
  const size_t L = layers.size();
  TORCH_CHECK(L > 0, "no layers");
  const int lb = b200_decode_mega_layer_bytes();
  TORCH_CHECK(lb == 10 * (int)sizeof(void*), "DmLayer layout mismatch");
  std::vector<const void*> table(L * 10, nullptr);
  std::vector<uint8_t> maps(L * 4 * 128, 0);
  at::Device dev(at::kCPU);
  for (size_t l = 0; l < L; ++l) {
    const auto& t = layers[l];
    TORCH_CHECK(t.size() == 14, "each layer needs 14 entries");
    auto ptr = [&](int i) -> const void* {
      if (!t[i].has_value()) return nullptr;
      CHECK_BF16(*t[i]);
      TORCH_CHECK(t[i]->is_contiguous(), "megakernel tensors must be contiguous");
      return t[i]->data_ptr();
    };
    const void* e[10] = {ptr(0), ptr(1), ptr(2), ptr(3), ptr(5), ptr(7), ptr(9), ptr(11), ptr(12), ptr(13)};
    for (int i = 0; i < 10; ++i) table[l * 10 + i] = e[i];
    const int widx[4] = {4, 6, 8, 10};
    for (int g = 0; g < 4; ++g) {
      TORCH_CHECK(t[widx[g]].has_value(), "weight missing");
      const Tensor& w = *t[widx[g]];
      CHECK_BF16(w);
      TORCH_CHECK(w.dim() == 2 && w.stride(1) == 1, "weights must be row-major [out, in]");
      dev = w.device();
      check(b200_decode_mega_make_map(maps.data() + (l * 4 + g) * 128, w.data_ptr(), w.size(0), w.size(1), w.stride(0)),
            "decode_mega_make_map");
    }
  }
  c10::cuda::CUDAGuard guard(dev);
  auto opts = torch::TensorOptions().dtype(torch::kUInt8);
  Tensor table_t = torch::from_blob(table.data(), {(int64_t)(table.size() * sizeof(void*))}, opts).clone().to(dev);
  // tensor maps must be 64-byte aligned: over-allocate and hand back an aligned view
  Tensor raw = torch::empty({(int64_t)maps.size() + 128}, opts.device(dev));
  const uintptr_t base = reinterpret_cast<uintptr_t>(raw.data_ptr());
  const int64_t off = (int64_t)(((base + 127) & ~uintptr_t(127)) - base);
  Tensor maps_t = raw.narrow(0, off, (int64_t)maps.size());
  maps_t.copy_(torch::from_blob(maps.data(), {(int64_t)maps.size()}, opts));
  return {table_t, maps_t};
}

void decode_mega(Tensor& x, Tensor& a, Tensor& mid, const Tensor& block_table, const Tensor& seq_lens, const Tensor& table,
                 const Tensor& maps, int64_t nh, int64_t L, const std::string& act, bool rms, double eps, double scale,
                 int64_t page_size, const OptTensor& trunk_out, const OptTensor& step, int64_t branch, const OptTensor& alibi,
                 const OptTensor& timing, int64_t cluster_size) {
  CHECK_BF16(x); CHECK_BF16(a); CHECK_BF16(mid);
  TORCH_CHECK(x.is_contiguous() && a.is_contiguous() && mid.is_contiguous() && block_table.is_contiguous());
  TORCH_CHECK(block_table.scalar_type() == at::kInt && seq_lens.scalar_type() == at::kInt);
  c10::cuda::CUDAGuard guard(x.device());
  const int B = (int)x.size(0), H = (int)x.size(1), F = (int)mid.size(1);
  void* tr = nullptr;
  long long tr_stride = 0;
  const long long* step_ptr = nullptr;
  if (trunk_out.has_value()) {
    CHECK_BF16(*trunk_out);
    TORCH_CHECK(trunk_out->dim() == 3 && trunk_out->size(2) == H && trunk_out->stride(2) == 1 && trunk_out->stride(1) == H);
    tr = trunk_out->data_ptr();
    tr_stride = trunk_out->stride(0);
    TORCH_CHECK(step.has_value() && step->scalar_type() == at::kLong);
    step_ptr = reinterpret_cast<const long long*>(step->data_ptr<int64_t>());
  }
  check(b200_decode_mega(B, H, F, (int)nh, (int)L, act_code(act), rms ? 1 : 0, (float)eps, (float)scale, (int)page_size,
                         (int)block_table.size(1), block_table.data_ptr<int>(), seq_lens.data_ptr<int>(), x.data_ptr(),
                         a.data_ptr(), mid.data_ptr(), table.data_ptr(), maps.data_ptr(), tr, tr_stride, step_ptr, (int)branch,
                         (const float*)optptr(alibi),
                         timing.has_value() ? reinterpret_cast<long long*>(timing->data_ptr<int64_t>()) : nullptr,
                         (int)cluster_size, stream()),
        "decode_mega");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("umma_probe", [](int64_t M, int64_t N, int64_t n_mma, int64_t nacc) {
    Tensor out = torch::zeros({2}, torch::TensorOptions().dtype(torch::kLong).device(torch::kCUDA));
    check(b200_umma_probe((int)M, (int)N, (int)n_mma, (int)nacc, reinterpret_cast<long long*>(out.data_ptr<int64_t>()), stream()),
          "umma_probe");
    return out;
  });
  m.def("decode_mega_build", &decode_mega_build);
  m.def("decode_mega", &decode_mega, py::arg("x"), py::arg("a"), py::arg("mid"), py::arg("block_table"), py::arg("seq_lens"),
        py::arg("table"), py::arg("maps"), py::arg("nh"), py::arg("L"), py::arg("act"), py::arg("rms"), py::arg("eps"),
        py::arg("scale"), py::arg("page_size"), py::arg("trunk_out") = py::none(), py::arg("step") = py::none(),
        py::arg("branch") = -1, py::arg("alibi") = py::none(), py::arg("timing") = py::none(), py::arg("cluster_size") = 0);
  m.def("decode_mega_stages", [](int64_t H, int64_t F) { return b200_decode_mega_stages((int)H, (int)F); });
  m.def("decode_mega_max_clusters", [](int64_t H, int64_t F, int64_t cs) { return b200_decode_mega_max_clusters((int)H, (int)F, (int)cs); },
        py::arg("H"), py::arg("F"), py::arg("cluster_size") = 16);
  m.def("decode_mega_cluster_size", [](int64_t H, int64_t F, int64_t groups) { return b200_decode_mega_cluster_size((int)H, (int)F, (int)groups); });
  namespace py = pybind11;
  m.doc() = "trlx_b200 sm_100a kernels";
  m.def("gemm", &gemm, py::arg("x"), py::arg("w"), py::arg("bias") = py::none(), py::arg("residual") = py::none(),
        py::arg("act") = "none", py::arg("out_f32") = false, py::arg("out") = py::none(), py::arg("col_scale") = py::none(),
        py::arg("alpha") = 1.0, py::arg("force_bn") = 0);
  m.def("lmhead", &lmhead, py::arg("h"), py::arg("w"), py::arg("bias") = py::none(), py::arg("labels") = py::none(),
        py::arg("sample") = false, py::arg("temperature") = 1.0, py::arg("seed") = 0, py::arg("step") = py::none(),
        py::arg("suppress_col") = -1, py::arg("suppress_until") = 0, py::arg("workspace") = py::none(),
        py::arg("seed_tensor") = py::none());
  m.def("gemm_ln", &gemm_ln, py::arg("x"), py::arg("w"), py::arg("bias") = py::none(), py::arg("residual") = py::none(),
        py::arg("act") = "none", py::arg("ln_stats") = py::none(), py::arg("ln_c1") = py::none(), py::arg("ln_eps") = 1e-5,
        py::arg("ln_rms") = false, py::arg("stats_out") = py::none());
  m.def("gemm_fp8", &gemm_fp8, py::arg("a8"), py::arg("b8"), py::arg("row_scale"), py::arg("col_scale"), py::arg("bias") = py::none(),
        py::arg("residual") = py::none(), py::arg("act") = "none");
  m.def("quant_rows", &quant_rows);
  m.def("norm_quant", &norm_quant, py::arg("x"), py::arg("w"), py::arg("b") = py::none(), py::arg("eps") = 1e-5,
        py::arg("rms") = false);
  m.def("gemm_flagged", &gemm_flagged, py::arg("a"), py::arg("w"), py::arg("bias"), py::arg("act"), py::arg("flags"),
        py::arg("epoch"), py::arg("rows_per_flag"), py::arg("first_chunk"));
  m.def("gemm_stage_scatter", &gemm_stage_scatter, py::arg("x"), py::arg("w"), py::arg("stage_peers"), py::arg("rank"),
        py::arg("ldstage"), py::arg("bias") = py::none());
  m.def("mc_reduce_rows", &mc_reduce_rows, py::arg("mc_ptr"), py::arg("out"), py::arg("bias") = py::none(),
        py::arg("residual") = py::none(), py::arg("col0") = 0, py::arg("ncols") = 0, py::arg("ldp") = 0, py::arg("blocks") = 0);
  m.def("stage_reduce", &stage_reduce, py::arg("stage"), py::arg("bias") = py::none(), py::arg("residual") = py::none());
  m.def("gemm_ex", &gemm_ex, py::arg("a"), py::arg("b"), py::arg("a_mn") = false, py::arg("b_mn") = false,
        py::arg("out_f32") = false, py::arg("split_k") = -1, py::arg("out") = py::none(), py::arg("accumulate") = false);
  m.def("gemm_splitk_plan", [](int64_t m, int64_t n, int64_t k) { return (int64_t)b200_gemm_splitk_plan((int)m, (int)n, (int)k); },
        py::arg("m"), py::arg("n"), py::arg("k"));
  m.def("lmhead_tiles", [](int64_t n) { return (int64_t)b200_lmhead_tiles((int)n); }, py::arg("vocab"));
  m.def("lmhead_dlogits", &lmhead_dlogits, py::arg("h"), py::arg("w"), py::arg("bias"), py::arg("labels"), py::arg("lse"),
        py::arg("grad"));
  m.def("norm", &norm, py::arg("x"), py::arg("w"), py::arg("b") = py::none(), py::arg("eps") = 1e-5, py::arg("rms") = false,
        py::arg("out") = py::none());
  m.def("embed", &embed, py::arg("tokens"), py::arg("positions"), py::arg("wte"), py::arg("wpe") = py::none(),
        py::arg("pos_offset") = 0, py::arg("out") = py::none(), py::arg("stats_out") = py::none());
  m.def("decode_attention", &decode_attention, py::arg("qkv"), py::arg("kcache"), py::arg("vcache"), py::arg("block_table"),
        py::arg("seq_lens"), py::arg("positions"), py::arg("nq"), py::arg("nkv"), py::arg("d"), py::arg("scale"),
        py::arg("rot_dim") = 0, py::arg("rot_base") = 10000.0, py::arg("rot_interleaved") = false,
        py::arg("alibi") = py::none(), py::arg("window") = 0, py::arg("out") = py::none());
  m.def("rowdot", &rowdot, py::arg("x"), py::arg("w"), py::arg("bias") = py::none(), py::arg("out") = py::none());
  m.def("decode_step", &decode_step);
  m.def("paged_kv_write", &paged_kv_write);
  m.def("logprob_from_logits", &logprob_from_logits);
  m.def("logprob_backward_inplace", &logprob_backward_inplace);
  m.def("gae", &gae, py::arg("values"), py::arg("rewards"), py::arg("width"), py::arg("gamma"), py::arg("lam"),
        py::arg("whiten") = true, py::arg("unbiased") = true, py::arg("width_tensor") = py::none());
  m.def("whiten_", &whiten_, py::arg("adv"), py::arg("width"), py::arg("stats"), py::arg("unbiased"),
        py::arg("width_tensor") = py::none());
  m.def("ppo_loss", &ppo_loss, py::arg("logprobs"), py::arg("values"), py::arg("old_logprobs"), py::arg("old_values"),
        py::arg("adv"), py::arg("ret"), py::arg("mask"), py::arg("clip"), py::arg("clip_v"), py::arg("vf_coef"),
        py::arg("width_tensor") = py::none());
  m.def("adamw_flat", &adamw_flat);
  m.def("adam8bit", &adam8bit);
  m.def("sqnorm_", &sqnorm_);
  m.def("clip_coef_", &clip_coef_, py::arg("sqsum"), py::arg("max_norm"), py::arg("hyper"), py::arg("norm_out") = py::none());
  m.def("signal_barrier", &signal_barrier);
  m.def("rs_adamw_ag", &rs_adamw_ag);
  m.def("rs_adamw_ag_bucket", &rs_adamw_ag_bucket);
  m.def("rollout_rewards", &rollout_rewards);
  m.def("ilql_sample", &ilql_sample, py::arg("logits"), py::arg("q1"), py::arg("q2") = py::none(), py::arg("vs"), py::arg("V"),
        py::arg("beta"), py::arg("top_k"), py::arg("temperature"), py::arg("seed"), py::arg("step") = py::none(),
        py::arg("seed_tensor") = py::none(), py::arg("logit_mask") = py::none(), py::arg("last_tokens") = py::none());
  m.def("sample_filtered", &sample_filtered, py::arg("logits"), py::arg("V"), py::arg("top_k"), py::arg("top_p"),
        py::arg("temperature"), py::arg("seed"), py::arg("step") = py::none(), py::arg("suppress_col") = -1,
        py::arg("suppress_until") = 0, py::arg("seed_tensor") = py::none());
  m.def("clip_exchange", &clip_exchange);
  m.def("lerp_", &lerp_);
  m.def("gemm_allgather", &gemm_allgather, py::arg("peers"), py::arg("rows_per_rank"), py::arg("K"), py::arg("lda"),
        py::arg("w"), py::arg("bias") = py::none(), py::arg("act") = "none", py::arg("out") = py::none());
  m.def("gemm_reduce_scatter", &gemm_reduce_scatter, py::arg("x"), py::arg("w"), py::arg("acc_peers"), py::arg("ldacc"),
        py::arg("bias") = py::none());
  m.def("rs_finalize", &rs_finalize, py::arg("acc"), py::arg("bias") = py::none(), py::arg("residual") = py::none(),
        py::arg("out") = py::none());
  m.def("ppo_loss_num_outputs", [] { return b200_ppo_loss_num_outputs(); });
  m.def("set_static_weights", [](bool on) { b200_set_static_weights(on ? 1 : 0); },
        "GEMMs launched while set may prefetch their weight tiles ahead of the PDL wait (weights must be read-only)");
  m.def("get_static_weights", []() { return b200_get_static_weights() != 0; });
  m.def("attn_short_ok", [](int64_t tq, int64_t tk, int64_t d, bool backward) {
    return b200_attn_short_ok((int)tq, (int)tk, (int)d, backward ? 1 : 0) != 0; },
        py::arg("tq"), py::arg("tk"), py::arg("d"), py::arg("backward") = false);
  m.def("attn_tc_ok", [](int64_t tq, int64_t tk, int64_t d) { return b200_attn_tc_ok((int)tq, (int)tk, (int)d) != 0; });
  m.def("attn_tc_fwd", &attn_tc_fwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("bias") = py::none(),
        py::arg("causal") = false, py::arg("scale") = 1.0);
  m.def("attn_short_fwd", &attn_short_fwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("bias") = py::none(),
        py::arg("causal") = false, py::arg("scale") = 1.0);
  m.def("attn_short_bwd", &attn_short_bwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("bias"), py::arg("o"),
        py::arg("d_o"), py::arg("stats"), py::arg("causal") = false, py::arg("scale") = 1.0);
  m.def("ln_train_ok", [](int64_t h) { return b200_ln_train_ok((int)h) != 0; });
  m.def("ln_fwd", &ln_fwd, py::arg("x"), py::arg("w"), py::arg("b") = py::none(), py::arg("eps") = 1e-5, py::arg("rms") = false);
  m.def("ln_bwd", &ln_bwd, py::arg("x"), py::arg("w"), py::arg("stats"), py::arg("dy"), py::arg("rms") = false,
        py::arg("has_bias") = true);
  m.def("colsum", &colsum, py::arg("x"));
  m.def("set_pdl", [](bool on) { b200_set_pdl(on ? 1 : 0); }, "enable/disable programmatic dependent launch for the kernels");
  m.def("get_pdl", [] { return b200_get_pdl() != 0; });
  py::class_<PagedKVAllocator>(m, "PagedKVAllocator")
      .def(py::init<int64_t, int64_t>())
      .def("reserve", &PagedKVAllocator::reserve)
      .def("release", &PagedKVAllocator::release)
      .def("reset", &PagedKVAllocator::reset)
      .def("block_table", &PagedKVAllocator::block_table)
      .def_property_readonly("free_pages", &PagedKVAllocator::free_pages)
      .def_property_readonly("page_size", &PagedKVAllocator::page_size)
      .def_property_readonly("num_pages", &PagedKVAllocator::num_pages);
}
