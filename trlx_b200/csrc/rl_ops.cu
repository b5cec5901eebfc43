// Fused RL math (sm_100a).  Each of these replaces a chain of aten launches (and, in the reference, host syncs):
//   * logprob_from_logits : log_softmax + gather without materialising log-probs   (trlx/utils/modeling.py:213-219)
//   * gae + whiten        : reverse scan per row + global moments                  (trlx/models/modeling_ppo.py:161-173)
//   * ppo_loss            : clipped policy / value losses, ~20 statistics and both gradients in one pass
//                           (trlx/models/modeling_ppo.py:189-238 — every `.item()` there is a device sync)
//   * rollout_rewards     : per-token KL penalty, score placement, window slicing, k3 KL statistics
//                           (accelerate_ppo_trainer.py:455-504)
#include "ptx.cuh"

namespace b200 {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) r += sh[i];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = -INFINITY;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) r = fmaxf(r, sh[i]);
  __syncthreads();
  return r;
}

// ----------------------------------------------------------------------------- logprob of labels from logits
// One block per row; online softmax with one pass over V.  label < 0 -> 0.
template <typename T>
__global__ void __launch_bounds__(256) logprob_kernel(const T* __restrict__ logits, const long long* __restrict__ labels,
                                                      float* __restrict__ out, float* __restrict__ lse_out, int V,
                                                      long long ld) {
  __shared__ float sh[8];
  const long long row = blockIdx.x;
  const T* lr = logits + row * ld;
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float x = to_f32(lr[i]);
    if (x > m) { s = s * __expf(m - x) + 1.f; m = x; }
    else if (x > -INFINITY) s += __expf(x - m);
  }
  const float gm = block_max(m, sh);
  const float part = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
  const float gs = block_sum(part, sh);
  if (threadIdx.x == 0) {
    const float lse = gm + __logf(gs);
    if (lse_out) lse_out[row] = lse;
    const long long lab = labels[row];
    out[row] = (lab >= 0 && lab < V) ? to_f32(lr[lab]) - lse : 0.f;
  }
}

// d logits = (onehot(label) - softmax) * g   written in place over the (recomputed) logits; used by the fused
// LM-head backward so the [M, V] probability tensor is never a separate allocation.
template <typename T>
__global__ void __launch_bounds__(256) logprob_bwd_kernel(T* __restrict__ logits, const long long* __restrict__ labels,
                                                          const float* __restrict__ lse, const float* __restrict__ grad,
                                                          int V, long long ld) {
  const long long row = blockIdx.x;
  T* lr = logits + row * ld;
  const float g = grad[row], l = lse[row];
  const long long lab = labels[row];
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float p = __expf(to_f32(lr[i]) - l);
    float d = ((i == lab) ? 1.f : 0.f) - p;
    if (lab < 0) d = 0.f;
    lr[i] = (T)(d * g);
  }
}

// ----------------------------------------------------------------------------- GAE + whitening
// stats (double[3]) accumulates (count, sum, sum of squares) of the advantages over [B, width].
__global__ void gae_kernel(const float* __restrict__ values, const float* __restrict__ rewards, float* __restrict__ adv,
                           float* __restrict__ ret, int B, int R, int width_arg, const int* __restrict__ width_ptr,
                           long long ld, float gamma, float lam, double* __restrict__ stats) {
  // the effective width may live on the device so that one captured CUDA graph serves batches of different widths
  int width = width_ptr ? *width_ptr : width_arg;
  if (width > R) width = R;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0, ss = 0.0;
  if (b < B) {
    const float* v = values + (size_t)b * ld;
    const float* r = rewards + (size_t)b * ld;
    float last = 0.f, nextv = 0.f;
    for (int t = width - 1; t >= 0; --t) {
      const float vt = v[t];
      const float delta = r[t] + gamma * nextv - vt;
      last = delta + gamma * lam * last;
      adv[(size_t)b * ld + t] = last;
      ret[(size_t)b * ld + t] = last + vt;
      nextv = vt;
      s += last;
      ss += (double)last * last;
    }
    for (int t = width; t < R; ++t) { adv[(size_t)b * ld + t] = 0.f; ret[(size_t)b * ld + t] = 0.f; }
  }
  // warp-aggregate then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&stats[1], s);
    atomicAdd(&stats[2], ss);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[0], (double)B * width);
  }
}

// adv <- (adv - mean) * rsqrt(var + 1e-8); var is unbiased when `unbiased` (single-process torch.var_mean semantics)
__global__ void whiten_kernel(float* __restrict__ adv, int B, int width_arg, const int* __restrict__ width_ptr, long long ld,
                              const double* __restrict__ stats, int unbiased) {
  int width = width_ptr ? *width_ptr : width_arg;
  if (width > (int)ld) width = (int)ld;
  const double n = stats[0];
  const double mean = stats[1] / n;
  double var = stats[2] / n - mean * mean;
  if (var < 0) var = 0;
  if (unbiased && n > 1) var = var * n / (n - 1);
  const float fm = (float)mean, fr = rsqrtf((float)var + 1e-8f);
  const int span = width_ptr ? (int)ld : width;  // threads are laid out over the launch-time span
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * span) return;
  const int b = i / span, t = i % span;
  if (t >= width) return;
  float* p = adv + (size_t)b * ld + t;
  *p = (*p - fm) * fr;
}

// ----------------------------------------------------------------------------- PPO loss (forward + gradients + stats)
enum PpoAcc {
  A_N = 0, A_PG, A_VF, A_PGCLIP, A_VFCLIP, A_KL, A_RATIO, A_VAL, A_VAL2, A_OLD, A_OLD2, A_RET, A_RET2, A_VERR, A_VMAPE, A_COUNT
};
enum PpoOut {
  O_LOSS = 0, O_PG, O_VF, O_VAL_MEAN, O_VAL_MIN, O_VAL_MAX, O_VAL_STD, O_VERR, O_VMAPE, O_VFCLIP, O_OLD_MEAN, O_OLD_MIN,
  O_OLD_MAX, O_OLD_STD, O_RET_MEAN, O_RET_MIN, O_RET_MAX, O_RET_STD, O_KL, O_PGCLIP, O_RATIO, O_PAD, O_N, O_COUNT
};

// partial sums per block -> part[blockIdx][A_COUNT] ; mins/maxes -> ext[blockIdx][6]
__global__ void __launch_bounds__(256)
ppo_loss_partial_kernel(const float* __restrict__ logprobs, const float* __restrict__ values,
                        const float* __restrict__ old_logprobs, const float* __restrict__ old_values,
                        const float* __restrict__ adv, const float* __restrict__ ret, const float* __restrict__ mask,
                        int total, float clip, float clip_v, float* __restrict__ dlogprobs, float* __restrict__ dvalues,
                        float* __restrict__ part, float* __restrict__ ext) {
  __shared__ float sh[8];
  float acc[A_COUNT];
#pragma unroll
  for (int i = 0; i < A_COUNT; ++i) acc[i] = 0.f;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mxv[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const float m = mask[i], v = values[i], ov = old_values[i], R = ret[i], A = adv[i];
    const float lr = (logprobs[i] - old_logprobs[i]) * m;
    const float ratio = __expf(lr);
    const float vc = fminf(fmaxf(v, ov - clip_v), ov + clip_v);
    const float vf1 = (v - R) * (v - R), vf2 = (vc - R) * (vc - R);
    const float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
    const float pg1 = -A * ratio, pg2 = -A * rc;
    acc[A_N] += m;
    acc[A_PG] += fmaxf(pg1, pg2) * m;
    acc[A_VF] += fmaxf(vf1, vf2) * m;
    acc[A_PGCLIP] += (pg2 > pg1 ? 1.f : 0.f) * m;
    acc[A_VFCLIP] += (vf2 > vf1 ? 1.f : 0.f) * m;
    acc[A_KL] += (ratio - 1.f) - lr;
    acc[A_RATIO] += ratio * m;
    acc[A_VAL] += v * m;   acc[A_VAL2] += v * v * m;
    acc[A_OLD] += ov * m;  acc[A_OLD2] += ov * ov * m;
    acc[A_RET] += R * m;   acc[A_RET2] += R * R * m;
    acc[A_VERR] += (v - R) * (v - R) * m * m;
    acc[A_VMAPE] += fabsf(v - R) * m / fabsf(R * m + 1e-2f);
    if (m != 0.f) {
      mn[0] = fminf(mn[0], v); mxv[0] = fmaxf(mxv[0], v);
      mn[1] = fminf(mn[1], ov); mxv[1] = fmaxf(mxv[1], ov);
      mn[2] = fminf(mn[2], R); mxv[2] = fmaxf(mxv[2], R);
    }
    // un-normalised gradients (the finalize kernel / backward divides by n)
    dlogprobs[i] = (pg1 >= pg2) ? (-A * ratio * m * m) : 0.f;
    dvalues[i] = (vf1 >= vf2) ? (v - R) * m : 0.f;
  }
#pragma unroll
  for (int i = 0; i < A_COUNT; ++i) {
    const float r = block_sum(acc[i], sh);
    if (threadIdx.x == 0) part[blockIdx.x * A_COUNT + i] = r;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a = -block_max(-mn[i], sh), b = block_max(mxv[i], sh);
    if (threadIdx.x == 0) { ext[blockIdx.x * 6 + 2 * i] = a; ext[blockIdx.x * 6 + 2 * i + 1] = b; }
  }
}

// single block: reduce partials, emit the stats vector and scale the gradients by 1/n (policy) and vf_coef/n (value)
__global__ void __launch_bounds__(256)
ppo_loss_finalize_kernel(const float* __restrict__ part, const float* __restrict__ ext, int nblocks, int total_arg,
                         int rows, const int* __restrict__ width_ptr, float vf_coef, float* __restrict__ out) {
  // number of (row, position) cells the unmasked means (approx_kl, padding_percentage) are taken over
  const int total = width_ptr ? rows * (*width_ptr) : total_arg;
  __shared__ double acc[A_COUNT];
  __shared__ float mm[6];
  if (threadIdx.x < A_COUNT) {
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += part[b * A_COUNT + threadIdx.x];
    acc[threadIdx.x] = s;
  } else if (threadIdx.x >= 32 && threadIdx.x < 38) {
    const int k = threadIdx.x - 32;
    float r = (k & 1) ? -INFINITY : INFINITY;
    for (int b = 0; b < nblocks; ++b) r = (k & 1) ? fmaxf(r, ext[b * 6 + k]) : fminf(r, ext[b * 6 + k]);
    mm[k] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = acc[A_N] > 0 ? acc[A_N] : 1.0;  // an all-padding batch yields zero loss/gradients, not NaN
    const double pg = acc[A_PG] / n, vf = 0.5 * acc[A_VF] / n;
    out[O_LOSS] = (float)(pg + vf_coef * vf);
    out[O_PG] = (float)pg;
    out[O_VF] = (float)vf;
    auto stdv = [&](double s, double s2) { const double mu = s / n; double v = s2 / n - mu * mu; return (float)sqrt(v > 0 ? v : 0); };
    out[O_VAL_MEAN] = (float)(acc[A_VAL] / n); out[O_VAL_MIN] = mm[0]; out[O_VAL_MAX] = mm[1]; out[O_VAL_STD] = stdv(acc[A_VAL], acc[A_VAL2]);
    out[O_VERR] = (float)(acc[A_VERR] / n);
    out[O_VMAPE] = (float)(acc[A_VMAPE] / n);
    out[O_VFCLIP] = (float)(acc[A_VFCLIP] / n);
    out[O_OLD_MEAN] = (float)(acc[A_OLD] / n); out[O_OLD_MIN] = mm[2]; out[O_OLD_MAX] = mm[3]; out[O_OLD_STD] = stdv(acc[A_OLD], acc[A_OLD2]);
    out[O_RET_MEAN] = (float)(acc[A_RET] / n); out[O_RET_MIN] = mm[4]; out[O_RET_MAX] = mm[5]; out[O_RET_STD] = stdv(acc[A_RET], acc[A_RET2]);
    out[O_KL] = (float)(acc[A_KL] / total);
    out[O_PGCLIP] = (float)(acc[A_PGCLIP] / n);
    out[O_RATIO] = (float)(acc[A_RATIO] / n);
    out[O_PAD] = (float)(1.0 - n / total);
    out[O_N] = (float)n;
  }
}

__global__ void ppo_grad_scale_kernel(float* __restrict__ dlogprobs, float* __restrict__ dvalues, int total,
                                      const float* __restrict__ out, float vf_coef) {
  const float inv_n = 1.f / fmaxf(out[O_N], 1.f);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) { dlogprobs[i] *= inv_n; dvalues[i] *= vf_coef * inv_n; }
}

// ----------------------------------------------------------------------------- rollout post-processing
// Everything `make_experience` derives from a scored rollout, in one launch (reference: accelerate_ppo_trainer.py:455-504,
// a Python loop over samples; previously ~12 aten launches here).  One block per row.
//   lp / ref_lp / values : [B, T-1] over all positions;  mask : [B, T] (1 = real token);  start = Q - 1;  R = T - 1 - start
//   slice_len[b] = min(sum(mask[b, start:]) + 1, R)          scored response positions (the reference includes the EOS step)
//   rewards[b, c] = -kl_coef * (lp - ref_lp)[b, start + c] * mask   for c < slice_len, + score[b] at c = slice_len - 1
//   lp_out / v_out = lp / values on the same window, zero past slice_len
//   kl_sum        += sum over ALL positions of k3 = exp(d) - 1 - d,  d = (lp - ref_lp) * mask     (mean_kl statistics)
__global__ void __launch_bounds__(128)
rollout_rewards_kernel(const float* __restrict__ lp, const float* __restrict__ ref_lp, const float* __restrict__ values,
                       const long long* __restrict__ mask, const float* __restrict__ scores, int Tm1, int start, float kl_coef,
                       float* __restrict__ rewards, float* __restrict__ lp_out, float* __restrict__ v_out,
                       int* __restrict__ slice_len, double* __restrict__ kl_sum) {
  __shared__ float sh[4];
  __shared__ int sh_len;
  const int b = blockIdx.x, T = Tm1 + 1, R = Tm1 - start;
  const long long* mrow = mask + (size_t)b * T;
  float cnt = 0.f;
  for (int t = start + threadIdx.x; t < T; t += blockDim.x) cnt += mrow[t] ? 1.f : 0.f;
  cnt = block_sum(cnt, sh);
  if (threadIdx.x == 0) {
    sh_len = min((int)(cnt + 0.5f) + 1, R);
    slice_len[b] = sh_len;
  }
  __syncthreads();
  const int n = sh_len;
  float kl = 0.f;
  for (int t = threadIdx.x; t < Tm1; t += blockDim.x) {
    const float d = mrow[t] ? lp[(size_t)b * Tm1 + t] - ref_lp[(size_t)b * Tm1 + t] : 0.f;
    kl += __expf(d) - 1.f - d;
    const int c = t - start;
    if (c >= 0) {
      const bool valid = c < n;
      float r = valid ? -kl_coef * d : 0.f;
      if (c == n - 1) r += scores[b];
      rewards[(size_t)b * R + c] = r;
      lp_out[(size_t)b * R + c] = valid ? lp[(size_t)b * Tm1 + t] : 0.f;
      v_out[(size_t)b * R + c] = valid ? values[(size_t)b * Tm1 + t] : 0.f;
    }
  }
  kl = block_sum(kl, sh);
  if (threadIdx.x == 0) atomicAdd(kl_sum, (double)kl);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_rollout_rewards(const float* lp, const float* ref_lp, const float* values, const long long* mask,
                                    const float* scores, int B, int Tm1, int start, float kl_coef, float* rewards,
                                    float* lp_out, float* v_out, int* slice_len, double* kl_sum, cudaStream_t stream) {
  if (B <= 0 || Tm1 - start <= 0) return 0;
  rollout_rewards_kernel<<<B, 128, 0, stream>>>(lp, ref_lp, values, mask, scores, Tm1, start, kl_coef, rewards, lp_out, v_out,
                                                slice_len, kl_sum);
  return (int)cudaGetLastError();
}

// dtype: 0 = fp32, 1 = bf16, 2 = fp16
extern "C" int b200_logprob_from_logits(const void* logits, const long long* labels, float* out, float* lse, long long rows,
                                        int V, long long ld, int dtype, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (dtype == 0) logprob_kernel<float><<<rows, 256, 0, stream>>>((const float*)logits, labels, out, lse, V, ld);
  else if (dtype == 1) logprob_kernel<__nv_bfloat16><<<rows, 256, 0, stream>>>((const __nv_bfloat16*)logits, labels, out, lse, V, ld);
  else logprob_kernel<__half><<<rows, 256, 0, stream>>>((const __half*)logits, labels, out, lse, V, ld);
  return (int)cudaGetLastError();
}

extern "C" int b200_logprob_backward_inplace(void* logits, const long long* labels, const float* lse, const float* grad,
                                             long long rows, int V, long long ld, int dtype, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (dtype == 0) logprob_bwd_kernel<float><<<rows, 256, 0, stream>>>((float*)logits, labels, lse, grad, V, ld);
  else if (dtype == 1) logprob_bwd_kernel<__nv_bfloat16><<<rows, 256, 0, stream>>>((__nv_bfloat16*)logits, labels, lse, grad, V, ld);
  else logprob_bwd_kernel<__half><<<rows, 256, 0, stream>>>((__half*)logits, labels, lse, grad, V, ld);
  return (int)cudaGetLastError();
}

// stats must be zeroed by the caller (double[3]).
extern "C" int b200_gae(const float* values, const float* rewards, float* adv, float* ret, int B, int R, int width,
                        const int* width_ptr, long long ld, float gamma, float lam, double* stats, cudaStream_t stream) {
  if (B <= 0) return 0;
  gae_kernel<<<(B + 63) / 64, 64, 0, stream>>>(values, rewards, adv, ret, B, R, width, width_ptr, ld, gamma, lam, stats);
  return (int)cudaGetLastError();
}

extern "C" int b200_whiten(float* adv, int B, int width, const int* width_ptr, long long ld, const double* stats,
                           int unbiased, cudaStream_t stream) {
  const int total = B * (width_ptr ? (int)ld : width);  // launch for the widest case when the width is device-side
  if (total <= 0) return 0;
  whiten_kernel<<<(total + 255) / 256, 256, 0, stream>>>(adv, B, width, width_ptr, ld, stats, unbiased);
  return (int)cudaGetLastError();
}

extern "C" int b200_ppo_loss_num_outputs() { return O_COUNT; }
extern "C" int b200_ppo_loss_workspace_floats(int nblocks) { return nblocks * (A_COUNT + 6); }

// All tensors are contiguous fp32 with `total` elements.  workspace: nblocks * (A_COUNT + 6) floats.
extern "C" int b200_ppo_loss(const float* logprobs, const float* values, const float* old_logprobs, const float* old_values,
                             const float* adv, const float* ret, const float* mask, int total, float clip, float clip_v,
                             float vf_coef, float* dlogprobs, float* dvalues, float* workspace, int nblocks, float* out,
                             int rows, const int* width_ptr, cudaStream_t stream) {
  if (total <= 0) return 0;
  float* part = workspace;
  float* ext = workspace + (size_t)nblocks * A_COUNT;
  ppo_loss_partial_kernel<<<nblocks, 256, 0, stream>>>(logprobs, values, old_logprobs, old_values, adv, ret, mask, total, clip,
                                                       clip_v, dlogprobs, dvalues, part, ext);
  ppo_loss_finalize_kernel<<<1, 256, 0, stream>>>(part, ext, nblocks, total, rows, width_ptr, vf_coef, out);
  ppo_grad_scale_kernel<<<(total + 255) / 256, 256, 0, stream>>>(dlogprobs, dvalues, total, out, vf_coef);
  return (int)cudaGetLastError();
}

