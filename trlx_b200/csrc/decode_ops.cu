// Rollout / scoring support kernels (sm_100a): normalisation, embedding gather, paged-KV decode attention with fused
// rotary + cache append, value-head row-dot, and the per-step sequence bookkeeping that keeps the whole decode loop on
// the device (so it can live in one CUDA graph).  Replaces the per-token Python loop + dozens of aten launches of HF
// generate() used by the reference (trlx/trainer/accelerate_base_trainer.py:256-269).
#include <cuda_fp8.h>

#include "ptx.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------ norms
// One warp per row when H <= 2048 would do, but a 128-thread block per row handles any H with 16-byte loads.
template <bool RMS>
__global__ void __launch_bounds__(128) norm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                   const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                   int H, long long ldx, long long ldy, float eps) {
  griddep_wait();
  griddep_launch();
  const int row = blockIdx.x;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  __nv_bfloat16* yr = y + (size_t)row * ldy;
  __shared__ float red[2][4];
  float s = 0.f, ss = 0.f;
  const int nvec = H >> 3;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 v = *reinterpret_cast<const uint4*>(xr + i * 8);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { float f = __bfloat162float(h[j]); s += f; ss += f * f; }
  }
  for (int i = (nvec << 3) + threadIdx.x; i < H; i += blockDim.x) { float f = __bfloat162float(xr[i]); s += f; ss += f * f; }
  s = warp_sum(s); ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  s = red[0][0] + red[0][1] + red[0][2] + red[0][3];
  ss = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  const float mean = RMS ? 0.f : s / H;
  const float var = RMS ? ss / H : fmaxf(ss / H - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 v = *reinterpret_cast<const uint4*>(xr + i * 8);
    uint4 wv = *reinterpret_cast<const uint4*>(w + i * 8);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
    const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(&wv);
    uint4 o;
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(&o);
    if (b) {
      uint4 bv = *reinterpret_cast<const uint4*>(b + i * 8);
      const __nv_bfloat16* bh = reinterpret_cast<const __nv_bfloat16*>(&bv);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        oh[j] = __float2bfloat16((__bfloat162float(h[j]) - mean) * rstd * __bfloat162float(wh[j]) + __bfloat162float(bh[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) oh[j] = __float2bfloat16((__bfloat162float(h[j]) - mean) * rstd * __bfloat162float(wh[j]));
    }
    *reinterpret_cast<uint4*>(yr + i * 8) = o;
  }
  for (int i = (nvec << 3) + threadIdx.x; i < H; i += blockDim.x) {
    float f = (__bfloat162float(xr[i]) - mean) * rstd * __bfloat162float(w[i]);
    if (b) f += __bfloat162float(b[i]);
    yr[i] = __float2bfloat16(f);
  }
}

// LayerNorm / RMSNorm fused with per-row e4m3 quantisation (rollout_dtype = fp8): y8[row, :] = e4m3(norm(x) / scale[row]),
// scale[row] = max|norm(x)| / 448.  Three passes over a row that stays in L1/L2: moments, amax, quantise.
template <bool RMS>
__global__ void __launch_bounds__(128) norm_quant_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                         const __nv_bfloat16* __restrict__ b, uint8_t* __restrict__ y8,
                                                         float* __restrict__ scale_out, int H, long long ldx, long long ldy,
                                                         float eps) {
  griddep_wait();
  griddep_launch();
  const int row = blockIdx.x;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  uint8_t* yr = y8 + (size_t)row * ldy;
  __shared__ float red[3][4];
  float s = 0.f, ss = 0.f;
  const int nvec = H >> 3;  // H % 16 == 0 is required by the fp8 GEMM
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 v = *reinterpret_cast<const uint4*>(xr + i * 8);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { float f = __bfloat162float(h[j]); s += f; ss += f * f; }
  }
  s = warp_sum(s); ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  s = red[0][0] + red[0][1] + red[0][2] + red[0][3];
  ss = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  const float mean = RMS ? 0.f : s / H;
  const float var = RMS ? ss / H : fmaxf(ss / H - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  auto normed = [&](int i, float (&o)[8]) {
    uint4 v = *reinterpret_cast<const uint4*>(xr + i * 8);
    uint4 wv = *reinterpret_cast<const uint4*>(w + i * 8);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
    const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(&wv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (__bfloat162float(h[j]) - mean) * rstd * __bfloat162float(wh[j]);
    if (b) {
      uint4 bv = *reinterpret_cast<const uint4*>(b + i * 8);
      const __nv_bfloat16* bh = reinterpret_cast<const __nv_bfloat16*>(&bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += __bfloat162float(bh[j]);
    }
  };
  float amax = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float o[8];
    normed(i, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(o[j]));
  }
  amax = warp_max(amax);
  if ((threadIdx.x & 31) == 0) red[2][threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = fmaxf(fmaxf(red[2][0], red[2][1]), fmaxf(red[2][2], red[2][3]));
  const float scale = fmaxf(amax, 1e-12f) / 448.f, inv = 1.f / scale;
  if (threadIdx.x == 0) scale_out[row] = scale;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float o[8];
    normed(i, o);
    uint2 q;
    uint8_t* qb = reinterpret_cast<uint8_t*>(&q);
#pragma unroll
    for (int j = 0; j < 8; ++j) qb[j] = (uint8_t)__nv_cvt_float_to_fp8(o[j] * inv, __NV_SATFINITE, __NV_E4M3);
    *reinterpret_cast<uint2*>(yr + i * 8) = q;
  }
}

// ------------------------------------------------------------------------------------------------ embedding
// x[b,:] = wte[token[b]] + wpe[pos[b] + pos_offset]   (wpe optional)
__global__ void embed_kernel(const long long* __restrict__ tokens, const int* __restrict__ positions,
                             const __nv_bfloat16* __restrict__ wte, const __nv_bfloat16* __restrict__ wpe, int pos_offset,
                             __nv_bfloat16* __restrict__ x, int H, float* __restrict__ stats_out) {
  griddep_wait();
  griddep_launch();
  const int b = blockIdx.x;
  const __nv_bfloat16* te = wte + (size_t)tokens[b] * H;
  const __nv_bfloat16* pe = wpe ? wpe + (size_t)(positions[b] + pos_offset) * H : nullptr;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < (H >> 3); i += blockDim.x) {
    uint4 a = *reinterpret_cast<const uint4*>(te + i * 8);
    __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&a);
    if (pe) {
      uint4 p = *reinterpret_cast<const uint4*>(pe + i * 8);
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&p);
#pragma unroll
      for (int j = 0; j < 4; ++j) a2[j] = __hadd2(a2[j], p2[j]);
    }
    if (stats_out) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(a2[j]);
        s1 += f.x + f.y;
        s2 += f.x * f.x + f.y * f.y;
      }
    }
    *reinterpret_cast<uint4*>(x + (size_t)b * H + i * 8) = a;
  }
  if (stats_out) {  // row moments for the first block's folded LayerNorm (see gemm_sm100.cu: StoreEpilogue::ln_stats)
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(stats_out + 2 * (size_t)b, s1);
      atomicAdd(stats_out + 2 * (size_t)b + 1, s2);
    }
  }
}

// ------------------------------------------------------------------------------------------------ decode attention
// One block per (sequence, query head); 4 warps; each 8-lane group owns one key at a time (16-byte loads).
// qkv       : [B, (nq + 2 nkv) * d]   current-token projections (bias already added by the GEMM epilogue)
// k/v cache : [num_pages, page_size, nkv, d]
// The block for query head h also appends K/V of kv-head (h / group) when h % group == 0, after applying rotary.
// seq_lens[b] is the number of valid keys INCLUDING the current token; rows with active[b] == 0 are skipped.
template <int MAX_D>
__global__ void __launch_bounds__(128)
decode_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ kcache, __nv_bfloat16* __restrict__ vcache,
                   const int* __restrict__ block_table, const int* __restrict__ seq_lens, const int* __restrict__ positions,
                   __nv_bfloat16* __restrict__ out, int nq, int nkv, int d, int page_size, int max_pages, float scale,
                   int rot_dim, float rot_base, int rot_interleaved, const float* __restrict__ alibi_slopes, int window) {
  griddep_wait();
  griddep_launch();
  const int b = blockIdx.x, h = blockIdx.y;
  const int group = nq / nkv, kvh = h / group;
  const int len = seq_lens[b];
  if (len <= 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane >> 3, l8 = lane & 7;       // 4 key-slots per warp, 8 lanes per key
  const int row_stride = (nq + 2 * nkv) * d;
  const __nv_bfloat16* qp = qkv + (size_t)b * row_stride + (size_t)h * d;
  const __nv_bfloat16* kp = qkv + (size_t)b * row_stride + (size_t)(nq + kvh) * d;
  const __nv_bfloat16* vp = qkv + (size_t)b * row_stride + (size_t)(nq + nkv + kvh) * d;
  const int* bt = block_table + (size_t)b * max_pages;
  const int pos = positions ? positions[b] : len - 1;

  __shared__ float q_s[MAX_D];
  __shared__ float knew_s[MAX_D];
  __shared__ float red_m[4 * 4], red_l[4 * 4];
  __shared__ float acc_s[16][MAX_D + 4];

  // --- load q (and the new k), apply rotary
  for (int i = tid; i < d; i += blockDim.x) { q_s[i] = __bfloat162float(qp[i]); knew_s[i] = __bfloat162float(kp[i]); }
  __syncthreads();
  if (rot_dim > 0) {
    const int half = rot_dim >> 1;
    float qn = 0.f, kn = 0.f;
    int i0 = -1, i1 = -1;
    if (tid < half) {
      const float inv = powf(rot_base, -(float)tid / (float)half);
      float sn, cs;
      sincosf((float)pos * inv, &sn, &cs);
      i0 = rot_interleaved ? 2 * tid : tid;
      i1 = rot_interleaved ? 2 * tid + 1 : tid + half;
      const float q0 = q_s[i0], q1 = q_s[i1], k0 = knew_s[i0], k1 = knew_s[i1];
      qn = q0 * cs - q1 * sn; kn = k0 * cs - k1 * sn;
      const float qn1 = q1 * cs + q0 * sn, kn1 = k1 * cs + k0 * sn;
      q_s[i0] = qn; q_s[i1] = qn1; knew_s[i0] = kn; knew_s[i1] = kn1;
    }
    __syncthreads();
  }
  // --- append the new K/V to the paged cache (one query head per kv head does it)
  const int last = len - 1;
  if (h % group == 0) {
    const size_t slot = ((size_t)bt[last / page_size] * page_size + (last % page_size)) * nkv + kvh;
    for (int i = tid; i < d; i += blockDim.x) {
      kcache[slot * d + i] = __float2bfloat16(knew_s[i]);
      vcache[slot * d + i] = vp[i];
    }
  }
  // --- online-softmax over cached keys [lo, last) plus the new key from shared memory
  const int lo = (window > 0 && len > window) ? len - window : 0;
  const float slope = alibi_slopes ? alibi_slopes[h] : 0.f;
  float m = -INFINITY, l = 0.f;
  float acc[MAX_D / 8];
#pragma unroll
  for (int i = 0; i < MAX_D / 8; ++i) acc[i] = 0.f;
  // lane (sub, l8) accumulates output dims { l8*8 + 64*c + j } for its keys
  // NOTE: the trip count is warp-uniform (`base`), lanes whose key index falls past `len` stay in the loop so the
  // full-mask shuffles below are executed by every lane.
  for (int base_t = lo + warp * 4; base_t < len; base_t += 16) {
    const int t = base_t + sub;
    const bool valid = t < len;
    float part = 0.f;
    const __nv_bfloat16* kptr = nullptr;
    const __nv_bfloat16* vptr = nullptr;
    if (valid && t < last) {
      const size_t slot = ((size_t)bt[t / page_size] * page_size + (t % page_size)) * nkv + kvh;
      kptr = kcache + slot * d;
      vptr = vcache + slot * d;
#pragma unroll
      for (int c = 0; c < MAX_D / 64; ++c) {
        const int base = l8 * 8 + 64 * c;
        if (base < d) {
          uint4 kv = *reinterpret_cast<const uint4*>(kptr + base);
          const __nv_bfloat16* kh = reinterpret_cast<const __nv_bfloat16*>(&kv);
#pragma unroll
          for (int j = 0; j < 8; ++j) part += q_s[base + j] * __bfloat162float(kh[j]);
        }
      }
    } else if (valid) {
#pragma unroll
      for (int c = 0; c < MAX_D / 64; ++c) {
        const int base = l8 * 8 + 64 * c;
        if (base < d) {
#pragma unroll
          for (int j = 0; j < 8; ++j) part += q_s[base + j] * knew_s[base + j];
        }
      }
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (!valid) continue;
    float sc = part * scale + slope * (float)t;
    const float mn = fmaxf(m, sc);
    const float corr = __expf(m - mn), p = __expf(sc - mn);
    l = l * corr + p;
    m = mn;
#pragma unroll
    for (int c = 0; c < MAX_D / 64; ++c) {
      const int base = l8 * 8 + 64 * c;
      if (base < d) {
        if (t < last) {
          uint4 vv = *reinterpret_cast<const uint4*>(vptr + base);
          const __nv_bfloat16* vh = reinterpret_cast<const __nv_bfloat16*>(&vv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c * 8 + j] = acc[c * 8 + j] * corr + p * __bfloat162float(vh[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c * 8 + j] = acc[c * 8 + j] * corr + p * __bfloat162float(vp[base + j]);
        }
      }
    }
  }
  // --- combine the 16 (warp, sub) partial softmaxes
  const int g = warp * 4 + sub;
  if (l8 == 0) { red_m[g] = m; red_l[g] = l; }
  __syncthreads();
  float gm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) gm = fmaxf(gm, red_m[i]);
  float gl = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) gl += red_l[i] * ((red_m[i] == -INFINITY) ? 0.f : __expf(red_m[i] - gm));
  const float my = (m == -INFINITY) ? 0.f : __expf(m - gm);
#pragma unroll
  for (int c = 0; c < MAX_D / 64; ++c) {
    const int base = l8 * 8 + 64 * c;
    if (base < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_s[g][base + j] = acc[c * 8 + j] * my;
    }
  }
  __syncthreads();
  const float inv = 1.f / gl;
  for (int i = tid; i < d; i += blockDim.x) {
    float o = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) o += acc_s[k][i];
    out[(size_t)b * nq * d + (size_t)h * d + i] = __float2bfloat16(o * inv);
  }
}

// ------------------------------------------------------------------------------------------------ value head tail
// out[m] = dot(x[m,:], w) + bias    (the N=1 projection of the value MLP; fp32 result)
__global__ void rowdot_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                              const __nv_bfloat16* __restrict__ bias, float* __restrict__ out, int M, int K, long long ldx) {
  griddep_wait();
  griddep_launch();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  float s = 0.f;
  for (int i = lane * 8; i + 8 <= K; i += 256) {
    uint4 a = *reinterpret_cast<const uint4*>(xr + i), b = *reinterpret_cast<const uint4*>(w + i);
    const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(&a);
    const __nv_bfloat16* bh = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __bfloat162float(ah[j]) * __bfloat162float(bh[j]);
  }
  for (int i = (K & ~7) + lane; i < K; i += 32) s += __bfloat162float(xr[i]) * __bfloat162float(w[i]);
  s = warp_sum(s);
  if (lane == 0) out[row] = s + (bias ? __bfloat162float(bias[0]) : 0.f);
}

// ------------------------------------------------------------------------------------------------ step bookkeeping
// After sampling at decode step `*step_ptr`: record outputs for still-running rows, retire rows that emitted EOS or
// hit their budget, set up the next step's inputs and bump the device-side step counter.  Everything stays on the
// device, so ONE captured CUDA graph serves every decode step.  Launched as a single block.
//   tokens_out/logprobs_out/ref_logprobs_out/values_out : [B, max_new]
//   finished[b] : 0/1 ; seq_lens/positions advanced for rows that keep running ; n_running counts live rows.
__global__ void decode_step_kernel(const long long* __restrict__ sampled, const float* __restrict__ lp,
                                   const float* __restrict__ ref_lp, const float* __restrict__ value, int* __restrict__ step_ptr,
                                   int max_new, int B, long long eos_id, long long pad_id, long long* __restrict__ tokens_out,
                                   float* __restrict__ logprobs_out, float* __restrict__ ref_logprobs_out,
                                   float* __restrict__ values_out, int* __restrict__ finished, int* __restrict__ resp_lens,
                                   int* __restrict__ seq_lens, int* __restrict__ positions, long long* __restrict__ next_tokens,
                                   int* __restrict__ n_running) {
  griddep_wait();
  griddep_launch();
  const int step = *step_ptr;
  int retired = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    if (step >= max_new) break;
    const size_t o = (size_t)b * max_new + step;
    if (finished[b]) {
      tokens_out[o] = pad_id;
      logprobs_out[o] = 0.f;
      if (ref_logprobs_out) ref_logprobs_out[o] = 0.f;
      if (values_out) values_out[o] = 0.f;
      next_tokens[b] = pad_id;
      continue;
    }
    const long long tok = sampled[b];
    tokens_out[o] = tok;
    logprobs_out[o] = lp[b];
    if (ref_logprobs_out) ref_logprobs_out[o] = ref_lp ? ref_lp[b] : 0.f;
    if (values_out) values_out[o] = value ? value[b] : 0.f;
    resp_lens[b] = step + 1;
    const bool done = (tok == eos_id) || (step + 1 >= max_new);
    if (done) {
      finished[b] = 1;
      seq_lens[b] = 0;  // attention blocks for retired rows exit immediately
      ++retired;
    } else {
      seq_lens[b] += 1;
      positions[b] += 1;
    }
    next_tokens[b] = tok;
  }
  if (retired) atomicSub(n_running, retired);
  __syncthreads();
  if (threadIdx.x == 0) *step_ptr = step + 1;
}

// Scatter prefill K/V ([B, T, nkv, d], left- or right-padded) into the paged cache: token t of row b (valid when
// first[b] <= t < first[b] + lens[b]) goes to logical slot t - first[b].
__global__ void paged_kv_write_kernel(const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                                      __nv_bfloat16* __restrict__ kcache, __nv_bfloat16* __restrict__ vcache,
                                      const int* __restrict__ block_table, const int* __restrict__ first,
                                      const int* __restrict__ lens, int T, int nkv, int d, int page_size, int max_pages,
                                      long long k_stride_b, long long k_stride_t) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int s = t - first[b];
  if (s < 0 || s >= lens[b]) return;
  const size_t slot = (size_t)block_table[(size_t)b * max_pages + s / page_size] * page_size + (s % page_size);
  const int row = nkv * d;
  const __nv_bfloat16* ks = k + (size_t)b * k_stride_b + (size_t)t * k_stride_t;
  const __nv_bfloat16* vs = v + (size_t)b * k_stride_b + (size_t)t * k_stride_t;
  for (int i = threadIdx.x * 8; i < row; i += blockDim.x * 8) {
    *reinterpret_cast<uint4*>(kcache + slot * row + i) = *reinterpret_cast<const uint4*>(ks + i);
    *reinterpret_cast<uint4*>(vcache + slot * row + i) = *reinterpret_cast<const uint4*>(vs + i);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_norm_bf16(const void* x, const void* w, const void* b, void* y, int rows, int H, long long ldx,
                              long long ldy, float eps, int rms, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (rms)
    return (int)launch_kernel(norm_kernel<true>, dim3(rows), dim3(128), 0, stream, (const __nv_bfloat16*)x,
                              (const __nv_bfloat16*)w, (const __nv_bfloat16*)nullptr, (__nv_bfloat16*)y, H, ldx, ldy, eps);
  return (int)launch_kernel(norm_kernel<false>, dim3(rows), dim3(128), 0, stream, (const __nv_bfloat16*)x,
                            (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, H, ldx, ldy, eps);
}

// Per-row e4m3 quantisation without a norm (fp8 rollouts: the attention output feeding the out-projection and the MLP activation
// feeding the down-projection): y8[row, :] = e4m3(x / scale[row]), scale[row] = max|x| / 448.  One CTA per row, two passes (the row
// stays in L1 between them).
__global__ void __launch_bounds__(128) quant_rows_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ y8,
                                                         float* __restrict__ scale_out, int H, long long ldx, long long ldy) {
  griddep_wait();
  griddep_launch();
  const int row = blockIdx.x;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  uint8_t* yr = y8 + (size_t)row * ldy;
  __shared__ float red[4];
  const int nvec = H >> 3;
  float amax = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i * 8);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(__bfloat162float(h[j])));
  }
  amax = warp_max(amax);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  const float scale = fmaxf(amax, 1e-12f) / 448.f, inv = 1.f / scale;
  if (threadIdx.x == 0) scale_out[row] = scale;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i * 8);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
    uint2 q;
    uint8_t* qb = reinterpret_cast<uint8_t*>(&q);
#pragma unroll
    for (int j = 0; j < 8; ++j) qb[j] = (uint8_t)__nv_cvt_float_to_fp8(__bfloat162float(h[j]) * inv, __NV_SATFINITE, __NV_E4M3);
    *reinterpret_cast<uint2*>(yr + i * 8) = q;
  }
}

extern "C" int b200_quant_rows_fp8(const void* x, void* y8, float* scale, int rows, int H, long long ldx, long long ldy,
                                   cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (H % 16) return -2;
  return (int)launch_kernel(quant_rows_kernel, dim3(rows), dim3(128), 0, stream, (const __nv_bfloat16*)x, (uint8_t*)y8, scale, H, ldx,
                            ldy);
}

extern "C" int b200_norm_quant_fp8(const void* x, const void* w, const void* b, void* y8, float* scale, int rows, int H,
                                   long long ldx, long long ldy, float eps, int rms, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (H % 16) return -2;
  if (rms)
    return (int)launch_kernel(norm_quant_kernel<true>, dim3(rows), dim3(128), 0, stream, (const __nv_bfloat16*)x,
                              (const __nv_bfloat16*)w, (const __nv_bfloat16*)nullptr, (uint8_t*)y8, scale, H, ldx, ldy, eps);
  return (int)launch_kernel(norm_quant_kernel<false>, dim3(rows), dim3(128), 0, stream, (const __nv_bfloat16*)x,
                            (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, (uint8_t*)y8, scale, H, ldx, ldy, eps);
}

extern "C" int b200_embed_bf16(const long long* tokens, const int* positions, const void* wte, const void* wpe,
                               int pos_offset, void* x, int B, int H, float* stats_out, cudaStream_t stream) {
  if (B <= 0) return 0;
  return (int)launch_kernel(embed_kernel, dim3(B), dim3(128), 0, stream, tokens, positions, (const __nv_bfloat16*)wte,
                            (const __nv_bfloat16*)wpe, pos_offset, (__nv_bfloat16*)x, H, stats_out);
}

extern "C" int b200_decode_attention_bf16(const void* qkv, void* kcache, void* vcache, const int* block_table,
                                          const int* seq_lens, const int* positions, void* out, int B, int nq, int nkv, int d,
                                          int page_size, int max_pages, float scale, int rot_dim, float rot_base,
                                          int rot_interleaved, const float* alibi_slopes, int window, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (d % 8 != 0 || d > 256) return -2;
  dim3 grid(B, nq);
#define LAUNCH(MD)                                                                                                        \
  return (int)launch_kernel(decode_attn_kernel<MD>, grid, dim3(128), 0, stream, (const __nv_bfloat16*)qkv,                 \
                            (__nv_bfloat16*)kcache, (__nv_bfloat16*)vcache, block_table, seq_lens, positions,               \
                            (__nv_bfloat16*)out, nq, nkv, d, page_size, max_pages, scale, rot_dim, rot_base, rot_interleaved, \
                            alibi_slopes, window)
  if (d <= 64) { LAUNCH(64); }
  else if (d <= 128) { LAUNCH(128); }
  else { LAUNCH(256); }
#undef LAUNCH
}

// ------------------------------------------------------------------------------------------------ top-k / top-p sampling
// HF `generate(do_sample=True, temperature, top_k, top_p)` semantics on one row of fp32 logits per block:
//   temperature -> top-k (keep everything >= the k-th largest, ties included) -> top-p (smallest prefix of the sorted
//   distribution whose mass reaches top_p; the crossing token is kept) -> multinomial draw (Gumbel-max over the kept set).
// The thresholds are found EXACTLY by 4-pass radix selection on the order-preserving integer image of the logits
// (256-bin histograms of counts for top-k, of probability mass for top-p) instead of sorting 50k entries.
// Also returns the RAW log-probability of the drawn token (temperature 1, no filtering): what PPO scores.
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float samp_uniform(unsigned long long seed, unsigned int row, unsigned int col) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (((unsigned long long)row << 32) | col);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return ((float)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(1024)
sample_filtered_kernel(const float* __restrict__ logits, long long ld, int V, int top_k, float top_p, float inv_temp,
                       unsigned long long seed, const long long* __restrict__ seed_ptr, const int* __restrict__ step_ptr,
                       int suppress_col, int suppress_until, long long* __restrict__ tok_out, float* __restrict__ lp_out) {
  griddep_wait();
  griddep_launch();
  __shared__ float red_a[32], red_b[32];
  __shared__ int red_i[32];
  __shared__ float hist_m[256];
  __shared__ int hist_c[256];
  __shared__ uint32_t sh_prefix;
  __shared__ float sh_above;
  __shared__ int sh_kleft;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const float* x = logits + (size_t)row * ld;
  const int step = step_ptr ? *step_ptr : 0;
  const int sup = (suppress_col >= 0 && step < suppress_until) ? suppress_col : -1;
  const unsigned long long sd = seed + (seed_ptr ? (unsigned long long)(*seed_ptr) : 0ull) + 0x632BE59BD9B4E019ull * (unsigned long long)(step + 1);
  auto at = [&](int i) { return i == sup ? -INFINITY : x[i]; };
  auto bsum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red_a[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < nw; ++i) r += red_a[i];
    return r;
  };
  // ---- pass 1: max, then raw and tempered partition functions
  float mx = -INFINITY;
  for (int i = tid; i < V; i += blockDim.x) mx = fmaxf(mx, at(i));
  mx = warp_max(mx);
  if (lane == 0) red_b[warp] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < nw; ++i) mx = fmaxf(mx, red_b[i]);
  float z1 = 0.f;
  for (int i = tid; i < V; i += blockDim.x) z1 += __expf(at(i) - mx);
  z1 = bsum(z1);
  const float lse_raw = mx + __logf(z1);

  // ---- top-k: key of the k-th largest logit
  uint32_t kth = 0u;  // keep keys >= kth
  if (top_k > 0 && top_k < V) {
    if (tid == 0) { sh_prefix = 0u; sh_kleft = top_k; }
    for (int level = 0; level < 4; ++level) {
      const int shift = 24 - 8 * level;
      if (tid < 256) hist_c[tid] = 0;
      __syncthreads();
      const uint32_t prefix = sh_prefix;
      const uint32_t pmask = level == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
      for (int i = tid; i < V; i += blockDim.x) {
        const uint32_t key = float_key(at(i));
        if ((key & pmask) == prefix) atomicAdd(&hist_c[(key >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int left = sh_kleft, b = 255;
        for (; b > 0; --b) {
          if (hist_c[b] >= left) break;
          left -= hist_c[b];
        }
        sh_kleft = left;
        sh_prefix = prefix | ((uint32_t)b << shift);
      }
      __syncthreads();
    }
    kth = sh_prefix;
  }
  // ---- tempered mass of the kept set
  float zt = 0.f;
  for (int i = tid; i < V; i += blockDim.x) {
    const float v = at(i);
    if (float_key(v) >= kth) zt += __expf((v - mx) * inv_temp);
  }
  zt = bsum(zt);
  // ---- top-p: key of the token at which the sorted cumulative mass first reaches top_p
  uint32_t pth = kth;
  if (top_p < 1.f && top_p > 0.f) {
    const float target = top_p * zt;
    if (tid == 0) { sh_prefix = 0u; sh_above = 0.f; }
    for (int level = 0; level < 4; ++level) {
      const int shift = 24 - 8 * level;
      if (tid < 256) hist_m[tid] = 0.f;
      __syncthreads();
      const uint32_t prefix = sh_prefix;
      const uint32_t pmask = level == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
      for (int i = tid; i < V; i += blockDim.x) {
        const float v = at(i);
        const uint32_t key = float_key(v);
        if (key >= kth && (key & pmask) == prefix) atomicAdd(&hist_m[(key >> shift) & 255u], __expf((v - mx) * inv_temp));
      }
      __syncthreads();
      if (tid == 0) {
        float above = sh_above;
        int b = 255;
        for (; b > 0; --b) {
          if (above + hist_m[b] >= target) break;
          above += hist_m[b];
        }
        sh_above = above;
        sh_prefix = prefix | ((uint32_t)b << shift);
      }
      __syncthreads();
    }
    pth = sh_prefix > kth ? sh_prefix : kth;
  }
  // ---- draw: Gumbel-max over the kept set (inv_temp <= 0: greedy)
  float best = -INFINITY, best_logit = 0.f;
  int best_i = 0x7fffffff;
  for (int i = tid; i < V; i += blockDim.x) {
    const float v = at(i);
    if (float_key(v) < pth || v == -INFINITY) continue;
    float key = v;
    if (inv_temp > 0.f) key = v * inv_temp - __logf(-__logf(samp_uniform(sd, (unsigned)row, (unsigned)i)));
    if (key > best || (key == best && i < best_i)) { best = key; best_i = i; best_logit = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o), ol = __shfl_xor_sync(0xffffffffu, best_logit, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; best_logit = ol; }
  }
  __syncthreads();
  if (lane == 0) { red_a[warp] = best; red_b[warp] = best_logit; red_i[warp] = best_i; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < nw; ++w)
      if (red_a[w] > best || (red_a[w] == best && red_i[w] < best_i)) { best = red_a[w]; best_i = red_i[w]; best_logit = red_b[w]; }
    tok_out[row] = best_i;
    lp_out[row] = best_logit - lse_raw;
  }
}

extern "C" int b200_sample_filtered(const float* logits, long long ld, int B, int V, int top_k, float top_p, float temperature,
                                    unsigned long long seed, const long long* seed_ptr, const int* step_ptr, int suppress_col,
                                    int suppress_until, long long* tok_out, float* lp_out, cudaStream_t stream) {
  if (B <= 0) return 0;
  const float inv_temp = temperature > 0.f ? 1.f / temperature : 0.f;
  return (int)launch_kernel(sample_filtered_kernel, dim3(B), dim3(1024), 0, stream, logits, ld, V, top_k, top_p, inv_temp, seed,
                            seed_ptr, step_ptr, suppress_col, suppress_until, tok_out, lp_out);
}

// ------------------------------------------------------------------------------------------------ ILQL advantage-shifted sampling
// One decode step of the reference's ILQL sampler (trlx/models/modeling_ilql.py:360-412) on one row per block:
//   pi_beta = log_softmax(mask(logits));  score = pi_beta + beta * (min(q1, q2) - v);  keep the top-k scores;
//   token ~ softmax(score / T)   (T == 0: argmax).   logits / q1 / q2: fp32 [B, >= V] with row pitch ld; q2 may be null.
// logit_mask: optional uint8 table [mask_rows, mask_cols]; entry [last_token, col] != 0 forbids `col` after `last_token`.
__global__ void __launch_bounds__(1024)
ilql_sample_kernel(const float* __restrict__ logits, const float* __restrict__ q1, const float* __restrict__ q2,
                   const float* __restrict__ vs, long long ld, int V, float beta, int top_k, float inv_temp,
                   unsigned long long seed, const long long* __restrict__ seed_ptr, const int* __restrict__ step_ptr,
                   const unsigned char* __restrict__ logit_mask, int mask_rows, int mask_cols,
                   const long long* __restrict__ last_tokens, long long* __restrict__ tok_out) {
  griddep_wait();
  griddep_launch();
  __shared__ float red_a[32], red_b[32];
  __shared__ int red_i[32];
  __shared__ int hist_c[256];
  __shared__ uint32_t sh_prefix;
  __shared__ int sh_kleft;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const float* x = logits + (size_t)row * ld;
  const float* a1 = q1 + (size_t)row * ld;
  const float* a2 = q2 ? q2 + (size_t)row * ld : nullptr;
  const float v = vs[row];
  const int step = step_ptr ? *step_ptr : 0;
  const unsigned long long sd = seed + (seed_ptr ? (unsigned long long)(*seed_ptr) : 0ull) + 0x632BE59BD9B4E019ull * (unsigned long long)(step + 1);
  const unsigned char* mrow = nullptr;
  if (logit_mask) {
    const long long last = last_tokens[row];
    if (last >= 0 && last < mask_rows) mrow = logit_mask + (size_t)last * mask_cols;
  }
  auto lg = [&](int i) { return (mrow && i < mask_cols && mrow[i]) ? -INFINITY : x[i]; };
  // ---- log-softmax normaliser of the (masked) behaviour policy
  float mx = -INFINITY;
  for (int i = tid; i < V; i += blockDim.x) mx = fmaxf(mx, lg(i));
  mx = warp_max(mx);
  if (lane == 0) red_b[warp] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < nw; ++i) mx = fmaxf(mx, red_b[i]);
  float z = 0.f;
  for (int i = tid; i < V; i += blockDim.x) z += __expf(lg(i) - mx);
  z = warp_sum(z);
  __syncthreads();
  if (lane == 0) red_a[warp] = z;
  __syncthreads();
  z = 0.f;
  for (int i = 0; i < nw; ++i) z += red_a[i];
  const float lse = mx + __logf(z);
  auto score = [&](int i) {
    const float q = a2 ? fminf(a1[i], a2[i]) : a1[i];
    return (lg(i) - lse) + beta * (q - v);
  };
  // ---- top-k on the shifted scores (exact: 4-pass radix selection of the k-th largest key)
  uint32_t kth = 0u;
  if (top_k > 0 && top_k < V) {
    if (tid == 0) { sh_prefix = 0u; sh_kleft = top_k; }
    for (int level = 0; level < 4; ++level) {
      const int shift = 24 - 8 * level;
      if (tid < 256) hist_c[tid] = 0;
      __syncthreads();
      const uint32_t prefix = sh_prefix;
      const uint32_t pmask = level == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
      for (int i = tid; i < V; i += blockDim.x) {
        const uint32_t key = float_key(score(i));
        if ((key & pmask) == prefix) atomicAdd(&hist_c[(key >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int left = sh_kleft, b = 255;
        for (; b > 0; --b) {
          if (hist_c[b] >= left) break;
          left -= hist_c[b];
        }
        sh_kleft = left;
        sh_prefix = prefix | ((uint32_t)b << shift);
      }
      __syncthreads();
    }
    kth = sh_prefix;
  }
  // ---- draw
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = tid; i < V; i += blockDim.x) {
    const float sc = score(i);
    if (float_key(sc) < kth || sc == -INFINITY) continue;
    float key = sc;
    if (inv_temp > 0.f) key = sc * inv_temp - __logf(-__logf(samp_uniform(sd, (unsigned)row, (unsigned)i)));
    if (key > best || (key == best && i < best_i)) { best = key; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  __syncthreads();
  if (lane == 0) { red_a[warp] = best; red_i[warp] = best_i; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < nw; ++w)
      if (red_a[w] > best || (red_a[w] == best && red_i[w] < best_i)) { best = red_a[w]; best_i = red_i[w]; }
    tok_out[row] = best_i == 0x7fffffff ? 0 : best_i;
  }
}

extern "C" int b200_ilql_sample(const float* logits, const float* q1, const float* q2, const float* vs, long long ld, int B, int V,
                                float beta, int top_k, float temperature, unsigned long long seed, const long long* seed_ptr,
                                const int* step_ptr, const unsigned char* logit_mask, int mask_rows, int mask_cols,
                                const long long* last_tokens, long long* tok_out, cudaStream_t stream) {
  if (B <= 0) return 0;
  const float inv_temp = temperature > 0.f ? 1.f / temperature : 0.f;
  return (int)launch_kernel(ilql_sample_kernel, dim3(B), dim3(1024), 0, stream, logits, q1, q2, vs, ld, V, beta, top_k, inv_temp,
                            seed, seed_ptr, step_ptr, logit_mask, mask_rows, mask_cols, last_tokens, tok_out);
}

extern "C" int b200_rowdot_bf16(const void* x, const void* w, const void* bias, float* out, int M, int K, long long ldx,
                                cudaStream_t stream) {
  if (M <= 0) return 0;
  return (int)launch_kernel(rowdot_kernel, dim3((M + 3) / 4), dim3(128), 0, stream, (const __nv_bfloat16*)x,
                            (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, out, M, K, ldx);
}

extern "C" int b200_decode_step(const long long* sampled, const float* lp, const float* ref_lp, const float* value,
                                int* step_ptr, int max_new, int B, long long eos_id, long long pad_id, long long* tokens_out,
                                float* logprobs_out, float* ref_logprobs_out, float* values_out, int* finished, int* resp_lens,
                                int* seq_lens, int* positions, long long* next_tokens, int* n_running, cudaStream_t stream) {
  if (B <= 0) return 0;
  return (int)launch_kernel(decode_step_kernel, dim3(1), dim3(256), 0, stream, sampled, lp, ref_lp, value, step_ptr, max_new,
                            B, eos_id, pad_id, tokens_out, logprobs_out, ref_logprobs_out, values_out, finished, resp_lens,
                            seq_lens, positions, next_tokens, n_running);
}

// k / v: [B, T, nkv*d] views with arbitrary batch / time strides (elements); nkv*d % 8 == 0.
extern "C" int b200_paged_kv_write(const void* k, const void* v, void* kcache, void* vcache, const int* block_table,
                                   const int* first, const int* lens, int B, int T, int nkv, int d, int page_size,
                                   int max_pages, long long stride_b, long long stride_t, cudaStream_t stream) {
  if (B <= 0 || T <= 0) return 0;
  dim3 grid(T, B);
  paged_kv_write_kernel<<<grid, 64, 0, stream>>>((const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)kcache,
                                                 (__nv_bfloat16*)vcache, block_table, first, lens, T, nkv, d, page_size,
                                                 max_pages, stride_b, stride_t);
  return (int)cudaGetLastError();
}
