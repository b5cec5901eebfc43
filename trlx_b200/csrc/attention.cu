// Short-sequence attention for the training step and the rollout prefill (sm_100a).
//
// The PPO update of the benchmark runs attention over 56-token rows (32 x 12 heads of 56 x 56 scores): per (batch, head) that
// is 0.4 MFLOP — far too small for a tensor-core pipeline to amortise its fill, and the library path (cuDNN SDPA) spends
// 20 us forward / 50 us backward per layer on it.  Here ONE CTA owns one (batch, head): Q, K, V (and dO) live in shared memory
// for the whole kernel, a warp owns a query row (scores, softmax and P.V for that row never leave the warp), and the backward
// keeps the full P and dS tiles on chip so dQ, dK and dV come out of a single launch without atomics.
//
//   o[b, i, h, :] = softmax_j(scale * q[b,h,i,:].k[b,h,j,:] + bias[b,h,i,j]) . v[b,h,j,:]
//
// q / k / v: bf16 [B, H, T, d] with arbitrary (B, H, T) strides and a contiguous head dimension (the views the fused QKV
// projection produces), d % 8 == 0, d <= 128, Tq, Tk <= 128.  bias: fp32, broadcastable strides, or null with `causal`
// (key j visible to query i iff j <= i + Tk - Tq).  Outputs are written [B, T, H, d] contiguous, which is the layout the output
// projection consumes, so the caller's transpose + reshape is a free view.  Rows save (max, 1 / sum) instead of a fused
// log-sum-exp so that rows whose keys are all at the finite mask value reproduce the forward's uniform weights exactly.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx.cuh"

namespace b200 {

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int ATT_MAX_T = 128;

struct AttnDims {
  int B, H, Tq, Tk, d;
  long long q_sb, q_sh, q_st;   // element strides of q (batch, head, token); the head dimension is contiguous
  long long k_sb, k_sh, k_st;
  long long v_sb, v_sh, v_st;
  long long b_sb, b_sh, b_sq;   // bias strides (0 = broadcast); last dimension contiguous
  float scale;
  int causal;
};

// rows of `d` bf16 padded to d + 2 elements: row stride (d/2 + 1) 32-bit words is odd, so a warp reading one word of 32
// different rows hits 32 different banks
__device__ __forceinline__ int row_words(int d) { return d / 2 + 1; }

__device__ __forceinline__ void load_rows(uint32_t* dst, const __nv_bfloat16* src, long long row_stride, int rows, int d) {
  const int vec_per_row = d / 8;
  const int rw = row_words(d);
  for (int idx = threadIdx.x; idx < rows * vec_per_row; idx += ATT_THREADS) {
    const int r = idx / vec_per_row, c = idx - r * vec_per_row;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (long long)r * row_stride + c * 8);
    uint32_t* p = dst + r * rw + c * 4;
    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
  }
}

__device__ __forceinline__ float2 bf2(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

// dot product of two padded shared-memory rows (a: broadcast row, b: this lane's row)
__device__ __forceinline__ float dot_rows(const uint32_t* a, const uint32_t* b, int half_d) {
  float acc = 0.f;
#pragma unroll 8
  for (int p = 0; p < half_d; ++p) {
    const float2 x = bf2(a[p]), y = bf2(b[p]);
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_short_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                      const __nv_bfloat16* __restrict__ v, const float* __restrict__ bias, __nv_bfloat16* __restrict__ o,
                      float* __restrict__ stats, AttnDims D) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int b = blockIdx.x / D.H, h = blockIdx.x % D.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rw = row_words(D.d), half_d = D.d / 2;
  uint32_t* Ks = reinterpret_cast<uint32_t*>(att_smem);
  uint32_t* Vs = Ks + D.Tk * rw;
  uint32_t* Qs = Vs + D.Tk * rw;
  float* Ps = reinterpret_cast<float*>(Qs + D.Tq * rw);  // [ATT_WARPS][Tk]
  griddep_wait();
  griddep_launch();
  load_rows(Ks, k + b * D.k_sb + h * D.k_sh, D.k_st, D.Tk, D.d);
  load_rows(Vs, v + b * D.v_sb + h * D.v_sh, D.v_st, D.Tk, D.d);
  load_rows(Qs, q + b * D.q_sb + h * D.q_sh, D.q_st, D.Tq, D.d);
  __syncthreads();
  float* P = Ps + warp * D.Tk;
  const int shift = D.Tk - D.Tq;
  for (int i = warp; i < D.Tq; i += ATT_WARPS) {
    const uint32_t* qi = Qs + i * rw;
    const float* brow = bias ? bias + b * D.b_sb + h * D.b_sh + (long long)i * D.b_sq : nullptr;
    float s[ATT_MAX_T / 32];
    float m = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < ATT_MAX_T / 32; ++jj) {
      const int j = jj * 32 + lane;
      s[jj] = -INFINITY;
      if (j < D.Tk && !(D.causal && j > i + shift)) {
        s[jj] = dot_rows(qi, Ks + j * rw, half_d) * D.scale + (brow ? brow[j] : 0.f);
        m = fmaxf(m, s[jj]);
      }
    }
    m = warp_max(m);
    if (m == -INFINITY) m = 0.f;  // every key masked with a true -inf: all weights zero
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < ATT_MAX_T / 32; ++jj) {
      s[jj] = __expf(s[jj] - m);
      sum += s[jj];
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
    for (int jj = 0; jj < ATT_MAX_T / 32; ++jj) {
      const int j = jj * 32 + lane;
      if (j < D.Tk) P[j] = s[jj] * inv;
    }
    __syncwarp();
    __nv_bfloat16* orow = o + (((long long)b * D.Tq + i) * D.H + h) * D.d;
    for (int p = lane; p < half_d; p += 32) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < D.Tk; ++j) {
        const float w = P[j];
        const float2 vv = bf2(Vs[j * rw + p]);
        a0 = fmaf(w, vv.x, a0);
        a1 = fmaf(w, vv.y, a1);
      }
      *reinterpret_cast<__nv_bfloat162*>(orow + 2 * p) = __floats2bfloat162_rn(a0, a1);
    }
    if (lane == 0 && stats) {
      float* st = stats + (((long long)b * D.H + h) * D.Tq + i) * 2;
      st[0] = m;
      st[1] = inv;
    }
    __syncwarp();
  }
}

// dq / dk / dv: bf16 [B, T, H, d] contiguous.  o / d_o: [B, Tq, H, d] with token stride `o_st`, batch stride `o_sb` (head
// stride d).
__global__ void __launch_bounds__(ATT_THREADS)
attn_short_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                      const __nv_bfloat16* __restrict__ v, const float* __restrict__ bias,
                      const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, long long do_sb,
                      long long do_sh, long long do_st, const float* __restrict__ stats, __nv_bfloat16* __restrict__ dq,
                      __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, AttnDims D) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int b = blockIdx.x / D.H, h = blockIdx.x % D.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rw = row_words(D.d), half_d = D.d / 2;
  const int pw = D.Tk + 1;  // padded row of the P / dS tiles
  uint32_t* Ks = reinterpret_cast<uint32_t*>(att_smem);
  uint32_t* Vs = Ks + D.Tk * rw;
  uint32_t* Qs = Vs + D.Tk * rw;
  uint32_t* Gs = Qs + D.Tq * rw;                           // dO
  float* P = reinterpret_cast<float*>(Gs + D.Tq * rw);     // [Tq][Tk + 1]
  float* dS = P + D.Tq * pw;                               // [Tq][Tk + 1], already multiplied by `scale`
  griddep_wait();
  griddep_launch();
  load_rows(Ks, k + b * D.k_sb + h * D.k_sh, D.k_st, D.Tk, D.d);
  load_rows(Vs, v + b * D.v_sb + h * D.v_sh, D.v_st, D.Tk, D.d);
  load_rows(Qs, q + b * D.q_sb + h * D.q_sh, D.q_st, D.Tq, D.d);
  load_rows(Gs, d_o + b * do_sb + h * do_sh, do_st, D.Tq, D.d);
  __syncthreads();
  const int shift = D.Tk - D.Tq;
  // ---- rows: P, dS and dQ
  for (int i = warp; i < D.Tq; i += ATT_WARPS) {
    const uint32_t* qi = Qs + i * rw;
    const uint32_t* gi = Gs + i * rw;
    const __nv_bfloat16* orow = o + (((long long)b * D.Tq + i) * D.H + h) * D.d;
    float di = 0.f;
    for (int p = lane; p < half_d; p += 32) {
      const float2 g = bf2(gi[p]);
      const float2 oo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(orow + 2 * p));
      di = fmaf(g.x, oo.x, fmaf(g.y, oo.y, di));
    }
    di = warp_sum(di);
    const float* st = stats + (((long long)b * D.H + h) * D.Tq + i) * 2;
    const float m = st[0], inv = st[1];
    const float* brow = bias ? bias + b * D.b_sb + h * D.b_sh + (long long)i * D.b_sq : nullptr;
    for (int j = lane; j < D.Tk; j += 32) {
      float p = 0.f, ds = 0.f;
      if (!(D.causal && j > i + shift)) {
        const float s = dot_rows(qi, Ks + j * rw, half_d) * D.scale + (brow ? brow[j] : 0.f);
        p = __expf(s - m) * inv;
        const float dp = dot_rows(gi, Vs + j * rw, half_d);
        ds = p * (dp - di) * D.scale;
      }
      P[i * pw + j] = p;
      dS[i * pw + j] = ds;
    }
    __syncwarp();
    __nv_bfloat16* dqrow = dq + (((long long)b * D.Tq + i) * D.H + h) * D.d;
    for (int p = lane; p < half_d; p += 32) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < D.Tk; ++j) {
        const float w = dS[i * pw + j];
        const float2 kk = bf2(Ks[j * rw + p]);
        a0 = fmaf(w, kk.x, a0);
        a1 = fmaf(w, kk.y, a1);
      }
      *reinterpret_cast<__nv_bfloat162*>(dqrow + 2 * p) = __floats2bfloat162_rn(a0, a1);
    }
  }
  __syncthreads();
  // ---- columns: dK and dV
  for (int j = warp; j < D.Tk; j += ATT_WARPS) {
    __nv_bfloat16* dkrow = dk + (((long long)b * D.Tk + j) * D.H + h) * D.d;
    __nv_bfloat16* dvrow = dv + (((long long)b * D.Tk + j) * D.H + h) * D.d;
    for (int p = lane; p < half_d; p += 32) {
      float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
      for (int i = 0; i < D.Tq; ++i) {
        const float w = P[i * pw + j], ws = dS[i * pw + j];
        const float2 g = bf2(Gs[i * rw + p]);
        const float2 qq = bf2(Qs[i * rw + p]);
        v0 = fmaf(w, g.x, v0);
        v1 = fmaf(w, g.y, v1);
        k0 = fmaf(ws, qq.x, k0);
        k1 = fmaf(ws, qq.y, k1);
      }
      *reinterpret_cast<__nv_bfloat162*>(dkrow + 2 * p) = __floats2bfloat162_rn(k0, k1);
      *reinterpret_cast<__nv_bfloat162*>(dvrow + 2 * p) = __floats2bfloat162_rn(v0, v1);
    }
  }
}

// ---------------------------------------------------------------------------------------------- tcgen05 forward
// The same forward on the tensor cores: S = Q.K^T and O = P.V are two tcgen05.mma groups (UMMA M = 128) with the accumulators
// in TMEM — S in columns [0, 128), O in [128, 128 + d).  TMEM lane = query row, so after tcgen05.ld every thread holds one
// whole row of scores: bias, masking, max, exp and the row sum are thread-local (no shuffles), and the un-normalised
// probabilities go back to shared memory as the bf16 K-major A operand of the second MMA (128-byte swizzled rows, the layout
// TMA would have produced).  V is consumed as an MN-major B operand straight from its [token, d] layout.  One CTA of 128 threads
// per (batch, head); Tq, Tk <= 128, d in {64, 128}.
constexpr int ATC_THREADS = 128;

// [rows_total x 64] bf16 tile, 128-byte rows, 16-byte units XOR-swizzled with (row & 7); rows >= rows_valid are zero
__device__ __forceinline__ void atc_load_tile(uint32_t tile_u32, const __nv_bfloat16* src, long long row_stride, int rows_valid,
                                              int rows_total) {
  for (int idx = threadIdx.x; idx < rows_total * 8; idx += ATC_THREADS) {
    const int r = idx >> 3, u = idx & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows_valid) v = *reinterpret_cast<const uint4*>(src + (long long)r * row_stride + u * 8);
    uint32_t off = (uint32_t)r * 128u + (uint32_t)u * 16u;
    off ^= ((off >> 7) & 7u) << 4;
    st_shared_v4(tile_u32 + off, v);
  }
}

__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_tc_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ o, float* __restrict__ stats, AttnDims D) {
  extern __shared__ __align__(1024) uint8_t atc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~uintptr_t(1023));
  const int b = blockIdx.x / D.H, h = blockIdx.x % D.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dch = D.d / 64;                       // 64-column chunks of the head dimension (k-blocks of the first MMA)
  const int n_s = (D.Tk + 15) & ~15;              // UMMA N of S = Q.K^T (and K extent of P.V), multiple of 16
  const int kb_tok = (n_s + 63) / 64;             // 64-token k-blocks of the second MMA
  // shared memory: Q [dch][128 x 128 B] | K [dch][128 x 128 B] | V [kb_tok][dch][64 x 128 B] | P [kb_tok][128 x 128 B] | bar, slot
  const uint32_t q_u32 = smem_u32(smem);
  const uint32_t k_u32 = q_u32 + (uint32_t)dch * 16384u;
  const uint32_t v_u32 = k_u32 + (uint32_t)dch * 16384u;
  const uint32_t p_u32 = v_u32 + (uint32_t)kb_tok * dch * 8192u;
  uint8_t* tail = smem + (size_t)dch * 32768 + (size_t)kb_tok * dch * 8192 + (size_t)kb_tok * 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 16);

  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  griddep_wait();
  griddep_launch();
  const __nv_bfloat16* qb = q + b * D.q_sb + h * D.q_sh;
  const __nv_bfloat16* kb_ = k + b * D.k_sb + h * D.k_sh;
  const __nv_bfloat16* vb = v + b * D.v_sb + h * D.v_sh;
  for (int c = 0; c < dch; ++c) {
    atc_load_tile(q_u32 + c * 16384u, qb + c * 64, D.q_st, D.Tq, 128);
    atc_load_tile(k_u32 + c * 16384u, kb_ + c * 64, D.k_st, D.Tk, 128);
  }
  for (int t = 0; t < kb_tok; ++t)
    for (int c = 0; c < dch; ++c)
      atc_load_tile(v_u32 + (uint32_t)(t * dch + c) * 8192u, vb + (long long)t * 64 * D.v_st + c * 64, D.v_st,
                    max(min(D.Tk - t * 64, 64), 0), 64);
  fence_proxy_async();           // generic-proxy writes -> visible to the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {  // S = Q . K^T
    const uint32_t idesc = umma_idesc(1, 1, 128, (uint32_t)n_s);
    for (int c = 0; c < dch; ++c) {
      const uint64_t da = umma_desc_k_sw128(q_u32 + c * 16384u), db = umma_desc_k_sw128(k_u32 + c * 16384u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, da + 2 * kk, db + 2 * kk, idesc, (c > 0 || kk > 0) ? 1u : 0u);
    }
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after_sync();

  const int i = warp * 32 + lane;  // query row == TMEM lane
  const bool row_ok = i < D.Tq;
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const float* brow = (bias && row_ok) ? bias + b * D.b_sb + h * D.b_sh + (long long)i * D.b_sq : nullptr;
  const int shift = D.Tk - D.Tq;
  float m = -INFINITY;
  for (int c0 = 0; c0 < n_s; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(taddr + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = c0 + j;
      if (row_ok && col < D.Tk && !(D.causal && col > i + shift))
        m = fmaxf(m, __uint_as_float(r[j]) * D.scale + (brow ? brow[col] : 0.f));
    }
  }
  if (m == -INFINITY) m = 0.f;
  float sum = 0.f;
  for (int c0 = 0; c0 < kb_tok * 64; c0 += 16) {
    float e[16];
    if (c0 < n_s) {
      uint32_t r[16];
      tmem_ld16(taddr + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c0 + j;
        const bool ok = row_ok && col < D.Tk && !(D.causal && col > i + shift);
        e[j] = ok ? __expf(__uint_as_float(r[j]) * D.scale + (brow ? brow[col] : 0.f) - m) : 0.f;
        // the second MMA consumes bf16 probabilities: accumulate the row sum from the rounded values it will actually use
        e[j] = __bfloat162float(__float2bfloat16(e[j]));
        sum += e[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) e[j] = 0.f;
    }
    uint4 pk[2];
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
    for (int j = 0; j < 8; ++j) h2[j] = __floats2bfloat162_rn(e[2 * j], e[2 * j + 1]);
    const uint32_t tile = p_u32 + (uint32_t)(c0 >> 6) * 16384u;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t off = (uint32_t)i * 128u + (uint32_t)((c0 & 63) * 2 + hh * 16);
      off ^= ((off >> 7) & 7u) << 4;
      st_shared_v4(tile + off, pk[hh]);
    }
  }
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  if (threadIdx.x == 0) {  // O = P . V   (A: P K-major, B: V MN-major)
    const uint32_t idesc = umma_idesc(1, 1, 128, (uint32_t)D.d, 0, 1);
    bool first = true;
    for (int t = 0; t < kb_tok; ++t) {
      const uint64_t da = umma_desc_k_sw128(p_u32 + (uint32_t)t * 16384u);
      const uint64_t db = umma_desc_mn_sw128(v_u32 + (uint32_t)t * dch * 8192u, 8192u);
      const int steps = min((n_s - t * 64) / 16, 4);
      for (int kk = 0; kk < steps; ++kk) {
        umma_bf16(tmem_base + 128, da + 2 * kk, db + 128 * kk, idesc, first ? 0u : 1u);
        first = false;
      }
    }
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after_sync();
  // every lane runs the (warp-aligned) TMEM loads; only the stores are predicated on the row being real
  __nv_bfloat16* orow = o + (((long long)b * D.Tq + (row_ok ? i : 0)) * D.H + h) * D.d;
  for (int c0 = 0; c0 < D.d; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(taddr + 128 + c0, r);
    tmem_ld_wait();
    uint4 pk[2];
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
    for (int j = 0; j < 8; ++j) h2[j] = __floats2bfloat162_rn(__uint_as_float(r[2 * j]) * inv, __uint_as_float(r[2 * j + 1]) * inv);
    if (row_ok) {
      *reinterpret_cast<uint4*>(orow + c0) = pk[0];
      *reinterpret_cast<uint4*>(orow + c0 + 8) = pk[1];
    }
  }
  if (row_ok && stats) {
    float* st = stats + (((long long)b * D.H + h) * D.Tq + i) * 2;
    st[0] = m;
    st[1] = inv;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

static size_t attn_tc_smem(int Tk, int d) {
  const int dch = d / 64, kb_tok = (((Tk + 15) & ~15) + 63) / 64;
  return (size_t)dch * 32768 + (size_t)kb_tok * dch * 8192 + (size_t)kb_tok * 16384 + 64 + 1024;
}

static size_t attn_fwd_smem(int Tq, int Tk, int d) {
  return (size_t)(2 * Tk + Tq) * (d / 2 + 1) * 4 + (size_t)ATT_WARPS * Tk * 4;
}
static size_t attn_bwd_smem(int Tq, int Tk, int d) {
  return (size_t)(2 * Tk + 2 * Tq) * (d / 2 + 1) * 4 + (size_t)2 * Tq * (Tk + 1) * 4;
}

}  // namespace b200

using namespace b200;

// 1 when the shape can run on these kernels (the caller falls back to the library path otherwise)
extern "C" int b200_attn_short_ok(int Tq, int Tk, int d, int backward) {
  if (Tq < 1 || Tk < 1 || Tq > ATT_MAX_T || Tk > ATT_MAX_T || d % 8 || d < 8 || d > 128) return 0;
  const size_t need = backward ? attn_bwd_smem(Tq, Tk, d) : attn_fwd_smem(Tq, Tk, d);
  return need <= 220 * 1024 ? 1 : 0;
}

extern "C" int b200_attn_short_fwd(const void* q, const void* k, const void* v, const float* bias, void* o, float* stats,
                                   int B, int H, int Tq, int Tk, int d, const long long* qs, const long long* ks,
                                   const long long* vs, const long long* bs, float scale, int causal, cudaStream_t stream) {
  if (!b200_attn_short_ok(Tq, Tk, d, 0)) return -1;
  AttnDims D{B, H, Tq, Tk, d, qs[0], qs[1], qs[2], ks[0], ks[1], ks[2], vs[0], vs[1], vs[2],
             bias ? bs[0] : 0, bias ? bs[1] : 0, bias ? bs[2] : 0, scale, causal};
  const size_t smem = attn_fwd_smem(Tq, Tk, d);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_short_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
      return -5;
    configured = true;
  }
  return (int)launch_kernel(attn_short_fwd_kernel, dim3((unsigned)(B * H)), dim3(ATT_THREADS), smem, stream,
                            (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, bias,
                            (__nv_bfloat16*)o, stats, D);
}

extern "C" int b200_attn_short_bwd(const void* q, const void* k, const void* v, const float* bias, const void* o,
                                   const void* d_o, const long long* dos, const float* stats, void* dq, void* dk, void* dv,
                                   int B, int H, int Tq, int Tk, int d, const long long* qs, const long long* ks,
                                   const long long* vs, const long long* bs, float scale, int causal, cudaStream_t stream) {
  if (!b200_attn_short_ok(Tq, Tk, d, 1)) return -1;
  AttnDims D{B, H, Tq, Tk, d, qs[0], qs[1], qs[2], ks[0], ks[1], ks[2], vs[0], vs[1], vs[2],
             bias ? bs[0] : 0, bias ? bs[1] : 0, bias ? bs[2] : 0, scale, causal};
  const size_t smem = attn_bwd_smem(Tq, Tk, d);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_short_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
      return -5;
    configured = true;
  }
  return (int)launch_kernel(attn_short_bwd_kernel, dim3((unsigned)(B * H)), dim3(ATT_THREADS), smem, stream,
                            (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, bias,
                            (const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o, dos[0], dos[1], dos[2], stats,
                            (__nv_bfloat16*)dq, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, D);
}

// tcgen05 forward (Tq, Tk <= 128, d in {64, 128}); same contract as b200_attn_short_fwd
extern "C" int b200_attn_tc_ok(int Tq, int Tk, int d) {
  return (Tq >= 1 && Tk >= 1 && Tq <= 128 && Tk <= 128 && (d == 64 || d == 128)) ? 1 : 0;
}

extern "C" int b200_attn_tc_fwd(const void* q, const void* k, const void* v, const float* bias, void* o, float* stats, int B,
                                int H, int Tq, int Tk, int d, const long long* qs, const long long* ks, const long long* vs,
                                const long long* bs, float scale, int causal, cudaStream_t stream) {
  if (!b200_attn_tc_ok(Tq, Tk, d)) return -1;
  AttnDims D{B, H, Tq, Tk, d, qs[0], qs[1], qs[2], ks[0], ks[1], ks[2], vs[0], vs[1], vs[2],
             bias ? bs[0] : 0, bias ? bs[1] : 0, bias ? bs[2] : 0, scale, causal};
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -5;
    configured = true;
  }
  return (int)launch_kernel(attn_tc_fwd_kernel, dim3((unsigned)(B * H)), dim3(ATC_THREADS), attn_tc_smem(Tk, d), stream,
                            (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, bias, (__nv_bfloat16*)o,
                            stats, D);
}
