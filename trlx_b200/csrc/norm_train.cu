// Training-side LayerNorm / RMSNorm (forward that keeps the row statistics, one-pass backward) and the bias-gradient column sum.
//
// The PPO update runs five norms and nine bias gradients per optimizer step on [1792, 768..3072] activations; the stock
// kernels take 13.7 us (gamma/beta backward) + 4.4 us (input gradient) per norm and 9 us per column sum
// (profiles/torchprof_train_step_v3_graph.txt).  All three are memory-trivial (a few MB, L2 resident), so the cost is launch
// count and reduction structure:
//   * ln_fwd_kernel   : one 128-thread CTA per row, 16-byte loads, writes y and (mean, rstd)
//   * ln_bwd_kernel   : persistent CTAs stride over rows; dx is finished per row, the d-gamma / d-beta contributions stay in
//                       registers (each thread owns fixed columns) and leave as ONE fp32 partial per CTA — no atomics
//   * ln_bwd_finalize : sums the per-CTA partials into bf16 d-gamma / d-beta
//   * colsum_kernel   : bias gradient; a CTA owns 64 columns, its 8 warps stride over the rows, one shared-memory fold
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx.cuh"

namespace b200 {

constexpr int LN_THREADS = 128;
constexpr int LN_MAX_VPT = 4;  // 16-byte vectors per thread: H <= 128 * 4 * 8 = 4096

__device__ __forceinline__ void block_sum2(float& a, float& b, float (*red)[LN_THREADS / 32]) {
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();  // `red` may still be read from the previous use
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  a = red[0][0] + red[0][1] + red[0][2] + red[0][3];
  b = red[1][0] + red[1][1] + red[1][2] + red[1][3];
}

template <bool RMS>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ b,
              __nv_bfloat16* __restrict__ y, float* __restrict__ stats, int H, long long ldx, float eps) {
  __shared__ float red[2][LN_THREADS / 32];
  griddep_wait();
  griddep_launch();
  const int row = blockIdx.x;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  __nv_bfloat16* yr = y + (size_t)row * H;
  const int nvec = H >> 3;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < nvec; i += LN_THREADS) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i * 8);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float f = __bfloat162float(h[j]); s += f; ss += f * f; }
  }
  block_sum2(s, ss, red);
  const float mean = RMS ? 0.f : s / H;
  const float var = RMS ? ss / H : fmaxf(ss / H - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) { stats[2 * (size_t)row] = mean; stats[2 * (size_t)row + 1] = rstd; }
  for (int i = threadIdx.x; i < nvec; i += LN_THREADS) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i * 8);
    const uint4 wv = *reinterpret_cast<const uint4*>(w + i * 8);
    uint4 bv = make_uint4(0, 0, 0, 0);
    if (b) bv = *reinterpret_cast<const uint4*>(b + i * 8);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
    const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(&wv);
    const __nv_bfloat16* bh = reinterpret_cast<const __nv_bfloat16*>(&bv);
    uint4 o;
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      oh[j] = __float2bfloat16((__bfloat162float(h[j]) - mean) * rstd * __bfloat162float(wh[j]) + __bfloat162float(bh[j]));
    *reinterpret_cast<uint4*>(yr + i * 8) = o;
  }
}

// partial: [gridDim.x, 2, H] fp32 (d-gamma, d-beta contributions of the rows this CTA processed)
template <bool RMS>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const float* __restrict__ stats,
              const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, float* __restrict__ partial, int rows, int H,
              long long ldx, long long lddy) {
  __shared__ float red[2][LN_THREADS / 32];
  griddep_wait();
  griddep_launch();
  const int nvec = H >> 3;
  float dg[LN_MAX_VPT][8], db[LN_MAX_VPT][8], wf[LN_MAX_VPT][8];
#pragma unroll
  for (int v = 0; v < LN_MAX_VPT; ++v) {
    const int i = threadIdx.x + v * LN_THREADS;
    uint4 wv = make_uint4(0, 0, 0, 0);
    if (i < nvec) wv = *reinterpret_cast<const uint4*>(w + i * 8);
    const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(&wv);
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[v][j] = 0.f; db[v][j] = 0.f; wf[v][j] = __bfloat162float(wh[j]); }
  }
  const float inv_h = 1.f / H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mean = stats[2 * (size_t)row], rstd = stats[2 * (size_t)row + 1];
    const __nv_bfloat16* xr = x + (size_t)row * ldx;
    const __nv_bfloat16* gr = dy + (size_t)row * lddy;
    float xh[LN_MAX_VPT][8], g[LN_MAX_VPT][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int v = 0; v < LN_MAX_VPT; ++v) {
      const int i = threadIdx.x + v * LN_THREADS;
      if (i < nvec) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xr + i * 8);
        const uint4 gv = *reinterpret_cast<const uint4*>(gr + i * 8);
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(&xv);
        const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(&gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = __bfloat162float(gb[j]);
          xh[v][j] = (__bfloat162float(xb[j]) - mean) * rstd;
          g[v][j] = d * wf[v][j];
          c1 += g[v][j];
          c2 += g[v][j] * xh[v][j];
          dg[v][j] += d * xh[v][j];
          db[v][j] += d;
        }
      }
    }
    block_sum2(c1, c2, red);
    c1 = RMS ? 0.f : c1 * inv_h;
    c2 *= inv_h;
    __nv_bfloat16* dr = dx + (size_t)row * H;
#pragma unroll
    for (int v = 0; v < LN_MAX_VPT; ++v) {
      const int i = threadIdx.x + v * LN_THREADS;
      if (i < nvec) {
        uint4 o;
        __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
        for (int j = 0; j < 8; ++j) oh[j] = __float2bfloat16(rstd * (g[v][j] - c1 - xh[v][j] * c2));
        *reinterpret_cast<uint4*>(dr + i * 8) = o;
      }
    }
  }
  float* pg = partial + (size_t)blockIdx.x * 2 * H;
  float* pb = pg + H;
#pragma unroll
  for (int v = 0; v < LN_MAX_VPT; ++v) {
    const int i = threadIdx.x + v * LN_THREADS;
    if (i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { pg[i * 8 + j] = dg[v][j]; pb[i * 8 + j] = db[v][j]; }
    }
  }
}

// Same contract, one WARP per row (H <= 1024): no block-wide barrier inside the row loop and four rows in flight per CTA, so at
// 1792 rows x 768 nearly every row of the batch is being processed at once (the CTA-per-row variant above serialises ~6 rows
// behind two __syncthreads each and measured no faster than the stock kernels, run40).
constexpr int LNW_VPT = 4;      // 16-byte vectors per lane
constexpr int LNW_MAX_H = 32 * LNW_VPT * 8;
template <bool RMS>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_warp_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const float* __restrict__ stats,
                   const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, float* __restrict__ partial, int rows,
                   int H, long long ldx, long long lddy) {
  __shared__ float fold[LN_THREADS / 32][2][LNW_MAX_H];  // 32 KB
  griddep_wait();
  griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  float dg[LNW_VPT][8], db[LNW_VPT][8], wf[LNW_VPT][8];
#pragma unroll
  for (int v = 0; v < LNW_VPT; ++v) {
    const int i = lane + v * 32;
    uint4 wv = make_uint4(0, 0, 0, 0);
    if (i < nvec) wv = *reinterpret_cast<const uint4*>(w + i * 8);
    const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(&wv);
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[v][j] = 0.f; db[v][j] = 0.f; wf[v][j] = __bfloat162float(wh[j]); }
  }
  const float inv_h = 1.f / H;
  for (int row = blockIdx.x * (LN_THREADS / 32) + warp; row < rows; row += gridDim.x * (LN_THREADS / 32)) {
    const float mean = stats[2 * (size_t)row], rstd = stats[2 * (size_t)row + 1];
    const __nv_bfloat16* xr = x + (size_t)row * ldx;
    const __nv_bfloat16* gr = dy + (size_t)row * lddy;
    float xh[LNW_VPT][8], g[LNW_VPT][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int v = 0; v < LNW_VPT; ++v) {
      const int i = lane + v * 32;
      if (i < nvec) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xr + i * 8);
        const uint4 gv = *reinterpret_cast<const uint4*>(gr + i * 8);
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(&xv);
        const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(&gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = __bfloat162float(gb[j]);
          xh[v][j] = (__bfloat162float(xb[j]) - mean) * rstd;
          g[v][j] = d * wf[v][j];
          c1 += g[v][j];
          c2 += g[v][j] * xh[v][j];
          dg[v][j] += d * xh[v][j];
          db[v][j] += d;
        }
      }
    }
    c1 = RMS ? 0.f : warp_sum(c1) * inv_h;
    c2 = warp_sum(c2) * inv_h;
    __nv_bfloat16* dr = dx + (size_t)row * H;
#pragma unroll
    for (int v = 0; v < LNW_VPT; ++v) {
      const int i = lane + v * 32;
      if (i < nvec) {
        uint4 o;
        __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
        for (int j = 0; j < 8; ++j) oh[j] = __float2bfloat16(rstd * (g[v][j] - c1 - xh[v][j] * c2));
        *reinterpret_cast<uint4*>(dr + i * 8) = o;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < LNW_VPT; ++v) {
    const int i = lane + v * 32;
    if (i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { fold[warp][0][i * 8 + j] = dg[v][j]; fold[warp][1][i * 8 + j] = db[v][j]; }
    }
  }
  __syncthreads();
  float* pg = partial + (size_t)blockIdx.x * 2 * H;
  for (int c = threadIdx.x; c < 2 * H; c += LN_THREADS) {
    const int which = c >= H ? 1 : 0, col = c - which * H;
    pg[c] = fold[0][which][col] + fold[1][which][col] + fold[2][which][col] + fold[3][which][col];
  }
}

// out[0:H] = d-gamma, out[H:2H] = d-beta (bf16); one thread per output column
// CTA = 64 of the 2H columns (d-gamma | d-beta), 8 warps stride over the per-CTA partial rows with four loads in flight, one
// shared-memory fold.  (One thread per column walking all ~300 partial rows serially took 17 us, five times per optimizer step.)
__global__ void __launch_bounds__(256)
ln_bwd_finalize_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ dgamma,
                       __nv_bfloat16* __restrict__ dbeta, int n_part, int H) {
  __shared__ float red[8][64];
  griddep_wait();
  griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 64 + lane * 2;
  const int W = 2 * H;
  float a0 = 0.f, a1 = 0.f;
  if (c0 < W) {  // H is even: the pair is inside or outside together
    const float* q = partial + c0;
    int p = warp;
    for (; p + 24 < n_part; p += 32) {
      const float2 v0 = *reinterpret_cast<const float2*>(q + (size_t)p * W);
      const float2 v1 = *reinterpret_cast<const float2*>(q + (size_t)(p + 8) * W);
      const float2 v2 = *reinterpret_cast<const float2*>(q + (size_t)(p + 16) * W);
      const float2 v3 = *reinterpret_cast<const float2*>(q + (size_t)(p + 24) * W);
      a0 += (v0.x + v1.x) + (v2.x + v3.x);
      a1 += (v0.y + v1.y) + (v2.y + v3.y);
    }
    for (; p < n_part; p += 8) {
      const float2 v = *reinterpret_cast<const float2*>(q + (size_t)p * W);
      a0 += v.x;
      a1 += v.y;
    }
  }
  red[warp][lane * 2] = a0;
  red[warp][lane * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < H) dgamma[c] = __float2bfloat16(s);
    else if (c < W && dbeta) dbeta[c - H] = __float2bfloat16(s);
  }
}

// out[n] = sum_m x[m, n].  CTA (bx, by) = 64 columns x the by-th slice of the rows; its 8 warps stride over the slice (lanes own a
// bf16 pair, four independent loads in flight), one shared-memory fold, then fp32 atomics into `acc`; the LAST slice to finish a
// column block (ticket counter) converts it to bf16 and clears accumulator + ticket for the next call.  (A 64-column CTA over ALL
// rows — 12 CTAs for a 768-wide bias — was a chain of ~56 dependent L2 round trips: 26 us x 9 bias gradients per optimizer step.)
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int M, int N, long long ldx,
              float* __restrict__ acc, unsigned int* __restrict__ tickets) {
  __shared__ float red[8][64];
  __shared__ bool last;
  griddep_wait();
  griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 64 + lane * 2;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_lo = blockIdx.y * rows_per, m_hi = min(M, m_lo + rows_per);
  float a0 = 0.f, a1 = 0.f;
  if (col < N) {  // N is even (checked on the host), so the pair is either fully inside or fully outside
    const __nv_bfloat16* p = x + col;
    int m = m_lo + warp;
    for (; m + 24 < m_hi; m += 32) {  // four independent loads in flight per lane
      const float2 v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + (size_t)m * ldx));
      const float2 v1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + (size_t)(m + 8) * ldx));
      const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + (size_t)(m + 16) * ldx));
      const float2 v3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + (size_t)(m + 24) * ldx));
      a0 += (v0.x + v1.x) + (v2.x + v3.x);
      a1 += (v0.y + v1.y) + (v2.y + v3.y);
    }
    for (; m < m_hi; m += 8) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + (size_t)m * ldx));
      a0 += v.x;
      a1 += v.y;
    }
  }
  red[warp][lane * 2] = a0;
  red[warp][lane * 2 + 1] = a1;
  __syncthreads();
  const int c = blockIdx.x * 64 + threadIdx.x;
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
    if (gridDim.y == 1) {
      if (c < N) out[c] = __float2bfloat16(s);
    } else if (c < N) {
      atomicAdd(acc + c, s);
    }
  }
  if (gridDim.y == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1;
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < 64 && c < N) {
      out[c] = __float2bfloat16(__ldcg(acc + c));
      acc[c] = 0.f;
    }
    if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_ln_train_ok(int H) { return (H % 8 == 0 && H >= 8 && H <= LN_THREADS * LN_MAX_VPT * 8) ? 1 : 0; }

// grid size of the backward (the caller sizes the partial buffer [n, 2, H] with it)
extern "C" int b200_ln_bwd_ctas(int rows) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = 2 * sms;
  return rows < want ? (rows < 1 ? 1 : rows) : want;
}

extern "C" int b200_ln_fwd(const void* x, const void* w, const void* b, void* y, float* stats, int rows, int H, long long ldx,
                           float eps, int rms, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (!b200_ln_train_ok(H)) return -2;
  if (rms)
    return (int)launch_kernel(ln_fwd_kernel<true>, dim3(rows), dim3(LN_THREADS), 0, stream, (const __nv_bfloat16*)x,
                              (const __nv_bfloat16*)w, (const __nv_bfloat16*)nullptr, (__nv_bfloat16*)y, stats, H, ldx, eps);
  return (int)launch_kernel(ln_fwd_kernel<false>, dim3(rows), dim3(LN_THREADS), 0, stream, (const __nv_bfloat16*)x,
                            (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, stats, H, ldx, eps);
}

extern "C" int b200_ln_bwd(const void* x, const void* w, const float* stats, const void* dy, void* dx, float* partial,
                           void* dgamma, void* dbeta, int rows, int H, long long ldx, long long lddy, int rms,
                           cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (!b200_ln_train_ok(H)) return -2;
  const int ctas = b200_ln_bwd_ctas(rows);
  cudaError_t e;
  if (H <= LNW_MAX_H) {
    const int want = (rows + LN_THREADS / 32 - 1) / (LN_THREADS / 32);
    const int grid = want < ctas ? want : ctas;  // never more CTAs than the partial buffer was sized for
    if (rms)
      e = launch_kernel(ln_bwd_warp_kernel<true>, dim3(grid), dim3(LN_THREADS), 0, stream, (const __nv_bfloat16*)x,
                        (const __nv_bfloat16*)w, stats, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, partial, rows, H, ldx, lddy);
    else
      e = launch_kernel(ln_bwd_warp_kernel<false>, dim3(grid), dim3(LN_THREADS), 0, stream, (const __nv_bfloat16*)x,
                        (const __nv_bfloat16*)w, stats, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, partial, rows, H, ldx, lddy);
    if (e != cudaSuccess) return (int)e;
    return (int)launch_kernel(ln_bwd_finalize_kernel, dim3((2 * H + 63) / 64), dim3(256), 0, stream, (const float*)partial,
                              (__nv_bfloat16*)dgamma, (__nv_bfloat16*)dbeta, grid, H);
  }
  if (rms)
    e = launch_kernel(ln_bwd_kernel<true>, dim3(ctas), dim3(LN_THREADS), 0, stream, (const __nv_bfloat16*)x,
                      (const __nv_bfloat16*)w, stats, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, partial, rows, H, ldx, lddy);
  else
    e = launch_kernel(ln_bwd_kernel<false>, dim3(ctas), dim3(LN_THREADS), 0, stream, (const __nv_bfloat16*)x,
                      (const __nv_bfloat16*)w, stats, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, partial, rows, H, ldx, lddy);
  if (e != cudaSuccess) return (int)e;
  return (int)launch_kernel(ln_bwd_finalize_kernel, dim3((2 * H + 63) / 64), dim3(256), 0, stream, (const float*)partial,
                            (__nv_bfloat16*)dgamma, (__nv_bfloat16*)dbeta, ctas, H);
}

// `workspace`: N fp32 accumulators + ceil(N / 64) tickets, ZERO on entry (the kernel leaves it zero again); may be null for
// short matrices (single row slice).
extern "C" int b200_colsum_rows(int M, int N) {
  const int col_blocks = (N + 63) / 64;
  int r = (2 * 148 + col_blocks - 1) / col_blocks;   // ~2 CTAs per SM in total
  const int max_r = (M + 63) / 64;                   // at least 64 rows per CTA
  if (r > max_r) r = max_r;
  return r < 1 ? 1 : r;
}

extern "C" int b200_colsum_bf16(const void* x, void* out, int M, int N, long long ldx, float* workspace, cudaStream_t stream) {
  if (N <= 0) return 0;
  if (N % 2 || ldx % 2) return -2;
  const int r = workspace ? b200_colsum_rows(M, N) : 1;
  return (int)launch_kernel(colsum_kernel, dim3((N + 63) / 64, r), dim3(256), 0, stream, (const __nv_bfloat16*)x,
                            (__nv_bfloat16*)out, M, N, ldx, workspace,
                            reinterpret_cast<unsigned int*>(workspace ? workspace + N : nullptr));
}
