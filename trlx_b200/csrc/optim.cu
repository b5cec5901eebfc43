// Partitioned AdamW for sm_100a, optionally fused with the data-parallel gradient reduce-scatter and the parameter
// all-gather over NVLink peer memory.
//
// Layout: all trainable parameters of a rank live in ONE flat bf16 buffer (`param`), gradients in a matching flat bf16
// buffer (`grad`).  Optimizer state (fp32 master copy, exp_avg, exp_avg_sq) exists only for the shard a rank owns.
// With world > 1 the flat grad/param buffers are symmetric-memory allocations, so every rank holds device pointers to
// every peer's copy:
//
//   fused_rs_adamw_ag:   g[i]   = (1/world) * sum_r peer_grad[r][i]          <- P2P loads over NVLink (reduce-scatter)
//                        m,v,w  = AdamW(g[i])                                 <- fp32 master update of the owned shard
//                        peer_param[r][i] = bf16(w)   for every r             <- P2P stores over NVLink (all-gather)
//
// which is the reference's  DDP all-reduce -> (clip) -> optimizer.step  /  ZeRO-2 reduce-scatter -> step -> all-gather
// (accelerate_base_trainer.py:574-587) done in a single pass over the shard.  `signal_barrier` is the system-scope
// flag handshake that orders "every rank finished backward" before and "every rank sees new params" after.
#include "ptx.cuh"

namespace b200 {

constexpr int MAX_PEERS = 16;
struct PeerPtrs { void* p[MAX_PEERS]; };

// hyper[0]=lr, hyper[1]=1-beta1^t, hyper[2]=1-beta2^t, hyper[3]=grad scale (clip coefficient / loss-scale inverse)
struct AdamArgs {
  float beta1, beta2, eps, weight_decay;
  int decoupled;  // 1 = AdamW (decoupled decay), 0 = Adam (L2 added to the gradient)
};

__device__ __forceinline__ float adam_update(float g, float& w, float& m, float& v, const AdamArgs& a, float lr, float bc1,
                                             float bc2) {
  if (!a.decoupled) g += a.weight_decay * w;
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  const float denom = sqrtf(v / bc2) + a.eps;
  if (a.decoupled) w *= (1.f - lr * a.weight_decay);
  w -= lr * (m / bc1) / denom;
  return w;
}

// Single-GPU / already-reduced path.  grad may be bf16 (grad_f32 == 0) or fp32.
__global__ void __launch_bounds__(256)
adamw_flat_kernel(__nv_bfloat16* __restrict__ param, float* __restrict__ master, const void* __restrict__ grad, int grad_f32,
                  float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, long long n, AdamArgs a,
                  const float* __restrict__ hyper) {
  const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2], gs = hyper[3];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = (grad_f32 ? reinterpret_cast<const float*>(grad)[i]
                              : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(grad)[i])) * gs;
    float w = master[i], m = exp_avg[i], v = exp_avg_sq[i];
    adam_update(g, w, m, v, a, lr, bc1, bc2);
    master[i] = w; exp_avg[i] = m; exp_avg_sq[i] = v;
    param[i] = __float2bfloat16(w);
  }
}

// ---- 8-bit-state AdamW (bitsandbytes' Adam8bit / AdamW8bit role) ----------------------------------------------------------------
// One CTA per 256-element block: decode both moments (log-domain codes x per-block fp32 absmax), apply the update in fp32,
// reduce the new block maxima in shared memory, re-encode.  First moment: sign + 1/4-octave magnitude (codes +-1..127, 0 = exact
// zero); second moment: 1/8-octave magnitude (codes 1..255).  Same code book as the PyTorch fallback in parallel/optim.py, so
// optimizer state moves freely between the two.  Parameters / gradients are bf16 or fp32.
__device__ __forceinline__ float block_max_256(float x, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
  __syncthreads();
  float r = sh[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = fmaxf(r, sh[i]);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256)
adam8bit_kernel(void* __restrict__ param, const void* __restrict__ grad, int is_f32, signed char* __restrict__ mq,
                float* __restrict__ mscale, unsigned char* __restrict__ vq, float* __restrict__ vscale, long long n, AdamArgs a,
                float lr, float bc1, float bc2) {
  __shared__ float sh[8];
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const bool live = i < n;
  float w = 0.f, g = 0.f;
  if (live) {
    w = is_f32 ? reinterpret_cast<float*>(param)[i] : __bfloat162float(reinterpret_cast<__nv_bfloat16*>(param)[i]);
    g = is_f32 ? reinterpret_cast<const float*>(grad)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(grad)[i]);
  }
  const int cm = mq[i], cv = vq[i];  // state buffers are padded to whole blocks
  const float ms = mscale[blockIdx.x], vs = vscale[blockIdx.x];
  float m = cm == 0 ? 0.f : exp2f(((float)abs(cm) - 127.f) * 0.25f) * (cm < 0 ? -ms : ms);
  float v = cv == 0 ? 0.f : exp2f(((float)cv - 255.f) * 0.125f) * vs;
  if (live) {
    if (!a.decoupled && a.weight_decay != 0.f) g += a.weight_decay * w;
    m = a.beta1 * m + (1.f - a.beta1) * g;
    v = a.beta2 * v + (1.f - a.beta2) * g * g;
    if (a.decoupled && a.weight_decay != 0.f) w *= 1.f - lr * a.weight_decay;
    w -= lr * (m / bc1) / (sqrtf(v / bc2) + a.eps);
    if (is_f32) reinterpret_cast<float*>(param)[i] = w;
    else reinterpret_cast<__nv_bfloat16*>(param)[i] = __float2bfloat16(w);
  } else {
    m = 0.f; v = 0.f;
  }
  const float nms = fmaxf(block_max_256(fabsf(m), sh), 1e-12f);
  const float nvs = fmaxf(block_max_256(v, sh), 1e-12f);
  int qm = 0, qv = 0;
  if (m != 0.f) {
    const float lg = log2f(fmaxf(fabsf(m) / nms, 9.094947e-13f));  // 2^-40
    qm = (int)fminf(fmaxf(rintf(127.f + 4.f * lg), 1.f), 127.f);
    if (m < 0.f) qm = -qm;
  }
  if (v > 0.f) {
    const float lg = log2f(fmaxf(v / nvs, 9.094947e-13f));
    qv = (int)fminf(fmaxf(rintf(255.f + 8.f * lg), 1.f), 255.f);
  }
  mq[i] = (signed char)qm;
  vq[i] = (unsigned char)qv;
  if (threadIdx.x == 0) { mscale[blockIdx.x] = nms; vscale[blockIdx.x] = nvs; }
}

// Same update, one WARP per 256-element block and 8 elements per lane: 16-byte parameter / gradient accesses, 8-byte code accesses,
// block maxima by warp shuffles (no __syncthreads), fast log2.  (The one-element-per-thread kernel above is issue-bound: ncu shows
// 84 % SM throughput at 1.0 TB/s of HBM traffic, profiles/ncu_optim_summary.csv.)  Requires 16-byte aligned param / grad.
template <bool F32>
__global__ void __launch_bounds__(256)
adam8bit_warp_kernel(void* __restrict__ param, const void* __restrict__ grad, signed char* __restrict__ mq, float* __restrict__ mscale,
                     unsigned char* __restrict__ vq, float* __restrict__ vscale, long long n, AdamArgs a, float lr, float bc1,
                     float bc2) {
  // decode tables (code -> magnitude relative to the block scale): 127 first-moment and 255 second-moment magnitudes
  __shared__ float tab_m[128], tab_v[256];
  tab_v[threadIdx.x] = threadIdx.x == 0 ? 0.f : exp2f(((float)threadIdx.x - 255.f) * 0.125f);
  if (threadIdx.x < 128) tab_m[threadIdx.x] = threadIdx.x == 0 ? 0.f : exp2f(((float)threadIdx.x - 127.f) * 0.25f);
  __syncthreads();
  const long long blk = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (blk >= (n + 255) / 256) return;  // whole warps leave together; nothing below synchronises across warps
  const int lane = threadIdx.x & 31;
  const long long i0 = blk * 256 + lane * 8;
  const bool full = i0 + 8 <= n;
  float w[8], g[8], m[8], v[8];
  if (full) {
    if (F32) {
      const float4* pw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(param) + i0);
      const float4* pg = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grad) + i0);
      const float4 w0 = pw[0], w1 = pw[1], g0 = pg[0], g1 = pg[1];
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
      g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    } else {
      const uint4 wv = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(param) + i0);
      const uint4 gv = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(grad) + i0);
      const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wv);
      const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fw = __bfloat1622float2(wh[j]), fg = __bfloat1622float2(gh[j]);
        w[2 * j] = fw.x; w[2 * j + 1] = fw.y; g[2 * j] = fg.x; g[2 * j + 1] = fg.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool live = i0 + j < n;
      w[j] = !live ? 0.f : F32 ? reinterpret_cast<const float*>(param)[i0 + j]
                               : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(param)[i0 + j]);
      g[j] = !live ? 0.f : F32 ? reinterpret_cast<const float*>(grad)[i0 + j]
                               : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(grad)[i0 + j]);
    }
  }
  const uint2 cmv = *reinterpret_cast<const uint2*>(mq + i0);  // state buffers are padded to whole blocks
  const uint2 cvv = *reinterpret_cast<const uint2*>(vq + i0);
  const signed char* cm = reinterpret_cast<const signed char*>(&cmv);
  const unsigned char* cv = reinterpret_cast<const unsigned char*>(&cvv);
  const float ms = mscale[blk], vs = vscale[blk];
  float amax = 0.f, vmax = 0.f;
  const float step = lr / bc1, inv_bc2 = 1.f / bc2, decay = 1.f - lr * a.weight_decay;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cm[j], d = cv[j];
    m[j] = tab_m[abs(c)] * (c < 0 ? -ms : ms);
    v[j] = tab_v[d] * vs;
    if (i0 + j < n) {
      float gj = g[j];
      if (!a.decoupled && a.weight_decay != 0.f) gj += a.weight_decay * w[j];
      m[j] = a.beta1 * m[j] + (1.f - a.beta1) * gj;
      v[j] = a.beta2 * v[j] + (1.f - a.beta2) * gj * gj;
      if (a.decoupled && a.weight_decay != 0.f) w[j] *= decay;
      w[j] -= __fdividef(step * m[j], sqrtf(v[j] * inv_bc2) + a.eps);
    } else {
      m[j] = 0.f; v[j] = 0.f;
    }
    amax = fmaxf(amax, fabsf(m[j]));
    vmax = fmaxf(vmax, v[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  const float nms = fmaxf(amax, 1e-12f), nvs = fmaxf(vmax, 1e-12f);
  const float inv_m = 1.f / nms, inv_v = 1.f / nvs;
  uint2 qmv, qvv;
  signed char* qm = reinterpret_cast<signed char*>(&qmv);
  unsigned char* qv = reinterpret_cast<unsigned char*>(&qvv);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int cq = 0, dq = 0;
    if (m[j] != 0.f) {
      cq = (int)fminf(fmaxf(rintf(127.f + 4.f * __log2f(fmaxf(fabsf(m[j]) * inv_m, 9.094947e-13f))), 1.f), 127.f);
      if (m[j] < 0.f) cq = -cq;
    }
    if (v[j] > 0.f) dq = (int)fminf(fmaxf(rintf(255.f + 8.f * __log2f(fmaxf(v[j] * inv_v, 9.094947e-13f))), 1.f), 255.f);
    qm[j] = (signed char)cq;
    qv[j] = (unsigned char)dq;
  }
  *reinterpret_cast<uint2*>(mq + i0) = qmv;
  *reinterpret_cast<uint2*>(vq + i0) = qvv;
  if (lane == 0) { mscale[blk] = nms; vscale[blk] = nvs; }
  if (full) {
    if (F32) {
      float4* pw = reinterpret_cast<float4*>(reinterpret_cast<float*>(param) + i0);
      pw[0] = make_float4(w[0], w[1], w[2], w[3]);
      pw[1] = make_float4(w[4], w[5], w[6], w[7]);
    } else {
      uint4 o;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(w[2 * j], w[2 * j + 1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(param) + i0) = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (i0 + j < n) {
        if (F32) reinterpret_cast<float*>(param)[i0 + j] = w[j];
        else reinterpret_cast<__nv_bfloat16*>(param)[i0 + j] = __float2bfloat16(w[j]);
      }
    }
  }
}

extern "C" int b200_adam8bit(void* param, const void* grad, int is_f32, void* mq, float* mscale, void* vq, float* vscale,
                             long long n, float beta1, float beta2, float eps, float weight_decay, int decoupled, float lr,
                             float bc1, float bc2, cudaStream_t stream) {
  if (n <= 0) return 0;
  AdamArgs a{beta1, beta2, eps, weight_decay, decoupled};
  const long long blocks = (n + 255) / 256;
  const char* wv = getenv("TRLX_B200_ADAM8BIT_WARP");  // default on; "0" selects the one-element-per-thread kernel (read per call: tests flip it)
  const bool warp_variant = wv == nullptr || wv[0] != '0';
  const bool aligned = ((uintptr_t)param % 16 == 0) && ((uintptr_t)grad % 16 == 0) && ((uintptr_t)mq % 8 == 0) && ((uintptr_t)vq % 8 == 0);
  if (warp_variant && aligned) {
    const unsigned ctas = (unsigned)((blocks + 7) / 8);
    if (is_f32)
      adam8bit_warp_kernel<true><<<ctas, 256, 0, stream>>>(param, grad, (signed char*)mq, mscale, (unsigned char*)vq, vscale, n, a, lr,
                                                           bc1, bc2);
    else
      adam8bit_warp_kernel<false><<<ctas, 256, 0, stream>>>(param, grad, (signed char*)mq, mscale, (unsigned char*)vq, vscale, n, a,
                                                            lr, bc1, bc2);
    return (int)cudaGetLastError();
  }
  adam8bit_kernel<<<(unsigned)blocks, 256, 0, stream>>>(param, grad, is_f32, (signed char*)mq, mscale, (unsigned char*)vq, vscale, n,
                                                        a, lr, bc1, bc2);
  return (int)cudaGetLastError();
}

// sum of squares of a flat bf16/fp32 buffer -> out[0] (double, atomically accumulated; caller zeroes)
__global__ void __launch_bounds__(256) sqnorm_kernel(const void* __restrict__ x, int is_f32, long long n, double* __restrict__ out) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float f = is_f32 ? reinterpret_cast<const float*>(x)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
    s += f * f;
  }
  s = warp_sum(s);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    atomicAdd(out, t);
  }
}

// hyper[3] = min(1, max_norm / (sqrt(sqsum) + 1e-6))
__global__ void clip_coef_kernel(const double* __restrict__ sqsum, float max_norm, float* __restrict__ hyper,
                                 float* __restrict__ norm_out) {
  const float norm = (float)sqrt(*sqsum);
  if (norm_out) *norm_out = norm;
  hyper[3] = fminf(1.f, max_norm / (norm + 1e-6f));
}

// System-scope barrier across `world` ranks: rank r bumps slot [r] in every peer's signal pad, then waits until all
// slots of its own pad reach `epoch`.  One block, >= world threads.
__global__ void signal_barrier_kernel(PeerPtrs pads, int rank, int world, uint32_t* __restrict__ epoch_ptr) {
  // the epoch lives on the device and is bumped by every launch, so the barrier can sit inside a replayed CUDA graph
  // (all ranks launch barriers in the same order, which keeps their counters in lock-step)
  const int t = threadIdx.x;
  const uint32_t epoch = *epoch_ptr + 1;
  __syncthreads();
  if (t == 0) *epoch_ptr = epoch;
  __threadfence_system();
  if (t < world) {
    uint32_t* remote = reinterpret_cast<uint32_t*>(pads.p[t]) + rank;
    st_release_sys(remote, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(pads.p[rank]) + t;
    const long long t0 = clock64();
    while ((int32_t)(ld_relaxed_sys(mine) - epoch) < 0) { __nanosleep(64); spin_guard(t0); }
    fence_acq_rel_sys();
  }
  __syncthreads();
  __threadfence_system();
}

// Cross-rank readiness of ONE gradient bucket, folded into the reduce kernel itself (no separate barrier launch): block 0
// tells every peer "my gradients of this bucket are final" (flag slot [rank] in the peer's flag buffer := epoch), every block
// waits until all peers have said so.  The epoch lives on the device and is advanced by the last block to finish, so the
// kernel can sit inside a replayed CUDA graph and be launched per bucket from autograd hooks while backward is still running.
struct BucketSync {
  void* flags[MAX_PEERS];   // peers' flag buffers, already offset to this bucket's row of `world` uint32 slots
  uint32_t* epoch;          // local: last completed epoch of this bucket
  uint32_t* done;           // local: blocks finished (wraps to 0)
  int rank, enabled;
};

__device__ __forceinline__ uint32_t bucket_sync_begin(const BucketSync& sync, int world) {
  if (!sync.enabled) return 0;
  const uint32_t target = *sync.epoch + 1;
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(sync.flags[threadIdx.x]) + sync.rank, target);
  }
  if ((int)threadIdx.x < world) {
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(sync.flags[sync.rank]) + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_relaxed_sys(mine) - target) < 0) { __nanosleep(100); spin_guard(t0); }
    fence_acq_rel_sys();
  }
  __syncthreads();
  return target;
}
__device__ __forceinline__ void bucket_sync_end(const BucketSync& sync, uint32_t target) {
  if (!sync.enabled) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(sync.done, gridDim.x - 1) == gridDim.x - 1) *sync.epoch = target;
  }
}

// Fused reduce-scatter + AdamW + all-gather over the owned shard [lo, lo + n).  8 bf16 (16 B) per thread-iteration.
// lo and n must be multiples of 8.  If sq_out != null the squared norm of the averaged gradient is accumulated instead
// of updating (phase 1 of clipped updates: the reduced gradient is parked in fp32 `gshard`).
template <bool UPDATE>
__global__ void __launch_bounds__(256)
rs_adamw_ag_kernel(PeerPtrs grads, PeerPtrs params, int world, long long lo, long long n, float* __restrict__ master,
                   float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, float* __restrict__ gshard, int use_gshard,
                   AdamArgs a, const float* __restrict__ hyper, double* __restrict__ sq_out, BucketSync sync,
                   const __nv_bfloat16* __restrict__ mc_grad, __nv_bfloat16* __restrict__ mc_param, float grad_scale) {
  const uint32_t sync_target = bucket_sync_begin(sync, world);
  const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2], gs = hyper[3];
  // the reduced gradient is (sum over the group) * grad_scale: 1 / world for data parallelism; 1 / dp for parameters that are
  // replicated inside a tensor-parallel group under sequence parallelism (partial gradients per TP rank are SUMMED)
  const float inv_world = grad_scale > 0.f ? grad_scale : 1.f / (float)world;
  float sq = 0.f;
  const long long nvec = n >> 3;
  for (long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec; vi += (long long)gridDim.x * blockDim.x) {
    const long long i = vi << 3;
    float g[8];
    if (use_gshard && UPDATE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = gshard[i + j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = 0.f;
      if (mc_grad) {   // NVLS: ONE load, the NVSwitch adds the `world` copies (fp32 accumulate) on the way
        const uint4 raw = multimem_ld_reduce_add_bf16x8(mc_grad + lo + i);
        const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = __bfloat162float(h[j]);
      } else {
        for (int r = 0; r < world; ++r) {
          const int4 raw = ld_nc_v4(reinterpret_cast<const int4*>(reinterpret_cast<const __nv_bfloat16*>(grads.p[r]) + lo + i));
          const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += __bfloat162float(h[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= inv_world;
    }
    if (!UPDATE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { gshard[i + j] = g[j]; sq += g[j] * g[j]; }
      continue;
    }
    uint4 packed;
    __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(&packed);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float w = master[i + j], m = exp_avg[i + j], v = exp_avg_sq[i + j];
      adam_update(g[j] * gs, w, m, v, a, lr, bc1, bc2);
      master[i + j] = w; exp_avg[i + j] = m; exp_avg_sq[i + j] = v;
      ph[j] = __float2bfloat16(w);
    }
    if (mc_param) {   // NVLS: one multicast store updates every rank's copy of the parameters
      multimem_st_bf16x8(mc_param + lo + i, packed);
    } else {
      for (int r = 0; r < world; ++r)
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(params.p[r]) + lo + i) = packed;
    }
  }
  if (!UPDATE && sq_out) {
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) atomicAdd(sq_out, (double)sq);
  }
  bucket_sync_end(sync, sync_target);
}

// Global gradient norm for clipping without a library collective: every rank writes its partial sum of squares into slot
// [parity][rank] of every peer's buffer (P2P store), a flag round publishes them, and each rank adds the `world` slots
// (same order everywhere -> bit-identical coefficient on every rank).  One block of >= world threads.
__global__ void clip_exchange_kernel(PeerPtrs sqbufs, PeerPtrs flags, int rank, int world, const double* __restrict__ sq_local,
                                     uint32_t* __restrict__ epoch_ptr, float max_norm, float* __restrict__ hyper,
                                     float* __restrict__ norm_out) {
  const int t = threadIdx.x;
  const uint32_t epoch = *epoch_ptr + 1;
  const int parity = epoch & 1;
  __syncthreads();
  if (t < world) {
    double* slot = reinterpret_cast<double*>(sqbufs.p[t]) + parity * world + rank;
    *reinterpret_cast<volatile double*>(slot) = *sq_local;
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(flags.p[t]) + rank, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.p[rank]) + t;
    const long long t0 = clock64();
    while ((int32_t)(ld_relaxed_sys(mine) - epoch) < 0) { __nanosleep(64); spin_guard(t0); }
    fence_acq_rel_sys();
  }
  __syncthreads();
  if (t == 0) {
    *epoch_ptr = epoch;
    const volatile double* mine = reinterpret_cast<const volatile double*>(sqbufs.p[rank]) + parity * world;
    double tot = 0.0;
    for (int r = 0; r < world; ++r) tot += mine[r];
    const float norm = (float)sqrt(tot);
    if (norm_out) *norm_out = norm;
    hyper[3] = fminf(1.f, max_norm / (norm + 1e-6f));
  }
}

// Polyak / EMA update of target networks:  tgt = alpha * src + (1 - alpha) * tgt   (ILQL target-Q sync)
__global__ void lerp_kernel(__nv_bfloat16* __restrict__ tgt, const __nv_bfloat16* __restrict__ src, long long n, float alpha) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    tgt[i] = __float2bfloat16(alpha * __bfloat162float(src[i]) + (1.f - alpha) * __bfloat162float(tgt[i]));
}

static int grid_for(long long n, int per_thread = 1) {
  long long b = (n / per_thread + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_adamw_flat(void* param, float* master, const void* grad, int grad_f32, float* exp_avg, float* exp_avg_sq,
                               long long n, float beta1, float beta2, float eps, float weight_decay, int decoupled,
                               const float* hyper, cudaStream_t stream) {
  if (n <= 0) return 0;
  AdamArgs a{beta1, beta2, eps, weight_decay, decoupled};
  adamw_flat_kernel<<<grid_for(n), 256, 0, stream>>>((__nv_bfloat16*)param, master, grad, grad_f32, exp_avg, exp_avg_sq, n, a, hyper);
  return (int)cudaGetLastError();
}

extern "C" int b200_sqnorm(const void* x, int is_f32, long long n, double* out, cudaStream_t stream) {
  if (n <= 0) return 0;
  sqnorm_kernel<<<grid_for(n), 256, 0, stream>>>(x, is_f32, n, out);
  return (int)cudaGetLastError();
}

extern "C" int b200_clip_coef(const double* sqsum, float max_norm, float* hyper, float* norm_out, cudaStream_t stream) {
  clip_coef_kernel<<<1, 1, 0, stream>>>(sqsum, max_norm, hyper, norm_out);
  return (int)cudaGetLastError();
}

extern "C" int b200_signal_barrier(void* const* pads, int rank, int world, unsigned int* epoch_ptr, cudaStream_t stream) {
  if (world > MAX_PEERS) return -3;
  PeerPtrs p{};
  for (int i = 0; i < world; ++i) p.p[i] = pads[i];
  signal_barrier_kernel<<<1, 32, 0, stream>>>(p, rank, world, epoch_ptr);
  return (int)cudaGetLastError();
}

// mode 0: fused reduce + update + gather ; mode 1: reduce into gshard + squared norm ; mode 2: update from gshard + gather
extern "C" int b200_rs_adamw_ag(void* const* grads, void* const* params, int world, long long lo, long long n, float* master,
                                float* exp_avg, float* exp_avg_sq, float* gshard, int mode, float beta1, float beta2,
                                float eps, float weight_decay, int decoupled, const float* hyper, double* sq_out,
                                cudaStream_t stream) {
  if (n <= 0) return 0;
  if (world > MAX_PEERS || (lo & 7) || (n & 7)) return -3;
  PeerPtrs g{}, p{};
  for (int i = 0; i < world; ++i) { g.p[i] = grads[i]; p.p[i] = params[i]; }
  AdamArgs a{beta1, beta2, eps, weight_decay, decoupled};
  const int grid = grid_for(n, 8);
  BucketSync sync{};
  if (mode == 1)
    rs_adamw_ag_kernel<false><<<grid, 256, 0, stream>>>(g, p, world, lo, n, master, exp_avg, exp_avg_sq, gshard, 1, a, hyper, sq_out, sync, nullptr, nullptr, 0.f);
  else
    rs_adamw_ag_kernel<true><<<grid, 256, 0, stream>>>(g, p, world, lo, n, master, exp_avg, exp_avg_sq, gshard, mode == 2, a, hyper, nullptr, sync, nullptr, nullptr, 0.f);
  return (int)cudaGetLastError();
}

// One gradient bucket: same modes as above, with the cross-rank "gradients final" handshake folded into the kernel.
// flags[r] = rank r's flag buffer (uint32), flag_offset = first slot of this bucket's row; epoch / done: local counters.
extern "C" int b200_rs_adamw_ag_bucket(void* const* grads, void* const* params, int world, int rank, long long lo, long long n,
                                       float* master, float* exp_avg, float* exp_avg_sq, float* gshard, int mode, float beta1,
                                       float beta2, float eps, float weight_decay, int decoupled, const float* hyper,
                                       double* sq_out, void* const* flags, long long flag_offset, unsigned int* epoch,
                                       unsigned int* done, int max_blocks, const void* mc_grad, void* mc_param,
                                       float grad_scale, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (world > MAX_PEERS || (lo & 7) || (n & 7)) return -3;
  PeerPtrs g{}, p{};
  BucketSync sync{};
  for (int i = 0; i < world; ++i) {
    g.p[i] = grads[i];
    p.p[i] = params[i];
    sync.flags[i] = reinterpret_cast<uint32_t*>(flags[i]) + flag_offset;
  }
  sync.epoch = epoch; sync.done = done; sync.rank = rank; sync.enabled = mode != 2;  // phase 2 reads the parked gradient
  AdamArgs a{beta1, beta2, eps, weight_decay, decoupled};
  int grid = grid_for(n, 8);
  if (max_blocks > 0 && grid > max_blocks) grid = max_blocks;
  const __nv_bfloat16* mg = reinterpret_cast<const __nv_bfloat16*>(mc_grad);
  __nv_bfloat16* mp = reinterpret_cast<__nv_bfloat16*>(mc_param);
  if (mode == 1)
    rs_adamw_ag_kernel<false><<<grid, 256, 0, stream>>>(g, p, world, lo, n, master, exp_avg, exp_avg_sq, gshard, 1, a, hyper, sq_out, sync, mg, mp, grad_scale);
  else
    rs_adamw_ag_kernel<true><<<grid, 256, 0, stream>>>(g, p, world, lo, n, master, exp_avg, exp_avg_sq, gshard, mode == 2, a, hyper, nullptr, sync,
                                                       mode == 2 ? nullptr : mg, mp, grad_scale);
  return (int)cudaGetLastError();
}

extern "C" int b200_clip_exchange(void* const* sqbufs, void* const* flags, long long flag_offset, int rank, int world,
                                  const double* sq_local, unsigned int* epoch, float max_norm, float* hyper, float* norm_out,
                                  cudaStream_t stream) {
  if (world > MAX_PEERS) return -3;
  PeerPtrs s{}, f{};
  for (int i = 0; i < world; ++i) {
    s.p[i] = sqbufs[i];
    f.p[i] = reinterpret_cast<uint32_t*>(flags[i]) + flag_offset;
  }
  clip_exchange_kernel<<<1, 32, 0, stream>>>(s, f, rank, world, sq_local, epoch, max_norm, hyper, norm_out);
  return (int)cudaGetLastError();
}

extern "C" int b200_lerp_bf16(void* tgt, const void* src, long long n, float alpha, cudaStream_t stream) {
  if (n <= 0) return 0;
  lerp_kernel<<<grid_for(n), 256, 0, stream>>>((__nv_bfloat16*)tgt, (const __nv_bfloat16*)src, n, alpha);
  return (int)cudaGetLastError();
}
