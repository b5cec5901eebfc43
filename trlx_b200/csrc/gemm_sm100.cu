// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = A[M,K] · B[N,K]^T  (both operands K-major bf16, fp32 accumulate)
//
// One CTA computes a 128 x BN output tile:
//   warp 0 (one lane)  : TMA producer  — cp.async.bulk.tensor tiles of A (128x64) and B (BNx64), 128B-swizzled,
//                        into a multi-stage shared-memory ring guarded by full/empty mbarriers
//   warp 1 (one lane)  : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, accumulator in TMEM (BN fp32 columns),
//                        tcgen05.commit releases ring slots and finally signals the epilogue
//   warps 2..5         : epilogue      — tcgen05.ld 32x32b (one output row per thread), fused epilogue, global store
//
// Two epilogues share the mainloop:
//   * STORE   : out = act(alpha * acc [* col_scale[n]] + bias[n]) + residual[m,n]      (bf16 or fp32 out)
//   * LMHEAD  : never writes logits.  Per (row, N-tile) it emits online-softmax partials (max, sum-exp), the
//               logit of a given label, and (optionally) a Gumbel-max sampling candidate; a tiny second kernel
//               merges the partials into  logsumexp / log p(label) / sampled token + its log-prob.
//               This replaces the reference's  logits = lm_head(h); log_softmax; gather  chain
//               (trlx/utils/modeling.py:213-219) and HF generate's softmax+multinomial.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "ptx.cuh"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;        // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int STG_BYTES = 4096;   // per epilogue warp: 32 rows x 128 B staging for the TMA-store epilogue

enum Act { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2, ACT_RELU = 3, ACT_SILU = 4 };

struct StoreEpilogue {
  void* out;                 // bf16 or fp32 [M, ldo]
  const __nv_bfloat16* bias; // [N] or null
  const __nv_bfloat16* residual;  // [M, ldr] or null
  const float* col_scale;    // [N] or null
  long long ldo, ldr;
  float alpha;
  int act;
  int out_f32;
  // d-logits mode (LM-head backward, all three set): out = ((col == label) - exp(x - lse[row])) * grad[row]
  const float* dl_lse;
  const float* dl_grad;
  const long long* dl_labels;
  int debug_nostore;
  // LayerNorm / RMSNorm folded into this GEMM (decode path): the A operand is the RAW residual stream x and the weights are
  // pre-multiplied by gamma, so  LN(x).W^T = rstd*(x.(gamma*W)^T) - rstd*mu*c1 + (W.beta + b)  with per-row (mu, rstd) from
  // ln_stats = [M, 2] (sum x, sum x^2) that the PRODUCER of x accumulated, c1[n] = sum_k gamma_k W[n,k]; the constant term
  // is passed as the ordinary bias.  ln_rms: RMSNorm (mu = 0).
  const float* ln_stats;
  const float* ln_c1;
  float ln_inv_k, ln_eps;
  int ln_rms;
  // ... and the producer side: accumulate (sum, sum of squares) of the bf16-rounded output rows for the NEXT folded norm
  float* stats_out;
  const float* row_scale;    // [M] or null: per-row dequantisation scale of an fp8 A operand (col_scale carries B's)
  int tma_store;              // bf16 output goes through shared memory + cp.async.bulk.tensor (needs ldo % 8 == 0)
};

constexpr int MAX_TP = 8;
// A-operand tensor maps, one per tensor-parallel peer (all-gather -> GEMM reads row-block r of A from peer r's HBM)
struct MapArray { CUtensorMap m[MAX_TP]; };

// GEMM -> reduce-scatter epilogue: rows [r*rows_per_rank, (r+1)*rows_per_rank) of the partial product are added
// (NVLink red.add.f32) into peer r's fp32 accumulation buffer [rows_per_rank, ldacc].
struct ReduceScatterEpilogue {
  float* acc[MAX_TP];
  long long ldacc;
  int rows_per_rank;
  // staged variant (tensor-parallel GEMM -> reduce-scatter): instead of fp32 atomics into the owner's accumulator, each rank
  // writes its bf16 partial tile into slot `src_rank` of the owner's staging buffer [world, rows_per_rank, ldstage] with plain
  // 16-byte stores (half the NVLink bytes, no remote atomics); the owner sums the slots afterwards.
  __nv_bfloat16* stage[MAX_TP];
  long long ldstage;
  int src_rank;
};

// Row-chunk readiness of the A operand (all-gather -> GEMM): rows [c*rows_per_flag, (c+1)*rows_per_flag) of A may be read once
// flags[c] == epoch (set by the stream that copies peer shards into the local gathered buffer).  m_rot rotates the M-tile order
// so every rank starts on its own (already local) shard while the copies of the others are in flight.
struct AReady {
  const uint32_t* flags;
  uint32_t epoch;
  int rows_per_flag;
  int m_rot;
  // B (the weights) is not written by the kernel that precedes this one in the stream: its first tiles are fetched BEFORE
  // griddepcontrol.wait, i.e. while the producer of A is still running (decode graphs; see b200_set_static_weights)
  int b_static;
};

struct LMHeadEpilogue {
  const __nv_bfloat16* bias;   // [N] or null
  const long long* labels;     // [M] or null  (label < 0 -> ignored)
  float* part_max;             // [M, n_tiles]
  float* part_sum;             // [M, n_tiles]
  float* label_logit;          // [M]   (written by the tile that owns the label)
  // Gumbel-max sampling (enabled when samp_key != null)
  float* samp_key;             // [M, n_tiles] best perturbed key in tile
  float* samp_logit;           // [M, n_tiles] raw logit of that candidate
  int* samp_idx;               // [M, n_tiles]
  float inv_temperature;      // <= 0 -> greedy (no Gumbel noise)
  unsigned long long seed;
  const long long* seed_ptr;   // optional device-side seed offset (lets a captured CUDA graph draw fresh noise per call)
  const int* step_ptr;         // optional device step counter: mixed into the seed, gates `suppress_col`
  int suppress_col;            // column forced to -inf while step < suppress_until (EOS before min_new_tokens), -1 = none
  int suppress_until;
  int n_tiles;
};

// Ring-buffer position kept incrementally: `it % stages` / `(it / stages) & 1` with a run-time `stages` cost two integer
// divisions (~150 cycles of dependent ALU work) per k-block in the single producer / MMA-issuer threads — more than the
// tensor-core time of a narrow tile's k-block.
__device__ __forceinline__ void ring_next(int& s, uint32_t& phase, int stages) {
  if (++s == stages) { s = 0; phase ^= 1u; }
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_GELU_TANH: return gelu_tanh(x);
    case ACT_GELU_ERF: return gelu_erf(x);
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_SILU: return silu(x);
    default: return x;
  }
}

// counter-based uniform in (0,1): splitmix64 of (seed, row, col)
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned int row, unsigned int col) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (((unsigned long long)row << 32) | col);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return ((float)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------------- epilogues
// Code size matters here: the epilogue warps of all 148 CTAs stream this code continuously, and an earlier version that
// unrolled per-element bounds / bias / activation branches grew to ~190 KB of SASS and spent half of its issue slots in
// `stall_no_inst` (instruction-cache misses; profiles/ncu_gemm_v3_icache.md).  So: one branch-free hot path for full
// 16-column chunks (vector loads of bias / scale / residual, the activation switch hoisted out of the element loop) and a
// rolled scalar path for the ragged last chunk of a row.

__device__ __forceinline__ void load16_bf16(const __nv_bfloat16* p, float (&o)[16]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 8);
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(x[j]), g = __bfloat1622float2(y[j]);
    o[2 * j] = f.x; o[2 * j + 1] = f.y; o[8 + 2 * j] = g.x; o[8 + 2 * j + 1] = g.y;
  }
}

__device__ __forceinline__ void activate16(float (&v)[16], int act) {
  switch (act) {  // hoisted: one compact loop per activation, only the one in use is ever fetched
    case ACT_GELU_TANH:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = gelu_tanh(v[j]);
      break;
    case ACT_GELU_ERF:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
      break;
    case ACT_RELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case ACT_SILU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = silu(v[j]);
      break;
    default: break;
  }
}

struct RowCtx {      // per-thread (= per output row) constants of the store epilogue
  bool dl;
  float dl_l, dl_g;
  long long dl_lab;
  float ln_rstd, ln_murstd;
  float row_scale;
};

// v = act(alpha * acc [* col_scale] [+ bias])  (or the d-logits transform)  [+ residual]   for a FULL 16-column chunk
__device__ __forceinline__ void store_values_full(const uint32_t (&r)[16], float (&v)[16], int row, int col0,
                                                  const StoreEpilogue& se, const RowCtx& rc, bool vec_in, bool row_ok) {
#pragma unroll
  const float a_scale = se.alpha * rc.row_scale;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * a_scale;
  if (se.col_scale) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 c = *reinterpret_cast<const float4*>(se.col_scale + col0 + j);  // col0 % 16 == 0, fp32 [N] 16B-aligned
      v[j] *= c.x; v[j + 1] *= c.y; v[j + 2] *= c.z; v[j + 3] *= c.w;
    }
  }
  if (se.ln_stats) {  // folded normalisation: per-row scale, per-row x per-column mean correction
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 c = *reinterpret_cast<const float4*>(se.ln_c1 + col0 + j);
      v[j] = v[j] * rc.ln_rstd - rc.ln_murstd * c.x;
      v[j + 1] = v[j + 1] * rc.ln_rstd - rc.ln_murstd * c.y;
      v[j + 2] = v[j + 2] * rc.ln_rstd - rc.ln_murstd * c.z;
      v[j + 3] = v[j + 3] * rc.ln_rstd - rc.ln_murstd * c.w;
    }
  }
  if (se.bias) {
    float b[16];
    if (vec_in) {
      load16_bf16(se.bias + col0, b);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) b[j] = __bfloat162float(se.bias[col0 + j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += b[j];
  }
  if (rc.dl) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = -__expf(v[j] - rc.dl_l) * rc.dl_g;
    const long long o = rc.dl_lab - col0;
    if (o >= 0 && o < 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += (j == (int)o) ? rc.dl_g : 0.f;
    }
    if (rc.dl_lab < 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
  }
  activate16(v, se.act);
  if (se.residual) {
    const __nv_bfloat16* rp = se.residual + (size_t)row * se.ldr + col0;
    float b[16];
    if ((se.ldr & 7) == 0 && vec_in) {
      load16_bf16(rp, b);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) b[j] = __bfloat162float(rp[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += b[j];
  }
  if (se.stats_out && row_ok) {  // moments of what the consumer will actually read (the bf16-rounded row)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float f = __bfloat162float(__float2bfloat16(v[j]));
      s1 += f;
      s2 += f * f;
    }
    atomicAdd(se.stats_out + 2 * (size_t)row, s1);
    atomicAdd(se.stats_out + 2 * (size_t)row + 1, s2);
  }
}

// ragged last chunk of a row (col0 + 16 > N): rolled, scalar, correct for any alignment; columns >= N come out as 0
__device__ __noinline__ void tail_values(const uint32_t* r, float* v, int row, int col0, int N, const StoreEpilogue& se, RowCtx rc) {
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    const int col = col0 + j;
    float x = 0.f;
    if (col < N) {
      x = __uint_as_float(r[j]) * se.alpha * rc.row_scale;
      if (se.col_scale) x *= se.col_scale[col];
      if (se.bias) x += __bfloat162float(se.bias[col]);
      if (rc.dl) x = (rc.dl_lab < 0) ? 0.f : (((long long)col == rc.dl_lab ? 1.f : 0.f) - __expf(x - rc.dl_l)) * rc.dl_g;
      x = apply_act(x, se.act);
      if (se.residual) x += __bfloat162float(se.residual[(size_t)row * se.ldr + col]);
    }
    v[j] = x;
  }
}

__device__ __forceinline__ void pack16(const float (&v)[16], uint4 (&pk)[2]) {
  __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
  for (int j = 0; j < 8; ++j) p2[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
}

// Epilogue bodies: one output row per thread, columns [c_lo, c_hi) of the current tile's accumulator.
//   stg      : this warp's 4 KB shared-memory staging buffer (TMA-store path)
//   map_out  : tensor map of the bf16 output (TMA-store path; box = CW columns x 32 rows, swizzle = CW*2 bytes)
//   CSK      : cluster split-K leader — `csk_parts` fp32 partial tiles written by the other CTAs of the cluster into this
//              CTA's shared memory ([chunk of 16 columns][4 x 16-byte unit][lane], `csk_stride` bytes per partial, base
//              `csk_base` already offset to this warp) are added to the accumulator before the epilogue math
template <int EPI, int CW, int CSK = 0>
__device__ __forceinline__ void epilogue_cols(uint32_t taddr_row, int row, bool row_ok, int n0, int c_lo, int c_hi, int N,
                                              int part_idx, const StoreEpilogue& se, const LMHeadEpilogue& le,
                                              const ReduceScatterEpilogue& re, uint8_t* stg, const CUtensorMap* map_out,
                                              int row0_warp, int lane, uint32_t csk_base = 0, int csk_parts = 0,
                                              uint32_t csk_stride = 0) {
  if constexpr (EPI == 0) {
    RowCtx rc{se.dl_lse != nullptr, 0.f, 0.f, -1, 1.f, 0.f, 1.f};
    if (se.row_scale && row_ok) rc.row_scale = se.row_scale[row];
    if (rc.dl && row_ok) { rc.dl_l = se.dl_lse[row]; rc.dl_g = se.dl_grad[row]; rc.dl_lab = se.dl_labels[row]; }
    if (se.ln_stats && row_ok) {
      const float s1 = se.ln_stats[2 * (size_t)row], s2 = se.ln_stats[2 * (size_t)row + 1];
      const float mean = se.ln_rms ? 0.f : s1 * se.ln_inv_k;
      const float var = fmaxf(s2 * se.ln_inv_k - mean * mean, 0.f);
      rc.ln_rstd = rsqrtf(var + se.ln_eps);
      rc.ln_murstd = mean * rc.ln_rstd;
    }
    // 16-byte vector access to bias / residual / output needs 16-byte aligned bases (col0 is a multiple of 16 elements)
    const bool vec_in = ((reinterpret_cast<uintptr_t>(se.bias) | reinterpret_cast<uintptr_t>(se.residual)) & 15) == 0;
    const bool vec_out = ((se.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(se.out) & 15) == 0);
    // Two ways out of registers, one shared body:
    //  * TMA store: registers -> (swizzled) shared memory -> one bulk tensor store per 32 x CW chunk (TMA clips at M / N)
    //  * direct   : each thread writes its row's 16 columns (32 B) straight to global memory
    constexpr int ROWB = CW * 2;                                  // bytes per staged row: 128 / 64 / 32
    constexpr uint32_t SW_MASK = ROWB == 128 ? 7u : (ROWB == 64 ? 3u : 1u);
    const uint32_t stg_u32 = smem_u32(stg);
    const bool tma = se.tma_store != 0;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 16) {
      const int col0 = n0 + c;
      if (col0 >= N) break;
      const int cc = (c - c_lo) % CW;  // offset inside the staged chunk
      if (tma && cc == 0) {
        if (lane == 0) tma_store_wait_read();   // the previous chunk's bulk store has drained the staging buffer
        __syncwarp();
      }
      uint32_t r[16];
      tmem_ld16(taddr_row + c, r);
      tmem_ld_wait();
      if constexpr (CSK) {
        const uint32_t chunk = csk_base + (uint32_t)((c - c_lo) >> 4) * 2048u + (uint32_t)lane * 16u;
#pragma unroll 1
        for (int p = 0; p < csk_parts; ++p) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 t = ld_shared_v4f(chunk + (uint32_t)p * csk_stride + (uint32_t)j * 512u);
            r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + t.x);
            r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + t.y);
            r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + t.z);
            r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + t.w);
          }
        }
      }
      float v[16];
      const bool full = col0 + 16 <= N;
      if (full) store_values_full(r, v, row_ok ? row : 0, col0, se, rc, vec_in, row_ok);
      else {  // rare: keep the address-taken copies out of the hot path's registers
        uint32_t tr[16];
        float tv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) tr[j] = r[j];
        tail_values(tr, tv, row_ok ? row : 0, col0, N, se, rc);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tv[j];
      }
      if (tma) {
        uint4 pk[2];
        pack16(v, pk);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t off = (uint32_t)lane * ROWB + (uint32_t)(cc * 2 + h * 16);
          off ^= ((off >> 7) & SW_MASK) << 4;                     // the TMA swizzle: 16-byte unit ^= (128-byte row index)
          st_shared_v4(stg_u32 + off, pk[h]);
        }
        const bool last_of_chunk = (cc + 16 == CW) || (c + 16 >= c_hi) || (col0 + 16 >= N);
        if (last_of_chunk) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {  // the box is clipped at M rows / N columns by the tensor map
            tma_store_2d(map_out, stg, col0 - cc, row0_warp);
            tma_store_commit();
          }
        }
        continue;
      }
      if (!row_ok || se.debug_nostore) continue;
      if (!full) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < N - col0) {
            if (se.out_f32) reinterpret_cast<float*>(se.out)[(size_t)row * se.ldo + col0 + j] = v[j];
            else reinterpret_cast<__nv_bfloat16*>(se.out)[(size_t)row * se.ldo + col0 + j] = __float2bfloat16(v[j]);
          }
        }
        continue;
      }
      if (se.out_f32) {
        float* op = reinterpret_cast<float*>(se.out) + (size_t)row * se.ldo + col0;
        if ((se.ldo & 3) == 0 && vec_out) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) op[j] = v[j];
        }
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(se.out) + (size_t)row * se.ldo + col0;
        if (vec_out) {
          uint4 pk[2];
          pack16(v, pk);
          *reinterpret_cast<uint4*>(op) = pk[0];
          *reinterpret_cast<uint4*>(op + 8) = pk[1];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) op[j] = __float2bfloat16(v[j]);
        }
      }
    }
  } else if constexpr (EPI == 2) {
    const int owner = row_ok ? row / re.rows_per_rank : 0;
    const bool staged = re.stage[0] != nullptr;
    float* dst_row = (row_ok && !staged) ? re.acc[owner] + (size_t)(row - owner * re.rows_per_rank) * re.ldacc : nullptr;
    __nv_bfloat16* stg_row = (row_ok && staged)
        ? re.stage[owner] + ((size_t)re.src_rank * re.rows_per_rank + (row - owner * re.rows_per_rank)) * re.ldstage : nullptr;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 16) {
      uint32_t r[16];
      tmem_ld16(taddr_row + c, r);
      tmem_ld_wait();
      const int col0 = n0 + c;
      if (!row_ok || col0 >= N) continue;
      if (staged) {
        if (col0 + 16 <= N && (re.ldstage & 7) == 0) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          if (se.bias) {  // the rank that holds the (row-parallel) bias folds it into its partial
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __bfloat162float(se.bias[col0 + j]);
          }
          uint4 pk[2];
          pack16(v, pk);
          *reinterpret_cast<uint4*>(stg_row + col0) = pk[0];
          *reinterpret_cast<uint4*>(stg_row + col0 + 8) = pk[1];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col0 + j < N)
              stg_row[col0 + j] = __float2bfloat16(__uint_as_float(r[j]) + (se.bias ? __bfloat162float(se.bias[col0 + j]) : 0.f));
        }
        continue;
      }
      if (col0 + 16 <= N && (re.ldacc & 3) == 0) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (se.bias) {  // only the rank that owns the bias passes it
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += __bfloat162float(se.bias[col0 + j]);
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) red_add_v4_f32(dst_row + col0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (col0 + j < N) {
            float x = __uint_as_float(r[j]);
            if (se.bias) x += __bfloat162float(se.bias[col0 + j]);
            atomicAdd(dst_row + col0 + j, x);
          }
        }
      }
    }
  } else {
    const long long label = (row_ok && le.labels) ? le.labels[row] : -1;
    float mx = -INFINITY, sum = 0.f;
    float best_key = -INFINITY, best_logit = 0.f;
    int best_idx = -1;
    const bool sampling = le.samp_key != nullptr;
    const int step = le.step_ptr ? *le.step_ptr : 0;
    const int suppress = (le.suppress_col >= 0 && step < le.suppress_until) ? le.suppress_col : -1;
    const unsigned long long seed = le.seed + (le.seed_ptr ? (unsigned long long)(*le.seed_ptr) : 0ull) +
                                    0x632BE59BD9B4E019ull * (unsigned long long)(step + 1);
    const bool greedy = le.inv_temperature <= 0.f;
    const bool vec_bias = (reinterpret_cast<uintptr_t>(le.bias) & 15) == 0;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 16) {
      uint32_t r[16];
      tmem_ld16(taddr_row + c, r);
      tmem_ld_wait();
      const int col0 = n0 + c;
      if (!row_ok || col0 >= N) continue;
      float z[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = __uint_as_float(r[j]);
      const int valid = min(16, N - col0);  // < 16 only in the last chunk of a row
      if (le.bias) {
        if (valid == 16 && vec_bias) {
          float b[16];
          load16_bf16(le.bias + col0, b);
#pragma unroll
          for (int j = 0; j < 16; ++j) z[j] += b[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < valid) z[j] += __bfloat162float(le.bias[col0 + j]);
        }
      }
      // rare per-chunk fix-ups, resolved with one range test each instead of per-element compares
      const int so = suppress - col0, lo = (int)(label - col0);
      if (valid < 16 || (so >= 0 && so < 16) || (label >= col0 && lo < 16)) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j == lo && label >= col0 && j < valid && j != so) le.label_logit[row] = z[j];
          if (j >= valid || j == so) z[j] = -INFINITY;
        }
      }
      float cmax = z[0];
#pragma unroll
      for (int j = 1; j < 16; ++j) cmax = fmaxf(cmax, z[j]);
      if (cmax > mx) { sum *= __expf(mx - cmax); mx = cmax; }   // exp(-inf - finite) = 0 handles the first chunk
      if (mx > -INFINITY) {
#pragma unroll
        for (int j = 0; j < 16; ++j) sum += __expf(z[j] - mx);
      }
      if (sampling) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float key = z[j];
          if (!greedy && key > -INFINITY)
            key = z[j] * le.inv_temperature - __logf(-__logf(uniform01(seed, (unsigned)row, (unsigned)(col0 + j))));
          if (key > best_key) { best_key = key; best_logit = z[j]; best_idx = col0 + j; }
        }
      }
    }
    if (row_ok) {
      const size_t o = (size_t)row * le.n_tiles + part_idx;
      le.part_max[o] = mx;
      le.part_sum[o] = sum;
      if (sampling) { le.samp_key[o] = best_key; le.samp_logit[o] = best_logit; le.samp_idx[o] = best_idx; }
    }
  }
}

// Persistent, warp-specialised kernel.  Each CTA walks output tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...
// (M-fastest order: CTAs that run concurrently share the same B panel in L2).  The accumulator is double-buffered in
// TMEM (2 x BN fp32 columns), so the epilogue of tile i overlaps the TMA/MMA mainloop of tile i+1.
// AMN / BMN: the operand is MN-major in global memory (A given as [K, M], B as [K, N], row-major) — the layouts the backward
// GEMMs need (dX = dY · W reads W [N, K] as an MN-major B; dW = dYᵀ · X reads both operands MN-major), so no transposes are
// ever materialised.  Such a tile is fetched as 64-column TMA boxes ([64 k rows] x [64 MN elements]) laid out chunk by chunk.
// TBM: rows per tile, 128 or 64.  UMMA M = 64 keeps its accumulator in lanes 0-15 of each 32-lane TMEM quadrant
// (row m -> lane (m % 16) + 32 * (m / 16)), so the epilogue warps use half of their lanes; it exists for the decode path,
// where M = batch is a single 128-row tile and halving the A panel per CTA halves each CTA's (redundant) operand traffic.
// FP8: both operands are e4m3 bytes (kind::f8f6f4, UMMA K = 32): a k-block is still one 128-byte swizzle row = 128 elements.
template <int BN, int EPI, int AMN = 0, int BMN = 0, int TBM = 128, int FP8 = 0>  // EPI: 0 store, 1 lm-head, 2 reduce-scatter
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ MapArray maps_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_out, int M, int N, int K, int stages, int rows_per_map,
               StoreEpilogue se, LMHeadEpilogue le, ReduceScatterEpilogue re, int k_splits, AReady ar) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int BKE = FP8 ? 2 * BK : BK;            // elements per k-block (128 bytes per row either way)
  static_assert(!FP8 || (!AMN && !BMN), "fp8 operands are K-major");
  constexpr uint32_t A_BYTES = TBM * BK * 2;
  static_assert(TBM == 128 || (TBM == 64 && !AMN), "64-row tiles: K-major A only");
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t MN_CHUNK = BK * 64 * 2;          // one [64 k] x [64 MN] box of an MN-major operand (8 KB)
  static_assert(!(BMN && BN < 64), "MN-major B needs BN >= 64");
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;   // TMEM columns of one accumulator buffer
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  constexpr int HALF = BN / 2;                        // columns per epilogue warp (two warps share a TMEM lane quadrant)

  // 1024-byte aligned base (dynamic smem alignment is only guaranteed to 16 B)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int CW = HALF < 64 ? HALF : 64;           // columns per staged output chunk (TMA-store epilogue)
  uint8_t* stage_out = smem + (size_t)stages * STAGE_BYTES;  // NUM_EPI_WARPS x 4 KB, 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_out + NUM_EPI_WARPS * STG_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + BKE - 1) / BKE;
  const int m_tiles = (M + TBM - 1) / TBM;
  const int n_tiles = (N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  // split-K (k_splits > 1, EPI 2 only): work item w = (split, tile) covers k-blocks [split*per, min(nkb, (split+1)*per)) and
  // its partial product is added into an fp32 accumulator, so a skinny output with a very long contraction still fills the GPU
  const int total_work = total_tiles * k_splits;
  const int kb_per = (nkb + k_splits - 1) / k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps_a.m[0]);
    tma_prefetch_desc(&map_b);
    if (EPI == 0 && se.tma_store) tma_prefetch_desc(&map_out);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  uint32_t b_pre = 0;  // leading k-blocks of this CTA's first tile whose B boxes are already in flight (producer thread only)
  if constexpr (!BMN) {
    if (ar.b_static && warp == 0 && (int)blockIdx.x < total_work) {   // whole warp: see the note on warp-uniform issue loops below
      const int tile = blockIdx.x % total_tiles, split = blockIdx.x / total_tiles;
      const int kb_lo = split * kb_per, kb_hi = min(nkb, kb_lo + kb_per);
      const int n0 = (tile / m_tiles) * BN;
      for (int kb = kb_lo; kb < kb_hi && (int)b_pre < stages; ++kb, ++b_pre) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[b_pre], STAGE_BYTES);
          tma_load_2d(smem + (size_t)b_pre * STAGE_BYTES + A_BYTES, &map_b, &full_bar[b_pre], kb * BKE, n0);
        }
        __syncwarp();
      }
    }
  }
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch, static-weight tiles) overlapped with the
  // tail of the previous kernel; from here on we touch memory it produced.  Let our own successor start its prologue.
  griddep_wait();
  griddep_launch();

  // The producer and MMA-issuer loops are executed by ALL 32 lanes of their warp; only the tcgen05 / TMA instructions sit
  // under `elect_one()`.  When the whole loop ran in one thread (`if (lane == 0)`), every descriptor lived in that thread's
  // vector registers and ptxas wrapped each UTCHMMA / UTMALDG (which take uniform registers) in an R2UR + ELECT + BRA.U.ANY
  // "waterfall": ~103 cycles per MMA issue regardless of its shape (scripts/umma_probe.py) — more than the tensor-core time of
  // any atom narrower than N = 256.  With warp-uniform control flow the operands stay in uniform registers and consecutive
  // UTCHMMAs issue back to back.
  if (warp == 0) {
    {
      uint32_t it = 0;  // k-block counter across all of this CTA's tiles (ring position)
      int rs = 0;
      uint32_t rph = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int tile = w % total_tiles, split = w / total_tiles;
        const int kb_lo = split * kb_per, kb_hi = min(nkb, kb_lo + kb_per);
        const int m0 = (((tile % m_tiles) + ar.m_rot) % m_tiles) * TBM, n0 = (tile / m_tiles) * BN;
        if (ar.flags) {  // the shard holding these rows has landed in the local gathered buffer
          if (lane == 0) {
            const uint32_t* f = ar.flags + m0 / ar.rows_per_flag;
            { const long long t0 = clock64(); while (ld_relaxed_sys(f) != ar.epoch) { __nanosleep(32); spin_guard(t0); } }   // relaxed polls + one fence (no CCTL.IVALL per poll)
            fence_acq_rel_sys();
            __threadfence();
          }
          __syncwarp();
        }
        // which peer's copy of A holds this M-tile (all-gather -> GEMM); plain GEMMs have a single map
        const int a_map = m0 / rows_per_map;
        const int a_row = m0 - a_map * rows_per_map;
        const CUtensorMap* map_a_ptr = &maps_a.m[a_map];
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it, ring_next(rs, rph, stages)) {
          const int s = rs;
          const uint32_t phase = rph;
          uint8_t* a_dst = smem + (size_t)s * STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_BYTES;
          const bool b_inflight = it < b_pre;  // armed + B issued before the PDL wait (first pass over fresh stages)
          if (!b_inflight) mbar_wait(&empty_bar[s], phase ^ 1);
          if (elect_one()) {
            if (!b_inflight) mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
            if constexpr (AMN) {
#pragma unroll
              for (int ch = 0; ch < TBM / 64; ++ch) tma_load_2d(a_dst + ch * MN_CHUNK, map_a_ptr, &full_bar[s], m0 + ch * 64, kb * BK);
            } else {
              tma_load_2d(a_dst, map_a_ptr, &full_bar[s], kb * BKE, a_row);
            }
            if constexpr (BMN) {
#pragma unroll
              for (int ch = 0; ch < BN / 64; ++ch) tma_load_2d(b_dst + ch * MN_CHUNK, &map_b, &full_bar[s], n0 + ch * 64, kb * BK);
            } else {
              if (!b_inflight) tma_load_2d(b_dst, &map_b, &full_bar[s], kb * BKE, n0);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = FP8 ? umma_idesc(0, 0, TBM, BN) : umma_idesc(1, 1, TBM, BN, AMN, BMN);  // fp8: e4m3 x e4m3
      uint32_t it = 0, tcount = 0, rph = 0;
      int rs = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tcount) {
        const int split = w / total_tiles;
        const int kb_lo = split * kb_per, kb_hi = min(nkb, kb_lo + kb_per);
        const uint32_t as = tcount & 1, aphase = (tcount >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);  // epilogue drained this accumulator buffer
        tc_fence_after_sync();
        const uint32_t tmem_acc = tmem_base + as * ACC_COLS;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it, ring_next(rs, rph, stages)) {
          const int s = rs;
          const uint32_t phase = rph;
          mbar_wait(&full_bar[s], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint64_t da = AMN ? umma_desc_mn_sw128(a_addr, MN_CHUNK) : umma_desc_k_sw128(a_addr);
          const uint64_t db = BMN ? umma_desc_mn_sw128(a_addr + A_BYTES, MN_CHUNK) : umma_desc_k_sw128(a_addr + A_BYTES);
          // one UMMA consumes 16 k: K-major = 32 bytes inside the 128-byte swizzle row (+2 in the addr>>4 field);
          // MN-major = 16 rows of 128 bytes = 2048 bytes (+128)
          constexpr uint32_t A_STEP = AMN ? (16 * 128) >> 4 : 2, B_STEP = BMN ? (16 * 128) >> 4 : 2;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if constexpr (FP8) umma_fp8(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb > kb_lo || k > 0) ? 1u : 0u);
              else umma_bf16(tmem_acc, da + A_STEP * k, db + B_STEP * k, idesc, (kb > kb_lo || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);
            if (kb == kb_hi - 1) umma_commit(&tmem_full_bar[as]);   // same thread as the MMAs it tracks
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; warp w may only touch TMEM lanes [32*(w&3), 32*(w&3)+32), and the two warps of
    // a quadrant split the tile's columns
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int c_lo = half * HALF, c_hi = c_lo + HALF;
    uint8_t* stg = stage_out + (size_t)(warp - 2) * STG_BYTES;
    uint32_t tcount = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tcount) {
      const int tile = w % total_tiles;
      const uint32_t as = tcount & 1, aphase = (tcount >> 1) & 1;
      const int m_idx = ((tile % m_tiles) + ar.m_rot) % m_tiles, n_idx = tile / m_tiles;
      constexpr int ROWS_PER_WARP = TBM / 4;  // 32, or 16 valid lanes per quadrant for 64-row tiles
      const int row = m_idx * TBM + q * ROWS_PER_WARP + lane;
      const bool row_ok = row < M && lane < ROWS_PER_WARP;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after_sync();
      const uint32_t taddr_row = tmem_base + as * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      epilogue_cols<EPI, CW>(taddr_row, row, row_ok, n_idx * BN, c_lo, c_hi, N, n_idx * 2 + half, se, le, re, stg, &map_out,
                             m_idx * TBM + q * ROWS_PER_WARP, lane);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
    }
    if (EPI == 0 && se.tma_store && lane == 0) tma_store_wait_all();  // bulk stores must land before the CTA retires
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- cluster split-K GEMM
// Decode-shaped GEMMs (M <= 128: ONE row tile, a few dozen column tiles, K = 768 ... 8192) are latency bound: every CTA of
// the plain kernel pulls the whole [128, K] activation panel through the ~50 B/clk L2 -> SMEM path while most SMs idle.
// Here a CLUSTER of S CTAs owns one 128 x BN output tile and CTA r multiplies k-blocks [r*per, (r+1)*per): each SM moves
// 1/S of the panel.  The fp32 partial tiles of ranks 1..S-1 go straight from TMEM (tcgen05.ld) into rank 0's shared memory
// through distributed shared memory (st.shared::cluster), one split cluster barrier publishes them, and rank 0 adds them to
// its own accumulator inside the ordinary store epilogue (bias / activation / residual / folded norm / TMA store).
// One tile per cluster, no persistence: the grid is (tiles * S) CTAs and the host only picks S so that all clusters are
// co-resident (cudaOccupancyMaxActiveClusters).
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_csk_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_out, int M, int N, int K, int stages, StoreEpilogue se,
                int b_static) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;
  constexpr int HALF = BN / 2;
  constexpr int CW = HALF < 64 ? HALF : 64;
  constexpr uint32_t PART_BYTES = BN * 512;            // one fp32 128 x BN partial tile
  constexpr uint32_t WARP_PART = (HALF / 16) * 2048;   // the slice of it one epilogue warp owns

  const uint32_t S = cluster_nctarank(), rank = cluster_ctarank();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + (size_t)stages * STAGE_BYTES;
  uint8_t* parts = stage_out + NUM_EPI_WARPS * STG_BYTES;       // (S - 1) partial tiles (used in rank 0 only)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(parts + (size_t)(S - 1) * PART_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + BK - 1) / BK;
  const int kb_per = (nkb + (int)S - 1) / (int)S;
  const int kb_lo = (int)rank * kb_per, kb_hi = min(nkb, kb_lo + kb_per);   // host guarantees kb_lo < kb_hi
  const int n_idx = blockIdx.x / S, n0 = n_idx * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (se.tma_store && rank == 0) tma_prefetch_desc(&map_out);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ACC_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  int b_pre = 0;  // weight tiles already in flight when the PDL wait returns (see AReady::b_static)
  if (b_static && warp == 0) {   // warp-uniform loop, elected lane issues (see gemm_tn_kernel)
    for (int kb = kb_lo; kb < kb_hi && b_pre < stages; ++kb, ++b_pre) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[b_pre], STAGE_BYTES);
        tma_load_2d(smem + (size_t)b_pre * STAGE_BYTES + A_BYTES, &map_b, &full_bar[b_pre], kb * BK, n0);
      }
      __syncwarp();
    }
  }
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    {
      int rs = 0;
      uint32_t rph = 0;
      for (int kb = kb_lo, it = 0; kb < kb_hi; ++kb, ++it, ring_next(rs, rph, stages)) {
        const int s = rs;
        uint8_t* a_dst = smem + (size_t)s * STAGE_BYTES;
        if (it >= b_pre) mbar_wait(&empty_bar[s], rph ^ 1);
        if (elect_one()) {
          if (it >= b_pre) {
            mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_load_2d(a_dst + A_BYTES, &map_b, &full_bar[s], kb * BK, n0);
          }
          tma_load_2d(a_dst, &map_a, &full_bar[s], kb * BK, 0);
        }
        __syncwarp();
      }
    }
    __syncwarp();
    cluster_arrive_release();
    cluster_wait_acquire();
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc(1, 1, BM, BN);
      int rs = 0;
      uint32_t rph = 0;
      for (int kb = kb_lo, it = 0; kb < kb_hi; ++kb, ++it, ring_next(rs, rph, stages)) {
        const int s = rs;
        mbar_wait(&full_bar[s], rph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t da = umma_desc_k_sw128(a_addr), db = umma_desc_k_sw128(a_addr + A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (kb == kb_hi - 1) umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
    }
    __syncwarp();
    cluster_arrive_release();
    cluster_wait_acquire();
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int c_lo = half * HALF, c_hi = c_lo + HALF;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t my_part = smem_u32(parts) + (uint32_t)(warp - 2) * WARP_PART;
    if (rank != 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after_sync();
      const uint32_t dst = mapa_shared(my_part + (rank - 1) * PART_BYTES + (uint32_t)lane * 16u, 0);
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr_row + c, r);
        tmem_ld_wait();
        const uint32_t chunk = dst + (uint32_t)((c - c_lo) >> 4) * 2048u;
#pragma unroll
        for (int j = 0; j < 4; ++j) st_cluster_v4(chunk + (uint32_t)j * 512u, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      }
      tc_fence_before_sync();
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
    } else {
      cluster_arrive_release();
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after_sync();
      cluster_wait_acquire();  // every other rank's partial tile has landed in `parts`
      const int row = q * 32 + lane;
      LMHeadEpilogue le{};
      ReduceScatterEpilogue re{};
      epilogue_cols<0, CW, 1>(taddr_row, row, row < M, n0, c_lo, c_hi, N, 0, se, le, re, stage_out + (size_t)(warp - 2) * STG_BYTES,
                              &map_out, q * 32, lane, my_part, (int)S - 1, PART_BYTES);
      tc_fence_before_sync();
      if (se.tma_store && lane == 0) tma_store_wait_all();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, ACC_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- CTA-pair GEMM
// 256 x BN output tiles computed by a CLUSTER OF TWO CTAs with tcgen05.mma.cta_group::2 (UMMA M = 256).  CTA r of the pair
// stages rows [128 r, 128 r + 128) of the A tile and rows [BN/2 r, BN/2 r + BN/2) of the B tile; the leader's single MMA
// thread multiplies the full 256 x BN x 64 block from both CTAs' shared memory, each CTA's TMEM receives its own 128 rows.
// Per SM and k-block that is 16 KB of A + 16 KB of B for 128 x 256 x 64 MACs: 128 flop per byte from L2 instead of the 85
// of the single-CTA 128 x 256 tile — the large-GEMM regime of this part is bound by exactly that L2 -> SMEM path.
//
// Barriers (offsets identical in both CTAs):
//   full[s]        leader only — expect_tx(2 x stage bytes); both CTAs' TMA loads complete_tx on it (peer-bit-cleared address)
//   empty[s]       one per CTA  — armed by the leader's tcgen05.commit ... multicast::cluster (mask 0b11)
//   tmem_full[a]   one per CTA  — same multicast commit after the last k-block of a tile
//   tmem_empty[a]  leader only — 2 x NUM_EPI_WARPS arrivals (the peer's epilogue warps arrive remotely via mapa)
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_out, int M, int N, int K, int stages, StoreEpilogue se, AReady ar) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int HB = BN / 2;                          // B rows staged by each CTA
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = HB * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t ACC_COLS = BN;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;        // double-buffered accumulator
  constexpr int HALF = BN / 2;
  constexpr int CW = HALF < 64 ? HALF : 64;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + (size_t)stages * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_out + NUM_EPI_WARPS * STG_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int nkb = (K + BK - 1) / BK;
  const int m_tiles = (M + 2 * BM - 1) / (2 * BM);   // 256-row tiles
  const int n_tiles = (N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (se.tma_store) tma_prefetch_desc(&map_out);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything remote targets them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    {
      uint32_t it = 0, rph = 0;
      int rs = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int m0 = (((tile % m_tiles) + ar.m_rot) % m_tiles) * 2 * BM + (int)cta * BM;  // this CTA's 128 rows of A
        const int n0 = (tile / m_tiles) * BN + (int)cta * HB;                               // ... and its half of the B tile
        if (ar.flags) {  // all-gather -> GEMM: the shard holding these rows has landed (see AReady)
          if (lane == 0) {
            const uint32_t* f = ar.flags + m0 / ar.rows_per_flag;
            { const long long t0 = clock64(); while (ld_relaxed_sys(f) != ar.epoch) { __nanosleep(32); spin_guard(t0); } }   // relaxed polls + one fence (no CCTL.IVALL per poll)
            fence_acq_rel_sys();
            __threadfence();
          }
          __syncwarp();
        }
        for (int kb = 0; kb < nkb; ++kb, ++it, ring_next(rs, rph, stages)) {
          const int s = rs;
          const uint32_t phase = rph;
          mbar_wait(&empty_bar[s], phase ^ 1);
          uint8_t* a_dst = smem + (size_t)s * STAGE_BYTES;
          if (elect_one()) {   // warp-uniform loop, elected lane issues (see gemm_tn_kernel)
            if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
            tma_load_2d_2sm(a_dst, &map_a, &full_bar[s], kb * BK, m0);
            tma_load_2d_2sm(a_dst + A_BYTES, &map_b, &full_bar[s], kb * BK, n0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc(1, 1, 2 * BM, BN);
      uint32_t it = 0, tcount = 0, rph = 0;
      int rs = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
        const uint32_t as = tcount & 1, aphase = (tcount >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);  // both CTAs' epilogues drained this accumulator buffer
        tc_fence_after_sync();
        const uint32_t tmem_acc = tmem_base + as * ACC_COLS;
        for (int kb = 0; kb < nkb; ++kb, ++it, ring_next(rs, rph, stages)) {
          const int s = rs;
          const uint32_t phase = rph;
          mbar_wait(&full_bar[s], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(a_addr);
          const uint64_t db = umma_desc_k_sw128(a_addr + A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              umma_bf16_2sm(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[s], 0b11);      // frees slot s in BOTH CTAs
            if (kb == nkb - 1) umma_commit_2sm(&tmem_full_bar[as], 0b11);   // both epilogues may read their half
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int c_lo = half * HALF, c_hi = c_lo + HALF;
    uint8_t* stg = stage_out + (size_t)(warp - 2) * STG_BYTES;
    LMHeadEpilogue le{};
    ReduceScatterEpilogue re{};
    // the leader's tmem_empty barrier, as seen from this CTA
    const uint32_t empty_remote0 = mapa_shared(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t empty_remote1 = mapa_shared(smem_u32(&tmem_empty_bar[1]), 0);
    uint32_t tcount = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
      const uint32_t as = tcount & 1, aphase = (tcount >> 1) & 1;
      const int m_idx = ((tile % m_tiles) + ar.m_rot) % m_tiles, n_idx = tile / m_tiles;
      const int row0 = m_idx * 2 * BM + (int)cta * BM + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < M;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after_sync();
      const uint32_t taddr_row = tmem_base + as * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      epilogue_cols<0, CW>(taddr_row, row, row_ok, n_idx * BN, c_lo, c_hi, N, 0, se, le, re, stg, &map_out, row0, lane);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(as ? empty_remote1 : empty_remote0);
    }
    if (se.tma_store && lane == 0) tma_store_wait_all();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // nobody frees TMEM / shared memory the other CTA may still touch
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// Merge the per-tile partials of the LM-head epilogue.  One warp per row.
__global__ void lmhead_reduce_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                     const float* __restrict__ label_logit, const long long* __restrict__ labels,
                                     const float* __restrict__ samp_key, const float* __restrict__ samp_logit,
                                     const int* __restrict__ samp_idx, int M, int n_tiles, float* __restrict__ lse_out,
                                     float* __restrict__ logprob_out, long long* __restrict__ token_out,
                                     float* __restrict__ token_logprob_out) {
  griddep_wait();
  griddep_launch();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const size_t base = (size_t)row * n_tiles;
  // ONE pass, four independent loads in flight per lane: online (max, sum-exp) merge per lane, then across the warp (the
  // two-pass version was a chain of ~50 dependent L2 round trips per row: ~100 us per optimizer step at 1280 x 786 partials)
  float mx = -INFINITY, s = 0.f;
  for (int t0 = lane; t0 < n_tiles; t0 += 128) {
    float pm[4], ps[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + 32 * u;
      pm[u] = t < n_tiles ? part_max[base + t] : -INFINITY;
      ps[u] = t < n_tiles ? part_sum[base + t] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pm[u] == -INFINITY) continue;
      const float nm = fmaxf(mx, pm[u]);
      s = s * __expf(mx - nm) + ps[u] * __expf(pm[u] - nm);
      mx = nm;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(mx, om);
    s = (mx == -INFINITY ? 0.f : s * __expf(mx - nm)) + (om == -INFINITY ? 0.f : os * __expf(om - nm));
    mx = nm;
  }
  const float lse = mx + __logf(s);
  if (lane == 0) {
    if (lse_out) lse_out[row] = lse;
    if (logprob_out) logprob_out[row] = (labels && labels[row] >= 0) ? label_logit[row] - lse : 0.f;
  }
  if (samp_key) {
    float bk = -INFINITY, bl = 0.f;
    int bi = -1;
    for (int t = lane; t < n_tiles; t += 32) {
      const float k = samp_key[base + t];
      if (k > bk) { bk = k; bl = samp_logit[base + t]; bi = samp_idx[base + t]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const float ol = __shfl_xor_sync(0xffffffffu, bl, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ok > bk || (ok == bk && oi >= 0 && (bi < 0 || oi < bi))) { bk = ok; bl = ol; bi = oi; }
    }
    if (lane == 0) {
      token_out[row] = bi;
      token_logprob_out[row] = bl - lse;
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Row-major bf16 matrix [rows, cols] with row pitch ld (elements); box = box_rows x box_cols elements, swizzle span =
// box_cols * 2 bytes (128 / 64 / 32).  Operand loads use 64-column (128-byte) boxes.
static bool make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                     int box_cols = BK, int elem_bytes = 2) {
  using Key = std::tuple<const void*, long long, long long, long long, int, int, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key{ptr, rows, cols, ld, box_rows, box_cols, elem_bytes};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *map = it->second; return true; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const int row_bytes = box_cols * elem_bytes;
  const CUtensorMapSwizzle sw = row_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *map;
  return true;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static int pick_bn(int M, int N) {
  // Operand-traffic model: every SM streams (BM + BN) rows per k-block for each of its tiles, and the L2 -> SMEM path is
  // the limiter, so time ~ rounds x (BM + BN) with rounds = ceil(tiles / SMs).  Pick the width that minimises it (ties go
  // to the wider tile: fewer, fatter MMAs and less epilogue per flop).  E.g. 1792 x 768: 128-wide gives 84 tiles in one
  // round (cost 256) where 64-wide needs two rounds of 168 tiles (cost 384).
  const long long m_tiles = (M + BM - 1) / BM;
  const int sms = num_sms();
  int best = 32;
  long long best_cost = -1;
  for (int bn : {256, 128, 64, 32}) {
    if (bn > 32 && bn / 2 >= N) continue;  // do not pad a narrow output into a much wider tile
    const long long tiles = m_tiles * ((N + bn - 1) / bn);
    const long long rounds = (tiles + sms - 1) / sms;
    const long long cost = rounds * (BM + bn);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// While set, every K-major GEMM launched from this process may fetch its B operand ahead of the programmatic-dependent-launch
// wait.  Only valid when the kernel that precedes each GEMM in its stream never writes B: the rollout engine sets it around
// the capture of its decode / prefill graphs (weights are read-only there) and clears it afterwards.
static bool g_static_b = false;

template <int BN, int EPI, int AMN = 0, int BMN = 0, int TBM = 128, int FP8 = 0>
static cudaError_t launch(const MapArray& ma, const CUtensorMap& mb, const CUtensorMap& mo, int M, int N, int K,
                          int rows_per_map, const StoreEpilogue& se, const LMHeadEpilogue& le,
                          const ReduceScatterEpilogue& re, cudaStream_t stream, int k_splits = 1, AReady ar = AReady{}) {
  constexpr int stage_bytes = TBM * BK * 2 + BN * BK * 2;
  constexpr int fixed_bytes = NUM_EPI_WARPS * STG_BYTES + 1024 /*alignment slack*/ + 512 /*barriers*/;
  const int nkb = (K + (FP8 ? 2 * BK : BK) - 1) / (FP8 ? 2 * BK : BK);
  int stages = (227 * 1024 - fixed_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > nkb) stages = nkb < 2 ? 2 : nkb;
  const size_t smem = (size_t)stages * stage_bytes + fixed_bytes;
  auto kern = gemm_tn_kernel<BN, EPI, AMN, BMN, TBM, FP8>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long tiles = (long long)((N + BN - 1) / BN) * ((M + TBM - 1) / TBM) * k_splits;
  ar.b_static = (g_static_b && !AMN && !BMN && ar.flags == nullptr) ? 1 : 0;
  dim3 grid((unsigned)(tiles < num_sms() ? tiles : num_sms()));  // persistent: one CTA per SM walks the tile list
  return launch_kernel(kern, grid, dim3(NUM_THREADS), smem, stream, ma, mb, mo, M, N, K, stages, rows_per_map, se, le, re, k_splits, ar);
}

// Route a bf16 [M, N] output (row pitch ldo) through the TMA-store epilogue when the pitch allows a tensor map.
static bool setup_tma_store(StoreEpilogue& se, CUtensorMap* mo, int M, int N, int bn) {
  static const bool disabled = getenv("B200_GEMM_DIRECT_STORE") != nullptr;
  se.tma_store = 0;
  if (disabled || se.out_f32 || se.debug_nostore || (se.ldo & 7) || (reinterpret_cast<uintptr_t>(se.out) & 15)) return true;
  const int cw = bn / 2 < 64 ? bn / 2 : 64;
  if (!make_map(mo, se.out, M, N, se.ldo, 32, cw)) return false;
  se.tma_store = 1;
  return true;
}

static bool g_pdl = true;
bool pdl_enabled() { return g_pdl; }
void set_pdl_enabled(bool on) { g_pdl = on; }

}  // namespace b200

using namespace b200;

extern "C" void b200_set_pdl(int on) { set_pdl_enabled(on != 0); }
extern "C" void b200_set_static_weights(int on) { g_static_b = on != 0; }
extern "C" int b200_get_static_weights() { return g_static_b ? 1 : 0; }
extern "C" int b200_get_pdl() { return pdl_enabled() ? 1 : 0; }

// act: 0 none, 1 gelu_tanh, 2 gelu_erf, 3 relu, 4 silu.  Requirements: K % 8 == 0, lda/ldb % 8 == 0, 16B-aligned A/B.
// ---- CTA-pair launch (shared by the plain and the all-gather-flagged entry points)
static bool cta_pair_eligible(int M, int N, int K, const void* out, long long ldo, bool forced) {
  static const bool allow = getenv("B200_GEMM_NO_2CTA") == nullptr;
  const long long t2 = (long long)((M + 255) / 256) * ((N + 255) / 256);
  if (!(allow || forced) || (ldo % 8) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  // measured: +6 % at K >= 4096, neutral at K = 768 (run22)
  return forced || getenv("B200_GEMM_2CTA") != nullptr || (t2 >= num_sms() && K >= 1024);
}

static int launch_cta_pair(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb, StoreEpilogue s2,
                           AReady ar, cudaStream_t stream) {
  CUtensorMap m2a, m2b, m2o{};
  if (!make_map(&m2a, A, M, K, lda, BM) || !make_map(&m2b, B, N, K, ldb, 128)) return -1;
  if (!setup_tma_store(s2, &m2o, M, N, 256)) return -1;
  constexpr int BN2 = 256;
  constexpr int stage_bytes = BM * BK * 2 + (BN2 / 2) * BK * 2;
  constexpr int fixed_bytes = NUM_EPI_WARPS * STG_BYTES + 1024 + 512;
  int stages = (227 * 1024 - fixed_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  const int nkb2 = (K + BK - 1) / BK;
  if (stages > nkb2) stages = nkb2 < 2 ? 2 : nkb2;
  auto kern = gemm_2cta_kernel<BN2>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -5;
    configured = true;
  }
  const long long t2 = (long long)((M + 255) / 256) * ((N + 255) / 256);
  long long pairs = t2 < num_sms() / 2 ? t2 : num_sms() / 2;
  dim3 grid((unsigned)(2 * pairs));
  const size_t smem = (size_t)stages * stage_bytes + fixed_bytes;
  return (int)launch_kernel_cluster(kern, grid, dim3(NUM_THREADS), smem, stream, 2u, m2a, m2b, m2o, M, N, K, stages, s2, ar);
}

// ---- cluster split-K launch (decode shapes, see gemm_csk_kernel)
struct CskPlan {
  int bn = 0, S = 0, stages = 0;
  size_t smem = 0;
};

template <int BN>
static int csk_max_clusters(int S, size_t smem) {
  static std::map<std::pair<int, size_t>, int> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find({S, smem});
  if (it != cache.end()) return it->second;
  auto kern = gemm_csk_kernel<BN>;
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(S * 64));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
  }
  cudaGetLastError();  // a failed query must not poison the stream's error state
  cache[{S, smem}] = n;
  return n;
}

// Cost model: kilobytes one CTA pulls through the L2 -> SMEM path (the A panel dominates: 16 KB per k-block) plus a fixed
// charge for the cluster barrier and the partial-tile exchange.  `forced` ignores the "must save 30 %" threshold.
static bool csk_plan(int M, int N, int K, bool forced, CskPlan* plan) {
  const bool off = getenv("B200_GEMM_NO_CSK") != nullptr;  // read per launch: the A/B scripts toggle it inside one process
  if ((off && !forced) || M > BM) return false;
  const int nkb = (K + BK - 1) / BK;
  if (nkb < 4) return false;
  const int sms = num_sms();
  const int base_bn = pick_bn(M, N);
  double best = forced ? 1e30 : 0.7 * nkb * (16.0 + base_bn / 8.0);
  bool found = false;
  int want_bn = 0, want_s = 0;  // B200_CSK_FORCE="bn,S": tuning override (scripts/bench_chain.py sweeps it)
  if (const char* f = getenv("B200_CSK_FORCE")) {
    if (sscanf(f, "%d,%d", &want_bn, &want_s) == 2) best = 1e30;
    else want_bn = want_s = 0;
  }
  for (int bn : {32, 64}) {
    const int tiles = (N + bn - 1) / bn;
    if (want_bn && bn != want_bn) continue;
    for (int S = 8; S >= 2; --S) {
      if (want_s && S != want_s) continue;
      if (tiles * S > sms) continue;
      const int per = (nkb + S - 1) / S;
      if ((nkb + per - 1) / per != S) continue;  // no empty split: the cluster size IS the split count
      // measured in-chain (run35): the 64-wide tile halves the number of leaders, whose serial epilogue then costs more than
      // the narrower tile's extra A traffic (QKV: 6.5 us at 32x2 vs 7.3 at 64x3; fc2: 9.3 vs 10.7) — hence the surcharge
      const double cost = per * (16.0 + bn / 8.0) + 14.0 + (S - 1) + (bn == 64 ? 30.0 : 0.0);
      if (cost >= best) continue;
      const size_t stage_bytes = (size_t)BM * BK * 2 + (size_t)bn * BK * 2;
      const size_t fixed = (size_t)NUM_EPI_WARPS * STG_BYTES + (size_t)(S - 1) * bn * 512 + 1024 + 512;
      int stages = (int)((227 * 1024 - fixed) / stage_bytes);
      if (stages > 8) stages = 8;
      if (stages > per) stages = per;
      if (stages < (per < 2 ? per : 2)) continue;
      const size_t smem = (size_t)stages * stage_bytes + fixed;
      const int resident = bn == 32 ? csk_max_clusters<32>(S, smem) : csk_max_clusters<64>(S, smem);
      if (resident < tiles) continue;  // every cluster must be co-resident: a second wave would double the latency
      best = cost;
      *plan = CskPlan{bn, S, stages, smem};
      found = true;
    }
  }
  return found;
}

template <int BN>
static int launch_csk(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb, StoreEpilogue se,
                      const CskPlan& p, cudaStream_t stream) {
  CUtensorMap ma, mb, mo{};
  if (!make_map(&ma, A, M, K, lda, BM) || !make_map(&mb, B, N, K, ldb, BN)) return -1;
  if (!setup_tma_store(se, &mo, M, N, BN)) return -1;
  const int tiles = (N + BN - 1) / BN;
  return (int)launch_kernel_cluster(gemm_csk_kernel<BN>, dim3((unsigned)(tiles * p.S)), dim3(NUM_THREADS), p.smem, stream,
                                    (unsigned)p.S, ma, mb, mo, M, N, K, p.stages, se, g_static_b ? 1 : 0);
}

struct LnFold {
  const float* stats_in = nullptr;  // [M, 2] (sum, sum sq) of the raw A rows
  const float* c1 = nullptr;        // [N]
  float eps = 1e-5f;
  int rms = 0;
  float* stats_out = nullptr;       // [M, 2] accumulated for the next folded norm
};
static int gemm_bf16_impl(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                          long long ldo, const void* bias, const void* residual, long long ldr, const float* col_scale,
                          float alpha, int act, int out_f32, int force_bn, const LnFold& ln, cudaStream_t stream);

extern "C" int b200_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                              long long ldo, const void* bias, const void* residual, long long ldr, const float* col_scale,
                              float alpha, int act, int out_f32, int force_bn, cudaStream_t stream) {
  return gemm_bf16_impl(A, B, out, M, N, K, lda, ldb, ldo, bias, residual, ldr, col_scale, alpha, act, out_f32, force_bn, LnFold{},
                        stream);
}

// GEMM with a folded LayerNorm/RMSNorm on its input and/or row-moment accumulation on its output (see StoreEpilogue).
// Requires N % 16 == 0 (decode shapes); ln_stats_in / ln_c1 / stats_out may each be null.
extern "C" int b200_gemm_bf16_ln(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                                 long long ldo, const void* bias, const void* residual, long long ldr, int act,
                                 const float* ln_stats_in, const float* ln_c1, float ln_eps, int ln_rms, float* stats_out,
                                 cudaStream_t stream) {
  if (N % 16) return -4;
  LnFold ln;
  ln.stats_in = ln_stats_in; ln.c1 = ln_c1; ln.eps = ln_eps; ln.rms = ln_rms; ln.stats_out = stats_out;
  return gemm_bf16_impl(A, B, out, M, N, K, lda, ldb, ldo, bias, residual, ldr, nullptr, 1.0f, act, 0, 0, ln, stream);
}

static int gemm_bf16_impl(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                          long long ldo, const void* bias, const void* residual, long long ldr, const float* col_scale,
                          float alpha, int act, int out_f32, int force_bn, const LnFold& ln, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int bn = force_bn > 0 ? force_bn : pick_bn(M, N);
  // CTA-pair kernel for the large-GEMM regime: plenty of 256x256 tiles and nothing but a plain bf16 store epilogue
  {
    const bool force2 = force_bn == -2;
    if ((force_bn == 0 || force2) && !out_f32 && !col_scale && !ln.stats_in && !ln.stats_out &&
        cta_pair_eligible(M, N, K, out, ldo, force2)) {
      StoreEpilogue s2{out, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual, nullptr, ldo, ldr, alpha, act, 0};
      return launch_cta_pair(A, B, M, N, K, lda, ldb, s2, AReady{}, stream);
    }
  }
  // cluster split-K for single-row-tile (decode) shapes: force_bn == -3 requests it explicitly
  if (force_bn == 0 || force_bn == -3) {
    CskPlan cp;
    if (csk_plan(M, N, K, force_bn == -3, &cp)) {
      StoreEpilogue sc{out, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual, col_scale, ldo, ldr, alpha, act, out_f32};
      sc.ln_stats = ln.stats_in;
      sc.ln_c1 = ln.c1;
      sc.ln_inv_k = 1.0f / (float)K;
      sc.ln_eps = ln.eps;
      sc.ln_rms = ln.rms;
      sc.stats_out = ln.stats_out;
      return cp.bn == 32 ? launch_csk<32>(A, B, M, N, K, lda, ldb, sc, cp, stream)
                         : launch_csk<64>(A, B, M, N, K, lda, ldb, sc, cp, stream);
    }
  }
  // 64-row tiles when even 128x32 tiles leave more than half of the SMs idle (decode: M = batch <= 128)
  static const bool allow_bm64 = getenv("B200_GEMM_BM64") != nullptr;  // opt-in: measured neutral on the decode path (run21)
  const long long tiles128 = (long long)((M + 127) / 128) * ((N + bn - 1) / bn);
  const long long tiles64 = (long long)((M + 63) / 64) * ((N + bn - 1) / bn);
  const bool bm64 = allow_bm64 && bn <= 64 && M > 64 && tiles128 * 2 <= num_sms() && tiles64 <= num_sms() && force_bn >= 0;
  if (force_bn < 0) return -6;  // an explicitly requested CTA-pair launch was not possible
  MapArray ma{};
  CUtensorMap mb;
  if (!make_map(&ma.m[0], A, M, K, lda, bm64 ? 64 : BM) || !make_map(&mb, B, N, K, ldb, bn)) return -1;
  StoreEpilogue se{out, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual, col_scale, ldo, ldr, alpha, act, out_f32};
  static const bool nostore = getenv("B200_GEMM_NOSTORE") != nullptr;
  se.debug_nostore = nostore ? 1 : 0;
  se.ln_stats = ln.stats_in;
  se.ln_c1 = ln.c1;
  se.ln_inv_k = 1.0f / (float)K;
  se.ln_eps = ln.eps;
  se.ln_rms = ln.rms;
  se.stats_out = ln.stats_out;
  CUtensorMap mo{};
  if (bm64) se.tma_store = 0;  // 64-row tiles keep 16 rows per epilogue warp: direct stores
  else if (!setup_tma_store(se, &mo, M, N, bn)) return -1;
  LMHeadEpilogue le{};
  ReduceScatterEpilogue re{};
  const int rpm = 1 << 30;
  cudaError_t e;
  if (bm64) {
    if (bn == 64) e = launch<64, 0, 0, 0, 64>(ma, mb, mo, M, N, K, rpm, se, le, re, stream);
    else e = launch<32, 0, 0, 0, 64>(ma, mb, mo, M, N, K, rpm, se, le, re, stream);
    return (int)e;
  }
  switch (bn) {
    case 256: e = launch<256, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 128: e = launch<128, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 64: e = launch<64, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    default: e = launch<32, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
  }
  return (int)e;
}

// fp8 (e4m3) GEMM for the rollout path: out[M, N] (bf16) = act((A8[M, K] . B8[N, K]^T) * row_scale[m] * col_scale[n] + bias[n])
// + residual.  A8 / B8 are e4m3 bytes, K-major, K % 16 == 0 and row pitches % 16 == 0 (TMA).
extern "C" int b200_gemm_fp8(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                             long long ldo, const float* row_scale, const float* col_scale, const void* bias,
                             const void* residual, long long ldr, int act, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int bn = pick_bn(M, N);
  MapArray ma{};
  CUtensorMap mb;
  if (!make_map(&ma.m[0], A, M, K, lda, BM, 128, 1) || !make_map(&mb, B, N, K, ldb, bn, 128, 1)) return -1;
  StoreEpilogue se{out, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual, col_scale, ldo, ldr, 1.0f, act, 0};
  se.row_scale = row_scale;
  se.ln_inv_k = 1.0f;
  CUtensorMap mo{};
  if (!setup_tma_store(se, &mo, M, N, bn)) return -1;
  LMHeadEpilogue le{};
  ReduceScatterEpilogue re{};
  const int rpm = 1 << 30;
  cudaError_t e;
  switch (bn) {
    case 256: e = launch<256, 0, 0, 0, 128, 1>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 128: e = launch<128, 0, 0, 0, 128, 1>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 64: e = launch<64, 0, 0, 0, 128, 1>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    default: e = launch<32, 0, 0, 0, 128, 1>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
  }
  return (int)e;
}

// Split-K plan for skinny outputs with a long contraction (dX of the LM head: [1280 x 768] with K = 50257).  Small tiles
// would fill the SMs but are L2-bandwidth bound (a 128x64 tile moves 42 flop per byte); instead keep the widest tile and cut
// K so that there are about two work items per SM.  Returns the number of splits (1 = none) and the tile width.
static int splitk_plan(int M, int N, int K, int* bn_out) {
  int bn = pick_bn(M, N);
  if (bn < 64) bn = 64;
  const int nkb = (K + BK - 1) / BK;
  const int sms = num_sms();
  const long long m_tiles = (M + BM - 1) / BM;
  int splits = 1;
  // K >= 8192 only: at K = 2304 / 3072 (the dX GEMMs of a GPT-2 block) the memset + red.add + finalize sequence cost more
  // than the split saved (38 / 42 us against 26 us unsplit and 21 us for cuBLAS, run33)
  if (nkb >= 128 && m_tiles * ((N + bn - 1) / bn) < 2LL * sms) {
    bn = N >= 256 ? 256 : (N >= 128 ? 128 : 64);
    const long long tiles = m_tiles * ((N + bn - 1) / bn);
    long long s = (2LL * sms + tiles - 1) / tiles;
    if (s > nkb / 8) s = nkb / 8;  // at least 8 k-blocks per work item
    if (s > 64) s = 64;
    if (s >= 2) {
      const int per = (nkb + (int)s - 1) / (int)s;  // no empty trailing split
      splits = (nkb + per - 1) / per;
    }
    if (splits < 2) { splits = 1; bn = pick_bn(M, N) < 64 ? 64 : pick_bn(M, N); }
  }
  if (bn_out) *bn_out = bn;
  return splits;
}
extern "C" int b200_gemm_splitk_plan(int M, int N, int K) { return splitk_plan(M, N, K, nullptr); }

__global__ void splitk_finalize_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ out, long long rows, int N,
                                       long long ldo, int accumulate) {
  const long long total = rows * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;
    __nv_bfloat16* o = out + r * ldo + (i - r * N);
    *o = __float2bfloat16(accumulate ? acc[i] + __bfloat162float(*o) : acc[i]);  // accumulate: out += product (wgrad into .grad)
  }
}

// General-layout GEMM for the backward pass: out[m, n] = sum_k A(m, k) · B(n, k) where
//   a_mn = 0: A is [M, K] row-major (K contiguous);   a_mn = 1: A is given as [K, M] row-major (M contiguous)
//   b_mn = 0: B is [N, K] row-major (K contiguous);   b_mn = 1: B is given as [K, N] row-major (N contiguous)
// lda / ldb are the row pitches of the arrays as stored.  bf16 (or fp32) output.
// k_splits > 1 (from b200_gemm_splitk_plan) needs `ws`: a ZEROED fp32 [M, N] buffer the partial products are added into
// (red.global.add.v4.f32 from the epilogue); it is then converted into `out` (or is the result itself when out_f32).
// accumulate != 0 (bf16 out only): out += product — the weight-gradient GEMM adds straight into the parameter's .grad (Apex's
// gradient_accumulation_fusion): the existing value rides in as the epilogue's residual operand, or is added by the split-K finalize.
extern "C" int b200_gemm_bf16_ex(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                                 long long ldo, int a_mn, int b_mn, int out_f32, int k_splits, float* ws, int accumulate,
                                 cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (accumulate && out_f32) return -3;
  if (!a_mn && !b_mn && k_splits <= 1)
    return b200_gemm_bf16(A, B, out, M, N, K, lda, ldb, ldo, nullptr, accumulate ? out : nullptr, accumulate ? ldo : 0, nullptr,
                          1.0f, ACT_NONE, out_f32, 0, stream);
  int bn = pick_bn(M, N);
  if (bn < 64) bn = 64;  // MN-major tiles are fetched in 64-element boxes
  if (k_splits > 1 && ws != nullptr) {
    int planned_bn = bn;
    if (splitk_plan(M, N, K, &planned_bn) > 1) bn = planned_bn;  // split-K keeps wide tiles
  }
  if (k_splits > 1) {  // normalise: every split must own at least one k-block
    const int nkb = (K + BK - 1) / BK;
    if (k_splits > nkb) k_splits = nkb;
    const int per = (nkb + k_splits - 1) / k_splits;
    k_splits = (nkb + per - 1) / per;
  }
  const bool split = k_splits > 1 && ws != nullptr;
  MapArray ma{};
  CUtensorMap mb;
  const bool ok_a = a_mn ? make_map(&ma.m[0], A, K, M, lda, 64, 64) : make_map(&ma.m[0], A, M, K, lda, BM);
  const bool ok_b = b_mn ? make_map(&mb, B, K, N, ldb, 64, 64) : make_map(&mb, B, N, K, ldb, bn);
  if (!ok_a || !ok_b) return -1;
  StoreEpilogue se{out, nullptr, accumulate ? (const __nv_bfloat16*)out : nullptr, nullptr, ldo, accumulate ? ldo : 0, 1.0f,
                   ACT_NONE, out_f32};
  CUtensorMap mo{};
  LMHeadEpilogue le{};
  ReduceScatterEpilogue re{};
  if (split) {
    re.acc[0] = ws;
    re.ldacc = N;
    re.rows_per_rank = 1 << 30;
    se = StoreEpilogue{};
  } else {
    k_splits = 1;
    if (!setup_tma_store(se, &mo, M, N, bn)) return -1;
  }
  const int rpm = 1 << 30;
  cudaError_t e = cudaErrorInvalidValue;
#define B200_EX_LAUNCH(BN_, EPI_)                                                                                    \
  (a_mn && b_mn ? launch<BN_, EPI_, 1, 1>(ma, mb, mo, M, N, K, rpm, se, le, re, stream, k_splits)                   \
   : a_mn       ? launch<BN_, EPI_, 1, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream, k_splits)                   \
   : b_mn       ? launch<BN_, EPI_, 0, 1>(ma, mb, mo, M, N, K, rpm, se, le, re, stream, k_splits)                   \
                : launch<BN_, EPI_, 0, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream, k_splits))
  if (split) {
    switch (bn) {
      case 256: e = B200_EX_LAUNCH(256, 2); break;
      case 128: e = B200_EX_LAUNCH(128, 2); break;
      default: e = B200_EX_LAUNCH(64, 2); break;
    }
    if (e == cudaSuccess && !out_f32) {
      long long blocks = ((long long)M * N + 255) / 256;
      if (blocks > num_sms() * 8) blocks = num_sms() * 8;
      splitk_finalize_kernel<<<(int)blocks, 256, 0, stream>>>(ws, (__nv_bfloat16*)out, M, N, ldo, accumulate);
      e = cudaGetLastError();
    }
  } else {
    switch (bn) {
      case 256: e = B200_EX_LAUNCH(256, 0); break;
      case 128: e = B200_EX_LAUNCH(128, 0); break;
      default: e = B200_EX_LAUNCH(64, 0); break;
    }
  }
#undef B200_EX_LAUNCH
  return (int)e;
}

// LM-head backward, fused: out[M, N] (bf16, row pitch ldo) = ((n == labels[m]) - exp(A.B^T + bias - lse[m])) * grad[m].
// The logits are recomputed tile by tile and turned into d-logits in the epilogue (never stored as logits).
extern "C" int b200_lmhead_dlogits_bf16(const void* A, const void* B, void* out, int M, int N, int K, long long lda,
                                        long long ldb, long long ldo, const void* bias, const long long* labels,
                                        const float* lse, const float* grad, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int bn = pick_bn(M, N);
  MapArray ma{};
  CUtensorMap mb;
  if (!make_map(&ma.m[0], A, M, K, lda, BM) || !make_map(&mb, B, N, K, ldb, bn)) return -1;
  StoreEpilogue se{out, (const __nv_bfloat16*)bias, nullptr, nullptr, ldo, 0, 1.0f, ACT_NONE, 0, lse, grad, labels};
  CUtensorMap mo{};
  if (!setup_tma_store(se, &mo, M, N, bn)) return -1;
  LMHeadEpilogue le{};
  ReduceScatterEpilogue re{};
  const int rpm = 1 << 30;
  cudaError_t e;
  switch (bn) {
    case 256: e = launch<256, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 128: e = launch<128, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 64: e = launch<64, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    default: e = launch<32, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
  }
  return (int)e;
}

// GEMM over a LOCAL gathered A [M, K] whose row shards become valid one by one (all-gather overlapped with the GEMM): see
// AReady.  flags: uint32 [M / rows_per_flag] in device memory; rows_per_flag % 128 == 0.
extern "C" int b200_gemm_flagged_bf16(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                                      long long ldo, const void* bias, int act, const uint32_t* flags, uint32_t epoch,
                                      int rows_per_flag, int first_chunk, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (rows_per_flag % BM) return -3;
  if (rows_per_flag % (2 * BM) == 0 && cta_pair_eligible(M, N, K, out, ldo, false)) {
    StoreEpilogue s2{out, (const __nv_bfloat16*)bias, nullptr, nullptr, ldo, 0, 1.0f, act, 0};
    return launch_cta_pair(A, B, M, N, K, lda, ldb, s2, AReady{flags, epoch, rows_per_flag, first_chunk * (rows_per_flag / (2 * BM))},
                           stream);
  }
  const int bn = pick_bn(M, N);
  MapArray ma{};
  CUtensorMap mb;
  if (!make_map(&ma.m[0], A, M, K, lda, BM) || !make_map(&mb, B, N, K, ldb, bn)) return -1;
  StoreEpilogue se{out, (const __nv_bfloat16*)bias, nullptr, nullptr, ldo, 0, 1.0f, act, 0};
  CUtensorMap mo{};
  if (!setup_tma_store(se, &mo, M, N, bn)) return -1;
  LMHeadEpilogue le{};
  ReduceScatterEpilogue re{};
  AReady ar{flags, epoch, rows_per_flag, first_chunk * (rows_per_flag / BM)};
  const int rpm = 1 << 30;
  cudaError_t e;
  switch (bn) {
    case 256: e = launch<256, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream, 1, ar); break;
    case 128: e = launch<128, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream, 1, ar); break;
    case 64: e = launch<64, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream, 1, ar); break;
    default: e = launch<32, 0>(ma, mb, mo, M, N, K, rpm, se, le, re, stream, 1, ar); break;
  }
  return (int)e;
}

// all-gather -> GEMM: A is sharded by rows over `world` peers (A_peers[r] = peer r's [M/world, K] shard, any of them may
// be a remote NVLink-mapped pointer); out[M, N] = act(concat_r(A_r) . B^T + bias).  rows_per_rank % 128 == 0.
extern "C" int b200_gemm_allgather_bf16(void* const* A_peers, int world, const void* B, void* out, int M, int N, int K,
                                        long long lda, long long ldb, long long ldo, const void* bias, int act,
                                        cudaStream_t stream) {
  if (world < 1 || world > MAX_TP || M % world) return -3;
  const int rows = M / world;
  if (rows % BM) return -3;
  const int bn = pick_bn(M, N);
  MapArray ma{};
  CUtensorMap mb;
  for (int r = 0; r < world; ++r)
    if (!make_map(&ma.m[r], A_peers[r], rows, K, lda, BM)) return -1;
  if (!make_map(&mb, B, N, K, ldb, bn)) return -1;
  StoreEpilogue se{out, (const __nv_bfloat16*)bias, nullptr, nullptr, ldo, 0, 1.0f, act, 0};
  CUtensorMap mo{};
  if (!setup_tma_store(se, &mo, M, N, bn)) return -1;
  LMHeadEpilogue le{};
  ReduceScatterEpilogue re{};
  cudaError_t e;
  switch (bn) {
    case 256: e = launch<256, 0>(ma, mb, mo, M, N, K, rows, se, le, re, stream); break;
    case 128: e = launch<128, 0>(ma, mb, mo, M, N, K, rows, se, le, re, stream); break;
    case 64: e = launch<64, 0>(ma, mb, mo, M, N, K, rows, se, le, re, stream); break;
    default: e = launch<32, 0>(ma, mb, mo, M, N, K, rows, se, le, re, stream); break;
  }
  return (int)e;
}

// GEMM -> reduce-scatter: every rank computes its partial product A[M, K_local] . B[N, K_local]^T and adds rows
// [r*M/world, (r+1)*M/world) into peer r's fp32 accumulator acc_peers[r] ([M/world, ldacc]) straight from the epilogue.
extern "C" int b200_gemm_reduce_scatter_bf16(const void* A, const void* B, float* const* acc_peers, int world, int M, int N,
                                             int K, long long lda, long long ldb, long long ldacc, const void* bias,
                                             cudaStream_t stream) {
  if (world < 1 || world > MAX_TP || M % world) return -3;
  int bn = pick_bn(M, N);
  if (bn > 128) bn = 128;
  MapArray ma{};
  CUtensorMap mb;
  if (!make_map(&ma.m[0], A, M, K, lda, BM) || !make_map(&mb, B, N, K, ldb, bn)) return -1;
  StoreEpilogue se{};
  CUtensorMap mo{};
  se.bias = (const __nv_bfloat16*)bias;
  LMHeadEpilogue le{};
  ReduceScatterEpilogue re{};
  for (int r = 0; r < world; ++r) re.acc[r] = acc_peers[r];
  re.ldacc = ldacc;
  re.rows_per_rank = M / world;
  const int rpm = 1 << 30;
  cudaError_t e;
  switch (bn) {
    case 128: e = launch<128, 2>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 64: e = launch<64, 2>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    default: e = launch<32, 2>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
  }
  return (int)e;
}

// Sum the `world` bf16 partial slots of a staged reduce-scatter (+ bias + residual) -> bf16.  8 elements per thread.
__global__ void stage_reduce_kernel(const __nv_bfloat16* __restrict__ stage, int world, long long slot_elems,
                                    const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* __restrict__ residual,
                                    __nv_bfloat16* __restrict__ out, long long rows, int N, long long ldstage, long long ldr,
                                    long long ldo) {
  const int nv = N >> 3;
  const long long total = rows * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i - r * nv) * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int w = 0; w < world; ++w) {
      const uint4 v = *reinterpret_cast<const uint4*>(stage + w * slot_elems + r * ldstage + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
    }
    if (bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += __bfloat162float(bias[c + j]);
    }
    if (residual) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += __bfloat162float(residual[r * ldr + c + j]);
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + r * ldo + c) = o;
  }
}

extern "C" int b200_stage_reduce(const void* stage, int world, const void* bias, const void* residual, void* out, long long rows,
                                 int N, long long ldstage, long long ldr, long long ldo, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (N % 8 || ldstage % 8 || ldo % 8) return -3;
  long long blocks = (rows * (N / 8) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  stage_reduce_kernel<<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)stage, world, rows * ldstage,
                                                       (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual,
                                                       (__nv_bfloat16*)out, rows, N, ldstage, ldr, ldo);
  return (int)cudaGetLastError();
}

// NVLS reduce-scatter tail of GEMM -> reduce-scatter: every rank's partial product [M, N] sits in a symmetric buffer that is also
// mapped as one multicast address range; the owner of a row block issues multimem.ld_reduce over it — the NVSwitch sums the
// `world` copies in fp32 and returns one bf16x8 vector, so each output byte crosses this GPU's links once and no staging slots or
// per-peer loops exist.  `mc` already points at this rank's first row (col0 selects a column window for the split variant).
__global__ void mc_reduce_rows_kernel(const __nv_bfloat16* mc, const __nv_bfloat16* __restrict__ bias,
                                      const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out, long long rows,
                                      int ncols, int col0, long long ldp, long long ldr, long long ldo) {
  const int nv = ncols >> 3;
  const long long total = rows * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = col0 + (int)(i - r * nv) * 8;
    const uint4 v = multimem_ld_reduce_add_bf16x8(mc + r * ldp + c);
    uint4 o = v;
    if (bias || residual) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] = f.x; acc[2 * j + 1] = f.y; }
      if (bias) {
        const uint4 bv = *reinterpret_cast<const uint4*>(bias + c);
        const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&bv);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(bh[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
      }
      if (residual) {
        const uint4 rv = *reinterpret_cast<const uint4*>(residual + r * ldr + c);
        const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(rh[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
      }
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    }
    *reinterpret_cast<uint4*>(out + r * ldo + c) = o;
  }
}

extern "C" int b200_mc_reduce_rows(const void* mc, const void* bias, const void* residual, void* out, long long rows, int ncols,
                                   int col0, long long ldp, long long ldr, long long ldo, int blocks, cudaStream_t stream) {
  if (rows <= 0 || ncols <= 0) return 0;
  if (ncols % 8 || col0 % 8 || ldp % 8 || ldo % 8 || (residual && ldr % 8)) return -3;
  long long want = (rows * (ncols / 8) + 255) / 256;
  const long long cap = blocks > 0 ? blocks : 148 * 4;
  if (want > cap) want = cap;
  mc_reduce_rows_kernel<<<(int)want, 256, 0, stream>>>((const __nv_bfloat16*)mc, (const __nv_bfloat16*)bias,
                                                       (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, rows, ncols, col0, ldp,
                                                       ldr, ldo);
  return (int)cudaGetLastError();
}

// GEMM -> staged reduce-scatter: out partial tiles (bf16) are stored into slot `rank` of every owner's staging buffer
// stage_peers[owner] ([world, M/world, ldstage] bf16, peer-mapped).  Follow with a barrier and b200_stage_reduce on the owner.
extern "C" int b200_gemm_stage_scatter_bf16(const void* A, const void* B, void* const* stage_peers, int world, int rank, int M,
                                            int N, int K, long long lda, long long ldb, long long ldstage, const void* bias,
                                            cudaStream_t stream) {
  if (world < 1 || world > MAX_TP || M % world) return -3;
  int bn = pick_bn(M, N);
  MapArray ma{};
  CUtensorMap mb, mo{};
  if (!make_map(&ma.m[0], A, M, K, lda, BM) || !make_map(&mb, B, N, K, ldb, bn)) return -1;
  StoreEpilogue se{};
  se.bias = (const __nv_bfloat16*)bias;
  LMHeadEpilogue le{};
  ReduceScatterEpilogue re{};
  for (int r = 0; r < world; ++r) re.stage[r] = (__nv_bfloat16*)stage_peers[r];
  re.ldstage = ldstage;
  re.src_rank = rank;
  re.rows_per_rank = M / world;
  const int rpm = 1 << 30;
  cudaError_t e;
  switch (bn) {
    case 256: e = launch<256, 2>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 128: e = launch<128, 2>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    case 64: e = launch<64, 2>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
    default: e = launch<32, 2>(ma, mb, mo, M, N, K, rpm, se, le, re, stream); break;
  }
  return (int)e;
}

// fp32 accumulator -> bf16 (+ bias + residual) after the reduce-scatter has completed on every peer
__global__ void rs_finalize_kernel(const float* __restrict__ acc, const __nv_bfloat16* __restrict__ bias,
                                   const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out, long long rows,
                                   int N, long long ldacc, long long ldr, long long ldo) {
  const long long total = rows * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;
    const int c = (int)(i - r * N);
    float v = acc[r * ldacc + c];
    if (bias) v += __bfloat162float(bias[c]);
    if (residual) v += __bfloat162float(residual[r * ldr + c]);
    out[r * ldo + c] = __float2bfloat16(v);
  }
}

extern "C" int b200_rs_finalize(const float* acc, const void* bias, const void* residual, void* out, long long rows, int N,
                                long long ldacc, long long ldr, long long ldo, cudaStream_t stream) {
  if (rows <= 0) return 0;
  long long blocks = (rows * N + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  rs_finalize_kernel<<<(int)blocks, 256, 0, stream>>>(acc, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual,
                                                     (__nv_bfloat16*)out, rows, N, ldacc, ldr, ldo);
  return (int)cudaGetLastError();
}

// number of (row, column-range) partials the fused LM head emits per row: two epilogue warps split every 128-wide tile
extern "C" int b200_lmhead_tiles(int N) { return 2 * ((N + 127) / 128); }

// Fused LM head.  Workspace (fp32/int32) of 5 * M * n_tiles + M elements is supplied by the caller:
//   part_max | part_sum | samp_key | samp_logit | samp_idx(int) | label_logit[M]
// Outputs (any may be null): lse[M], logprob[M] (of labels), token[M] + token_logprob[M] (when sample != 0).
extern "C" int b200_lmhead_bf16(const void* H, const void* W, int M, int N, int K, long long ldh, long long ldw,
                                const void* bias, const long long* labels, float* workspace, float* lse, float* logprob,
                                int sample, float temperature, unsigned long long seed, const long long* seed_ptr,
                                const int* step_ptr,
                                int suppress_col, int suppress_until, long long* token, float* token_logprob,
                                cudaStream_t stream) {
  if (M <= 0) return 0;
  constexpr int BN = 128;
  const int n_tiles = b200_lmhead_tiles(N);
  MapArray ma{};
  CUtensorMap mb;
  if (!make_map(&ma.m[0], H, M, K, ldh, BM) || !make_map(&mb, W, N, K, ldw, BN)) return -1;
  const size_t mt = (size_t)M * n_tiles;
  LMHeadEpilogue le{};
  le.bias = (const __nv_bfloat16*)bias;
  le.labels = labels;
  le.part_max = workspace;
  le.part_sum = workspace + mt;
  le.samp_key = sample ? workspace + 2 * mt : nullptr;
  le.samp_logit = workspace + 3 * mt;
  le.samp_idx = reinterpret_cast<int*>(workspace + 4 * mt);
  le.label_logit = workspace + 5 * mt;
  le.inv_temperature = temperature > 0.f ? 1.0f / temperature : 0.f;
  le.seed = seed;
  le.seed_ptr = seed_ptr;
  le.step_ptr = step_ptr;
  le.suppress_col = suppress_col;
  le.suppress_until = suppress_until;
  le.n_tiles = n_tiles;
  StoreEpilogue se{};
  CUtensorMap mo{};
  ReduceScatterEpilogue re{};
  cudaError_t e = launch<BN, 1>(ma, mb, mo, M, N, K, 1 << 30, se, le, re, stream);
  if (e != cudaSuccess) return (int)e;
  const int warps_per_block = 8;
  return (int)launch_kernel(lmhead_reduce_kernel, dim3((M + warps_per_block - 1) / warps_per_block),
                            dim3(warps_per_block * 32), 0, stream, (const float*)le.part_max, (const float*)le.part_sum,
                            (const float*)le.label_logit, labels, (const float*)le.samp_key, (const float*)le.samp_logit,
                            (const int*)le.samp_idx, M, n_tiles, lse, logprob, token, token_logprob);
}
