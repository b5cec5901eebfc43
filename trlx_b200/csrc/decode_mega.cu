// Persistent decode "megakernel" for sm_100a: ALL transformer blocks of one decode step in ONE launch.
//
// The kernel-per-op decode graph (norm -> QKV GEMM -> paged attention -> out-proj -> norm -> MLP up -> MLP down, x L) is bound
// by ~90 dependent launches per token, each a few microseconds of latency for almost no data.  Here the batch is cut into
// groups of 16 rows and every group is owned by ONE THREAD-BLOCK CLUSTER of 16 CTAs (one per SM, 8 clusters = 128 rows) that
// walks the whole layer stack without ever leaving the SMs:
//
//   * "swap-AB" tcgen05 tiles: the WEIGHT slice is the 64-row UMMA M operand (TMA, 128-byte swizzle, streamed through a deep
//     shared-memory ring by a producer thread that runs arbitrarily far ahead of the compute — weights do not depend on the
//     activations, so their HBM/L2 latency never sits on the token's critical path), the 16 batch rows are the UMMA N operand,
//     resident in shared memory for the whole GEMM; accumulators ([64 features x 16 rows] fp32) live in TMEM;
//   * CTA r of the cluster owns feature tiles r, r+16, ... of every GEMM; for QKV the tile is exactly one head's q / k / v, so
//     the paged attention of (16 rows x head) runs locally out of shared memory right after the three tiles of a head, with
//     the KV-cache append fused;
//   * LayerNorm / RMSNorm is computed by each CTA for its cluster's 16 rows while it builds the swizzled activation operand
//     (no norm kernels, no extra pass), bias / activation / residual live in the accumulator epilogue;
//   * phases are separated by a CLUSTER-scope barrier (one remote mbarrier arrive per CTA pair, ~0.2 us) instead of a kernel
//     boundary; clusters never synchronise with each other, so there is no grid barrier and no co-residency requirement.
//
// Reference being replaced: HF `generate()` driving one eager forward per token
// (trlx/trainer/accelerate_base_trainer.py:256-269, trlx/trainer/accelerate_ppo_trainer.py:277-290).
//
// Supported (everything else keeps the kernel-per-op graph): head_dim 64, hidden / ffn multiples of 64, serial residual,
// non-gated MLP, learned or no positional embedding inside the blocks (no rotary), full attention (no window),
// 16 * max(hidden, ffn) * 2 bytes of activation operand in shared memory (GPT-2 family, OPT).
#include <cstdio>
#include <mutex>

#include "ptx.cuh"

namespace b200 {

constexpr int DM_MAX_CS = 16;    // CTAs per cluster: chosen at launch (largest size with enough co-resident clusters)
constexpr int DM_ROWS = 16;      // batch rows per cluster = UMMA N
constexpr int DM_TM = 64;        // features per tile = UMMA M
constexpr int DM_BK = 64;        // k-block (one 128-byte swizzle row of bf16)
constexpr int DM_WORKERS = 8;    // worker warps (LN / copy / epilogue / attention)
constexpr int DM_THREADS = 64 + 32 * DM_WORKERS;
constexpr int DM_WTILE = DM_TM * DM_BK * 2;     // 8 KB weight tile
constexpr int DM_ATILE = DM_ROWS * DM_BK * 2;   // 2 KB activation k-block
constexpr int DM_SLOTS = 4;      // TMEM accumulator slots
constexpr int DM_NACC = 4;       // independent sub-accumulators per tile (k-step j of every k-block accumulates into sub-accumulator j):
                                 // consecutive tcgen05.mma on ONE accumulator serialise at the MMA latency (~150 clk measured for these
                                 // 64x16x16 atoms), four independent chains hide it; the epilogue adds the four
constexpr int DM_SLOT_COLS = DM_NACC * DM_ROWS;
constexpr int DM_TMEM_COLS = DM_SLOTS * DM_SLOT_COLS;   // 256
constexpr int DM_MAX_STAGES = 24;

static_assert(DM_NACC == 4 && DM_BK / 16 == DM_NACC, "one sub-accumulator per k-step of a k-block");

struct DmLayer {
  const __nv_bfloat16 *ln1_w, *ln1_b, *ln2_w, *ln2_b, *qkv_b, *out_b, *fc_b, *fc2_b;
  __nv_bfloat16 *kcache, *vcache;
};

struct DmParams {
  int B, H, F, nh, L;
  int act, rms;
  float eps, scale;
  int page_size, max_pages;
  const int* block_table;
  const int* seq_lens;
  __nv_bfloat16* x;       // [B, H] residual stream (in / out)
  __nv_bfloat16* a;       // [B, H] attention output
  __nv_bfloat16* mid;     // [B, F] MLP hidden
  const DmLayer* layers;
  const CUtensorMap* maps;  // [L * 4]: qkv, out, fc, fc2
  __nv_bfloat16* trunk_out; // optional: copy of x entering block `branch`  (row stride trunk_stride, column offset step * H)
  long long trunk_stride;
  const long long* step_ptr;
  int branch;
  int stages;
  const float* alibi;     // optional per-head slopes
  long long* timing;      // optional [16 ranks][L][16] clock64 stamps of cluster 0 (bring-up / profiling aid)
  int cluster_size;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(2000u)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * DM_WORKERS) : "memory"); }
__device__ __forceinline__ uint4 ldcg16(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldcg8(const void* p) {
  uint2 v;
  asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float dm_act(float v, int act) {
  switch (act) {
    case 1: return gelu_tanh(v);
    case 2: return gelu_erf(v);
    case 3: return fmaxf(v, 0.f);
    case 4: return silu(v);
    default: return v;
  }
}
// byte offset of the 16-byte chunk holding elements [8c, 8c+8) of row r inside the swizzled [16 x K] activation operand
__device__ __forceinline__ uint32_t act_chunk_off(int r, int c) {
  return (uint32_t)(c >> 3) * DM_ATILE + (uint32_t)r * 128u + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
}

// MAXC: 16-byte chunks of a residual row each lane keeps in registers while normalising it (hidden <= 256 * MAXC)
template <int MAXC>
__global__ void __launch_bounds__(DM_THREADS, 1) decode_mega_kernel(const DmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int H = p.H, F = p.F, nh = p.nh, stages = p.stages;
  const int KMAX = H > F ? H : F;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act_s = smem;                                                   // [KMAX/64][16][64] bf16, swizzled
  uint8_t* ring = act_s + (size_t)(KMAX / DM_BK) * DM_ATILE;              // stages x 8 KB (1024-aligned: ATILE = 2 KB, K/64 even)
  ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ring) + 1023) & ~uintptr_t(1023));
  float* qkv_s = reinterpret_cast<float*>(ring + (size_t)stages * DM_WTILE);  // [3][16][64] fp32
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(qkv_s + 3 * DM_ROWS * DM_TM);
  uint64_t* empty_bar = full_bar + DM_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + DM_MAX_STAGES;   // [DM_SLOTS]
  uint64_t* tempty_bar = tfull_bar + DM_SLOTS;       // [DM_SLOTS]
  uint64_t* act_bar = tempty_bar + DM_SLOTS;         // workers -> MMA: activation operand ready
  uint64_t* cl_bar = act_bar + 1;                    // cluster phase barrier (16 remote arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cl_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int DM_CS = (int)cluster_nctarank();
  const int group = blockIdx.x / DM_CS;              // which 16 rows of the batch
  const int row0 = group * DM_ROWS;
  const int rows_here = min(DM_ROWS, p.B - row0);
  const int nkH = H / DM_BK, nkF = F / DM_BK;
  const int tH = H / DM_TM, tF = F / DM_TM;          // feature tiles of the H- and F-wide outputs

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < DM_SLOTS; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], DM_WORKERS); }
    mbar_init(act_bar, DM_WORKERS);
    mbar_init(cl_bar, DM_CS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, DM_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // every CTA's cluster barrier must be initialised before any peer can arrive on it
  cluster_sync_all();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: weight tiles in schedule order
    // (the loop is warp-uniform, the elected lane issues: a single-thread loop costs ~100 cycles per UTMALDG / UTCHMMA in
    // register-to-uniform-register marshalling, see gemm_sm100.cu)
    {
      int s = 0;
      uint32_t ph = 0;   // ring position / pass parity kept incrementally (a runtime `%` / `/` per k-block is ~150 cycles of ALU)
      auto load = [&](const CUtensorMap* m, int row, int kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], DM_WTILE);
          tma_load_2d(ring + (size_t)s * DM_WTILE, m, &full_bar[s], kb * DM_BK, row);
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1; }
      };
      for (int l = 0; l < p.L; ++l) {
        const CUtensorMap* m = p.maps + (size_t)l * 4;
        for (int h = rank; h < nh; h += DM_CS)
          for (int part = 0; part < 3; ++part)
            for (int kb = 0; kb < nkH; ++kb) load(m + 0, part * H + h * DM_TM, kb);
        for (int t = rank; t < tH; t += DM_CS)
          for (int kb = 0; kb < nkH; ++kb) load(m + 1, t * DM_TM, kb);
        for (int t = rank; t < tF; t += DM_CS)
          for (int kb = 0; kb < nkH; ++kb) load(m + 2, t * DM_TM, kb);
        for (int t = rank; t < tH; t += DM_CS)
          for (int kb = 0; kb < nkF; ++kb) load(m + 3, t * DM_TM, kb);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
    {
      constexpr uint32_t idesc = umma_idesc(1, 1, DM_TM, DM_ROWS);
      uint32_t tc = 0, fills = 0, ph = 0;
      int s = 0;
      const uint32_t act_addr = smem_u32(act_s), ring_addr = smem_u32(ring);
      auto tile = [&](int nkb) {
        const uint32_t slot = tc % DM_SLOTS;
        mbar_wait(&tempty_bar[slot], ((tc / DM_SLOTS) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t tacc = tmem_base + slot * DM_SLOT_COLS;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint64_t da = umma_desc_k_sw128(ring_addr + (uint32_t)s * DM_WTILE);
          const uint64_t db = umma_desc_k_sw128(act_addr + (uint32_t)kb * DM_ATILE);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < DM_BK / 16; ++k) umma_bf16(tacc + k * DM_ROWS, da + 2 * k, db + 2 * k, idesc, kb > 0 ? 1u : 0u);
            umma_commit(&empty_bar[s]);
            if (kb == nkb - 1) umma_commit(&tfull_bar[slot]);
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        ++tc;
      };
      auto wait_act = [&]() {
        mbar_wait(act_bar, fills & 1);
        tc_fence_after_sync();
        ++fills;
      };
      for (int l = 0; l < p.L; ++l) {
        if ((int)rank < nh) {
          wait_act();
          for (int h = rank; h < nh; h += DM_CS)
            for (int part = 0; part < 3; ++part) tile(nkH);
        }
        if ((int)rank < tH) {
          wait_act();
          for (int t = rank; t < tH; t += DM_CS) tile(nkH);
        }
        if ((int)rank < tF) {
          wait_act();
          for (int t = rank; t < tF; t += DM_CS) tile(nkH);
        }
        if ((int)rank < tH) {
          wait_act();
          for (int t = rank; t < tH; t += DM_CS) tile(nkF);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ workers: operand fill, epilogues, attention
    const int wi = warp - 2;                    // 0..7
    const int wt = threadIdx.x - 64;            // 0..255
    const int q = warp & 3;                     // TMEM lane quadrant this warp may read
    const int ch = wi >> 2;                     // column half: batch rows [8 ch, 8 ch + 8)
    const int feat = q * 16 + lane;             // feature inside a tile (valid for lane < 16)
    uint32_t tc = 0, cphase = 0;
    const uint32_t cl_local = smem_u32(cl_bar);

    int stamp_l = 0, stamp_i = 0;
    auto stamp = [&]() {
      if (p.timing && group == 0 && wt == 0 && stamp_i < 16)
        p.timing[((size_t)rank * p.L + stamp_l) * 16 + stamp_i] = clock64();
      ++stamp_i;
    };
    auto cluster_phase = [&]() {   // all CTAs of the cluster finished their global-memory writes of this phase
      __threadfence();
      worker_sync();
      if (wt < DM_CS) mbar_arrive_cluster(mapa_shared(cl_local, (uint32_t)wt));
      // ONE thread polls: every cluster-scope acquire makes ptxas emit CCTL.IVALL (an L1 invalidate), and 256 threads spinning
      // on it starved the SM's whole load / store path (ncu: millions of CCTL.IVALL, every phase ~4x slower); the others
      // sleep on the hardware barrier
      if (wt == 0) mbar_wait_cluster(cl_bar, cphase & 1);
      worker_sync();
      ++cphase;
    };
    auto act_ready = [&]() {
      fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(act_bar);
    };
    // swizzled bf16 operand <- LayerNorm / RMSNorm of x rows (two rows per warp)
    auto fill_norm = [&](const __nv_bfloat16* w, const __nv_bfloat16* b) {
      const int nchunk = H >> 3;
      uint4 raw[2][MAXC];
      float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = wi * 2 + rr;
        const __nv_bfloat16* xr = p.x + (size_t)(row0 + r) * H;
#pragma unroll
        for (int u = 0; u < MAXC; ++u) {
          const int c = lane + 32 * u;
          raw[rr][u] = make_uint4(0, 0, 0, 0);
          if (c < nchunk && r < rows_here) raw[rr][u] = ldcg16(xr + c * 8);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
        for (int u = 0; u < MAXC; ++u) {
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[rr][u]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h2[j]);
            s1[rr] += f.x + f.y;
            s2[rr] += f.x * f.x + f.y * f.y;
          }
        }
        s1[rr] = warp_sum(s1[rr]);
        s2[rr] = warp_sum(s2[rr]);
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = wi * 2 + rr;
        const float mean = p.rms ? 0.f : s1[rr] / (float)H;
        const float var = fmaxf(s2[rr] / (float)H - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.eps);
#pragma unroll
        for (int u = 0; u < MAXC; ++u) {
          const int c = lane + 32 * u;
          if (c >= nchunk) continue;
          uint4 outv = make_uint4(0, 0, 0, 0);
          if (r < rows_here) {
            const uint4 wraw = *reinterpret_cast<const uint4*>(w + c * 8);
            uint4 braw = make_uint4(0, 0, 0, 0);
            if (b) braw = *reinterpret_cast<const uint4*>(b + c * 8);
            const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&raw[rr][u]);
            const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&wraw);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&braw);
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 xf = __bfloat1622float2(x2[j]), wf = __bfloat1622float2(w2[j]), bf = __bfloat1622float2(b2[j]);
              o2[j] = __floats2bfloat162_rn((xf.x - mean) * rstd * wf.x + bf.x, (xf.y - mean) * rstd * wf.y + bf.y);
            }
          }
          *reinterpret_cast<uint4*>(act_s + act_chunk_off(r, c)) = outv;
        }
      }
      act_ready();
    };
    // swizzled bf16 operand <- rows of a [B, K] global buffer
    auto fill_copy = [&](const __nv_bfloat16* src, int K) {
      const int nchunk = K >> 3, total = DM_ROWS * nchunk;
      // 8 independent 16-byte loads in flight per thread, then 8 shared-memory stores
      for (int i0 = wt; i0 < total; i0 += 8 * 32 * DM_WORKERS) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 32 * DM_WORKERS;
          const int r = i / nchunk, c = i - r * nchunk;
          v[u] = make_uint4(0, 0, 0, 0);
          if (i < total && r < rows_here) v[u] = ldcg16(src + (size_t)(row0 + r) * K + c * 8);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 32 * DM_WORKERS;
          const int r = i / nchunk, c = i - r * nchunk;
          if (i < total) *reinterpret_cast<uint4*>(act_s + act_chunk_off(r, c)) = v[u];
        }
      }
      act_ready();
    };
    // wait for tile `tc`'s accumulator, return this thread's 8 values (feature `feat`, rows 8 ch .. 8 ch + 7)
    auto take = [&](float (&v)[8]) {
      const uint32_t slot = tc % DM_SLOTS;
      mbar_wait(&tfull_bar[slot], (tc / DM_SLOTS) & 1);
      tc_fence_after_sync();
      uint32_t r[DM_NACC][8];
#pragma unroll
      for (int a = 0; a < DM_NACC; ++a)
        tmem_ld8(tmem_base + slot * DM_SLOT_COLS + a * DM_ROWS + ch * 8 + (static_cast<uint32_t>(q * 32) << 16), r[a]);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = (__uint_as_float(r[0][j]) + __uint_as_float(r[1][j])) + (__uint_as_float(r[2][j]) + __uint_as_float(r[3][j]));
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[slot]);
      ++tc;
    };

    griddep_wait();   // x / seq_lens come from the kernels before us (weights above were prefetched regardless)
    griddep_launch();

    for (int l = 0; l < p.L; ++l) {
      const DmLayer Lw = p.layers[l];   // by value: the pointers stay in registers across the asm memory barriers below
      stamp_l = l; stamp_i = 0;
      stamp();  // 0: layer start
      // ---- optional capture of the trunk activation entering block `branch`
      if (l == p.branch && p.trunk_out && rank == 0) {
        const long long step = p.step_ptr ? *p.step_ptr : 0;
        const int nchunk = H >> 3;
        for (int i = wt; i < rows_here * nchunk; i += 32 * DM_WORKERS) {
          const int r = i / nchunk, c = i - r * nchunk;
          const uint4 v = ldcg16(p.x + (size_t)(row0 + r) * H + c * 8);
          *reinterpret_cast<uint4*>(p.trunk_out + (size_t)(row0 + r) * p.trunk_stride + (size_t)step * H + c * 8) = v;
        }
      }
      // ---- phase 1: h = norm1(x); per head: q, k, v tiles -> local paged attention -> a[:, head]
      if ((int)rank < nh) {
        fill_norm(Lw.ln1_w, Lw.ln1_b);
        stamp();  // 1: norm1 operand built
        for (int h = rank; h < nh; h += DM_CS) {
          for (int part = 0; part < 3; ++part) {
            float v[8];
            take(v);
            if (lane < 16) {
              const float bias = Lw.qkv_b ? __bfloat162float(Lw.qkv_b[part * H + h * DM_TM + feat]) : 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j)   // rounded to bf16: what the kernel-per-op path and the KV cache hold
                qkv_s[(part * DM_ROWS + ch * 8 + j) * DM_TM + feat] = __bfloat162float(__float2bfloat16(v[j] + bias));
            }
          }
          worker_sync();
          stamp();  // 2: q/k/v tiles of the head in shared memory
          // attention: half-warp per row (16 lanes: one key each while scoring, 4 output dims each while mixing)
          {
            const int r = wi * 2 + (lane >> 4);
            const int l16 = lane & 15;
            const unsigned hmask = (lane >> 4) ? 0xffff0000u : 0x0000ffffu;
            const int row = row0 + r;
            const int len = (r < rows_here) ? p.seq_lens[row] : 0;
            const float* qs = qkv_s + (0 * DM_ROWS + r) * DM_TM;
            const float* ks = qkv_s + (1 * DM_ROWS + r) * DM_TM;
            const float* vs = qkv_s + (2 * DM_ROWS + r) * DM_TM;
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            if (len > 0) {
              const int last = len - 1;
              const int* bt = p.block_table + (size_t)row * p.max_pages;
              const float slope = p.alibi ? p.alibi[h] : 0.f;
              // append the new K / V (rounded to bf16 like the cache holds them)
              {
                const size_t slot = ((size_t)bt[last / p.page_size] * p.page_size + (last % p.page_size)) * nh + h;
                __nv_bfloat162 k01 = __floats2bfloat162_rn(ks[l16 * 4], ks[l16 * 4 + 1]);
                __nv_bfloat162 k23 = __floats2bfloat162_rn(ks[l16 * 4 + 2], ks[l16 * 4 + 3]);
                __nv_bfloat162 v01 = __floats2bfloat162_rn(vs[l16 * 4], vs[l16 * 4 + 1]);
                __nv_bfloat162 v23 = __floats2bfloat162_rn(vs[l16 * 4 + 2], vs[l16 * 4 + 3]);
                uint2 kk, vv;
                kk.x = *reinterpret_cast<uint32_t*>(&k01); kk.y = *reinterpret_cast<uint32_t*>(&k23);
                vv.x = *reinterpret_cast<uint32_t*>(&v01); vv.y = *reinterpret_cast<uint32_t*>(&v23);
                *reinterpret_cast<uint2*>(Lw.kcache + slot * DM_TM + l16 * 4) = kk;
                *reinterpret_cast<uint2*>(Lw.vcache + slot * DM_TM + l16 * 4) = vv;
              }
              float m = -INFINITY, lsum = 0.f;
              // page_size is a multiple of 16 (engine: 16): the 16 keys of a chunk share one page-table entry
              int page = bt[0];
              for (int t0 = 0; t0 < len; t0 += 16) {
                const int t = t0 + l16;
                const int page_next = (t0 + 16 < len) ? bt[(t0 + 16) / p.page_size] : 0;   // prefetched for the next chunk
                const size_t base = (((size_t)page * p.page_size + (t0 % p.page_size)) * nh + h) * DM_TM;
                const int nkeys = min(16, len - t0);
                // issue every global load of the chunk up front: this lane's key row (8 x 16 B) and, for each of the <= 16
                // keys, this lane's 4 output dims of the value row (8 B) — one memory round trip per chunk instead of 17
                uint4 kraw[8];
                uint2 vraw[16];
                if (t < last) {
                  const __nv_bfloat16* kp = Lw.kcache + base + (size_t)l16 * nh * DM_TM;
#pragma unroll
                  for (int c = 0; c < 8; ++c) kraw[c] = ldcg16(kp + c * 8);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (t0 + j < last) vraw[j] = ldcg8(Lw.vcache + base + (size_t)j * nh * DM_TM + l16 * 4);
                float sc = -INFINITY;
                if (t < last) {
                  float dot = 0.f;
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kraw[c]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float2 f = __bfloat1622float2(h2[j]);
                      dot += qs[c * 8 + 2 * j] * f.x + qs[c * 8 + 2 * j + 1] * f.y;
                    }
                  }
                  sc = dot * p.scale + slope * (float)t;
                }
                if (t0 + 16 > last) {  // the chunk holding the new key: its q . k_new is computed by all 16 lanes (4 dims each)
                  float part = 0.f;
#pragma unroll
                  for (int d = 0; d < 4; ++d) part += qs[l16 * 4 + d] * ks[l16 * 4 + d];
#pragma unroll
                  for (int o2 = 8; o2 > 0; o2 >>= 1) part += __shfl_xor_sync(hmask, part, o2);
                  if (t == last) sc = part * p.scale + slope * (float)t;
                }
                float mx = sc;
#pragma unroll
                for (int o2 = 8; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(hmask, mx, o2));
                const float mn = fmaxf(m, mx);
                const float corr = (m == -INFINITY) ? 0.f : __expf(m - mn);
                const float pr = (sc == -INFINITY) ? 0.f : __expf(sc - mn);
                float ps = pr;
#pragma unroll
                for (int o2 = 8; o2 > 0; o2 >>= 1) ps += __shfl_xor_sync(hmask, ps, o2);
                lsum = lsum * corr + ps;
                m = mn;
#pragma unroll
                for (int d = 0; d < 4; ++d) o[d] *= corr;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float pj = __shfl_sync(hmask, pr, (lane & 16) | j);
                  if (j < nkeys) {
                    float v4[4];
                    if (t0 + j < last) {
                      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vraw[j]);
                      const float2 f0 = __bfloat1622float2(h2[0]), f1 = __bfloat1622float2(h2[1]);
                      v4[0] = f0.x; v4[1] = f0.y; v4[2] = f1.x; v4[3] = f1.y;
                    } else {
#pragma unroll
                      for (int d = 0; d < 4; ++d) v4[d] = vs[l16 * 4 + d];
                    }
#pragma unroll
                    for (int d = 0; d < 4; ++d) o[d] += pj * v4[d];
                  }
                }
                page = page_next;
              }
              const float inv = 1.f / lsum;
#pragma unroll
              for (int d = 0; d < 4; ++d) o[d] *= inv;
            }
            if (r < rows_here) {
              __nv_bfloat162 o01 = __floats2bfloat162_rn(o[0], o[1]), o23 = __floats2bfloat162_rn(o[2], o[3]);
              uint2 ov;
              ov.x = *reinterpret_cast<uint32_t*>(&o01); ov.y = *reinterpret_cast<uint32_t*>(&o23);
              *reinterpret_cast<uint2*>(p.a + (size_t)row * H + h * DM_TM + l16 * 4) = ov;
            }
          }
          __syncwarp();
          worker_sync();   // qkv_s is rewritten by the next head's tiles
          stamp();  // 3: attention done
        }
      }
      cluster_phase();
      stamp();  // 4 (or 2): cluster barrier 1
      // ---- phase 2: x += a . Wo^T + bo
      if ((int)rank < tH) {
        fill_copy(p.a, H);
        stamp();  // 5: attention-output operand copied
        for (int t = rank; t < tH; t += DM_CS) {
          float v[8];
          take(v);
          if (lane < 16) {
            const int f = t * DM_TM + feat;
            const float bias = Lw.out_b ? __bfloat162float(Lw.out_b[f]) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = ch * 8 + j;
              if (r < rows_here) {
                __nv_bfloat16* px = p.x + (size_t)(row0 + r) * H + f;
                *px = __float2bfloat16(__bfloat162float(*px) + v[j] + bias);
              }
            }
          }
        }
      }
      stamp();  // 6: out-proj tiles done
      cluster_phase();
      stamp();  // 7: cluster barrier 2
      // ---- phase 3: mid = act(norm2(x) . Wfc^T + b)
      if ((int)rank < tF) {
        fill_norm(Lw.ln2_w, Lw.ln2_b);
        stamp();  // 8: norm2 operand built
        for (int t = rank; t < tF; t += DM_CS) {
          float v[8];
          take(v);
          if (lane < 16) {
            const int f = t * DM_TM + feat;
            const float bias = Lw.fc_b ? __bfloat162float(Lw.fc_b[f]) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = ch * 8 + j;
              if (r < rows_here) p.mid[(size_t)(row0 + r) * F + f] = __float2bfloat16(dm_act(v[j] + bias, p.act));
            }
          }
        }
      }
      stamp();  // 9: fc tiles done
      cluster_phase();
      stamp();  // 10: cluster barrier 3
      // ---- phase 4: x += mid . Wfc2^T + b
      if ((int)rank < tH) {
        fill_copy(p.mid, F);
        stamp();  // 11: mid operand copied
        for (int t = rank; t < tH; t += DM_CS) {
          float v[8];
          take(v);
          if (lane < 16) {
            const int f = t * DM_TM + feat;
            const float bias = Lw.fc2_b ? __bfloat162float(Lw.fc2_b[f]) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = ch * 8 + j;
              if (r < rows_here) {
                __nv_bfloat16* px = p.x + (size_t)(row0 + r) * H + f;
                *px = __float2bfloat16(__bfloat162float(*px) + v[j] + bias);
              }
            }
          }
        }
      }
      stamp();  // 12: fc2 tiles done
      cluster_phase();
      stamp();  // 13: cluster barrier 4
    }
  }
  // no CTA may exit while a peer can still arrive on its cluster barrier
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, DM_TMEM_COLS);
  }
}

typedef CUresult (*DmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DmEncodeFn dm_encode_fn() {
  static DmEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DmEncodeFn>(q);
  });
  return fn;
}

static size_t dm_smem_bytes(int H, int F, int stages) {
  const int kmax = H > F ? H : F;
  return 1024 + (size_t)(kmax / DM_BK) * DM_ATILE + 1024 + (size_t)stages * DM_WTILE + 3 * DM_ROWS * DM_TM * 4 +
         (2 * DM_MAX_STAGES + 2 * DM_SLOTS + 2) * 8 + 64;
}

}  // namespace b200

using namespace b200;

// Weight tensor map for the megakernel: row-major bf16 [rows, cols] (row pitch ld elements), 64 x 64 boxes, 128-byte swizzle.
extern "C" int b200_decode_mega_make_map(void* map_out, const void* w, long long rows, long long cols, long long ld) {
  DmEncodeFn fn = dm_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {DM_BK, DM_TM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

extern "C" int b200_decode_mega_layer_bytes() { return (int)sizeof(DmLayer); }

// Largest number of ring stages that fits (0 = the shape does not fit at all)
extern "C" int b200_decode_mega_stages(int H, int F) {
  int dev = 0, max_optin = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
  for (int s = DM_MAX_STAGES; s >= 4; --s)
    if (dm_smem_bytes(H, F, s) <= (size_t)max_optin) return s;
  return 0;
}

// How many `cs`-CTA clusters of this kernel can be co-resident (clusters are independent, this only decides waves).
static int dm_max_clusters(int H, int F, int cs) {
  const int stages = b200_decode_mega_stages(H, F);
  if (!stages) return 0;
  const size_t smem = dm_smem_bytes(H, F, stages);
  auto kern = H <= 1024 ? decode_mega_kernel<4> : decode_mega_kernel<8>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cs * 8);
  cfg.blockDim = dim3(DM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return n;
}
extern "C" int b200_decode_mega_max_clusters(int H, int F, int cs) { return dm_max_clusters(H, F, cs > 0 ? cs : DM_MAX_CS); }

// Cluster size for `groups` row groups: the largest size <= 16 whose clusters all fit at once (one wave) — a GPC that lost
// SMs to yield caps 16-CTA clusters at 7 on some parts, 12-CTA clusters fit 8+ — falling back to the size with most CTAs.
extern "C" int b200_decode_mega_cluster_size(int H, int F, int groups) {
  static int cache[64][2];
  static int ncache = 0;
  for (int i = 0; i < ncache; ++i) if (cache[i][0] == groups) return cache[i][1];
  int best = 8, best_ctas = 0;
  for (int cs = DM_MAX_CS; cs >= 4; cs -= 2) {
    const int n = dm_max_clusters(H, F, cs);
    if (n >= groups) { best = cs; best_ctas = 1 << 30; break; }
    if (n > 0 && n * cs > best_ctas) { best = cs; best_ctas = n * cs; }
  }
  if (ncache < 64) { cache[ncache][0] = groups; cache[ncache][1] = best; ++ncache; }
  return best;
}

extern "C" int b200_decode_mega(int B, int H, int F, int nh, int L, int act, int rms, float eps, float scale, int page_size,
                                int max_pages, const int* block_table, const int* seq_lens, void* x, void* a, void* mid,
                                const void* layers, const void* maps, void* trunk_out, long long trunk_stride,
                                const long long* step_ptr, int branch, const float* alibi, long long* timing,
                                int cluster_size, cudaStream_t stream) {
  if (B <= 0 || L <= 0) return 0;
  if (H % 64 || F % 64 || nh * DM_TM != H || H > 2048) return -2;
  const int stages = b200_decode_mega_stages(H, F);
  if (!stages) return -2;
  const size_t smem = dm_smem_bytes(H, F, stages);
  auto kern = H <= 1024 ? decode_mega_kernel<4> : decode_mega_kernel<8>;
  static size_t smem_set[2] = {0, 0};
  const int ki = H <= 1024 ? 0 : 1;
  if (smem > smem_set[ki]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -4;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -4;
    smem_set[ki] = smem;
  }
  DmParams p{};
  p.B = B; p.H = H; p.F = F; p.nh = nh; p.L = L; p.act = act; p.rms = rms; p.eps = eps; p.scale = scale;
  p.page_size = page_size; p.max_pages = max_pages; p.block_table = block_table; p.seq_lens = seq_lens;
  p.x = (__nv_bfloat16*)x; p.a = (__nv_bfloat16*)a; p.mid = (__nv_bfloat16*)mid;
  p.layers = (const DmLayer*)layers; p.maps = (const CUtensorMap*)maps;
  p.trunk_out = (__nv_bfloat16*)trunk_out; p.trunk_stride = trunk_stride; p.step_ptr = step_ptr; p.branch = branch;
  p.stages = stages; p.alibi = alibi; p.timing = timing;
  const int groups = (B + DM_ROWS - 1) / DM_ROWS;
  const int cs = cluster_size > 0 ? cluster_size : b200_decode_mega_cluster_size(H, F, groups);
  p.cluster_size = cs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * cs);
  cfg.blockDim = dim3(DM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, p);
}
