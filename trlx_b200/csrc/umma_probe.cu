// Microbenchmark: cost of a stream of small tcgen05.mma instructions issued by one thread (operands resident in shared memory),
// as a function of the atom shape (M, N) and of how many independent TMEM accumulators the stream rotates over.
// Used to size the tiles of the decode megakernel (profiles/umma_probe.jsonl); not on any product path.
#include "ptx.cuh"

namespace b200 {

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(int M, int N, int n_mma, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_slot;
  // A: up to 128 rows x 128 B, B: up to 256 rows x 128 B (128-byte swizzle layout; contents irrelevant)
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); mbar_arrive(&bar2); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  if (nacc >= 100) {
    // warp-uniform issue loop: all 32 lanes of warp 0 run the control flow and the descriptor arithmetic (so the compiler keeps
    // them in uniform registers), only the elected lane executes the tcgen05 instructions
    if (threadIdx.x < 32) {
      const int mode = nacc - 100;
      const uint32_t idesc = umma_idesc(1, 1, (uint32_t)M, (uint32_t)N);
      const uint64_t da = umma_desc_k_sw128(smem_u32(smem)), db = umma_desc_k_sw128(smem_u32(smem + 128 * 128));
      const int acc_stride = mode == 1 ? 0 : N;
      const long long t0 = clock64();
      for (int i = 0; i < n_mma / 4; ++i) {
        if (mode == 24) mbar_wait(&bar2, 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + k * acc_stride, da + 2 * k, db + 2 * k, idesc, (i > 0 || (mode == 1 && k > 0)) ? 1u : 0u);
          if (mode >= 14) umma_commit(&bar3);
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(1, 1, (uint32_t)M, (uint32_t)N);
    const uint64_t da = umma_desc_k_sw128(smem_u32(smem)), db = umma_desc_k_sw128(smem_u32(smem + 128 * 128));
    const long long t0 = clock64();
    // mode (passed in nacc): 1 = one accumulator, 4 = four accumulators (k-step j -> accumulator j),
    // 14 = four accumulators + a tcgen05.commit after every 4 MMAs, 24 = additionally a try_wait on a completed mbarrier
    const int acc_stride = nacc == 1 ? 0 : N;
    for (int i = 0; i < n_mma / 4; ++i) {
      if (nacc == 24) mbar_wait(&bar2, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base + k * acc_stride, da + 2 * k, db + 2 * k, idesc, (i > 0 || (nacc == 1 && k > 0)) ? 1u : 0u);
      if (nacc >= 14) umma_commit(&bar3);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;   // issue time
    out[1] = t2 - t0;   // until every MMA has completed
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace b200

extern "C" int b200_umma_probe(int M, int N, int n_mma, int nacc, long long* out, cudaStream_t stream) {
  if (!(M == 64 || M == 128) || N % 8 || N < 8 || N > 256 || nacc < 1 || ((nacc % 100) == 1 ? 1 : 4) * N > 512) return -2;
  const size_t smem = (128 + 256) * 128 + 1024;
  cudaFuncSetAttribute(b200::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  b200::umma_probe_kernel<<<1, 128, smem, stream>>>(M, N, n_mma, nacc, out);
  return (int)cudaGetLastError();
}
