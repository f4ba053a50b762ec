// Probe 3: issue-blocking cost of cta_group::1 kind::f16 tcgen05.mma (K = 16, no-swizzle K-major, M = 128 / 64), issued the way
// the inverse kernel does it (whole warp convergent, elect.sync-guarded groups of G MMAs, descriptors precomputed):
//   nacc accumulators used round-robin (1 = every MMA depends on the previous one), operands at nops different addresses,
//   optional A operand from tensor memory.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred;
}
template <int NACC, bool TS>
__global__ void __launch_bounds__(128) probe(long long* out, int M, int N, int iters, int nops) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t mbar[2];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 2; i += 128) reinterpret_cast<__half*>(sm)[i] = __float2half(0.001f * (i & 15));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar[0]))); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t abytes = M * 32, bbytes = N * 32, base = smem_u32(sm);
    uint64_t da[4], db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t off = (uint32_t)(j % nops) * (abytes + bbytes);
      da[j] = make_desc(base + off, M * 16, 128);
      db[j] = make_desc(base + off + abytes, N * 16, 128);
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t d = (uint32_t)((j % NACC) * 128);
          if (TS) mma_ts(d, 384u + (j & 1) * 8u, db[j], idesc, 1u);
          else mma_ss(d, da[j], db[j], idesc, 1u);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar[0])) : "memory");
    __syncwarp();
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar[0])), "r"(0));
    const long long t2 = clock64();
    if ((tid & 31) == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(0), "r"(512));
}
template <int NACC, bool TS>
void run(long long* d, int M, int N, int nops) {
  long long h[296];
  const int iters = 4000;
  CHECK(cudaFuncSetAttribute(probe<NACC, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  probe<NACC, TS><<<148, 128, 96 * 1024>>>(d, M, N, iters, nops);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M=%d N=%d: CUDA error %s\n", M, N, cudaGetErrorString(e)); exit(0); }
  CHECK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  printf("M=%3d N=%3d nacc=%d nops=%d %s: issue %.1f, to completion %.1f cycles / MMA   (pipe floor %.1f)\n", M, N, NACC, nops, TS ? "A-TMEM" : "A-smem",
         (double)h[0] / iters, (double)h[1] / iters, 128.0 * N / 256);
}
int main() {
  long long* d;
  CHECK(cudaMalloc(&d, sizeof(long long) * 296));
  for (int M : {128, 64}) {
    for (int N : {48, 96, 144, 256}) {
      if (N > 128 && false) continue;
      run<1, false>(d, M, N, 1);
      run<1, false>(d, M, N, 4);
      run<2, false>(d, M, N, 4);
      run<4, false>(d, M, N, 4);
      if (M == 128) { run<1, true>(d, M, N, 4); run<2, true>(d, M, N, 4); }
    }
  }
  return 0;
}
