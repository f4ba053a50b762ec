// Probe: sustained cost of a cta_group::1 kind::f16 tcgen05.mma (K = 16, SS operands, no-swizzle K-major) as a
// function of M (64 / 128) and N, one CTA per SM, one issuing thread, `iters` back-to-back MMAs into the same D.
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
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128) probe(long long* out, int M, int N, int iters, int two_lane_halves, int nissuers) {
  __shared__ __align__(128) __half a[2 * 128 * 8];
  __shared__ __align__(128) __half bt[2 * 256 * 8];
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t mbar[2];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 8; i += 128) a[i] = __float2half(0.001f * (i & 15));
  for (int i = tid; i < 2 * 256 * 8; i += 128) bt[i] = __float2half(0.002f * (i & 7));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar[0]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar[1]))); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_s;
  long long t0 = 0, t1 = 0;
  if ((tid & 31) == 0 && warp < nissuers) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(a), M * 16, 128), db = make_desc(smem_u32(bt), N * 16, 128);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + warp * 128 + ((two_lane_halves && (i & 1)) ? (16u << 16) : 0u);
      mma(d, da, db, idesc, i > 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar[warp])) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar[warp])), "r"(0));
    t1 = clock64();
    if (warp == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256));
}
int main() {
  long long* d; long long h[148];
  CHECK(cudaMalloc(&d, sizeof(h)));
  const int iters = 2000;
  int Ns[] = {32, 56, 64, 80, 96, 128, 160, 256};
  for (int M : {64, 128}) {
    for (int N : Ns) {
      if (M == 128 && (N % 16)) continue;
      if (N > 128) continue;
      for (int halves = 0; halves < 2; ++halves) {
        const int nissuers = halves + 1;
        probe<<<148, 128>>>(d, M, N, iters, 0, nissuers);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d: CUDA error %s\n", M, N, cudaGetErrorString(e)); return 0; }
        CHECK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("M=%3d N=%3d K=16 %s: %.1f cycles / MMA  (floor max(M,128)*N/256 = %.1f)\n", M, N, halves ? "two issuing warps (per-warp cycles)" : "one issuing warp", (double)h[0] / iters, 128.0 * N / 256);
      }
    }
  }
  return 0;
}
