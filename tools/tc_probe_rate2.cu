// Probe 2: cost of back-to-back cta_group::1 kind::f16 tcgen05.mma (K = 16, no-swizzle K-major) under the conditions of the
// inverse kernel: operands at a DIFFERENT shared-memory address for every MMA (cycling through `span` bytes), optional
// A operand from tensor memory, optional alternation between two accumulators, groups of 3 MMAs separated by a commit-less gap.
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
// mode bit 0: A from TMEM; bit 1: alternate two accumulators; bit 2: same operand address every time (round-1 probe)
__global__ void __launch_bounds__(128) probe(long long* out, int M, int N, int iters, int mode, int span) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t mbar[2];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < span / 2; i += 128) reinterpret_cast<__half*>(sm)[i] = __float2half(0.001f * (i & 15));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar[0]))); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t abytes = M * 32, bbytes = N * 32;
    const uint32_t base = smem_u32(sm);
    uint32_t off = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (!(mode & 4)) { off += abytes + bbytes; if (off + abytes + bbytes > (uint32_t)span) off = 0; }
      const uint64_t da = make_desc(base + off, M * 16, 128), db = make_desc(base + off + abytes, N * 16, 128);
      const uint32_t d = tmem + ((mode & 2) && (i & 1) ? 256u : 0u);
      if (mode & 1) mma_ts(d, tmem + 480 + (i & 1) * 8, db, idesc, i > 1);
      else mma_ss(d, da, db, idesc, i > 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar[0])) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar[0])), "r"(0));
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
}
int main() {
  long long* d; long long h[148];
  CHECK(cudaMalloc(&d, sizeof(h)));
  const int iters = 2000, span = 160 * 1024;
  CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, span));
  for (int M : {64, 128}) {
    for (int N : {32, 48, 96, 144, 160}) {
      for (int mode : {4, 0, 2, 1, 3}) {
        if (M == 64 && (mode & 1)) continue;
        probe<<<148, 128, span>>>(d, M, N, iters, mode, span);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d mode=%d: CUDA error %s\n", M, N, mode, cudaGetErrorString(e)); return 0; }
        CHECK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("M=%3d N=%3d mode=%d (%s%s%s): %.1f cycles / MMA   (pipe floor %.1f, smem bytes/128 = %.1f)\n", M, N, mode,
               (mode & 4) ? "same operands" : "cycling operands", (mode & 1) ? ", A from TMEM" : "", (mode & 2) ? ", two accumulators" : "",
               (double)h[0] / iters, 128.0 * N / 256, ((mode & 1) ? 0 : M * 32.0 + 0) / 128 + N * 32.0 / 128);
      }
    }
  }
  return 0;
}
