// Probe: tcgen05.ld.16x32bx2 semantics on the lanes an M=64 MMA fills (which thread gets which lane/column),
// for lane sub-offset 0 and 16.
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
// D[r][n] = 100 (r + 1 + 64 * chain) + n
__global__ void __launch_bounds__(128) probe(float* out) {
  __shared__ __align__(128) __half a64[2][2 * 64 * 8];
  __shared__ __align__(128) __half bt[2 * 32 * 8];
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 64 * 8; i += 128) { a64[0][i] = __float2half(0.f); a64[1][i] = __float2half(0.f); }
  for (int i = tid; i < 2 * 32 * 8; i += 128) bt[i] = __float2half(0.f);
  __syncthreads();
  if (tid < 64) {
    a64[0][tid * 8] = __float2half((float)(tid + 1)); a64[0][tid * 8 + 1] = __float2half(1.f);
    a64[1][tid * 8] = __float2half((float)(tid + 1 + 64)); a64[1][tid * 8 + 1] = __float2half(1.f);
  }
  if (tid < 32) { bt[tid * 8] = __float2half(100.f); bt[tid * 8 + 1] = __float2half((float)tid); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_s)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    uint32_t id64 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((64u >> 4) << 24);
    mma(tmem, make_desc(smem_u32(a64[0]), 64 * 16, 128), make_desc(smem_u32(bt), 32 * 16, 128), id64, 0);
    mma(tmem + (16u << 16), make_desc(smem_u32(a64[1]), 64 * 16, 128), make_desc(smem_u32(bt), 32 * 16, 128), id64, 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0));
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  for (int sub = 0; sub < 2; ++sub) {
    uint32_t r[2];
    uint32_t ta = tmem + ((uint32_t)(warp * 32 + sub * 16) << 16) + 2;   // column base 2
    asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x2.b32 {%0,%1}, [%2], 8;\n" : "=r"(r[0]), "=r"(r[1]) : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    out[(sub * 128 + tid) * 2] = __uint_as_float(r[0]);
    out[(sub * 128 + tid) * 2 + 1] = __uint_as_float(r[1]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(64));
}
int main() {
  float* d; float h[512];
  CHECK(cudaMalloc(&d, sizeof(h)));
  probe<<<1, 128>>>(d);
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  for (int sub = 0; sub < 2; ++sub) {
    printf("lane sub-offset %d (expect rows %d..): thread -> (r0, r1)   [value = 100*row1based + col]\n", sub * 16, sub * 64 + 1);
    for (int t = 0; t < 128; ++t) printf("%s(%5.0f,%5.0f)", (t % 8 == 0) ? "\n  " : " ", h[(sub * 128 + t) * 2], h[(sub * 128 + t) * 2 + 1]);
    printf("\n");
  }
  return 0;
}
