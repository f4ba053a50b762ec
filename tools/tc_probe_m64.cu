// Probe: where does a cta_group::1 M=64 tcgen05.mma put its 64 accumulator rows in TMEM, and can the
// D address carry a lane offset so two independent 64-row tiles share the same columns?
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
// A tiles: [kc][rows][8]; A128 value row r -> 1000 + r (k = 0 only); A64 value row r -> r + 1 (k = 0 only); B: ones at k = 0.
__global__ void __launch_bounds__(128) probe(float* out, int lane_off, int N) {
  __shared__ __align__(128) __half a128[2 * 128 * 8];
  __shared__ __align__(128) __half a64[2 * 64 * 8];
  __shared__ __align__(128) __half bt[2 * 64 * 8];
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 128 * 8; i += 128) a128[i] = __float2half(0.f);
  for (int i = tid; i < 2 * 64 * 8; i += 128) { a64[i] = __float2half(0.f); bt[i] = __float2half(0.f); }
  __syncthreads();
  a128[tid * 8] = __float2half((float)(1000 + tid));
  if (tid < 64) a64[tid * 8] = __float2half((float)(tid + 1));
  if (tid < N) bt[tid * 8] = __float2half(1.f);
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
    uint32_t id128 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t id64 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((64u >> 4) << 24);
    mma(tmem, make_desc(smem_u32(a128), 128 * 16, 128), make_desc(smem_u32(bt), N * 16, 128), id128, 0);   // sentinel everywhere
    mma(tmem + ((uint32_t)lane_off << 16), make_desc(smem_u32(a64), 64 * 16, 128), make_desc(smem_u32(bt), N * 16, 128), id64, 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0));
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  uint32_t r[2];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  out[tid * 2] = __uint_as_float(r[0]);
  out[tid * 2 + 1] = __uint_as_float(r[1]);
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(64));
}
int main(int argc, char** argv) {
  float* d; float h[256];
  CHECK(cudaMalloc(&d, sizeof(h)));
  int offs[] = {0, 16, 32, 64};
  for (int t = 0; t < 4; ++t) {
    int lane_off = offs[t];
    CHECK(cudaMemset(d, 0, sizeof(h)));
    probe<<<1, 128>>>(d, lane_off, 16);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lane_off=%d: CUDA error %s\n", lane_off, cudaGetErrorString(e)); return 0; }
    CHECK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("lane_off=%d: lane -> value(col0)\n", lane_off);
    for (int l = 0; l < 128; ++l) printf("%s%4.0f", (l % 32 == 0) ? "\n  " : " ", h[l * 2]);
    printf("\n");
  }
  return 0;
}
