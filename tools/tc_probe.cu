// tcgen05 probe: D[128 x N] = A[128 x K] * B[N x K]^T with the fp16 hi/lo 3-MMA split, no-swizzle
// K-major operands in shared memory, fp32 accumulation in TMEM.  Verifies descriptor encodings and
// measures the accumulation error against an fp64 host reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Dout,
                                              int N, int K, int nsplit) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = K / 8;
  // layout: [kc][row][8 halves]
  __half* a_hi = reinterpret_cast<__half*>(smem);
  __half* a_lo = a_hi + (size_t)KC * M * 8;
  __half* b_hi = a_lo + (size_t)KC * M * 8;
  __half* b_lo = b_hi + (size_t)KC * N * 8;
  // fill A: thread = row
  for (int kc = 0; kc < KC; ++kc) {
    __half hi[8], lo[8];
    for (int e = 0; e < 8; ++e) {
      float v = A[(size_t)tid * K + kc * 8 + e];
      hi[e] = __float2half_rn(v);
      lo[e] = __float2half_rn(v - __half2float(hi[e]));
    }
    *reinterpret_cast<uint4*>(a_hi + ((size_t)kc * M + tid) * 8) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(a_lo + ((size_t)kc * M + tid) * 8) = *reinterpret_cast<uint4*>(lo);
  }
  for (int i = tid; i < N * KC; i += 128) {
    int n = i % N, kc = i / N;
    __half hi[8], lo[8];
    for (int e = 0; e < 8; ++e) {
      float v = B[(size_t)n * K + kc * 8 + e];
      hi[e] = __float2half_rn(v);
      lo[e] = __float2half_rn(v - __half2float(hi[e]));
    }
    *reinterpret_cast<uint4*>(b_hi + ((size_t)kc * N + n) * 8) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(b_lo + ((size_t)kc * N + n) * 8) = *reinterpret_cast<uint4*>(lo);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D=F32, A=B=F16, K-major both, N>>3 @17, M>>4 @24
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    const uint32_t lbo_a = M * 16, lbo_b = N * 16, sbo = 128;
    uint32_t acc = 0;
    for (int s = 0; s < nsplit; ++s) {
      const __half* ap = (s == 2) ? a_lo : a_hi;
      const __half* bp = (s == 1) ? b_lo : b_hi;
      for (int k = 0; k < K / 16; ++k) {
        uint64_t da = make_desc(smem_u32(ap) + k * 2 * lbo_a, lbo_a, sbo);
        uint64_t db = make_desc(smem_u32(bp) + k * 2 * lbo_b, lbo_b, sbo);
        mma_f16_ss(tmem, da, db, idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 16; ++j) Dout[(size_t)(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256));
}

int main() {
  const int N = 160, K = 160;
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& v : A) v = tanhf(((rand() / (float)RAND_MAX) * 2 - 1) * 2.f);
  for (auto& v : B) v = ((rand() / (float)RAND_MAX) * 2 - 1) * 0.16f;
  float *dA, *dB, *dD;
  CHECK(cudaMalloc(&dA, A.size() * 4)); CHECK(cudaMalloc(&dB, B.size() * 4)); CHECK(cudaMalloc(&dD, D.size() * 4));
  CHECK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  size_t smem = (size_t)(K / 8) * (M + N) * 8 * 2 * 2;
  CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int nsplit = 1; nsplit <= 3; nsplit += 2) {
    CHECK(cudaMemset(dD, 0, D.size() * 4));
    probe<<<1, 128, smem>>>(dA, dB, dD, N, K, nsplit);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_err32 = 0, max_ref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0; float r32 = 0;
        for (int k = 0; k < K; ++k) { ref += (double)A[m * K + k] * (double)B[n * K + k]; r32 = fmaf(A[m * K + k], B[n * K + k], r32); }
        max_err = fmax(max_err, fabs(D[m * N + n] - ref));
        max_err32 = fmax(max_err32, fabs((double)r32 - ref));
        max_ref = fmax(max_ref, fabs(ref));
      }
    printf("nsplit=%d  max|D-ref64|=%.3e   (fp32 fma chain error %.3e, max|ref|=%.3f)\n", nsplit, max_err, max_err32, max_ref);
  }
  return 0;
}
