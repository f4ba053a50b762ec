// Shared-memory tiled fp32 GEMM panel of the SIMT engine (used by flow_simt.cu and flow_grad.cu).
// Activations live in shared memory as [unit][point] (point-contiguous); the weight operand streams from global memory
// (k-major rows, row length a multiple of 4 floats) through a cp.async double-buffered panel.
#pragma once
#include <cuda_runtime.h>

namespace {

constexpr int kThreads = 256;
constexpr int kKC = 32;   // k-rows per weight panel chunk

// TN_ = 5 with P = 32 makes one pass 160 columns wide: a 150-unit hidden layer then takes one pass instead of two
template <int P, int TN_ = 4>
struct Tile {
  static constexpr int TM = 4;
  static constexpr int TR = P / TM;            // thread rows
  static constexpr int TC = kThreads / TR;     // thread cols
  static constexpr int TN = TN_;
  static constexpr int NPASS = TC * TN;        // columns per pass
  static constexpr int WCHUNK = kKC * NPASS;   // floats per staged panel chunk
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// dstT[n][p] = epi( sum_k actT[k][p] * Wt[k][n] + bias[n] )  for n in [c0, c1)
template <int P, bool TANH, int TN_ = 4>
__device__ __forceinline__ void gemm_panel(const float* __restrict__ actT, int K,
                                           const float* __restrict__ Wt, int ldw,
                                           const float* __restrict__ bias, int c0, int c1,
                                           float* __restrict__ dstT, float* __restrict__ wbuf) {
  using T = Tile<P, TN_>;
  const int tid = threadIdx.x;
  const int tr = tid % T::TR, tc = tid / T::TR;
  const int nchunks = (K + kKC - 1) / kKC;
  for (int pc = (c0 & ~3); pc < c1; pc += T::NPASS) {
    float acc[T::TM][T::TN];
#pragma unroll
    for (int i = 0; i < T::TM; ++i)
#pragma unroll
      for (int j = 0; j < T::TN; ++j) acc[i][j] = 0.f;
    const int ncols = min(T::NPASS, ldw - pc);   // multiple of 4
    auto stage = [&](int ch, int buf) {
      float* dst = wbuf + buf * T::WCHUNK;
      const int k0 = ch * kKC;
      for (int i = tid; i < kKC * (T::NPASS / 4); i += kThreads) {
        int kk = i / (T::NPASS / 4), c4 = (i % (T::NPASS / 4)) * 4;
        bool valid = (k0 + kk < K) && (c4 < ncols);
        const float* src = valid ? (Wt + (size_t)(k0 + kk) * ldw + pc + c4) : Wt;
        cp_async16(dst + kk * T::NPASS + c4, src, valid);
      }
      cp_async_commit();
    };
    if (nchunks > 0) stage(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
      if (ch + 1 < nchunks) {
        stage(ch + 1, (ch + 1) & 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const float* wb = wbuf + (ch & 1) * T::WCHUNK + tc * T::TN;
      const float* ab = actT + (size_t)ch * kKC * P + tr * T::TM;
      const int kmax = min(kKC, K - ch * kKC);
      if (tc * T::TN < ncols) {
#pragma unroll 4
        for (int kk = 0; kk < kmax; ++kk) {
          float4 a = *reinterpret_cast<const float4*>(ab + kk * P);
          float av[4] = {a.x, a.y, a.z, a.w}, wv[T::TN];
          if constexpr (T::TN == 4) {
            float4 w = *reinterpret_cast<const float4*>(wb + kk * T::NPASS);
            wv[0] = w.x; wv[1] = w.y; wv[2] = w.z; wv[3] = w.w;
          } else {
#pragma unroll
            for (int j = 0; j < T::TN; ++j) wv[j] = wb[kk * T::NPASS + j];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < T::TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < T::TN; ++j) {
      int n = pc + tc * T::TN + j;
      if (n >= c0 && n < c1) {
        float bj = bias[n];
        float4 v;
        float* vp = &v.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float t = acc[i][j] + bj;
          vp[i] = TANH ? tanhf(t) : t;
        }
        *reinterpret_cast<float4*>(dstT + (size_t)n * P + tr * T::TM) = v;
      }
    }
  }
  __syncthreads();
}

// Same product with an NSTAGE-deep cp.async ring and one barrier per k-chunk (flow_grad.cu: one 8-warp CTA per SM has
// nothing else to hide the L2 latency of the weight panel behind).  wbuf holds NSTAGE chunks.
template <int P, bool TANH, int TN_, int NSTAGE>
__device__ __forceinline__ void gemm_panel_ms(const float* __restrict__ actT, int K,
                                              const float* __restrict__ Wt, int ldw,
                                              const float* __restrict__ bias, int c0, int c1,
                                              float* __restrict__ dstT, float* __restrict__ wbuf) {
  using T = Tile<P, TN_>;
  const int tid = threadIdx.x;
  const int tr = tid % T::TR, tc = tid / T::TR;
  const int nchunks = (K + kKC - 1) / kKC;
  for (int pc = (c0 & ~3); pc < c1; pc += T::NPASS) {
    float acc[T::TM][T::TN];
#pragma unroll
    for (int i = 0; i < T::TM; ++i)
#pragma unroll
      for (int j = 0; j < T::TN; ++j) acc[i][j] = 0.f;
    const int ncols = min(T::NPASS, ldw - pc);   // multiple of 4
    auto stage = [&](int ch) {
      if (ch < nchunks) {
        float* dst = wbuf + (ch % NSTAGE) * T::WCHUNK;
        const int k0 = ch * kKC;
        for (int i = tid; i < kKC * (T::NPASS / 4); i += kThreads) {
          int kk = i / (T::NPASS / 4), c4 = (i % (T::NPASS / 4)) * 4;
          bool valid = (k0 + kk < K) && (c4 < ncols);
          const float* src = valid ? (Wt + (size_t)(k0 + kk) * ldw + pc + c4) : Wt;
          cp_async16(dst + kk * T::NPASS + c4, src, valid);
        }
      }
      cp_async_commit();   // one group per call (possibly empty) keeps the wait count uniform
    };
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) stage(s);
    for (int ch = 0; ch < nchunks; ++ch) {
      cp_async_wait<NSTAGE - 2>();   // chunk ch has landed
      __syncthreads();               // ... for every thread, and buffer (ch - 1) % NSTAGE is free again
      stage(ch + NSTAGE - 1);
      const float* wb = wbuf + (ch % NSTAGE) * T::WCHUNK + tc * T::TN;
      const float* ab = actT + (size_t)ch * kKC * P + tr * T::TM;
      const int kmax = min(kKC, K - ch * kKC);
      if (tc * T::TN < ncols) {
#pragma unroll 4
        for (int kk = 0; kk < kmax; ++kk) {
          float4 a = *reinterpret_cast<const float4*>(ab + kk * P);
          float av[4] = {a.x, a.y, a.z, a.w}, wv[T::TN];
          if constexpr (T::TN == 4) {
            float4 w = *reinterpret_cast<const float4*>(wb + kk * T::NPASS);
            wv[0] = w.x; wv[1] = w.y; wv[2] = w.z; wv[3] = w.w;
          } else {
#pragma unroll
            for (int j = 0; j < T::TN; ++j) wv[j] = wb[kk * T::NPASS + j];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < T::TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < T::TN; ++j) {
      int n = pc + tc * T::TN + j;
      if (n >= c0 && n < c1) {
        float bj = bias[n];
        float4 v;
        float* vp = &v.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float t = acc[i][j] + bj;
          vp[i] = TANH ? tanhf(t) : t;
        }
        *reinterpret_cast<float4*>(dstT + (size_t)n * P + tr * T::TM) = v;
      }
    }
    __syncthreads();   // the ring is rewritten by the next pass / panel; dstT is complete
  }
  cp_async_wait<0>();
}

}  // namespace
