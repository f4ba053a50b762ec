// Weight packing (mask / dropout folding) for the SIMT engine and the cross-draw reduction kernels
// (kernel group 4 of BASELINE.json's north_star).
//
// Reference behaviour replaced:
//   W * mask on every call                    src/naz/flows/bflow_jax_maf.py:74-77 (pyro MaskedLinear)
//   per-draw param.copy_                       src/naz/trainers/train_flows.py:65-71
//   Dropout after every hidden activation      src/naz/flows/transforms.py:38-43
//   host numpy mean_s exp(lp)                  examples/papers/2506.05657/plot.py:272-275
//   pyro Importance / ESS, compute_bic         src/naz/trainers/train_flows.py:358-380, bflow_jax_maf.py:474-475
#include "nazb_internal.h"

namespace {

// dst Wt[k][col(n)] = W[s][n][k] * mask[n][k] * (keep[s][k] / (1-p));  bias[col(n)] = b[s][n]
__global__ void pack_simt_kernel(const float* __restrict__ W, long long wst, const float* __restrict__ b,
                                 long long bst, const float* __restrict__ mask, const float* __restrict__ keep,
                                 long long keep_draw_stride, float inv_keep, int S, int out, int in, int ldw,
                                 int M, int D, const int* __restrict__ rank, float* __restrict__ dst,
                                 long long off_w, long long off_b, long long draw_stride,
                                 const float* __restrict__ baseW, const float* __restrict__ baseB, float dm_scale) {
  long long total = (long long)S * out * in;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % in);
    int n = (int)((i / in) % out);
    int s = (int)(i / ((long long)in * out));
    float v = W[(size_t)s * wst + (size_t)n * in + k];
    if (baseW) v = nazb_draw_map(baseW[(size_t)n * in + k], v, dm_scale);
    v *= mask[(size_t)n * in + k];
    if (keep) v *= keep[(size_t)s * keep_draw_stride + k] * inv_keep;
    int col = n;
    if (rank) { int m = n / D, d = n % D; col = rank[d] * M + m; }
    float* base = dst + (size_t)s * draw_stride;
    base[off_w + (size_t)k * ldw + col] = v;
    if (k == 0) {
      float bv = b[(size_t)s * bst + n];
      if (baseB) bv = nazb_draw_map(baseB[n], bv, dm_scale);
      base[off_b + col] = bv;
    }
  }
}

// lp [S][N] -> per-point running (max, sum exp) over S.  One thread per 4 points, float4 loads,
// coalesced along N; every byte of lp is read exactly once.
__global__ void lse_reduce_kernel(const float* __restrict__ lp, int S, int N, const float* __restrict__ log_w,
                                  float* __restrict__ omax, float* __restrict__ osum) {
  const int n4 = N >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4 + (N & 3); i += gridDim.x * blockDim.x) {
    if (i < n4) {
      float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, su[4] = {0.f, 0.f, 0.f, 0.f};
      const float4* p = reinterpret_cast<const float4*>(lp) + i;
      // N % 4 != 0 breaks the 16-byte alignment of later rows: handled by the scalar path below
      for (int s = 0; s < S; ++s) {
        float4 v4 = __ldg(p + (size_t)s * n4);
        float w = log_w ? log_w[s] : 0.f;
        float v[4] = {v4.x + w, v4.y + w, v4.z + w, v4.w + w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!(v[j] <= m[j])) { su[j] = su[j] * __expf(m[j] - v[j]) + 1.f; m[j] = v[j]; }
          else if (v[j] > -INFINITY) su[j] += __expf(v[j] - m[j]);
        }
      }
      reinterpret_cast<float4*>(omax)[i] = make_float4(m[0], m[1], m[2], m[3]);
      reinterpret_cast<float4*>(osum)[i] = make_float4(su[0], su[1], su[2], su[3]);
    }
  }
}

__global__ void lse_reduce_scalar_kernel(const float* __restrict__ lp, int S, int N, const float* __restrict__ log_w,
                                         float* __restrict__ omax, float* __restrict__ osum) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float m = -INFINITY, su = 0.f;
    for (int s = 0; s < S; ++s) {
      float v = lp[(size_t)s * N + n] + (log_w ? log_w[s] : 0.f);
      if (!(v <= m)) { su = su * expf(m - v) + 1.f; m = v; }
      else if (v > -INFINITY) su += expf(v - m);
    }
    omax[n] = m;
    osum[n] = su;
  }
}

__global__ void lse_finish_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, int G, int N,
                                  float log_norm, float* __restrict__ out) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float m = -INFINITY;
    bool nan = false;
    for (int g = 0; g < G; ++g) { float v = pmax[(size_t)g * N + n]; nan |= (v != v); m = fmaxf(m, v); }
    float s = 0.f;
    for (int g = 0; g < G; ++g) {
      float mg = pmax[(size_t)g * N + n];
      if (mg > -INFINITY) s += psum[(size_t)g * N + n] * expf(mg - m);
    }
    float r = (m > -INFINITY) ? m + logf(s) + log_norm : -INFINITY;
    out[n] = nan ? NAN : r;
  }
}

// One block; S is small (hundreds to thousands).
__global__ void importance_kernel(const double* __restrict__ sum_n, const float* __restrict__ log_prior,
                                  const float* __restrict__ log_q, int S, double* __restrict__ log_w_out,
                                  double* __restrict__ out3) {
  __shared__ double sh[3][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double mx = -INFINITY, mxs = -INFINITY;
  for (int s = tid; s < S; s += blockDim.x) {
    double lw = sum_n[s] + (log_prior ? (double)log_prior[s] : 0.0) - (log_q ? (double)log_q[s] : 0.0);
    if (log_w_out) log_w_out[s] = lw;
    mx = fmax(mx, lw);
    mxs = fmax(mxs, sum_n[s]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mxs = fmax(mxs, __shfl_xor_sync(0xffffffffu, mxs, o));
  }
  if (lane == 0) { sh[0][warp] = mx; sh[1][warp] = mxs; }
  __syncthreads();
  mx = -INFINITY; mxs = -INFINITY;
  for (int w = 0; w < (blockDim.x >> 5); ++w) { mx = fmax(mx, sh[0][w]); mxs = fmax(mxs, sh[1][w]); }
  __syncthreads();
  double s1 = 0.0, s2 = 0.0;
  for (int s = tid; s < S; s += blockDim.x) {
    double lw = sum_n[s] + (log_prior ? (double)log_prior[s] : 0.0) - (log_q ? (double)log_q[s] : 0.0);
    s1 += exp(lw - mx);
    s2 += exp(2.0 * (lw - mx));
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; }
  __syncthreads();
  if (tid == 0) {
    s1 = 0.0; s2 = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { s1 += sh[0][w]; s2 += sh[1][w]; }
    double lse1 = mx + log(s1), lse2 = 2.0 * mx + log(s2);
    out3[0] = lse1 - log((double)S);
    out3[1] = exp(2.0 * lse1 - lse2);
    out3[2] = mxs;
  }
}

}  // namespace

cudaError_t nazb_pack_simt(nazb_handle* h, const float* const* W, const float* const* b, const int64_t* wst,
                           const int64_t* bst, const float* const* mask, const float* keep, float p_drop,
                           cudaStream_t st, const DrawMap& dm) {
  const FlowGeom& g = h->geom;
  const int S = h->desc.S, n_lin = g.n_hidden + 1;
  cudaError_t e = cudaMemsetAsync(h->packed, 0, (size_t)S * g.draw_stride * sizeof(float), st);
  if (e != cudaSuccess) return e;
  int hk = 0;
  for (int j = 0; j < g.n_hidden; ++j) hk = hk > g.hidden[j] ? hk : g.hidden[j];
  const long long keep_draw_stride = (long long)g.L * g.n_hidden * hk;
  const int* rank_dev = h->perm_dev + (size_t)g.L * g.D;   // rank tables follow the perm tables
  for (int l = 0; l < g.L; ++l) {
    for (int j = 0; j < n_lin; ++j) {
      const int i = l * n_lin + j;
      const int out = g.ndim[j], in = g.kdim[j];
      const float* kp = nullptr;
      if (keep && j > 0) kp = keep + ((size_t)l * g.n_hidden + (j - 1)) * hk;
      long long total = (long long)S * out * in;
      int blocks = (int)((total + 255) / 256);
      if (blocks > 148 * 32) blocks = 148 * 32;
      pack_simt_kernel<<<blocks, 256, 0, st>>>(W[i], wst[i], b[i], bst[i], mask[i], kp, keep_draw_stride,
                                               1.f / (1.f - p_drop), S, out, in, g.ldw[j], g.M, g.D,
                                               (j == n_lin - 1) ? rank_dev + (size_t)l * g.D : nullptr,
                                               h->packed + (size_t)l * g.layer_stride, g.off_w[j], g.off_b[j],
                                               g.draw_stride, dm.baseW ? dm.baseW[i] : nullptr,
                                               dm.baseB ? dm.baseB[i] : nullptr, dm.scale);
      nazb_count_launch();
    }
  }
  return cudaGetLastError();
}

extern "C" int nazb_lse_reduce(const float* lp, int32_t S, int32_t N, const float* log_w, float* lse_max,
                               float* lse_sum, void* stream) {
  if (!lp || !lse_max || !lse_sum || S <= 0 || N <= 0) return NAZB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((N & 3) == 0 && ((uintptr_t)lp & 15) == 0 && ((uintptr_t)lse_max & 15) == 0 && ((uintptr_t)lse_sum & 15) == 0) {
    int n4 = N >> 2;
    int blocks = (n4 + 127) / 128;
    lse_reduce_kernel<<<blocks, 128, 0, st>>>(lp, S, N, log_w, lse_max, lse_sum);
  } else {
    lse_reduce_scalar_kernel<<<(N + 255) / 256, 256, 0, st>>>(lp, S, N, log_w, lse_max, lse_sum);
  }
  nazb_count_launch();
  return cudaGetLastError() == cudaSuccess ? NAZB_OK : NAZB_ERR_CUDA;
}

extern "C" int nazb_lse_finish(const float* lse_max, const float* lse_sum, int32_t G, int32_t N, float log_norm,
                               float* out, void* stream) {
  if (!lse_max || !lse_sum || !out || G <= 0 || N <= 0) return NAZB_ERR_BAD_ARG;
  lse_finish_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(lse_max, lse_sum, G, N, log_norm, out);
  nazb_count_launch();
  return cudaGetLastError() == cudaSuccess ? NAZB_OK : NAZB_ERR_CUDA;
}

extern "C" int nazb_importance(const double* sum_n, const float* log_prior, const float* log_q, int32_t S,
                               double* log_w_out, double* out3, void* stream) {
  if (!sum_n || !out3 || S <= 0) return NAZB_ERR_BAD_ARG;
  importance_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sum_n, log_prior, log_q, S, log_w_out, out3);
  nazb_count_launch();
  return cudaGetLastError() == cudaSuccess ? NAZB_OK : NAZB_ERR_CUDA;
}
