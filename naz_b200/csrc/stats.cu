// Consumers of the [S][N][D] sample tensor: per-draw N-D histograms against fixed bin edges and
// the highest-posterior-density interval across draws of every bin.  Reference:
//   calibrate():      A = np.histogram2d / jnp.histogramdd(this_ppd, bins=eq_bins, density=True) per draw
//                     src/naz/flows/bflow_jax_maf.py:436-441
//   hpd_vectorized(): src/naz/statutils.py:22-46
// Both kernels are HBM-bound index work: the histogram reads 4*S*N*D bytes once, the HPD reads 4*S*M bytes once.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <vector>
#include "nazb_internal.h"

namespace {

constexpr int kHistThreads = 256;
constexpr int kMaxHistDim = 8;
constexpr int kSmemBins = 8192;   // 32 KB of shared counters; larger histograms use global atomics directly

struct HistGeom {
  int D;
  int nb[kMaxHistDim];        // bins per dim
  int eoff[kMaxHistDim];      // offset of the dim's edges inside `edges`
  long long total_bins;
};

// numpy.histogramdd binning of one coordinate: searchsorted(edges, v, side="right") - 1, the right-most edge belongs to
// the last bin, anything outside (or NaN) is dropped.  numpy compares the fp32 samples with the fp64 edges in double;
// for an fp32 v that is EXACTLY  v >= e  <=>  v >= round_up_to_float(e)  and  v <= e  <=>  v <= round_down_to_float(e),
// so the kernel compares floats against pre-rounded edges: ce[j] = float_ru(e[j]) for j < nb, ce[nb] = float_rd(e[nb]).
__device__ __forceinline__ int bin_of(float v, const float* ce, int nb) {
  if (!(v >= ce[0]) || !(v <= ce[nb])) return -1;
  int lo = 1, hi = nb;               // number of interior edges ce[1..nb-1] that are <= v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (v >= ce[mid]) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

template <int DV>   // DV = 4 / 2: vector loads of a whole point; 0: generic
__global__ void histdd_kernel(const float* __restrict__ x, long long N, HistGeom g, const double* __restrict__ edges,
                              int n_edges, unsigned int* __restrict__ counts, int chunks_per_draw, bool use_smem) {
  extern __shared__ unsigned char hsm[];
  float* se = reinterpret_cast<float*>(hsm);                                      // [n_edges] pre-rounded edges
  unsigned int* sc = reinterpret_cast<unsigned int*>(hsm + (size_t)((n_edges * 4 + 15) & ~15));   // [total_bins] if use_smem
  const int s = blockIdx.x / chunks_per_draw, chunk = blockIdx.x % chunks_per_draw;
  for (int i = threadIdx.x; i < n_edges; i += blockDim.x) {
    bool last = false;
    for (int d = 0; d < g.D; ++d) last = last || (i == g.eoff[d] + g.nb[d]);
    se[i] = last ? __double2float_rd(edges[i]) : __double2float_ru(edges[i]);
  }
  if (use_smem)
    for (long long i = threadIdx.x; i < g.total_bins; i += blockDim.x) sc[i] = 0u;
  __syncthreads();
  const long long per = (N + chunks_per_draw - 1) / chunks_per_draw;
  const long long n0 = (long long)chunk * per, n1 = min(N, n0 + per);
  const float* xs = x + (size_t)s * N * g.D;
  unsigned int* gc = counts + (size_t)s * g.total_bins;
  for (long long n = n0 + threadIdx.x; n < n1; n += blockDim.x) {
    float v[kMaxHistDim];
    if (DV == 4) {
      const float4 q = *reinterpret_cast<const float4*>(xs + n * 4);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else if (DV == 2) {
      const float2 q = *reinterpret_cast<const float2*>(xs + n * 2);
      v[0] = q.x; v[1] = q.y;
    } else {
      for (int d = 0; d < g.D; ++d) v[d] = xs[n * g.D + d];
    }
    long long flat = 0;
    bool ok = true;
    const int Dn = DV ? DV : g.D;
#pragma unroll
    for (int d = 0; d < (DV ? DV : kMaxHistDim); ++d) {
      if (d < Dn) {
        const int b = bin_of(v[d], se + g.eoff[d], g.nb[d]);
        ok = ok && (b >= 0);
        flat = flat * g.nb[d] + (b >= 0 ? b : 0);      // C order, as numpy
      }
    }
    if (ok) {
      if (use_smem) atomicAdd(sc + flat, 1u);
      else atomicAdd(gc + flat, 1u);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (long long i = threadIdx.x; i < g.total_bins; i += blockDim.x) {
      const unsigned int c = sc[i];
      if (c) atomicAdd(gc + i, c);
    }
  }
}

// density[s][b] = counts[s][b] / (sum_b counts[s][b] * volume[b])   (numpy density=True)
__global__ void hist_density_kernel(const unsigned int* __restrict__ counts, HistGeom g, const double* __restrict__ edges,
                                    float* __restrict__ density) {
  __shared__ unsigned long long tot_s;
  const int s = blockIdx.x;
  const unsigned int* c = counts + (size_t)s * g.total_bins;
  unsigned long long t = 0;
  for (long long i = threadIdx.x; i < g.total_bins; i += blockDim.x) t += c[i];
  if (threadIdx.x == 0) tot_s = 0ull;
  __syncthreads();
  atomicAdd(&tot_s, t);
  __syncthreads();
  const double tot = (double)tot_s;
  for (long long i = threadIdx.x; i < g.total_bins; i += blockDim.x) {
    long long r = i;
    double vol = 1.0;
    for (int d = g.D - 1; d >= 0; --d) {
      const int b = (int)(r % g.nb[d]);
      r /= g.nb[d];
      vol *= edges[g.eoff[d] + b + 1] - edges[g.eoff[d] + b];
    }
    density[(size_t)s * g.total_bins + i] = (float)((double)c[i] / tot / vol);
  }
}

// HPD across draws: one CTA = kCols consecutive columns (bins) of v[S][M]; the S values of each column are sorted in
// shared memory (bitonic, padded with +inf), then the narrowest window holding floor((1 - alpha) S) + 1 order statistics
// is selected (first minimum, as numpy.argmin).
constexpr int kHpdCols = 4;
__global__ void hpd_kernel(const float* __restrict__ v, int S, long long M, int n_pad, int inc, int cols, float* __restrict__ lo,
                           float* __restrict__ hi) {
  extern __shared__ float hs[];   // [cols][n_pad], cols <= kHpdCols columns per CTA (fewer when S is large: shared-memory budget)
  __shared__ float best_w[kHpdCols][32];
  __shared__ int best_i[kHpdCols][32];
  const long long m0 = (long long)blockIdx.x * cols;
  for (int i = threadIdx.x; i < n_pad * cols; i += blockDim.x) {
    const int c = i % cols, s = i / cols;
    float val = INFINITY;
    if (s < S && m0 + c < M) val = v[(size_t)s * M + m0 + c];
    hs[c * n_pad + s] = val;
  }
  __syncthreads();
  for (int k = 2; k <= n_pad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pad * cols; i += blockDim.x) {
        const int c = i / n_pad, a = i % n_pad, b = a ^ j;
        if (b > a) {
          float* col = hs + c * n_pad;
          const float xa = col[a], xb = col[b];
          const bool up = ((a & k) == 0);
          // NaNs sort to the end like numpy: treat NaN as larger than everything
          const bool gt = (xa > xb) || (xa != xa && xb == xb);
          if (gt == up) { col[a] = xb; col[b] = xa; }
        }
      }
      __syncthreads();
    }
  const int n_int = S - inc;   // number of candidate windows (> 0, checked by the host)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = 0; c < cols; ++c) {
    const float* col = hs + c * n_pad;
    float bw = INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n_int; i += blockDim.x) {
      const float w = col[i + inc] - col[i];
      if (w < bw || (w == bw && i < bi)) { bw = w; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ow = __shfl_xor_sync(0xffffffffu, bw, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ow < bw || (ow == bw && oi < bi)) { bw = ow; bi = oi; }
    }
    if (lane == 0) { best_w[c][warp] = bw; best_i[c][warp] = bi; }
  }
  __syncthreads();
  if (threadIdx.x < cols && m0 + threadIdx.x < M) {
    const int c = threadIdx.x;
    float bw = best_w[c][0];
    int bi = best_i[c][0];
    for (int w = 1; w < nwarps; ++w)
      if (best_w[c][w] < bw || (best_w[c][w] == bw && best_i[c][w] < bi)) { bw = best_w[c][w]; bi = best_i[c][w]; }
    if (bi == 0x7fffffff) bi = 0;
    lo[m0 + c] = hs[c * n_pad + bi];
    hi[m0 + c] = hs[c * n_pad + bi + inc];
  }
}

// Truncated-normal guide (src/naz/priors/TruncatedNormal.py:14-60): y = loc + scale * sqrt(2) erfinv(2 u - 1) with
// u = x (cdf_high - cdf_low) + cdf_low, x ~ U(0, 1) supplied by the caller; log q(y) = log pdf(y) - log(cdf_high - cdf_low),
// summed over the P parameters of a draw.  HBM-bound elementwise work: 4 B read + 4 B written per element.
__device__ __forceinline__ float std_normal_cdf(float v) { return 0.5f * (1.f + erff(v / sqrtf(2.0f))); }   // :14-16

__global__ void truncnorm_kernel(const float* __restrict__ x, int S, long long P, const float* __restrict__ loc,
                                 const float* __restrict__ scale, const float* __restrict__ low, const float* __restrict__ high,
                                 int loc_n, int scale_n, int low_n, int high_n, float* __restrict__ y, double* __restrict__ log_q,
                                 int chunks_per_draw) {
  const int s = blockIdx.x / chunks_per_draw, chunk = blockIdx.x % chunks_per_draw;
  const long long per = (P + chunks_per_draw - 1) / chunks_per_draw;
  const long long p0 = (long long)chunk * per, p1 = min(P, p0 + per);
  double acc = 0.0;
  for (long long i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
    const float m = loc[loc_n == 1 ? 0 : i], sc = scale[scale_n == 1 ? 0 : i];
    const float lo = low[low_n == 1 ? 0 : i], hi = high[high_n == 1 ? 0 : i];
    const float cl = std_normal_cdf((lo - m) / sc), ch = std_normal_cdf((hi - m) / sc);        // :43-44
    const float u = x[(size_t)s * P + i] * (ch - cl) + cl;                                       // :48
    const float v = m + sc * (sqrtf(2.0f) * erfinvf(2.f * u - 1.f));                             // :49, :18-28
    y[(size_t)s * P + i] = v;
    const float r = (v - m) / sc;
    const float pdf = (1.f / (sc * sqrtf(2.0f * 3.14159265358979323846f))) * expf(-0.5f * r * r);   // :30-32
    acc += (double)(logf(pdf) - logf(ch - cl));                                                   // -(:58)
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(log_q + s, t);
  }
}

}  // namespace

extern "C" int nazb_truncnorm_sample(const float* x, int32_t S, int64_t P, const float* loc, int32_t loc_n, const float* scale,
                                     int32_t scale_n, const float* low, int32_t low_n, const float* high, int32_t high_n,
                                     float* y, double* log_q, void* stream) {
  if (!x || !loc || !scale || !low || !high || !y || !log_q || S < 1 || P < 1) return NAZB_ERR_BAD_ARG;
  for (int n : {loc_n, scale_n, low_n, high_n})
    if (n != 1 && n != P) return NAZB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(log_q, 0, sizeof(double) * (size_t)S, st) != cudaSuccess) return NAZB_ERR_CUDA;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long chunks = std::max<long long>(1, std::min<long long>((8LL * sms + S - 1) / S, (P + 4095) / 4096));
  truncnorm_kernel<<<(unsigned)(S * chunks), 256, 0, st>>>(x, S, (long long)P, loc, scale, low, high, loc_n, scale_n, low_n, high_n,
                                                           y, log_q, (int)chunks);
  nazb_count_launch();
  return cudaGetLastError() == cudaSuccess ? NAZB_OK : NAZB_ERR_CUDA;
}

extern "C" int nazb_histogramdd(const float* x, int32_t S, int64_t N, int32_t D, const double* edges, const int32_t* nbins,
                                uint32_t* counts, float* density, void* stream) {
  if (!x || !edges || !nbins || !counts || S < 1 || N < 1 || D < 1 || D > kMaxHistDim) return NAZB_ERR_BAD_ARG;
  HistGeom g{};
  g.D = D;
  long long total = 1;
  int n_edges = 0;
  for (int d = 0; d < D; ++d) {
    if (nbins[d] < 1 || nbins[d] > 4096) return NAZB_ERR_BAD_ARG;
    g.nb[d] = nbins[d];
    g.eoff[d] = n_edges;
    n_edges += nbins[d] + 1;
    total *= nbins[d];
    if (total > (1LL << 26)) return NAZB_ERR_UNSUPPORTED;
  }
  g.total_bins = total;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)S * total, st) != cudaSuccess) return NAZB_ERR_CUDA;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool use_smem = total <= kSmemBins;
  // ~8 CTAs per SM in total, at least one chunk per draw, chunks of >= 4096 points
  long long chunks = std::max<long long>(1, std::min<long long>((16LL * sms + S - 1) / S, (N + 4095) / 4096));
  const size_t smem = (size_t)((n_edges * 4 + 15) & ~15) + (use_smem ? (size_t)total * 4 : 0);
  const unsigned grid = (unsigned)(S * chunks);
  // vector loads need the per-draw base (s * N * D floats) aligned: always true for D = 4; for D = 2 when x is 8-byte aligned
  if (D == 4 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
    histdd_kernel<4><<<grid, kHistThreads, smem, st>>>(x, (long long)N, g, edges, n_edges, counts, (int)chunks, use_smem);
  else if (D == 2 && (reinterpret_cast<uintptr_t>(x) & 7) == 0)
    histdd_kernel<2><<<grid, kHistThreads, smem, st>>>(x, (long long)N, g, edges, n_edges, counts, (int)chunks, use_smem);
  else
    histdd_kernel<0><<<grid, kHistThreads, smem, st>>>(x, (long long)N, g, edges, n_edges, counts, (int)chunks, use_smem);
  nazb_count_launch();
  if (density) {
    hist_density_kernel<<<S, 256, 0, st>>>(counts, g, edges, density);
    nazb_count_launch();
  }
  return cudaGetLastError() == cudaSuccess ? NAZB_OK : NAZB_ERR_CUDA;
}

extern "C" int nazb_hpd(const float* v, int32_t S, int64_t M, double alpha, float* lo, float* hi, void* stream) {
  if (!v || !lo || !hi || S < 1 || M < 1 || !(alpha >= 0.0 && alpha <= 1.0)) return NAZB_ERR_BAD_ARG;
  const int inc = (int)floor((1.0 - alpha) * (double)S);   // statutils.py:29-30: int(np.floor((1.0 - alpha) * ns)) in double
  if (S - inc <= 0) return NAZB_ERR_BAD_ARG;                        // "Too few elements for interval calculation"
  int n_pad = 1;
  while (n_pad < S) n_pad <<= 1;
  // columns per CTA: as many as the shared-memory budget allows (4 up to 8192 draws, 2 up to 16384, 1 up to 32768)
  int cols = kHpdCols;
  while (cols > 1 && (size_t)cols * n_pad * sizeof(float) > 200 * 1024) cols >>= 1;
  const size_t smem = (size_t)cols * n_pad * sizeof(float);
  if (smem > 200 * 1024) return NAZB_ERR_UNSUPPORTED;               // S <= 32768 draws
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaFuncSetAttribute(hpd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return NAZB_ERR_CUDA;
  const long long blocks = (M + cols - 1) / cols;
  hpd_kernel<<<(unsigned)blocks, 512, smem, st>>>(v, S, (long long)M, n_pad, inc, cols, lo, hi);
  nazb_count_launch();
  return cudaGetLastError() == cudaSuccess ? NAZB_OK : NAZB_ERR_CUDA;
}
