// SIMT (fp32 CUDA-core) engine: draw-batched MADE conditioner + affine / rational-spline transform,
// both directions, any shape.  One CTA owns a tile of P points and loops over the draws of its
// group; every activation stays in shared memory ([unit][point], point-contiguous), weights stream
// through a cp.async double-buffered shared-memory panel, and the transform + log-det run as an
// in-kernel epilogue, so conditioner outputs never touch HBM.
//
// Reference semantics reproduced (file:line in the reference tree):
//   conditioner      src/naz/flows/bflow_jax_maf.py:135-165 (context first, tanh, [.., M, D] reshape)
//   forward (sample) src/naz/flows/bflow_jax_maf.py:173-179
//   inverse (lp)     src/naz/flows/bflow_jax_maf.py:181-194  — done here as ONE block-triangular pass
//                    (NAZB_INV_INCREMENTAL) or as the reference's D full passes (NAZB_INV_JACOBI)
//   log-prob         src/naz/flows/bflow_jax_maf.py:210-212, src/naz/flows/flow.py:66-79
#include "nazb_internal.h"
#include "transforms.cuh"
#include "simt_gemm.cuh"

namespace {

template <int P>
__global__ void __launch_bounds__(kThreads) flow_simt_kernel(FlowGeom g, const float* __restrict__ packed,
                                                              const int* __restrict__ perm_all, IoArgs io) {
  using T = Tile<P>;
  extern __shared__ __align__(16) float smem[];
  // carve
  const int kin_pad = (g.kin + 3) & ~3;
  float* xin = smem;                                   // [kin_pad][P]   rows: ctx (C), x (D)
  float* ycur = xin + kin_pad * P;                     // [D][P]
  float* hbuf = ycur + g.D * P;                        // [n_hidden][hmax][P]
  float* obuf = hbuf + (size_t)g.n_hidden * g.hmax * P;  // [md_pad][P]
  const int md_pad = (g.md + 3) & ~3;
  float* ldtmp = obuf + (size_t)md_pad * P;            // [D][P]
  float* ldacc = ldtmp + g.D * P;                      // [P]
  float* ljac = ldacc + P;                             // [P]
  float* wbuf = ljac + P;                              // [2][WCHUNK]
  float* red = wbuf + 2 * T::WCHUNK;                   // [kThreads/32] doubles (as float pairs)

  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * P;
  const int npts = min(P, io.N - n0);
  const int D = g.D, C = g.C, M = g.M;
  const bool inverse = io.dir == 0;
  const bool spline = g.kind != NAZB_KIND_AFFINE;

  float run_m = -INFINITY, run_s = 0.f;   // threads < P: running logsumexp over this group's draws

  for (int si = blockIdx.y; si < io.s_count; si += gridDim.y) {
    const float* wdraw = packed + (size_t)(io.s_begin + si) * g.draw_stride;
    // ---- load tile: context rows, x rows (bounding transform for the inverse direction) ----
    for (int i = tid; i < kin_pad * P; i += kThreads) xin[i] = 0.f;
    if (tid < P) { ldacc[tid] = 0.f; ljac[tid] = 0.f; }
    __syncthreads();
    for (int i = tid; i < npts * C; i += kThreads) {
      int p = i / C, c = i % C;
      size_t row = (io.ctx_rows == 1) ? 0 : (size_t)(n0 + p);
      xin[c * P + p] = io.ctx[row * C + c];
    }
    {
      const float* xs = io.x + (size_t)si * io.x_draw_stride;
      for (int i = tid; i < npts * D; i += kThreads) {
        int p = i / D, d = i % D;
        float v = xs[(size_t)(n0 + p) * D + d];
        if (inverse) ycur[d * P + p] = v; else xin[(C + d) * P + p] = v;
      }
      for (int i = tid + npts * D; i < P * D; i += kThreads) {   // padding points of a ragged tile
        int p = i / D, d = i % D;
        if (inverse) ycur[d * P + p] = 0.f;
      }
    }
    __syncthreads();
    if (inverse && io.lo != nullptr && tid < npts) {
      float lj = 0.f;
      for (int d = 0; d < D; ++d) ycur[d * P + tid] = nazb::bound_fwd(ycur[d * P + tid], io.lo[d], io.hi[d], lj);
      ljac[tid] = lj;
    }
    __syncthreads();

    for (int li = 0; li < g.L; ++li) {
      const int l = inverse ? (g.L - 1 - li) : li;
      const float* wl = wdraw + (size_t)l * g.layer_stride;
      const int* perm = perm_all + l * D;
      const int nstage = inverse ? D : 1;
      const float* af = io.aff ? io.aff + (size_t)l * (2 * D + 1) : nullptr;
      if (inverse && af) {
        // eval-mode BatchNorm behind flow layer l (nazb_set_layer_affine): undo y = a x + b before inverting the layer
        for (int i = tid; i < P * D; i += kThreads) {
          int d = i / P;
          ycur[i] = (ycur[i] - af[D + d]) / af[d];
        }
        if (tid < P) ldacc[tid] += af[2 * D];
        __syncthreads();
      }
      for (int r = 0; r < nstage; ++r) {
        const bool full = !inverse || g.inv_mode == NAZB_INV_JACOBI;
        // ---- conditioner ----
        for (int j = 0; j < g.n_hidden; ++j) {
          int c0 = full ? 0 : g.blk[j][r], c1 = full ? g.hidden[j] : g.blk[j][r + 1];
          int K = (j == 0) ? g.kin : (full ? g.hidden[j - 1] : g.blk[j - 1][r + 1]);
          const float* act = (j == 0) ? xin : hbuf + (size_t)(j - 1) * g.hmax * P;
          if (c1 > c0)
            gemm_panel<P, true>(act, K, wl + g.off_w[j], g.ldw[j], wl + g.off_b[j], c0, c1,
                                hbuf + (size_t)j * g.hmax * P, wbuf);
        }
        {
          int j = g.n_hidden;
          int c0 = full ? 0 : r * M, c1 = full ? g.md : (r + 1) * M;
          int K = full ? g.hidden[j - 1] : g.blk[j - 1][r + 1];
          gemm_panel<P, false>(hbuf + (size_t)(j - 1) * g.hmax * P, K, wl + g.off_w[j], g.ldw[j],
                               wl + g.off_b[j], c0, c1, obuf, wbuf);
        }
        // ---- transform ----
        if (!inverse) {
          for (int i = tid; i < P * D; i += kThreads) {
            int p = i % P, rr = i / P;
            int d = perm[rr];
            float xv = xin[(C + d) * P + p], yv, ld;
            if (!spline) {
              float mu = obuf[(rr * 2 + 0) * P + p];
              float s = fminf(fmaxf(obuf[(rr * 2 + 1) * P + p], g.clip_lo), g.clip_hi);
              yv = mu + xv * expf(s);
              ld = s;
            } else {
              float* o = obuf + (size_t)(rr * M) * P + p;
              auto raw = [&](int m) { return o[m * P]; };
              auto setw = [&](int m, float v) { o[m * P] = v; };
              if (g.kind == NAZB_KIND_RQS) nazb::rational_spline<false>(xv, g.K, g.bound, false, raw, setw, yv, ld);
              else nazb::rational_spline<true>(xv, g.K, g.bound, false, raw, setw, yv, ld);
            }
            xin[(C + d) * P + p] = yv;
            ldtmp[rr * P + p] = ld;
          }
          __syncthreads();
          if (tid < P) {
            float a = ldacc[tid];
            for (int rr = 0; rr < D; ++rr) a += ldtmp[rr * P + tid];
            ldacc[tid] = a;
          }
          __syncthreads();
        } else if (!full) {
          // incremental: finalise the dimension of rank r
          if (tid < P) {
            int p = tid, d = perm[r];
            float yv = ycur[d * P + p], xv, ld;
            if (!spline) {
              float mu = obuf[(r * 2 + 0) * P + p];
              float s = fminf(fmaxf(obuf[(r * 2 + 1) * P + p], g.clip_lo), g.clip_hi);
              xv = (yv - mu) * expf(-s);
              ld = s;
            } else {
              float* o = obuf + (size_t)(r * M) * P + p;
              auto raw = [&](int m) { return o[m * P]; };
              auto setw = [&](int m, float v) { o[m * P] = v; };
              if (g.kind == NAZB_KIND_RQS) nazb::rational_spline<false>(yv, g.K, g.bound, true, raw, setw, xv, ld);
              else nazb::rational_spline<true>(yv, g.K, g.bound, true, raw, setw, xv, ld);
            }
            xin[(C + d) * P + p] = xv;
            ldacc[p] += ld;
          }
          __syncthreads();
        } else {
          // Jacobi sweep r (the reference's schedule): affine updates x[perm[r]] only, splines update
          // every dimension; the log-det kept is the one of the LAST sweep over all dimensions.
          const bool last = (r == D - 1);
          for (int i = tid; i < P * D; i += kThreads) {
            int p = i % P, rr = i / P;
            int d = perm[rr];
            float yv = ycur[d * P + p], xv, ld;
            bool upd;
            if (!spline) {
              float mu = obuf[(rr * 2 + 0) * P + p];
              float s = fminf(fmaxf(obuf[(rr * 2 + 1) * P + p], g.clip_lo), g.clip_hi);
              xv = (yv - mu) * expf(-s);
              ld = s;
              upd = (rr == r);
            } else {
              float* o = obuf + (size_t)(rr * M) * P + p;
              auto raw = [&](int m) { return o[m * P]; };
              auto setw = [&](int m, float v) { o[m * P] = v; };
              if (g.kind == NAZB_KIND_RQS) nazb::rational_spline<false>(yv, g.K, g.bound, true, raw, setw, xv, ld);
              else nazb::rational_spline<true>(yv, g.K, g.bound, true, raw, setw, xv, ld);
              upd = true;
            }
            ldtmp[rr * P + p] = ld;
            // x is only read by the NEXT sweep's conditioner, so the in-place update is safe
            if (upd) xin[(C + d) * P + p] = xv;
          }
          __syncthreads();
          if (last && tid < P) {
            float a = ldacc[tid];
            for (int rr = 0; rr < D; ++rr) a += ldtmp[rr * P + tid];
            ldacc[tid] = a;
          }
          __syncthreads();
        }
      }
      if (!inverse && af) {
        for (int i = tid; i < P * D; i += kThreads) {
          int d = i / P;
          xin[(size_t)C * P + i] = fmaf(af[d], xin[(size_t)C * P + i], af[D + d]);
        }
        if (tid < P) ldacc[tid] += af[2 * D];
        __syncthreads();
      }
      if (inverse) {
        // x of this layer becomes y of the next (earlier) layer; x restarts from zero
        for (int i = tid; i < P * D; i += kThreads) {
          int p = i % P, d = i / P;
          ycur[d * P + p] = xin[(C + d) * P + p];
          xin[(C + d) * P + p] = 0.f;
        }
        __syncthreads();
      }
    }

    // ---- outputs ----
    float lp = 0.f;
    if (tid < P) {
      if (inverse) {
        float q = 0.f;
        for (int d = 0; d < D; ++d) { float z = ycur[d * P + tid]; q += 0.5f * z * z; }
        lp = -q - 0.5f * D * NAZB_LOG_2PI - ldacc[tid] + ljac[tid];
      } else {
        lp = ldacc[tid];
      }
      if (tid < npts && io.out_l) io.out_l[(size_t)si * io.N + n0 + tid] = lp;
      if (inverse && tid < npts && io.lse_max) {
        float v = lp + (io.log_w ? io.log_w[si] : 0.f);
        if (!(v <= run_m)) { run_s = run_s * expf(run_m - v) + 1.f; run_m = v; }   // also propagates NaN
        else if (v > -INFINITY) run_s += expf(v - run_m);
      }
    }
    if (io.out_x) {
      const float* src = inverse ? ycur : xin + C * P;
      float* dst = io.out_x + ((size_t)si * io.N + n0) * D;
      const bool unbound = (!inverse) && io.lo != nullptr;
      for (int i = tid; i < npts * D; i += kThreads) {
        int p = i / D, d = i % D;
        float v = src[d * P + p];
        if (unbound) v = nazb::bound_inv(v, io.lo[d], io.hi[d]);
        dst[i] = v;
      }
    }
    if (inverse && io.sum_n) {
      double v = (tid < npts) ? (double)lp : 0.0;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      double* redd = reinterpret_cast<double*>(red);
      if ((tid & 31) == 0) redd[tid >> 5] = v;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += redd[w];
        atomicAdd(io.sum_n + si, t);
      }
    }
    __syncthreads();
  }
  if (inverse && io.lse_max && tid < npts) {
    io.lse_max[(size_t)blockIdx.y * io.N + n0 + tid] = run_m;
    io.lse_sum[(size_t)blockIdx.y * io.N + n0 + tid] = run_s;
  }
}

}  // namespace

size_t nazb_simt_smem_bytes(const FlowGeom& g, int P) {
  const int kin_pad = (g.kin + 3) & ~3, md_pad = (g.md + 3) & ~3;
  size_t f = (size_t)kin_pad * P + (size_t)g.D * P + (size_t)g.n_hidden * g.hmax * P + (size_t)md_pad * P +
             (size_t)g.D * P + 2 * P;
  int TR = P / 4, TC = kThreads / TR, NPASS = TC * 4;
  f += 2 * (size_t)kKC * NPASS;
  f += 2 * (kThreads / 32) + 4;
  return f * sizeof(float) + 16;
}

int nazb_simt_pick_P(const FlowGeom& g) {
  const size_t cap = 227 * 1024;
  if (2 * (nazb_simt_smem_bytes(g, 32) + 1024) <= cap) return 32;   // two CTAs per SM
  if (nazb_simt_smem_bytes(g, 64) <= cap) return 64;
  if (nazb_simt_smem_bytes(g, 32) <= cap) return 32;
  if (nazb_simt_smem_bytes(g, 16) <= cap) return 16;
  return 0;
}

cudaError_t nazb_simt_launch(const nazb_handle* h, const IoArgs& io, int n_groups, cudaStream_t st) {
  const FlowGeom& g = h->geom;
  int P = nazb_simt_pick_P(g);
  if (P == 0) return cudaErrorInvalidConfiguration;
  size_t smem = nazb_simt_smem_bytes(g, P);
  dim3 grid((io.N + P - 1) / P, n_groups);
  cudaError_t e;
#define LAUNCH(PP)                                                                                      \
  e = cudaFuncSetAttribute(flow_simt_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e != cudaSuccess) return e;                                                                       \
  flow_simt_kernel<PP><<<grid, kThreads, smem, st>>>(g, h->packed, h->perm_dev, io);
  if (P == 64) { LAUNCH(64) } else if (P == 32) { LAUNCH(32) } else { LAUNCH(16) }
#undef LAUNCH
  nazb_count_launch();
  return cudaGetLastError();
}
