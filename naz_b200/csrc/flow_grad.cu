// Gradient of the draw-batched log-likelihood  sum_n log p(x_n | c_n; theta_s)  with respect to every weight and bias of
// draw s (and, optionally, to the points) for masked-affine AND neural-spline (quadratic and linear order) autoregressive flows — SURVEY §8 row f1: the inner loop of
// the reference's NUTS / SVI / MLE drivers, which call jax.grad / torch autograd on exactly this scalar
// (src/naz/flows/bflow_jax_maf.py:233-246 log_prob -> :321-327 NUTS, :344-348 SVI, :277-287 MLE;
//  src/naz/trainers/train_flows.py:195-213).
//
// fp32 CUDA-core engine (first cut of the row; same tiling as flow_simt.cu).  One CTA = one tile of P points of one draw:
//   phase A  the incremental inverse of flow_simt.cu, keeping x^(l) (the conditioner input of every flow layer) in
//            shared memory: L * D * P floats;
//   phase B  flow layers in reverse evaluation order.  With x = x^(l) known, one conditioner pass gives h_j, mu, s.
//            The layer is the implicit relation  x = (y - mu(x)) exp(-s(x)),  lp = G(x) - sum_d s_d(x), so with
//            g = dG/dx the adjoint lambda solves  lambda = g + J^T c(lambda),
//                 c_mu = -exp(-s) * lambda,   c_s = -(1 + x * lambda) * [lo <= s_raw <= hi],
//            where J is the conditioner's Jacobian, strictly triangular in permutation order: D-1 sweeps (each one
//            back-propagation to the input, restricted to the MADE blocks above the rank it finalises) make every rank exact.  A last back-propagation with the final
//            c accumulates  dW_j += h_{j-1} (x) delta_j,  db_j += delta_j  over the tile and adds them to the caller's
//            gradient arrays (reference layout [out][in], masked entries untouched) with fp32 atomics;
//            g <- lambda * exp(-s) is handed to the next layer.
// Neural-spline layers (train_flows.py:195-213 differentiates them through pyro's spline): the same recursion with the general
// form of the cotangent.  A layer is  y_r = T(x_r; p_r(x_<r)),  lp = G(x) - sum_r ld(x_r; p_r),  ld = log dT/dx, hence
//      lambda = g - d ld/d x + J^T c(lambda),    c_m = lambda_r * ca_m + cb_m,
//      ca_m = -(dT/dp_m) / (dT/dx)   (= d x_r / d p_m at fixed y),   cb_m = -d ld / d p_m,
// and  d lp / d y_r = lambda_r / (dT/dx).  ca, cb, 1/(dT/dx) and d ld/dx are evaluated once per layer at the solved x
// (spline_grad.cuh: forward-mode duals over the six numbers of the selected bin, then the softmax / softplus reverse steps)
// and kept in shared memory; the affine layer is the special case ca = (-e^-s, -x [clip]), cb = (0, -[clip]).
// Back-propagation to the input needs the weights transposed: a second image (rows = output units) is produced from
// the SIMT image once per pack.
#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include "nazb_internal.h"
#include "transforms.cuh"
#include "spline_grad.cuh"
#include "simt_gemm.cuh"

namespace {

struct GradGeom {
  long long offT[NAZB_MAX_LIN];   // float offset of W_j^T ([ldw_j rows][ldk_j]) inside one (draw, layer) block
  int ldk[NAZB_MAX_LIN];          // kdim_j rounded up to 4
  long long layer_strideT, draw_strideT;
};

struct GradArgs {
  const float* const* mask;   // device table [L * n_lin] of device pointers, [out][in] 0/1
  float* const* gW;           // ... gradient arrays, [S][out][in] (+=)
  float* const* gb;           // ... [S][out] (+=)
  const long long* gwst;      // floats between draws of gW[i]
  const long long* gbst;
  float* dx;                  // [s_count][N][D] or null
  float* dctx;                // [s_count][N][C] or null: cotangent of the context rows (a trainable embedding net upstream)
  float* stash;               // per resident CTA: [L][n_hidden * hmax * P + md_pad * P] conditioner activations / raw outputs of phase A
  long long stash_cta;        // floats per CTA (0 = no stash: phase B recomputes the conditioner)
  int tiles;                  // point tiles per draw; work item w = draw * tiles + tile, CTA b takes w = b, b + gridDim.x, ...
  const float* wgt;           // per-point cotangents of lp, [s_count][N] (draw stride wgt_stride, 0 = shared) or null (= 1)
  long long wgt_stride;
  int diag;                   // dev: 1 = skip the atomics (timing diagnosis only)
};

// dst[(draw, layer)][n][k] = src[(draw, layer)][k][n]
__global__ void transpose_image_kernel(const float* __restrict__ src, float* __restrict__ dst, FlowGeom g, GradGeom gg,
                                       int n_lin, long long blocks_dl) {
  const long long per = gg.layer_strideT;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < blocks_dl * per;
       i += (long long)gridDim.x * blockDim.x) {
    const long long dl = i / per;
    const long long f = i % per;
    int j = 0;
    while (j + 1 < n_lin && f >= gg.offT[j + 1]) ++j;
    const long long e = f - gg.offT[j];
    const int n = (int)(e / gg.ldk[j]), k = (int)(e % gg.ldk[j]);
    float v = 0.f;
    if (k < g.kdim[j] && n < g.ldw[j]) v = src[dl * g.layer_stride + g.off_w[j] + (long long)k * g.ldw[j] + n];
    dst[dl * per + f] = v;
  }
}

// gW[row(n)][k] += sum_p aT[k][p] * bT[n][p]  for unmasked (n, k);  gb[row(n)] += sum_p bT[n][p].
// row(n) = n for hidden layers; the output layer's columns are rank-major (n = rank * M + m) while the reference's rows
// are slot-major (m * D + dim)  (bflow_jax_maf.py:161-163 reshape [.., M, D]).
template <int P>
__device__ __forceinline__ void outer_acc(const float* __restrict__ aT, int K, const float* __restrict__ bT, int Nn,
                                          const float* __restrict__ mask, float* __restrict__ gW,
                                          float* __restrict__ gb, const int* __restrict__ perm, int M, int D,
                                          bool out_layer, int diag) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int kt = (K + 3) >> 2, nt = (Nn + 3) >> 2;
  for (int t = tid; t < kt * nt; t += kThreads) {
    const int k0 = (t % kt) * 4, n0 = (t / kt) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int pp = 0; pp < P / 4; ++pp) {
      const int p = ((pp + lane) & (P / 4 - 1)) * 4;   // rotated per lane: conflict-free float4 reads
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(aT + (size_t)(k0 + i) * P + p);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(bT + (size_t)(n0 + j) * P + p);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = fmaf(a[i].x, b[j].x, fmaf(a[i].y, b[j].y, fmaf(a[i].z, b[j].z, fmaf(a[i].w, b[j].w, acc[i][j]))));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + j;
      if (n >= Nn) continue;
      const int row = out_layer ? ((n % M) * D + perm[n / M]) : n;
      const float* mrow = mask + (size_t)row * K + k0;
      float* grow = gW + (size_t)row * K + k0;
      // pairs of neighbouring entries go out as one 8-byte reduction when both are unmasked and the address allows it
      // (MADE masks are block-structured, so pairs almost always share their mask bit): half the mask loads and atomics
      const bool pairs = ((K & 1) == 0) && ((reinterpret_cast<uintptr_t>(grow) & 7) == 0) && ((reinterpret_cast<uintptr_t>(mrow) & 7) == 0);
#pragma unroll
      for (int i = 0; i < 4; i += 2) {
        const int k = k0 + i;
        if (pairs && k + 1 < K) {
          const float2 m = *reinterpret_cast<const float2*>(mrow + i);
          const bool d0 = !diag || acc[i][j] == 12345.678f, d1 = !diag || acc[i + 1][j] == 12345.678f;
          if (m.x != 0.f && m.y != 0.f && d0 && d1) atomicAdd(reinterpret_cast<float2*>(grow + i), make_float2(acc[i][j], acc[i + 1][j]));
          else {
            if (m.x != 0.f && d0) atomicAdd(grow + i, acc[i][j]);
            if (m.y != 0.f && d1) atomicAdd(grow + i + 1, acc[i + 1][j]);
          }
        } else {
          if (k < K && mrow[i] != 0.f && (!diag || acc[i][j] == 12345.678f)) atomicAdd(grow + i, acc[i][j]);
          if (k + 1 < K && mrow[i + 1] != 0.f && (!diag || acc[i + 1][j] == 12345.678f)) atomicAdd(grow + i + 1, acc[i + 1][j]);
        }
      }
    }
  }
  for (int n = tid; n < Nn; n += kThreads) {
    float s = 0.f;
    for (int pp = 0; pp < P; ++pp) s += bT[(size_t)n * P + ((pp + lane) & (P - 1))];
    const int row = out_layer ? ((n % M) * D + perm[n / M]) : n;
    atomicAdd(gb + row, s);
  }
}

// P = 32, TN = 5 (160-column passes), one CTA per SM: the default;  P = 16, TN = 3 (192-column passes): half the state
template <int P, int NST, int TN, bool SPLINE>
__global__ void __launch_bounds__(kThreads, (P == 16) ? 2 : 1) flow_grad_kernel(FlowGeom g, GradGeom gg,
                                                                     const float* __restrict__ packed,
                                                                     const float* __restrict__ packedT,
                                                                     const int* __restrict__ perm_all, IoArgs io,
                                                                     GradArgs ga) {
  using T = Tile<P, TN>;
  extern __shared__ __align__(16) float smem[];
  const int D = g.D, C = g.C, nh = g.n_hidden, n_lin = nh + 1, M = g.M;
  const int kin_pad = (g.kin + 3) & ~3, md_pad = (g.md + 3) & ~3;
  float* xin = smem;                                   // [kin_pad][P]  rows: ctx (C), x (D)
  float* gcur = xin + kin_pad * P;                     // [D][P]        phase A: y;  phase B: g = d lp / d x
  float* hbuf = gcur + D * P;                          // [nh][hmax][P]
  float* obuf = hbuf + (size_t)nh * g.hmax * P;        // [md_pad][P]   conditioner output ((mu, s_raw) / 3K-1 spline slots), rank-major
  float* cbuf = obuf + (size_t)md_pad * P;             // [md_pad][P]   cotangent of the conditioner output
  float* dA = cbuf + (size_t)md_pad * P;               // [hmax][P]     hidden cotangents (ping-pong)
  float* dB = dA + (size_t)g.hmax * P;
  float* dxb = dB + (size_t)g.hmax * P;                // [kin_pad][P]  cotangent of the conditioner input
  float* chain = dxb + kin_pad * P;                    // [L][D][P]     x^(l)
  float* lam = chain + (size_t)g.L * D * P;            // [D][P]
  float* es = lam + D * P;                             // [D][P]  1 / (dT/dx) = exp(-ld), by rank
  float* ca = es + D * P;                              // [md_pad][P]  c = lambda * ca + cb  (see the header)
  float* cb = ca + (size_t)md_pad * P;                 // [md_pad][P]
  float* ldacc = cb + (size_t)md_pad * P;              // [P]
  float* ljac = ldacc + P;                             // [P]
  float* pwv = ljac + P;                               // [P]  cotangent of lp per point (nazb_inverse_vjp; 1 for the plain gradient)
  float* dcacc = pwv + P;                              // [C][P]  d (w lp) / d ctx, summed over the flow layers
  // zero bias of the transposed products: they read bias[n] for n < max(hidden widths, kin) (the back-propagation to the
  // conditioner input has kin = C + D output columns, which a wide context can make larger than every hidden layer)
  const int zb_n = (g.hmax > kin_pad) ? g.hmax : kin_pad;
  float* zb = dcacc + (size_t)C * P;                   // [zb_n]
  float* wbuf = zb + zb_n;                             // [NST][WCHUNK]
  float* red = wbuf + NST * T::WCHUNK;

  const int tid = threadIdx.x;
  // Phase A leaves every hidden activation and raw conditioner output of a flow layer final (block r at stage r): instead of
  // recomputing the conditioner at the solved x in phase B they are parked in a per-CTA scratch area in global memory
  // (L2-resident: ~1 MB per CTA written and read once per work item) — one conditioner pass less per layer.
  const bool use_stash = ga.stash_cta > 0 && g.inv_mode == NAZB_INV_INCREMENTAL;
  float* stash = use_stash ? ga.stash + (size_t)blockIdx.x * ga.stash_cta : nullptr;
  const size_t stash_h = (size_t)nh * g.hmax * P, stash_l = stash_h + (size_t)md_pad * P;

  for (long long item = blockIdx.x; item < (long long)ga.tiles * io.s_count; item += gridDim.x) {
    const int si = (int)(item / ga.tiles);
    const int n0 = (int)(item % ga.tiles) * P;
    const int npts = min(P, io.N - n0);
    const int sg = io.s_begin + si;
    const float* wdraw = packed + (size_t)sg * g.draw_stride;
    const float* wdrawT = packedT + (size_t)sg * gg.draw_strideT;
    // ---- load tile ----
    for (int i = tid; i < kin_pad * P; i += kThreads) xin[i] = 0.f;
    for (int i = tid; i < zb_n; i += kThreads) zb[i] = 0.f;
    for (int i = tid; i < C * P; i += kThreads) dcacc[i] = 0.f;
    if (tid < P) {
      ldacc[tid] = 0.f; ljac[tid] = 0.f;
      pwv[tid] = (tid < npts) ? (ga.wgt ? ga.wgt[(size_t)si * ga.wgt_stride + n0 + tid] : 1.f) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < npts * C; i += kThreads) {
      int p = i / C, c = i % C;
      size_t row = (io.ctx_rows == 1) ? 0 : (size_t)(n0 + p);
      xin[c * P + p] = io.ctx[row * C + c];
    }
    for (int i = tid; i < P * D; i += kThreads) {
      int p = i / D, d = i % D;
      gcur[d * P + p] = (p < npts) ? io.x[(size_t)(n0 + p) * D + d] : 0.f;
    }
    __syncthreads();
    if (io.lo != nullptr && tid < npts) {
      float lj = 0.f;
      for (int d = 0; d < D; ++d) gcur[d * P + tid] = nazb::bound_fwd(gcur[d * P + tid], io.lo[d], io.hi[d], lj);
      ljac[tid] = lj;
    }
    __syncthreads();

    // ================= phase A: inverse (log_prob direction), keeping x^(l) =================
    for (int li = 0; li < g.L; ++li) {
      const int l = g.L - 1 - li;
      const float* wl = wdraw + (size_t)l * g.layer_stride;
      const int* perm = perm_all + l * D;
      const bool full = g.inv_mode == NAZB_INV_JACOBI;
      if (io.aff != nullptr) {
        // eval-mode BatchNorm behind flow layer l (nazb_set_layer_affine): undo y_hat = a y + b before inverting the layer
        const float* af = io.aff + (size_t)l * (2 * D + 1);
        for (int i = tid; i < P * D; i += kThreads) { const int d = i / P; gcur[i] = (gcur[i] - af[D + d]) / af[d]; }
        if (tid < P) ldacc[tid] += af[2 * D];
        __syncthreads();
      }
      for (int r = 0; r < D; ++r) {
        for (int j = 0; j < nh; ++j) {
          int c0 = full ? 0 : g.blk[j][r], c1 = full ? g.hidden[j] : g.blk[j][r + 1];
          int K = (j == 0) ? g.kin : (full ? g.hidden[j - 1] : g.blk[j - 1][r + 1]);
          const float* act = (j == 0) ? xin : hbuf + (size_t)(j - 1) * g.hmax * P;
          if (c1 > c0)
            gemm_panel_ms<P, true, TN, NST>(act, K, wl + g.off_w[j], g.ldw[j], wl + g.off_b[j], c0, c1, hbuf + (size_t)j * g.hmax * P, wbuf);
        }
        {
          int K = full ? g.hidden[nh - 1] : g.blk[nh - 1][r + 1];
          gemm_panel_ms<P, false, TN, NST>(hbuf + (size_t)(nh - 1) * g.hmax * P, K, wl + g.off_w[nh], g.ldw[nh], wl + g.off_b[nh],
                               r * M, (r + 1) * M, obuf, wbuf);
        }
        if (tid < P) {
          const int p = tid, d = perm[r];
          if (use_stash) {
            float* so = stash + (size_t)l * stash_l + stash_h + (size_t)(r * M) * P + p;
            for (int m = 0; m < M; ++m) so[m * P] = obuf[(size_t)(r * M + m) * P + p];   // raw outputs, before the spline's scratch writes
          }
          if (!SPLINE) {
            float mu = obuf[(r * 2 + 0) * P + p];
            float s = fminf(fmaxf(obuf[(r * 2 + 1) * P + p], g.clip_lo), g.clip_hi);
            xin[(C + d) * P + p] = (gcur[d * P + p] - mu) * expf(-s);
            ldacc[p] += s;
          } else {
            float* o = obuf + (size_t)(r * M) * P + p;
            auto raw = [&](int m) { return o[m * P]; };
            auto setw = [&](int m, float v) { o[m * P] = v; };
            float xv, ld;
            if (g.kind == NAZB_KIND_RQS) nazb::rational_spline<false>(gcur[d * P + p], g.K, g.bound, true, raw, setw, xv, ld);
            else nazb::rational_spline<true>(gcur[d * P + p], g.K, g.bound, true, raw, setw, xv, ld);
            xin[(C + d) * P + p] = xv;
            ldacc[p] += ld;
          }
        }
        __syncthreads();
      }
      for (int i = tid; i < P * D; i += kThreads) {
        int p = i % P, d = i / P;
        float xv = xin[(C + d) * P + p];
        chain[((size_t)l * D + d) * P + p] = xv;
        gcur[d * P + p] = xv;
        xin[(C + d) * P + p] = 0.f;
      }
      if (use_stash) {
        float4* dst = reinterpret_cast<float4*>(stash + (size_t)l * stash_l);
        const float4* src = reinterpret_cast<const float4*>(hbuf);
        for (size_t i = tid; i < stash_h / 4; i += kThreads) dst[i] = src[i];
      }
      __syncthreads();
    }
    // ---- value: sum_n lp;  g = d lp / d z = -z ----
    {
      float lp = 0.f;
      if (tid < P) {
        float q = 0.f;
        for (int d = 0; d < D; ++d) { float z = gcur[d * P + tid]; q += 0.5f * z * z; }
        lp = -q - 0.5f * D * NAZB_LOG_2PI - ldacc[tid] + ljac[tid];
        if (tid < npts && io.out_l) io.out_l[(size_t)si * io.N + n0 + tid] = lp;
      }
      if (io.sum_n) {
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
      for (int i = tid; i < P * D; i += kThreads) {
        int p = i % P;
        gcur[i] = -gcur[i] * pwv[p];                   // d (w lp) / d z = -w z  (w = 0 on the padding points of a ragged tile)
      }
      __syncthreads();
    }

    // ================= phase B: adjoint, flow layers in reverse evaluation order =================
    for (int l = 0; l < g.L; ++l) {
      const float* wl = wdraw + (size_t)l * g.layer_stride;
      const float* wlT = wdrawT + (size_t)l * gg.layer_strideT;
      const int* perm = perm_all + l * D;
      for (int i = tid; i < P * D; i += kThreads) {
        int p = i % P, d = i / P;
        xin[(C + d) * P + p] = chain[((size_t)l * D + d) * P + p];
      }
      __syncthreads();
      // conditioner at the solved x: parked by phase A, or recomputed
      if (use_stash) {
        const float4* src = reinterpret_cast<const float4*>(stash + (size_t)l * stash_l);
        float4* dh = reinterpret_cast<float4*>(hbuf);
        float4* dobuf = reinterpret_cast<float4*>(obuf);
        for (size_t i = tid; i < stash_h / 4; i += kThreads) dh[i] = src[i];
        for (size_t i = tid; i < (size_t)md_pad * P / 4; i += kThreads) dobuf[i] = src[stash_h / 4 + i];
        __syncthreads();
      } else {
        for (int j = 0; j < nh; ++j)
          gemm_panel_ms<P, true, TN, NST>((j == 0) ? xin : hbuf + (size_t)(j - 1) * g.hmax * P, g.kdim[j], wl + g.off_w[j], g.ldw[j],
                              wl + g.off_b[j], 0, g.hidden[j], hbuf + (size_t)j * g.hmax * P, wbuf);
        gemm_panel_ms<P, false, TN, NST>(hbuf + (size_t)(nh - 1) * g.hmax * P, g.kdim[nh], wl + g.off_w[nh], g.ldw[nh],
                             wl + g.off_b[nh], 0, g.md, obuf, wbuf);
      }
      for (int i = tid; i < P * D; i += kThreads) {
        const int p = i % P, r = i / P, d = perm[r];
        const float xv = xin[(C + d) * P + p];
        float gv = gcur[d * P + p];
        if (!SPLINE) {
          const float sraw = obuf[(r * 2 + 1) * P + p];
          const float s = fminf(fmaxf(sraw, g.clip_lo), g.clip_hi);
          const float e = expf(-s), mk = (sraw >= g.clip_lo && sraw <= g.clip_hi) ? 1.f : 0.f;
          es[r * P + p] = e;
          ca[(r * 2 + 0) * P + p] = -e;      cb[(r * 2 + 0) * P + p] = 0.f;
          ca[(r * 2 + 1) * P + p] = -xv * mk; cb[(r * 2 + 1) * P + p] = -mk * pwv[p];
        } else {
          const float* o = obuf + (size_t)(r * M) * P + p;
          float* pa = ca + (size_t)(r * M) * P + p;
          float* pb = cb + (size_t)(r * M) * P + p;
          float itx, ldx;
          const float pw = pwv[p];
          auto rawf = [&](int m) { return o[m * P]; };
          auto putf = [&](int m, float a, float b) { pa[m * P] = a; pb[m * P] = b * pw; };
          if (g.kind == NAZB_KIND_RQS) nazb::spline_grad<false>(xv, g.K, g.bound, rawf, putf, itx, ldx);
          else nazb::spline_grad<true>(xv, g.K, g.bound, rawf, putf, itx, ldx);
          es[r * P + p] = itx;
          gv -= ldx * pw;                         // the direct dependence of ld on x joins g for the rest of this layer
          gcur[d * P + p] = gv;
        }
        lam[d * P + p] = gv;
      }
      __syncthreads();
      for (int it = 0; it < D; ++it) {
        const bool last = (it == D - 1);
        for (int i = tid; i < P * g.md; i += kThreads) {
          const int p = i % P, row = i / P;
          const float lm = lam[perm[row / M] * P + p];
          cbuf[i] = (p < npts) ? fmaf(lm, ca[i], cb[i]) : 0.f;
        }
        __syncthreads();
        if (!last && g.inv_mode == NAZB_INV_INCREMENTAL) {
          // Sweep `it` only has to make rank t = D-2-it exact, from the ranks above it: x of rank t reaches the outputs
          // of ranks > t through the hidden units of blocks > t only (block b = units that depend on ranks < b), so the
          // back-propagation is restricted to output columns >= M (t+1) and hidden units >= blk[.][t+1].
          const int t = D - 2 - it, r1 = t + 1, dt = perm[t];
          const float* src = cbuf + (size_t)(r1 * M) * P;
          int ksrc = g.md - r1 * M, woff = r1 * M;
          for (int j = nh; j >= 1 && ksrc > 0; --j) {
            float* dst = ((nh - j) & 1) ? dB : dA;
            const int c0 = g.blk[j - 1][r1];
            if (c0 < g.kdim[j]) {
              gemm_panel_ms<P, false, TN, NST>(src, ksrc, wlT + gg.offT[j] + (size_t)woff * gg.ldk[j], gg.ldk[j], zb, c0, g.kdim[j],
                                               dst, wbuf);
              const float* hj = hbuf + (size_t)(j - 1) * g.hmax * P;
              for (int i = tid + c0 * P; i < g.kdim[j] * P; i += kThreads) { float hv = hj[i]; dst[i] *= (1.f - hv * hv); }
              __syncthreads();
            }
            src = dst + (size_t)c0 * P;
            ksrc = g.kdim[j] - c0;
            woff = c0;
          }
          if (ksrc > 0) {
            gemm_panel_ms<P, false, TN, NST>(src, ksrc, wlT + gg.offT[0] + (size_t)woff * gg.ldk[0], gg.ldk[0], zb, 0, g.kin, dxb, wbuf);
            if (tid < P) lam[dt * P + tid] = gcur[dt * P + tid] + dxb[(C + dt) * P + tid];
          }
          __syncthreads();
          continue;
        }
        const int to = l * n_lin + nh;
        if (last)
          outer_acc<P>(hbuf + (size_t)(nh - 1) * g.hmax * P, g.kdim[nh], cbuf, g.md, ga.mask[to],
                       ga.gW[to] + (size_t)sg * ga.gwst[to], ga.gb[to] + (size_t)sg * ga.gbst[to], perm, M, D, true, ga.diag);
        const float* src = cbuf;
        for (int j = nh; j >= 1; --j) {
          float* dst = ((nh - j) & 1) ? dB : dA;
          // delta_{j-1}[k][p] = (sum_n W_j[n][k] delta_j[n][p]) * (1 - h_{j-1}[k][p]^2)
          gemm_panel_ms<P, false, TN, NST>(src, g.ndim[j], wlT + gg.offT[j], gg.ldk[j], zb, 0, g.kdim[j], dst, wbuf);
          const float* hj = hbuf + (size_t)(j - 1) * g.hmax * P;
          for (int i = tid; i < g.kdim[j] * P; i += kThreads) { float hv = hj[i]; dst[i] *= (1.f - hv * hv); }
          __syncthreads();
          if (last) {
            const int ti = l * n_lin + (j - 1);
            outer_acc<P>((j - 1 == 0) ? xin : hbuf + (size_t)(j - 2) * g.hmax * P, g.kdim[j - 1], dst, g.ndim[j - 1],
                         ga.mask[ti], ga.gW[ti] + (size_t)sg * ga.gwst[ti], ga.gb[ti] + (size_t)sg * ga.gbst[ti], perm, M, D,
                         false, ga.diag);
          }
          src = dst;
        }
        if (!last) {
          gemm_panel_ms<P, false, TN, NST>(src, g.ndim[0], wlT + gg.offT[0], gg.ldk[0], zb, 0, g.kin, dxb, wbuf);
          for (int i = tid; i < P * D; i += kThreads) {
            int p = i % P, d = i / P;
            lam[i] = gcur[i] + dxb[(C + d) * P + p];
          }
        } else if (ga.dctx != nullptr && C > 0) {
          // the context enters every layer's first linear: its cotangent is the first C columns of W_0^T delta_0, summed over layers
          gemm_panel_ms<P, false, TN, NST>(src, g.ndim[0], wlT + gg.offT[0], gg.ldk[0], zb, 0, C, dxb, wbuf);
          for (int i = tid; i < C * P; i += kThreads) dcacc[i] += dxb[i];
        }
        __syncthreads();
      }
      // g of the next layer: d lp / d y = lambda / (dT/dx)
      for (int i = tid; i < P * D; i += kThreads) {
        int p = i % P, r = i / P;
        int d = perm[r];
        float gy = lam[d * P + p] * es[r * P + p];
        if (io.aff != nullptr) gy /= io.aff[(size_t)l * (2 * D + 1) + d];   // through y_hat = a y + b (constants) to the next layer
        gcur[d * P + p] = gy;
      }
      __syncthreads();
    }
    if (ga.dctx != nullptr && C > 0) {
      float* dst = ga.dctx + ((size_t)si * io.N + n0) * C;
      for (int i = tid; i < npts * C; i += kThreads) dst[i] = dcacc[(i % C) * P + i / C];
    }
    if (ga.dx) {
      float* dst = ga.dx + ((size_t)si * io.N + n0) * D;
      for (int i = tid; i < npts * D; i += kThreads) {
        int p = i / D, d = i % D;
        float v = gcur[d * P + p];
        if (io.lo != nullptr) {
          // y = logit(u), u = (x - lo) / (hi - lo);  lp also carries -log u - log(1 - u) - log(hi - lo)
          float w = io.hi[d] - io.lo[d];
          float u = (io.x[(size_t)(n0 + p) * D + d] - io.lo[d]) / w;
          v = v / (w * u * (1.f - u)) - pwv[p] * (1.f / u - 1.f / (1.f - u)) / w;
        }
        dst[i] = v;
      }
    }
    __syncthreads();
  }
}

size_t grad_smem_bytes(const FlowGeom& g, int P, int nst, int tn) {
  const int kin_pad = (g.kin + 3) & ~3, md_pad = (g.md + 3) & ~3;
  size_t f = (size_t)2 * kin_pad * P + (size_t)3 * g.D * P + (size_t)(g.n_hidden + 2) * g.hmax * P +
             (size_t)4 * md_pad * P + (size_t)g.L * g.D * P + 3 * P + (size_t)g.C * P + (size_t)((g.hmax > kin_pad) ? g.hmax : kin_pad);
  int TR = P / 4, TC = kThreads / TR, NPASS = TC * tn;
  f += (size_t)nst * kKC * NPASS;
  f += 2 * (kThreads / 32) + 4;
  return f * sizeof(float) + 16;
}

GradGeom make_grad_geom(const FlowGeom& g) {
  GradGeom gg{};
  long long off = 0;
  for (int j = 0; j <= g.n_hidden; ++j) {
    gg.ldk[j] = (g.kdim[j] + 3) & ~3;
    gg.offT[j] = off;
    off += (long long)g.ldw[j] * gg.ldk[j];
  }
  gg.layer_strideT = off;
  gg.draw_strideT = off * g.L;
  return gg;
}

}  // namespace

// Gradient launcher.  `tabs` = device memory holding the five tables (mask, gW, gb pointers; gW, gb draw strides).
cudaError_t nazb_grad_launch(nazb_handle* h, const IoArgs& io, const void* tabs, float* dx, float* dctx, const float* wgt,
                             long long wgt_stride, cudaStream_t st) {
  const FlowGeom& g = h->geom;
  const GradGeom gg = make_grad_geom(g);
  const int n = g.L * (g.n_hidden + 1);
  cudaError_t e;
  if (!h->packed_T) {
    e = cudaMalloc(&h->packed_T, sizeof(float) * (size_t)h->desc.S * gg.draw_strideT);
    if (e != cudaSuccess) return e;
    h->packed_T_valid = false;
  }
  if (!h->packed_T_valid) {
    const long long blocks_dl = (long long)h->desc.S * g.L;
    const long long total = blocks_dl * gg.layer_strideT;
    int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    transpose_image_kernel<<<blocks, 256, 0, st>>>(h->packed, h->packed_T, g, gg, g.n_hidden + 1, blocks_dl);
    nazb_count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    h->packed_T_valid = true;
  }
  GradArgs ga{};
  const char* t = static_cast<const char*>(tabs);
  ga.mask = reinterpret_cast<const float* const*>(t);
  ga.gW = reinterpret_cast<float* const*>(t + sizeof(void*) * n);
  ga.gb = reinterpret_cast<float* const*>(t + sizeof(void*) * 2 * n);
  ga.gwst = reinterpret_cast<const long long*>(t + sizeof(void*) * 3 * n);
  ga.gbst = reinterpret_cast<const long long*>(t + sizeof(void*) * 4 * n);
  ga.dx = dx;
  ga.dctx = dctx;
  ga.wgt = wgt; ga.wgt_stride = wgt_stride;
  ga.diag = h->opt_grad_diag;
  // Measured (maf 2|2, 4 chains x 100 k points): one 32-point CTA per SM 245 ms, two 16-point CTAs per SM 274 ms — the
  // shared-memory wavefronts per point double with the smaller tile — so 16-point tiles only serve shapes whose 32-point
  // state does not fit.  The 4-deep weight ring costs nothing at one CTA per SM and is dropped first.
  const size_t cap = 227 * 1024;
  int P = 32, nst = 4;
  if (grad_smem_bytes(g, 32, 4, 5) > cap) nst = 2;
  if (grad_smem_bytes(g, 32, nst, 5) > cap) P = 16;
  if (h->opt_grad_tile == 16) P = 16;
  const size_t smem = (P == 16) ? grad_smem_bytes(g, 16, 2, 3) : grad_smem_bytes(g, 32, nst, 5);
  if (smem > cap) return cudaErrorInvalidConfiguration;
  const int tiles = (io.N + P - 1) / P;
  // persistent CTAs over (draw, tile) work items, draw-major (concurrent CTAs share a draw's weights in L2)
  const long long items = (long long)tiles * io.s_count;
  const int resident = h->sm_count * ((P == 16) ? 2 : 1);
  dim3 grid((unsigned)std::min<long long>(items, resident));
  ga.tiles = tiles;
  {
    const int md_pad = (g.md + 3) & ~3;
    const long long per_cta = (long long)g.L * ((long long)g.n_hidden * g.hmax + md_pad) * P;
    const size_t need = sizeof(float) * (size_t)per_cta * resident;
    if (h->opt_grad_stash && g.inv_mode == NAZB_INV_INCREMENTAL) {
      if (h->grad_stash_bytes < need) {
        if (h->grad_stash) { if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e; cudaFree(h->grad_stash); h->grad_stash = nullptr; h->grad_stash_bytes = 0; }
        if ((e = cudaMalloc(&h->grad_stash, need)) != cudaSuccess) return e;
        h->grad_stash_bytes = need;
      }
      ga.stash = static_cast<float*>(h->grad_stash);
      ga.stash_cta = per_cta;
    }
  }
#define NAZB_GRAD_LAUNCH_K(PP, NS, TT, SP)                                                                                 \
  e = cudaFuncSetAttribute(flow_grad_kernel<PP, NS, TT, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
  if (e != cudaSuccess) return e;                                                                                          \
  flow_grad_kernel<PP, NS, TT, SP><<<grid, kThreads, smem, st>>>(g, gg, h->packed, h->packed_T, h->perm_dev, io, ga);
#define NAZB_GRAD_LAUNCH(PP, NS, TT)                                                                                       \
  if (g.kind == NAZB_KIND_AFFINE) { NAZB_GRAD_LAUNCH_K(PP, NS, TT, false) } else { NAZB_GRAD_LAUNCH_K(PP, NS, TT, true) }
  if (P == 16) { NAZB_GRAD_LAUNCH(16, 2, 3) } else if (nst == 4) { NAZB_GRAD_LAUNCH(32, 4, 5) } else { NAZB_GRAD_LAUNCH(32, 2, 5) }
#undef NAZB_GRAD_LAUNCH
#undef NAZB_GRAD_LAUNCH_K
  nazb_count_launch();
  return cudaGetLastError();
}

bool nazb_grad_fits(const FlowGeom& g) {
  return grad_smem_bytes(g, 16, 2, 3) <= 227 * 1024 || grad_smem_bytes(g, 32, 2, 5) <= 227 * 1024;
}
