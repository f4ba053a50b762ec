// Elementwise transform math shared by the SIMT and tcgen05 engines.
// Arithmetic follows pyro's `_monotonic_rational_spline` / AffineAutoregressive as restated in
// oracle/flow_oracle.py; operation order mirrors the oracle so the
// fp32 results track the reference's own fp32 path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace nazb {

#define NAZB_LOG_2PI 1.8378770664093453f

__device__ __forceinline__ float softplus_f(float a) {
  // torch.nn.functional.softplus (beta = 1, threshold = 20)
  return a > 20.f ? a : log1pf(expf(a));
}
__device__ __forceinline__ float sigmoid_f(float a) { return 1.f / (1.f + expf(-a)); }

// Monotone rational spline on [-B, B] for ONE (point, dim).
//   raw(m): conditioner output slot m for this dim; slots [0,K) widths, [K,2K) heights,
//           [2K,3K-1) derivatives, [3K-1,4K-1) lambdas (linear order only).
//   setw(m, v): scratch write-back (the same storage as raw) used to keep exp() values.
// Returns the transformed value and the FORWARD log|dy/dx| at the solution (for the inverse this is
// what pyro caches: ConditionedSpline._inverse stores the negated inverse log-det).
template <bool LINEAR, class Raw, class SetW>
__device__ __forceinline__ void rational_spline(float in, int K, float B, bool inverse, Raw raw, SetW setw,
                                                float& out, float& ld_fwd) {
  const float min_w = 1e-3f, min_h = 1e-3f, min_d = 1e-3f, min_lam = 0.025f, eps = 1e-6f;
  if (!(in >= -B && in <= B)) {   // identity outside the box (NaN also lands here, like the oracle's where())
    out = in;
    ld_fwd = 0.f;
    return;
  }
  // softmax of widths and heights: exp(v - max) kept in scratch, normalised on the fly
  float mw = -INFINITY, mh = -INFINITY;
  for (int j = 0; j < K; ++j) {
    mw = fmaxf(mw, raw(j));
    mh = fmaxf(mh, raw(K + j));
  }
  float sw = 0.f, sh = 0.f;
  for (int j = 0; j < K; ++j) {
    float ew = expf(raw(j) - mw), eh = expf(raw(K + j) - mh);
    setw(j, ew);
    setw(K + j, eh);
    sw += ew;
    sh += eh;
  }
  const float scale_w = 1.f - min_w * K, scale_h = 1.f - min_h * K;
  // walk the knots: cumsum -> [-B,B] -> forced end points; select the bin on the fly
  float cw = 0.f, ch = 0.f;                 // running cumsum of (min + scale * softmax)
  float kx0 = -B, ky0 = -B;                 // left knot of the current bin
  float sel_w = 0.f, sel_h = 0.f, sel_x = -B, sel_y = -B;
  int sel = 0;
  for (int j = 0; j < K; ++j) {
    float wj = min_w + scale_w * (raw(j) / sw);
    float hj = min_h + scale_h * (raw(K + j) / sh);
    cw += wj;
    ch += hj;
    float kx1 = (j == K - 1) ? B : (2.f * B) * cw + (-B);
    float ky1 = (j == K - 1) ? B : (2.f * B) * ch + (-B);
    float ks = inverse ? ky0 : kx0;
    if (j == 0 || in >= ks + eps) {
      sel = j;
      sel_w = kx1 - kx0;
      sel_h = ky1 - ky0;
      sel_x = kx0;
      sel_y = ky0;
    }
    kx0 = kx1;
    ky0 = ky1;
  }
  const float d_edge = 1.f - min_d;
  float d0 = (sel == 0) ? d_edge : min_d + softplus_f(raw(2 * K + sel - 1));
  float d1 = (sel == K - 1) ? d_edge : min_d + softplus_f(raw(2 * K + sel));
  float delta = sel_h / sel_w;
  if (!LINEAR) {
    if (inverse) {
      float dy = in - sel_y;
      float t2 = d0 + d1 - 2.f * delta;
      float a = dy * t2 + sel_h * (delta - d0);
      float b = sel_h * d0 - dy * t2;
      float c = -delta * dy;
      float disc = b * b - 4.f * a * c;
      float root = (2.f * c) / (-b - sqrtf(disc));
      out = root * sel_w + sel_x;
      float tomt = root * (1.f - root);
      float den = delta + t2 * tomt;
      float omr = 1.f - root;
      float dnum = delta * delta * (d1 * root * root + 2.f * delta * tomt + d0 * omr * omr);
      ld_fwd = logf(dnum) - 2.f * logf(den);
    } else {
      float theta = (in - sel_x) / sel_w;
      float tomt = theta * (1.f - theta);
      float t2 = d0 + d1 - 2.f * delta;
      float num = sel_h * (delta * theta * theta + d0 * tomt);
      float den = delta + t2 * tomt;
      out = sel_y + num / den;
      float omt = 1.f - theta;
      float dnum = delta * delta * (d1 * theta * theta + 2.f * delta * tomt + d0 * omt * omt);
      ld_fwd = logf(dnum) - 2.f * logf(den);
    }
  } else {
    float lam = (1.f - 2.f * min_lam) * sigmoid_f(raw(3 * K - 1 + sel)) + min_lam;
    float wa = 1.f;
    float wb = sqrtf(d0 / d1) * wa;
    float wc = (lam * wa * d0 + (1.f - lam) * wb * d1) / delta;
    float ya = sel_y, yb = sel_h + sel_y;
    float yc = ((1.f - lam) * wa * ya + lam * wb * yb) / ((1.f - lam) * wa + lam * wb);
    if (inverse) {
      bool le = in <= yc;
      float numerator = le ? (lam * wa * (ya - in)) : ((wc - lam * wb) * in + lam * wb * yb - wc * yc);
      float denominator = le ? ((wc - wa) * in + wa * ya - wc * yc) : ((wc - wb) * in + wb * yb - wc * yc);
      float theta = numerator / denominator;
      out = theta * sel_w + sel_x;
      float dnum = (le ? wa * wc * lam * (yc - ya) : wb * wc * (1.f - lam) * (yb - yc)) * sel_w;
      // inverse log-det; the forward one is its negation
      ld_fwd = -(logf(dnum) - 2.f * logf(fabsf(denominator)));
    } else {
      float theta = (in - sel_x) / sel_w;
      bool le = theta <= lam;
      float numerator = le ? (wa * ya * (lam - theta) + wc * yc * theta)
                           : (wc * yc * (1.f - theta) + wb * yb * (theta - lam));
      float denominator = le ? (wa * (lam - theta) + wc * theta) : (wc * (1.f - theta) + wb * (theta - lam));
      out = numerator / denominator;
      float dnum = (le ? wa * wc * lam * (yc - ya) : wb * wc * (1.f - lam) * (yb - yc)) / sel_w;
      ld_fwd = logf(dnum) - 2.f * logf(fabsf(denominator));
    }
  }
}

// Register-resident rational-QUADRATIC spline for a compile-time bin count K, built on the MUFU approximations
// (ex2 / lg2 / rcp / sqrt .approx, each ~1-2 ulp) with every division turned into a reciprocal multiply.
// `r` holds the raw conditioner outputs of this (point, dim): w[0..K), h[K..2K), d[2K..3K-1).  Same
// operation order as rational_spline<false> above; used by the tcgen05 engine's transform epilogues.
template <int K>
__device__ __forceinline__ void rqs_fast(float in, float B, bool inverse, const float* r, float& out, float& ld_fwd) {
  const float min_w = 1e-3f, min_h = 1e-3f, min_d = 1e-3f, eps = 1e-6f;
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  auto ex2 = [](float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto lg2 = [](float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto rcp = [](float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto sqr = [](float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  if (!(in >= -B && in <= B)) { out = in; ld_fwd = 0.f; return; }
  float mw = r[0], mh = r[K];
#pragma unroll
  for (int j = 1; j < K; ++j) { mw = fmaxf(mw, r[j]); mh = fmaxf(mh, r[K + j]); }
  float ew[K], eh[K], sw = 0.f, sh = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    ew[j] = ex2((r[j] - mw) * LOG2E);
    eh[j] = ex2((r[K + j] - mh) * LOG2E);
    sw += ew[j];
    sh += eh[j];
  }
  const float iw = (1.f - min_w * K) * rcp(sw), ih = (1.f - min_h * K) * rcp(sh);
  float cw = 0.f, ch = 0.f, kx0 = -B, ky0 = -B;
  float sel_w = 0.f, sel_h = 0.f, sel_x = -B, sel_y = -B, dl_raw = 0.f, dr_raw = 0.f;
  bool first = true, lastb = false;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    cw += fmaf(ew[j], iw, min_w);
    ch += fmaf(eh[j], ih, min_h);
    float kx1 = (j == K - 1) ? B : fmaf(2.f * B, cw, -B);
    float ky1 = (j == K - 1) ? B : fmaf(2.f * B, ch, -B);
    float ks = inverse ? ky0 : kx0;
    if (j == 0 || in >= ks + eps) {
      sel_w = kx1 - kx0; sel_h = ky1 - ky0; sel_x = kx0; sel_y = ky0;
      first = (j == 0); lastb = (j == K - 1);
      dl_raw = (j == 0) ? 0.f : r[2 * K + j - 1];
      dr_raw = (j == K - 1) ? 0.f : r[2 * K + j];
    }
    kx0 = kx1; ky0 = ky1;
  }
  auto softplus = [&](float a) { return a > 20.f ? a : lg2(1.f + ex2(a * LOG2E)) * LN2; };
  const float d_edge = 1.f - min_d;
  const float d0 = first ? d_edge : min_d + softplus(dl_raw);
  const float d1 = lastb ? d_edge : min_d + softplus(dr_raw);
  const float delta = sel_h * rcp(sel_w);
  const float t2 = d0 + d1 - 2.f * delta;
  float th;
  if (inverse) {
    float dy = in - sel_y;
    float a = fmaf(dy, t2, sel_h * (delta - d0));
    float b = fmaf(-dy, t2, sel_h * d0);
    float c = -delta * dy;
    float disc = fmaf(b, b, -4.f * a * c);
    th = (2.f * c) * rcp(-b - sqr(fmaxf(disc, 0.f)));
    out = fmaf(th, sel_w, sel_x);
  } else {
    th = (in - sel_x) * rcp(sel_w);
  }
  const float tomt = th * (1.f - th), omt = 1.f - th;
  const float den = fmaf(t2, tomt, delta);
  if (!inverse) out = fmaf(sel_h * fmaf(delta, th * th, d0 * tomt), rcp(den), sel_y);
  const float dnum = delta * delta * fmaf(d1, th * th, fmaf(2.f * delta, tomt, d0 * omt * omt));
  ld_fwd = (lg2(dnum) - 2.f * lg2(den)) * LN2;
}

// Inverse rational-quadratic spline (K = 8) for one (point, dim) computed by a PAIR of lanes (lane, lane ^ 16)
// running the same instruction stream: lane half `hw` = 0 holds the 8 raw widths in `own`, half 1 the 8 raw
// heights; both hold the 7 raw derivatives in `dr`.  Each lane builds the knots of its own axis (softmax ->
// min-width -> cumsum -> [-B, B]); the bin index comes from the heights lane (inverse direction searches the
// y knots), each lane selects its own knot pair and one of the two derivatives, three shuffles exchange them
// and both lanes finish the solve redundantly.  Arithmetic and operation order are those of rqs_fast<8>.
__device__ __forceinline__ void rqs8_inv_pair(float in, float B, const float* own, const float* dr, int hw, float& out,
                                              float& ld_fwd) {
  constexpr int K = 8;
  const float min_bin = 1e-3f, min_d = 1e-3f, eps = 1e-6f;
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  auto ex2 = [](float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto lg2 = [](float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto rcp = [](float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto sqr = [](float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  float m = own[0];
#pragma unroll
  for (int j = 1; j < K; ++j) m = fmaxf(m, own[j]);
  float e[K], sum = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) { e[j] = ex2((own[j] - m) * LOG2E); sum += e[j]; }
  const float inv = (1.f - min_bin * K) * rcp(sum);
  float kn[K + 1], cum = 0.f;
  kn[0] = -B;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    cum += fmaf(e[j], inv, min_bin);
    kn[j + 1] = (j == K - 1) ? B : fmaf(2.f * B, cum, -B);
  }
  // bin = #{j in 1..K-1 : in >= knot_j + eps}, decided on the heights lane
  int cnt = 0;
#pragma unroll
  for (int j = 1; j < K; ++j) cnt += (in >= kn[j] + eps) ? 1 : 0;
  const int k = __shfl_sync(0xffffffffu, cnt, (threadIdx.x & 15) | 16);
  float lo = kn[0], hi = kn[1];
#pragma unroll
  for (int j = 1; j < K; ++j)
    if (k == j) { lo = kn[j]; hi = kn[j + 1]; }
  // derivative at the left (lane half 0) / right (lane half 1) knot of the bin; the end knots are fixed
  const int di = k - 1 + hw;
  float draw = dr[0];
#pragma unroll
  for (int j = 1; j < K - 1; ++j)
    if (di == j) draw = dr[j];
  const float sp = draw > 20.f ? draw : lg2(1.f + ex2(draw * LOG2E)) * LN2;
  const float dv = (di < 0 || di > K - 2) ? (1.f - min_d) : (min_d + sp);
  const float width = hi - lo;
  const float p_lo = __shfl_xor_sync(0xffffffffu, lo, 16), p_w = __shfl_xor_sync(0xffffffffu, width, 16);
  const float p_d = __shfl_xor_sync(0xffffffffu, dv, 16);
  const float sel_x = hw ? p_lo : lo, sel_w = hw ? p_w : width, d0 = hw ? p_d : dv;
  const float sel_y = hw ? lo : p_lo, sel_h = hw ? width : p_w, d1 = hw ? dv : p_d;
  const float delta = sel_h * rcp(sel_w);
  const float t2 = d0 + d1 - 2.f * delta;
  const float dy = in - sel_y;
  const float a = fmaf(dy, t2, sel_h * (delta - d0));
  const float b = fmaf(-dy, t2, sel_h * d0);
  const float c = -delta * dy;
  const float disc = fmaf(b, b, -4.f * a * c);
  const float th = (2.f * c) * rcp(-b - sqr(fmaxf(disc, 0.f)));
  const float tomt = th * (1.f - th), omt = 1.f - th;
  const float den = fmaf(t2, tomt, delta);
  const float dnum = delta * delta * fmaf(d1, th * th, fmaf(2.f * delta, tomt, d0 * omt * omt));
  const bool inside = (in >= -B && in <= B);
  out = inside ? fmaf(th, sel_w, sel_x) : in;
  ld_fwd = inside ? (lg2(dnum) - 2.f * lg2(den)) * LN2 : 0.f;
}

// Same, but returns the two factors of the forward derivative instead of its logarithm: ld_fwd = log(dnum) - 2 log(den)
// (dnum = den = 1 outside the box), so the caller can take the logarithms off the critical path.
__device__ __forceinline__ void rqs8_inv_pair_nolog(float in, float B, const float* own, const float* dr, int hw, float& out,
                                                    float& dnum_out, float& den_out) {
  constexpr int K = 8;
  const float min_bin = 1e-3f, min_d = 1e-3f, eps = 1e-6f;
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  (void)LN2;
  auto ex2 = [](float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto lg2 = [](float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto rcp = [](float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  auto sqr = [](float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  float m = own[0];
#pragma unroll
  for (int j = 1; j < K; ++j) m = fmaxf(m, own[j]);
  float e[K], sum = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) { e[j] = ex2((own[j] - m) * LOG2E); sum += e[j]; }
  const float inv = (1.f - min_bin * K) * rcp(sum);
  float kn[K + 1], cum = 0.f;
  kn[0] = -B;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    cum += fmaf(e[j], inv, min_bin);
    kn[j + 1] = (j == K - 1) ? B : fmaf(2.f * B, cum, -B);
  }
  // bin = #{j in 1..K-1 : in >= knot_j + eps}, decided on the heights lane
  int cnt = 0;
#pragma unroll
  for (int j = 1; j < K; ++j) cnt += (in >= kn[j] + eps) ? 1 : 0;
  const int k = __shfl_sync(0xffffffffu, cnt, (threadIdx.x & 15) | 16);
  float lo = kn[0], hi = kn[1];
#pragma unroll
  for (int j = 1; j < K; ++j)
    if (k == j) { lo = kn[j]; hi = kn[j + 1]; }
  // derivative at the left (lane half 0) / right (lane half 1) knot of the bin; the end knots are fixed
  const int di = k - 1 + hw;
  float draw = dr[0];
#pragma unroll
  for (int j = 1; j < K - 1; ++j)
    if (di == j) draw = dr[j];
  const float sp = draw > 20.f ? draw : lg2(1.f + ex2(draw * LOG2E)) * LN2;
  const float dv = (di < 0 || di > K - 2) ? (1.f - min_d) : (min_d + sp);
  const float width = hi - lo;
  const float p_lo = __shfl_xor_sync(0xffffffffu, lo, 16), p_w = __shfl_xor_sync(0xffffffffu, width, 16);
  const float p_d = __shfl_xor_sync(0xffffffffu, dv, 16);
  const float sel_x = hw ? p_lo : lo, sel_w = hw ? p_w : width, d0 = hw ? p_d : dv;
  const float sel_y = hw ? lo : p_lo, sel_h = hw ? width : p_w, d1 = hw ? dv : p_d;
  const float delta = sel_h * rcp(sel_w);
  const float t2 = d0 + d1 - 2.f * delta;
  const float dy = in - sel_y;
  const float a = fmaf(dy, t2, sel_h * (delta - d0));
  const float b = fmaf(-dy, t2, sel_h * d0);
  const float c = -delta * dy;
  const float disc = fmaf(b, b, -4.f * a * c);
  const float th = (2.f * c) * rcp(-b - sqr(fmaxf(disc, 0.f)));
  const float tomt = th * (1.f - th), omt = 1.f - th;
  const float den = fmaf(t2, tomt, delta);
  const float dnum = delta * delta * fmaf(d1, th * th, fmaf(2.f * delta, tomt, d0 * omt * omt));
  const bool inside = (in >= -B && in <= B);
  out = inside ? fmaf(th, sel_w, sel_x) : in;
  dnum_out = inside ? dnum : 1.f;
  den_out = inside ? den : 1.f;
}


// Logit bounding transform of one coordinate (transforms.py:20-23): returns y, adds to log_jac.
__device__ __forceinline__ float bound_fwd(float x, float lo, float hi, float& log_jac) {
  float u = (x - lo) / (hi - lo);
  float lu = logf(u), l1u = log1pf(-u);
  log_jac -= (lu + l1u) + logf(hi - lo);
  return lu - l1u;
}
__device__ __forceinline__ float bound_inv(float y, float lo, float hi) {
  return sigmoid_f(y) * (hi - lo) + lo;
}

}  // namespace nazb
