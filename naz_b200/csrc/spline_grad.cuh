// First derivatives of the monotone rational spline (pyro `_monotonic_rational_spline`, order = "quadratic" and "linear", as
// restated in oracle/pyro_style.py `monotonic_rational_spline`; naz call site src/naz/flows/transforms.py:180-190) with respect
// to its input and to the RAW conditioner outputs — what the gradient of the summed log-likelihood needs for neural-spline
// autoregressive flows (the reference gets them from torch autograd, train_flows.py:195-213).
//
// The spline y = T(x; p) and its log-derivative ld(x; p) depend on the raw outputs only through the six numbers of the
// selected bin (left knot X0, width W, bottom knot Y0, height H, end-point derivatives d0, d1).  The value pass runs on
// forward-mode duals carrying the seven partials (X0, W, Y0, H, d0, d1, x; plus the bin's lambda for the linear order); two small reverse steps then take a
// cotangent on the bin numbers to the raw slots: through the knot sums and the softmax for widths / heights, through the
// softplus for the derivatives.  The bin index is piecewise constant and carries no derivative.
//
// __host__ __device__: the same code is compiled for the CPU behind `nazb_host_spline_grad` (include/nazb.h) so that the
// CPU test-suite checks it against autograd of the oracle without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace nazb {

template <int N>
struct Dual {
  float v;
  float d[N];
};
template <int N>
__host__ __device__ __forceinline__ Dual<N> dconst(float c) {
  Dual<N> r; r.v = c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.f;
  return r;
}
template <int N>
__host__ __device__ __forceinline__ Dual<N> dvar(float c, int i) { Dual<N> r = dconst<N>(c); r.d[i] = 1.f; return r; }
template <int N>
__host__ __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N>
__host__ __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N>
__host__ __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = fmaf(a.v, b.d[i], a.d[i] * b.v);
  return r;
}
template <int N>
__host__ __device__ __forceinline__ Dual<N> operator*(float a, const Dual<N>& b) {
  Dual<N> r; r.v = a * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i];
  return r;
}
template <int N>
__host__ __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  const float ib = 1.f / b.v;
  r.v = a.v * ib;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <int N>
__host__ __device__ __forceinline__ Dual<N> dlog(const Dual<N>& a) {   // log |a|
  Dual<N> r; r.v = logf(fabsf(a.v));
  const float ia = 1.f / a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * ia;
  return r;
}
template <int N>
__host__ __device__ __forceinline__ Dual<N> dsqrt(const Dual<N>& a) {
  Dual<N> r; r.v = sqrtf(a.v);
  const float h = 0.5f / r.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * h;
  return r;
}

// One (point, dimension).  x: the spline INPUT (sampling-direction argument; in the log_prob direction the solved value),
// raw(m): the raw conditioner outputs — widths [0,K), heights [K,2K), derivatives [2K,3K-1) and, for the linear-rational order
// (LINEAR: pyro's order = "linear", Dolatabadi et al.), lambdas [3K-1,4K-1), an eighth partial of the dual pass.
// Outputs:  inv_tx = 1 / (dT/dx),  ldx = d ld / d x,  and through put(m, ca, cb) for every raw slot m
//           ca_m = -(dT/draw_m) / (dT/dx)   ( = d x / d raw_m at fixed y, implicit-function theorem ),
//           cb_m = -(d ld / d raw_m),
// so that the cotangent of the conditioner output in the adjoint recursion is  c_m = lambda * ca_m + cb_m.
// Outside [-B, B] the transform is the identity: inv_tx = 1, everything else 0.
template <bool LINEAR, class Raw, class Put>
__host__ __device__ __forceinline__ void spline_grad(float x, int K, float B, Raw raw, Put put, float& inv_tx, float& ldx) {
  const float min_w = 1e-3f, min_h = 1e-3f, min_d = 1e-3f, min_lam = 0.025f, eps = 1e-6f;
  const int M = (LINEAR ? 4 : 3) * K - 1;
  constexpr int NV = LINEAR ? 8 : 7, IX = NV - 1;   // partials: X0, W, Y0, H, d0, d1, [lambda,] x
  using Du = Dual<NV>;
  if (!(x >= -B && x <= B)) {
    inv_tx = 1.f; ldx = 0.f;
    for (int m = 0; m < M; ++m) put(m, 0.f, 0.f);
    return;
  }
  // ---- value pass of the knots (same operation order as rational_spline<false> in transforms.cuh) ----
  float mw = -INFINITY, mh = -INFINITY;
  for (int j = 0; j < K; ++j) { mw = fmaxf(mw, raw(j)); mh = fmaxf(mh, raw(K + j)); }
  float sw = 0.f, sh = 0.f;
  for (int j = 0; j < K; ++j) { sw += expf(raw(j) - mw); sh += expf(raw(K + j) - mh); }
  const float scale_w = 1.f - min_w * K, scale_h = 1.f - min_h * K;
  float cw = 0.f, ch = 0.f, kx0 = -B, ky0 = -B;
  float sel_w = 0.f, sel_h = 0.f, sel_x = -B, sel_y = -B;
  int sel = 0;
  for (int j = 0; j < K; ++j) {
    cw += min_w + scale_w * (expf(raw(j) - mw) / sw);
    ch += min_h + scale_h * (expf(raw(K + j) - mh) / sh);
    const float kx1 = (j == K - 1) ? B : (2.f * B) * cw + (-B);
    const float ky1 = (j == K - 1) ? B : (2.f * B) * ch + (-B);
    if (j == 0 || x >= kx0 + eps) { sel = j; sel_w = kx1 - kx0; sel_h = ky1 - ky0; sel_x = kx0; sel_y = ky0; }
    kx0 = kx1; ky0 = ky1;
  }
  const float rl = (sel == 0) ? 0.f : raw(2 * K + sel - 1), rr = (sel == K - 1) ? 0.f : raw(2 * K + sel);
  auto softplus = [](float a) { return a > 20.f ? a : log1pf(expf(a)); };
  auto sigmoid = [](float a) { return a > 20.f ? 1.f : 1.f / (1.f + expf(-a)); };
  const float d0v = (sel == 0) ? 1.f - min_d : min_d + softplus(rl);
  const float d1v = (sel == K - 1) ? 1.f - min_d : min_d + softplus(rr);
  // ---- dual pass: y = T(x), ld = log T'(x) as functions of (X0, W, Y0, H, d0, d1, [lambda,] x) ----
  const Du X0 = dvar<NV>(sel_x, 0), W = dvar<NV>(sel_w, 1), Y0 = dvar<NV>(sel_y, 2), H = dvar<NV>(sel_h, 3), d0 = dvar<NV>(d0v, 4),
           d1 = dvar<NV>(d1v, 5), xd = dvar<NV>(x, IX);
  const Du one = dconst<NV>(1.f);
  const Du th = (xd - X0) / W;
  const Du delta = H / W;
  Du y, ld;
  float rlam = 0.f, slam = 0.f;
  if (!LINEAR) {
    const Du omt = one - th;
    const Du tomt = th * omt;
    const Du t2 = d0 + d1 - 2.f * delta;
    const Du den = delta + t2 * tomt;
    y = Y0 + H * (delta * th * th + d0 * tomt) / den;
    ld = dlog(delta * delta * (d1 * th * th + 2.f * (delta * tomt) + d0 * omt * omt)) - 2.f * dlog(den);
  } else {
    rlam = raw(3 * K - 1 + sel);
    slam = 1.f / (1.f + expf(-rlam));
    const Du lam = dvar<NV>((1.f - 2.f * min_lam) * slam + min_lam, 6);
    const Du oml = one - lam;
    const Du wb = dsqrt(d0 / d1);                       // wa = 1
    const Du wc = (lam * d0 + oml * wb * d1) / delta;
    const Du ya = Y0, yb = H + Y0;
    const Du yc = (oml * ya + lam * wb * yb) / (oml + lam * wb);
    const bool le = th.v <= lam.v;
    const Du num = le ? (ya * (lam - th) + wc * yc * th) : (wc * yc * (one - th) + wb * yb * (th - lam));
    const Du den = le ? ((lam - th) + wc * th) : (wc * (one - th) + wb * (th - lam));
    y = num / den;
    const Du dnum = (le ? (wc * lam * (yc - ya)) : (wb * wc * oml * (yb - yc))) / W;
    ld = dlog(dnum) - 2.f * dlog(den);
  }
  const float itx = 1.f / y.d[IX];
  inv_tx = itx;
  ldx = ld.d[IX];
  // ---- cotangents on the bin numbers -> raw slots ----
  float ga[7], gb[7];
#pragma unroll
  for (int i = 0; i < IX; ++i) { ga[i] = -y.d[i] * itx; gb[i] = -ld.d[i]; }
  // widths: X0 = -B + 2B sum_{j<sel} w_j, W = 2B w_sel, w = min_w + scale_w softmax(raw): g_raw_k = scale_w s_k (g_k - sum_j g_j s_j)
  // (the forced end knot makes W of the last bin B - X0 instead; as sum_j w_j == 1 both forms have the same raw derivatives)
  {
    float dota = 0.f, dotb = 0.f;
    for (int j = 0; j <= sel; ++j) {
      const float s = expf(raw(j) - mw) / sw;
      dota += ((j < sel) ? ga[0] : ga[1]) * s;
      dotb += ((j < sel) ? gb[0] : gb[1]) * s;
    }
    for (int k = 0; k < K; ++k) {
      const float s = expf(raw(k) - mw) / sw;
      const float a = (k < sel) ? ga[0] : (k == sel ? ga[1] : 0.f), b = (k < sel) ? gb[0] : (k == sel ? gb[1] : 0.f);
      put(k, (2.f * B) * scale_w * s * (a - dota), (2.f * B) * scale_w * s * (b - dotb));
    }
  }
  {
    float dota = 0.f, dotb = 0.f;
    for (int j = 0; j <= sel; ++j) {
      const float s = expf(raw(K + j) - mh) / sh;
      dota += ((j < sel) ? ga[2] : ga[3]) * s;
      dotb += ((j < sel) ? gb[2] : gb[3]) * s;
    }
    for (int k = 0; k < K; ++k) {
      const float s = expf(raw(K + k) - mh) / sh;
      const float a = (k < sel) ? ga[2] : (k == sel ? ga[3] : 0.f), b = (k < sel) ? gb[2] : (k == sel ? gb[3] : 0.f);
      put(K + k, (2.f * B) * scale_h * s * (a - dota), (2.f * B) * scale_h * s * (b - dotb));
    }
  }
  const float sl = sigmoid(rl), sr = sigmoid(rr);
  for (int k = 0; k < K - 1; ++k) {
    float a = 0.f, b = 0.f;
    if (k == sel - 1) { a = ga[4] * sl; b = gb[4] * sl; }
    if (k == sel) { a = ga[5] * sr; b = gb[5] * sr; }
    put(2 * K + k, a, b);
  }
  if (LINEAR) {
    const float dl = (1.f - 2.f * min_lam) * slam * (1.f - slam);
    for (int k = 0; k < K; ++k) put(3 * K - 1 + k, (k == sel) ? ga[6] * dl : 0.f, (k == sel) ? gb[6] * dl : 0.f);
  }
}

template <class Raw, class Put>
__host__ __device__ __forceinline__ void rqs_grad(float x, int K, float B, Raw raw, Put put, float& inv_tx, float& ldx) {
  spline_grad<false>(x, K, B, raw, put, inv_tx, ldx);
}

}  // namespace nazb
