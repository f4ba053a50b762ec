// C-ABI glue of libnazb.so (see include/nazb.h for the contract and the reference mapping).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "nazb_internal.h"
#include "spline_grad.cuh"

static std::atomic<long long> g_launches{0};
void nazb_count_launch(int n) { g_launches += n; }
extern "C" int64_t nazb_launch_count(void) { return g_launches.load(); }

#define CK(h, call)                                    \
  do {                                                 \
    cudaError_t e_ = (call);                           \
    if (e_ != cudaSuccess) {                           \
      if (h) (h)->cuda_err = cudaGetErrorString(e_);   \
      return NAZB_ERR_CUDA;                            \
    }                                                  \
  } while (0)

extern "C" const char* nazb_strerror(int status) {
  switch (status) {
    case NAZB_OK: return "ok";
    case NAZB_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or inconsistent shape)";
    case NAZB_ERR_UNSUPPORTED: return "configuration not supported by the selected engine";
    case NAZB_ERR_CUDA: return "CUDA runtime error (see nazb_last_cuda_error)";
    case NAZB_ERR_NOT_PACKED: return "nazb_pack has not been called on this handle";
    case NAZB_ERR_NO_DEVICE: return "no usable CUDA device (libnazb needs an sm_100 GPU)";
    default: return "unknown nazb status";
  }
}

extern "C" const char* nazb_last_cuda_error(const nazb_handle* h) { return h ? h->cuda_err.c_str() : ""; }

static int round4(int v) { return (v + 3) & ~3; }

static int build_geom(const nazb_desc& d, FlowGeom& g) {
  memset(&g, 0, sizeof(g));
  if (d.kind < NAZB_KIND_AFFINE || d.kind > NAZB_KIND_RLS) return NAZB_ERR_BAD_ARG;
  if (d.D < 1 || d.D > NAZB_MAX_DIM || d.C < 0 || d.L < 1 || d.n_hidden < 1 || d.n_hidden > NAZB_MAX_HIDDEN_LAYERS ||
      d.S < 1)
    return NAZB_ERR_BAD_ARG;
  g.kind = d.kind; g.D = d.D; g.C = d.C; g.L = d.L; g.n_hidden = d.n_hidden;
  g.K = d.count_bins;
  if (d.kind == NAZB_KIND_AFFINE) g.M = 2;
  else {
    if (d.count_bins < 2 || d.count_bins > 64 || !(d.bound > 0.f)) return NAZB_ERR_BAD_ARG;
    g.M = (d.kind == NAZB_KIND_RQS) ? 3 * d.count_bins - 1 : 4 * d.count_bins - 1;
  }
  g.bound = d.bound; g.clip_lo = d.clip_lo; g.clip_hi = d.clip_hi;
  g.kin = d.C + d.D;
  g.md = g.M * d.D;
  g.hmax = 0;
  for (int j = 0; j < d.n_hidden; ++j) {
    if (d.hidden[j] < d.D) return NAZB_ERR_BAD_ARG;   // pyro raises for hidden < input_dim
    g.hidden[j] = d.hidden[j];
    g.hmax = g.hmax > round4(d.hidden[j]) ? g.hmax : round4(d.hidden[j]);
  }
  long long off = 0;
  const int n_lin = d.n_hidden + 1;
  for (int j = 0; j < n_lin; ++j) {
    g.kdim[j] = (j == 0) ? g.kin : d.hidden[j - 1];
    g.ndim[j] = (j == n_lin - 1) ? g.md : d.hidden[j];
    g.ldw[j] = round4(g.ndim[j]);
    g.off_w[j] = off;
    off += (long long)g.kdim[j] * g.ldw[j];
    g.off_b[j] = off;
    off += g.ldw[j];
  }
  g.layer_stride = off;            // multiple of 4 floats by construction
  g.draw_stride = off * d.L;
  g.inv_mode = d.inverse_mode;
  return NAZB_OK;
}

extern "C" int nazb_create(nazb_handle** out, const nazb_desc* desc) {
  if (!out || !desc) return NAZB_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return NAZB_ERR_NO_DEVICE;
  if (desc->device < 0 || desc->device >= ndev) return NAZB_ERR_BAD_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, desc->device) != cudaSuccess) return NAZB_ERR_NO_DEVICE;
  if (prop.major != 10) return NAZB_ERR_NO_DEVICE;   // sm_100a cubins only; no fallback
  nazb_handle* h = new (std::nothrow) nazb_handle();
  if (!h) return NAZB_ERR_BAD_ARG;
  h->desc = *desc;
  h->device = desc->device;
  h->sm_count = prop.multiProcessorCount;
  int rc = build_geom(*desc, h->geom);
  if (rc != NAZB_OK) { delete h; return rc; }
  if (desc->inverse_mode != NAZB_INV_INCREMENTAL && desc->inverse_mode != NAZB_INV_JACOBI) { delete h; return NAZB_ERR_BAD_ARG; }
  // engine resolution
  std::string why;
  bool tc_ok = nazb_tc_supported(h->geom, &why);
  if (desc->engine == NAZB_ENGINE_TCGEN05 && !tc_ok) { delete h; return NAZB_ERR_UNSUPPORTED; }
  if (desc->engine == NAZB_ENGINE_AUTO) h->engine = tc_ok ? NAZB_ENGINE_TCGEN05 : NAZB_ENGINE_SIMT;
  else if (desc->engine == NAZB_ENGINE_SIMT || desc->engine == NAZB_ENGINE_TCGEN05) h->engine = desc->engine;
  else { delete h; return NAZB_ERR_BAD_ARG; }
  if (h->engine == NAZB_ENGINE_SIMT && nazb_simt_pick_P(h->geom) == 0) { delete h; return NAZB_ERR_UNSUPPORTED; }
  cudaError_t e = cudaSetDevice(h->device);
  if (e == cudaSuccess) e = cudaMalloc(&h->perm_dev, sizeof(int) * 2 * (size_t)desc->L * desc->D);
  if (e == cudaSuccess && h->engine == NAZB_ENGINE_SIMT)
    e = cudaMalloc(&h->packed, sizeof(float) * (size_t)desc->S * h->geom.draw_stride);
  if (e == cudaSuccess && h->engine == NAZB_ENGINE_TCGEN05) e = nazb_tc_create(h);
  if (e != cudaSuccess) {
    fprintf(stderr, "nazb_create: %s\n", cudaGetErrorString(e));
    nazb_destroy(h);
    return NAZB_ERR_CUDA;
  }
  *out = h;
  return NAZB_OK;
}

extern "C" void nazb_destroy(nazb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->tc) nazb_tc_destroy(h);
  if (h->packed) cudaFree(h->packed);
  if (h->packed_T) cudaFree(h->packed_T);
  if (h->grad_tabs) cudaFree(h->grad_tabs);
  if (h->grad_stash) cudaFree(h->grad_stash);
  if (h->perm_dev) cudaFree(h->perm_dev);
  if (h->aff_dev) cudaFree(h->aff_dev);
  if (h->stage_host) cudaFreeHost(h->stage_host);
  if (h->stage_ev) cudaEventDestroy((cudaEvent_t)h->stage_ev);
  delete h;
}

extern "C" int nazb_engine_in_use(const nazb_handle* h) { return h ? h->engine : NAZB_ERR_BAD_ARG; }

bool nazb_tc_direction_ok(const nazb_handle* h, int dir);
int nazb_tc_rows_per_item(const nazb_handle* h, int dir);

// Engine that serves one direction (0 = inverse / log_prob, 1 = forward / sample) after nazb_pack.
extern "C" int nazb_engine_for_direction(const nazb_handle* h, int dir) {
  if (!h || dir < 0 || dir > 1) return NAZB_ERR_BAD_ARG;
  if (h->engine == NAZB_ENGINE_TCGEN05 && !h->aff_dev && nazb_tc_direction_ok(h, dir)) return NAZB_ENGINE_TCGEN05;
  return NAZB_ENGINE_SIMT;
}

int nazb_tc_set_option(nazb_handle* h, const char* name, int value);
int nazb_tc_get_option(const nazb_handle* h, const char* name, int* value);
unsigned int nazb_tc_watchdog(const nazb_handle* h);

// Tuning / A-B switches of the engines (they replace the round-1 NAZB_* environment overrides: nothing in this library
// reads the environment).  Options that change the packed program ("inv_kernel", "inv_merge_n") require a new nazb_pack.
extern "C" int nazb_set_option(nazb_handle* h, const char* name, int32_t value) {
  if (!h || !name) return NAZB_ERR_BAD_ARG;
  if (!strcmp(name, "grad_diag")) { h->opt_grad_diag = value ? 1 : 0; return NAZB_OK; }
  if (!strcmp(name, "grad_stash")) { h->opt_grad_stash = value ? 1 : 0; return NAZB_OK; }
  if (!strcmp(name, "grad_tile")) { if (value != 0 && value != 16 && value != 32) return NAZB_ERR_BAD_ARG; h->opt_grad_tile = value; return NAZB_OK; }
  if (h->engine != NAZB_ENGINE_TCGEN05) return NAZB_ERR_UNSUPPORTED;
  return nazb_tc_set_option(h, name, value);
}
extern "C" int nazb_get_option(const nazb_handle* h, const char* name, int32_t* value) {
  if (!h || !name || !value) return NAZB_ERR_BAD_ARG;
  if (!strcmp(name, "grad_diag")) { *value = h->opt_grad_diag; return NAZB_OK; }
  if (!strcmp(name, "grad_stash")) { *value = h->opt_grad_stash; return NAZB_OK; }
  if (!strcmp(name, "grad_tile")) { *value = h->opt_grad_tile; return NAZB_OK; }
  if (!strcmp(name, "watchdog")) { *value = (h->engine == NAZB_ENGINE_TCGEN05) ? (int32_t)nazb_tc_watchdog(h) : 0; return NAZB_OK; }
  if (h->engine != NAZB_ENGINE_TCGEN05) return NAZB_ERR_UNSUPPORTED;
  int v = 0;
  int rc = nazb_tc_get_option(h, name, &v);
  if (rc == NAZB_OK) *value = v;
  return rc;
}

extern "C" int64_t nazb_packed_bytes(const nazb_handle* h) {
  if (!h) return 0;
  int64_t n = 0;
  if (h->engine == NAZB_ENGINE_TCGEN05) n += nazb_tc_packed_bytes(h);
  if (h->packed) n += (int64_t)h->desc.S * h->geom.draw_stride * (int64_t)sizeof(float);
  return n;
}

// Host -> device upload of a small table without hidden synchronisation: the bytes are copied into a pinned staging
// buffer owned by the handle and sent with cudaMemcpyAsync on the caller's stream.  nazb_stage_begin waits (host side
// only) for the previous pack's copies to have left the staging buffer before it is reused.
static cudaError_t nazb_stage_begin(nazb_handle* h, size_t need) {
  cudaError_t e;
  if (!h->stage_ev) {
    cudaEvent_t ev;
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    h->stage_ev = ev;
  } else if ((e = cudaEventSynchronize((cudaEvent_t)h->stage_ev)) != cudaSuccess) {
    return e;
  }
  if (h->stage_cap < need) {
    if (h->stage_host) cudaFreeHost(h->stage_host);
    h->stage_host = nullptr; h->stage_cap = 0;
    if ((e = cudaHostAlloc(&h->stage_host, need, cudaHostAllocDefault)) != cudaSuccess) return e;
    h->stage_cap = need;
  }
  h->stage_off = 0;
  return cudaSuccess;
}
cudaError_t nazb_stage_upload(nazb_handle* h, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  const size_t off = (h->stage_off + 15) & ~(size_t)15;
  if (off + bytes > h->stage_cap) return cudaErrorMemoryAllocation;
  memcpy(static_cast<char*>(h->stage_host) + off, src, bytes);
  h->stage_off = off + bytes;
  cudaError_t e = cudaMemcpyAsync(dst, static_cast<char*>(h->stage_host) + off, bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaEventRecord((cudaEvent_t)h->stage_ev, st);
  return e;
}

static int pack_impl(nazb_handle* h, const float* const* W, const float* const* b, const int64_t* wst,
                     const int64_t* bst, const float* const* mask, const int64_t* perm, const int32_t* hid_deg,
                     const float* keep, float p_drop, void* stream, const DrawMap& dm) {
  if (!h || !W || !b || !wst || !bst || !mask || !perm) return NAZB_ERR_BAD_ARG;
  if (keep && !(p_drop >= 0.f && p_drop < 1.f)) return NAZB_ERR_BAD_ARG;
  FlowGeom& g = h->geom;
  const int n_lin = g.n_hidden + 1;
  for (int i = 0; i < g.L * n_lin; ++i)
    if (!W[i] || !b[i] || !mask[i]) return NAZB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CK(h, cudaSetDevice(h->device));
  h->is_packed = false;   // stays false when anything below fails (a half-packed handle must not be evaluated)
  CK(h, nazb_stage_begin(h, 64 + sizeof(int) * 2 * (size_t)g.L * g.D + 8 * (size_t)7 * g.L * n_lin + 64 +
                               sizeof(short) * NAZB_MAX_HIDDEN_LAYERS * 256 + 64));   // + the column map of the block-aligned layout
  // perm + rank tables
  std::vector<int> pr(2 * (size_t)g.L * g.D);
  for (int l = 0; l < g.L; ++l) {
    std::vector<char> seen(g.D, 0);
    for (int r = 0; r < g.D; ++r) {
      long long d = perm[(size_t)l * g.D + r];
      if (d < 0 || d >= g.D || seen[d]) return NAZB_ERR_BAD_ARG;
      seen[d] = 1;
      pr[(size_t)l * g.D + r] = (int)d;
      pr[(size_t)g.L * g.D + (size_t)l * g.D + d] = r;
    }
  }
  CK(h, nazb_stage_upload(h, h->perm_dev, pr.data(), pr.size() * sizeof(int), st));
  // inverse schedule from the MADE degrees
  if (g.inv_mode == NAZB_INV_INCREMENTAL) {
    if (!hid_deg) return NAZB_ERR_BAD_ARG;
    int hk = 0;
    for (int j = 0; j < g.n_hidden; ++j) hk = hk > g.hidden[j] ? hk : g.hidden[j];
    for (int j = 0; j < g.n_hidden; ++j) {
      const int32_t* dg = hid_deg + (size_t)j * hk;
      for (int u = 0; u < g.hidden[j]; ++u) {
        if (dg[u] < 0 || dg[u] >= g.D) return NAZB_ERR_BAD_ARG;
        if (u > 0 && dg[u] < dg[u - 1]) return NAZB_ERR_UNSUPPORTED;   // needs degree-sorted units
      }
      for (int r = 0; r <= g.D; ++r) {
        int cnt = 0;
        while (cnt < g.hidden[j] && dg[cnt] < r) ++cnt;
        g.blk[j][r] = (short)cnt;
      }
    }
  }
  cudaError_t e = cudaSuccess;
  bool need_simt = (h->engine == NAZB_ENGINE_SIMT);
  if (h->engine == NAZB_ENGINE_TCGEN05) {
    e = nazb_tc_pack(h, W, b, wst, bst, mask, keep, p_drop, st, dm);
    CK(h, e);
    // a direction the tensor-core programs cannot hold (TMEM budget) is served by the SIMT engine
    need_simt = h->aff_dev != nullptr || !nazb_tc_direction_ok(h, 0) || !nazb_tc_direction_ok(h, 1);
    if (need_simt && h->desc.engine == NAZB_ENGINE_TCGEN05) return NAZB_ERR_UNSUPPORTED;
    if (need_simt && nazb_simt_pick_P(h->geom) == 0) return NAZB_ERR_UNSUPPORTED;
  }
  if (need_simt) {
    if (!h->packed) CK(h, cudaMalloc(&h->packed, sizeof(float) * (size_t)h->desc.S * h->geom.draw_stride));
    e = nazb_pack_simt(h, W, b, wst, bst, mask, keep, p_drop, st, dm);
    CK(h, e);
  }
  h->is_packed = true;
  h->has_keep = (keep != nullptr);
  h->packed_T_valid = false;
  return NAZB_OK;
}

extern "C" int nazb_pack(nazb_handle* h, const float* const* W, const float* const* b, const int64_t* wst,
                         const int64_t* bst, const float* const* mask, const int64_t* perm, const int32_t* hid_deg,
                         const float* keep, float p_drop, void* stream) {
  return pack_impl(h, W, b, wst, bst, mask, perm, hid_deg, keep, p_drop, stream, DrawMap());
}

// Pack S draws given as STANDARD parameters: theta_s = theta_0 * (1 + scale * u_s)  (bflow_jax_maf.py:239-240), applied
// on the fly while packing, so neither the reference's [S, P] `params` array nor a host loop over draws is materialised.
extern "C" int nazb_pack_draw_map(nazb_handle* h, const float* const* W0, const float* const* b0, const float* const* uW,
                                  const float* const* ub, const int64_t* uwst, const int64_t* ubst, float scale,
                                  const float* const* mask, const int64_t* perm, const int32_t* hid_deg,
                                  const float* keep, float p_drop, void* stream) {
  if (!h || !W0 || !b0) return NAZB_ERR_BAD_ARG;
  const int n = h->geom.L * (h->geom.n_hidden + 1);
  for (int i = 0; i < n; ++i)
    if (!W0[i] || !b0[i]) return NAZB_ERR_BAD_ARG;
  DrawMap dm;
  dm.baseW = W0; dm.baseB = b0; dm.scale = scale;
  return pack_impl(h, uW, ub, uwst, ubst, mask, perm, hid_deg, keep, p_drop, stream, dm);
}

namespace {
struct HostRaw { const float* r; __host__ __device__ float operator()(int m) const { return r[m]; } };
struct HostPut { float* a; float* b; __host__ __device__ void operator()(int m, float x, float y) const { a[m] = x; b[m] = y; } };
}  // namespace
extern "C" int nazb_host_spline_grad(float x, int32_t K, float bound, int32_t linear_order, const float* raw, float* ca,
                                     float* cb, float* inv_tx, float* ldx) {
  if (!raw || !ca || !cb || !inv_tx || !ldx || K < 2 || K > 64) return NAZB_ERR_BAD_ARG;
  if (linear_order) nazb::spline_grad<true>(x, K, bound, HostRaw{raw}, HostPut{ca, cb}, *inv_tx, *ldx);
  else nazb::spline_grad<false>(x, K, bound, HostRaw{raw}, HostPut{ca, cb}, *inv_tx, *ldx);
  return NAZB_OK;
}

// BatchNorm (eval mode) as a per-layer element-wise affine; see include/nazb.h.
extern "C" int nazb_set_layer_affine(nazb_handle* h, const float* a, const float* b, void* stream) {
  if (!h) return NAZB_ERR_BAD_ARG;
  CK(h, cudaSetDevice(h->device));
  if (!a) {
    if (h->aff_dev) { CK(h, cudaStreamSynchronize((cudaStream_t)stream)); cudaFree(h->aff_dev); h->aff_dev = nullptr; }
    return NAZB_OK;
  }
  if (!b) return NAZB_ERR_BAD_ARG;
  if (h->desc.engine == NAZB_ENGINE_TCGEN05) return NAZB_ERR_UNSUPPORTED;
  const int D = h->geom.D, L = h->geom.L, row = 2 * D + 1;
  std::vector<float> tab((size_t)L * row);
  for (int l = 0; l < L; ++l) {
    double ls = 0.0;
    for (int d = 0; d < D; ++d) {
      const float av = a[(size_t)l * D + d];
      if (!(av > 0.f) || !std::isfinite(av) || !std::isfinite(b[(size_t)l * D + d])) return NAZB_ERR_BAD_ARG;
      tab[(size_t)l * row + d] = av;
      tab[(size_t)l * row + D + d] = b[(size_t)l * D + d];
      ls += std::log((double)av);
    }
    tab[(size_t)l * row + 2 * D] = (float)ls;
  }
  if (!h->aff_dev) {
    CK(h, cudaMalloc(&h->aff_dev, sizeof(float) * tab.size()));
    if (h->engine == NAZB_ENGINE_TCGEN05) h->is_packed = false;   // the SIMT image is built by the next nazb_pack
  }
  // pageable source: the runtime stages the bytes before returning
  CK(h, cudaMemcpyAsync(h->aff_dev, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return NAZB_OK;
}

static int check_io(const nazb_handle* h, int s_begin, int s_count, const void* x, const float* ctx, int ctx_rows,
                    int N, const float* lo, const float* hi) {
  if (!h || !x) return NAZB_ERR_BAD_ARG;
  if (!h->is_packed) return NAZB_ERR_NOT_PACKED;
  if (N <= 0 || s_count <= 0 || s_begin < 0 || s_begin + s_count > h->desc.S) return NAZB_ERR_BAD_ARG;
  if ((h->geom.C > 0) != (ctx != nullptr)) return NAZB_ERR_BAD_ARG;   // flow.py:75 asserts condition is given
  if (ctx && ctx_rows != 1 && ctx_rows != N) return NAZB_ERR_BAD_ARG;
  if ((lo == nullptr) != (hi == nullptr)) return NAZB_ERR_BAD_ARG;
  return NAZB_OK;
}

// Work items are (point tile, draw group); CTA b takes items b, b + #SMs, ... (group-major order).  `rows` = points per
// item of the kernel that will run (64 SIMT, 128 tcgen05 inverse / one-tile forward, 256 two-tile forward).
static int pick_groups(const nazb_handle* h, int N, int s_count, int requested, int rows) {
  if (requested > 0) return requested < s_count ? requested : s_count;
  const long long sms = h->sm_count;
  const long long tiles = (N + rows - 1) / rows;
  long long gq = (4 * sms + tiles - 1) / tiles;          // >= 4 items per SM if the draw axis allows it
  if (gq < 1) gq = 1;
  // large N: ~8 draws per group, so a group's weight images stay L2-resident while the CTAs sweep the point tiles
  if (tiles >= 8 * sms && gq < s_count / 8) gq = s_count / 8;
  if (gq > s_count) gq = s_count;
  if (gq > 65535) gq = 65535;
  // few items per SM: the tail matters.  Pick the group count (up to 6x the minimum) whose busiest CTA is closest to
  // the mean, counting draws per item exactly (group g holds ceil((s_count - g) / G) draws).
  if (tiles * gq < 16 * sms) {
    const long long g_hi = std::min<long long>(std::min<long long>(s_count, 65535), 6 * gq);
    double best_eff = -1.0;
    long long best_g = gq;
    std::vector<long long> cost((size_t)sms);
    for (long long G = gq; G <= g_hi; ++G) {
      std::fill(cost.begin(), cost.end(), 0LL);
      long long total = 0, item = 0;
      for (long long g = 0; g < G; ++g) {
        const long long nd = (s_count - g + G - 1) / G;
        for (long long t = 0; t < tiles; ++t, ++item) { cost[(size_t)(item % sms)] += nd; total += nd; }
      }
      const long long mx = *std::max_element(cost.begin(), cost.end());
      const double eff = (double)total / ((double)mx * (double)sms);
      if (eff > best_eff + 1e-3) { best_eff = eff; best_g = G; }
    }
    gq = best_g;
  }
  return (int)gq;
}

extern "C" int nazb_inverse(nazb_handle* h, int32_t s_begin, int32_t s_count, const float* x, const float* ctx,
                            int32_t ctx_rows, int32_t N, const float* lo, const float* hi, float* z, float* lp,
                            const float* log_w, float* lse_max, float* lse_sum, int32_t n_groups, double* sum_n,
                            void* stream) {
  int rc = check_io(h, s_begin, s_count, x, ctx, ctx_rows, N, lo, hi);
  if (rc != NAZB_OK) return rc;
  if ((lse_max == nullptr) != (lse_sum == nullptr)) return NAZB_ERR_BAD_ARG;
  if (lse_max && (n_groups < 1 || n_groups > s_count || n_groups > 65535)) return NAZB_ERR_BAD_ARG;
  if (!z && !lp && !lse_max && !sum_n) return NAZB_ERR_BAD_ARG;
  CK(h, cudaSetDevice(h->device));
  IoArgs io{};
  io.x = x; io.x_draw_stride = 0; io.ctx = ctx; io.ctx_rows = ctx_rows; io.N = N;
  io.s_begin = s_begin; io.s_count = s_count; io.lo = lo; io.hi = hi;
  io.out_x = z; io.out_l = lp; io.log_w = log_w; io.lse_max = lse_max; io.lse_sum = lse_sum; io.sum_n = sum_n;
  io.dir = 0; io.aff = h->aff_dev;
  const int rows_inv = (nazb_engine_for_direction(h, 0) == NAZB_ENGINE_TCGEN05) ? 128 : 64;
  int G = pick_groups(h, N, s_count, lse_max ? n_groups : 0, rows_inv);
  cudaError_t e = (nazb_engine_for_direction(h, 0) == NAZB_ENGINE_TCGEN05)
                      ? nazb_tc_launch(h, io, G, (cudaStream_t)stream)
                      : nazb_simt_launch(h, io, G, (cudaStream_t)stream);
  CK(h, e);
  return NAZB_OK;
}

extern "C" int nazb_forward(nazb_handle* h, int32_t s_begin, int32_t s_count, const float* z, int32_t z_shared,
                            const float* ctx, int32_t ctx_rows, int32_t N, const float* lo, const float* hi, float* x,
                            float* logdet, void* stream) {
  int rc = check_io(h, s_begin, s_count, z, ctx, ctx_rows, N, lo, hi);
  if (rc != NAZB_OK) return rc;
  if (!x) return NAZB_ERR_BAD_ARG;
  CK(h, cudaSetDevice(h->device));
  IoArgs io{};
  io.x = z; io.x_draw_stride = z_shared ? 0 : (long long)N * h->geom.D; io.ctx = ctx; io.ctx_rows = ctx_rows; io.N = N;
  io.s_begin = s_begin; io.s_count = s_count; io.lo = lo; io.hi = hi;
  io.out_x = x; io.out_l = logdet; io.dir = 1; io.aff = h->aff_dev;
  const int rows_fwd = (nazb_engine_for_direction(h, 1) == NAZB_ENGINE_TCGEN05) ? nazb_tc_rows_per_item(h, 1) : 64;
  int G = pick_groups(h, N, s_count, 0, rows_fwd);
  cudaError_t e = (nazb_engine_for_direction(h, 1) == NAZB_ENGINE_TCGEN05)
                      ? nazb_tc_launch(h, io, G, (cudaStream_t)stream)
                      : nazb_simt_launch(h, io, G, (cudaStream_t)stream);
  CK(h, e);
  return NAZB_OK;
}

// value and gradient of the draw-batched log-likelihood (what the reference obtains from jax.grad /
// autograd of bflow_jax_maf.py:233-235 `log_prob`).  Gradients are ADDED to the caller's arrays.
static int grad_impl(nazb_handle* h, int32_t s_begin, int32_t s_count, const float* x, const float* ctx,
                     int32_t ctx_rows, int32_t N, const float* lo, const float* hi,
                     const float* const* mask, float* const* gW, float* const* gb, const int64_t* gwst,
                     const int64_t* gbst, float* dx, float* dctx, float* lp, double* sum_n, const float* w,
                     int64_t w_draw_stride, void* stream) {
  int rc = check_io(h, s_begin, s_count, x, ctx, ctx_rows, N, lo, hi);
  if (rc != NAZB_OK) return rc;
  if (!mask || !gW || !gb || !gwst || !gbst) return NAZB_ERR_BAD_ARG;
  const FlowGeom& g = h->geom;
  // masked-affine and neural-spline flows (both orders) on the fp32 engine's image, no dropout
  if (h->engine != NAZB_ENGINE_SIMT || !h->packed || h->has_keep || !nazb_grad_fits(g))
    return NAZB_ERR_UNSUPPORTED;
  const int n = g.L * (g.n_hidden + 1);
  for (int i = 0; i < n; ++i)
    if (!mask[i] || !gW[i] || !gb[i] || gwst[i] < 0 || gbst[i] < 0) return NAZB_ERR_BAD_ARG;
  CK(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (!h->grad_tabs) CK(h, cudaMalloc(&h->grad_tabs, sizeof(void*) * 5 * (size_t)n));
  std::vector<unsigned long long> tabs(5 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    tabs[i] = (unsigned long long)(uintptr_t)mask[i];
    tabs[n + i] = (unsigned long long)(uintptr_t)gW[i];
    tabs[2 * n + i] = (unsigned long long)(uintptr_t)gb[i];
    tabs[3 * n + i] = (unsigned long long)gwst[i];
    tabs[4 * n + i] = (unsigned long long)gbst[i];
  }
  // pageable source: the runtime stages the bytes before returning, so the vector may die with this frame
  CK(h, cudaMemcpyAsync(h->grad_tabs, tabs.data(), tabs.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  IoArgs io{};
  io.x = x; io.x_draw_stride = 0; io.ctx = ctx; io.ctx_rows = ctx_rows; io.N = N;
  io.s_begin = s_begin; io.s_count = s_count; io.lo = lo; io.hi = hi;
  io.out_l = lp; io.sum_n = sum_n; io.dir = 0; io.aff = h->aff_dev;
  CK(h, nazb_grad_launch(h, io, h->grad_tabs, dx, dctx, w, (long long)w_draw_stride, st));
  return NAZB_OK;
}

extern "C" int nazb_inverse_grad(nazb_handle* h, int32_t s_begin, int32_t s_count, const float* x, const float* ctx,
                                 int32_t ctx_rows, int32_t N, const float* lo, const float* hi,
                                 const float* const* mask, float* const* gW, float* const* gb, const int64_t* gwst,
                                 const int64_t* gbst, float* dx, float* lp, double* sum_n, void* stream) {
  return grad_impl(h, s_begin, s_count, x, ctx, ctx_rows, N, lo, hi, mask, gW, gb, gwst, gbst, dx, nullptr, lp, sum_n, nullptr, 0, stream);
}

// Vector-Jacobian product of lp[s][n] with caller-given cotangents w: what `loss.backward()` of the reference's training loop
// needs (train_flows.py:195-213: loss = -flow.log_prob(x, condition=y).mean(); loss.backward()).
extern "C" int nazb_inverse_vjp(nazb_handle* h, int32_t s_begin, int32_t s_count, const float* x, const float* ctx,
                                int32_t ctx_rows, int32_t N, const float* lo, const float* hi,
                                const float* const* mask, float* const* gW, float* const* gb, const int64_t* gwst,
                                const int64_t* gbst, float* dx, float* dctx, float* lp, const float* w,
                                int64_t w_draw_stride, void* stream) {
  if (!w || w_draw_stride < 0) return NAZB_ERR_BAD_ARG;
  return grad_impl(h, s_begin, s_count, x, ctx, ctx_rows, N, lo, hi, mask, gW, gb, gwst, gbst, dx, dctx, lp, nullptr, w, w_draw_stride, stream);
}
