// Internal structures shared by the kernels and the C-ABI glue of libnazb.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/nazb.h"

#define NAZB_MAX_LIN (NAZB_MAX_HIDDEN_LAYERS + 1)

// Geometry + schedule of one flow; passed BY VALUE to kernels (fits the 4 KB parameter space).
struct FlowGeom {
  int kind, D, C, L, n_hidden, M, K;
  float bound, clip_lo, clip_hi;
  int kin;                      // C + D
  int hidden[NAZB_MAX_HIDDEN_LAYERS];
  int hmax;                     // max hidden width rounded up to 4
  int md;                       // M * D (output width)
  // SIMT packed layout: per draw, per flow layer, linear j is Wt[kdim[j]][ldw[j]] (k-major rows,
  // mask folded, output columns rank-major: col = rank(d)*M + m) followed by bias[ldw[j]].
  int kdim[NAZB_MAX_LIN];
  int ndim[NAZB_MAX_LIN];
  int ldw[NAZB_MAX_LIN];
  long long off_w[NAZB_MAX_LIN], off_b[NAZB_MAX_LIN];
  long long layer_stride, draw_stride;   // floats
  // Inverse schedule. Stage r (0..D-1) finalises the dimension of rank r.
  //   incremental: hidden layer j computes units [blk[j][r], blk[j][r+1]) from the first
  //                blk[j-1][r+1] units of the previous layer; output columns [r*M, (r+1)*M).
  //   jacobi:      every stage recomputes everything (the reference's D full passes).
  int inv_mode;
  short blk[NAZB_MAX_HIDDEN_LAYERS][NAZB_MAX_DIM + 1];
};

struct IoArgs {
  const float* x;          // inverse: points [N][D]; forward: base noise [S][N][D] or [N][D]
  long long x_draw_stride; // floats between draws of x (0 = shared)
  const float* ctx;        // [ctx_rows][C] or null
  int ctx_rows;
  int N;
  int s_begin, s_count;
  const float* lo;         // [D] or null
  const float* hi;
  float* out_x;            // [s_count][N][D] or null
  float* out_l;            // inverse: lp [s_count][N]; forward: logdet [s_count][N]; or null
  const float* log_w;      // [s_count] or null
  float* lse_max;          // [G][N] or null
  float* lse_sum;
  double* sum_n;           // [s_count] or null
  int dir;                 // 0 = inverse (log_prob direction), 1 = forward (sample direction)
  const float* aff;        // per-layer affine table [L][2 D + 1] (a[D], b[D], sum log a) or null  (nazb_set_layer_affine)
};

// Draw map of the Bayesian flow (src/naz/flows/bflow_jax_maf.py:239-240):  theta_s = theta_0 * (1 + scale * u_s)  in fp32,
// rounded after every operation exactly as the reference's un-fused jnp expression.  When `base*` tables are given to the
// pack functions, the per-draw tables hold the standard parameters u_s and the map is applied while packing.
struct DrawMap {
  const float* const* baseW = nullptr;   // host table [L * n_lin] of device pointers to theta_0 weights [out][in]
  const float* const* baseB = nullptr;   // ... biases [out]
  float scale = 0.f;
};
__device__ __forceinline__ float nazb_draw_map(float base, float u, float scale) {
  return __fmul_rn(base, __fadd_rn(1.0f, __fmul_rn(scale, u)));
}

struct nazb_handle {
  nazb_desc desc;
  FlowGeom geom;
  int engine;               // resolved engine
  int device;
  int sm_count;
  float* packed = nullptr;  // SIMT image
  int* perm_dev = nullptr;  // [L][D] int32
  bool is_packed = false;
  bool has_keep = false;    // dropout keep-masks were folded in at pack time
  float* aff_dev = nullptr; // [L][2 D + 1] per-layer affine (BatchNorm in eval mode) or null; SIMT engine only
  // gradient engine (flow_grad.cu): transposed copy of the SIMT image, device tables of the caller's gradient arrays
  float* packed_T = nullptr;
  bool packed_T_valid = false;
  void* grad_tabs = nullptr;
  void* grad_stash = nullptr;   // per-CTA scratch of nazb_inverse_grad (activations parked by phase A), grow-only
  size_t grad_stash_bytes = 0;
  int opt_grad_stash = 1;   // nazb_set_option "grad_stash": 0 = recompute the conditioner in the adjoint phase instead
  int opt_grad_diag = 0;    // nazb_set_option "grad_diag": 1 = skip the gradient atomics (timing diagnosis only)
  int opt_grad_tile = 0;    // nazb_set_option "grad_tile": 16 = force 16-point tiles (0 = by shared-memory fit)
  std::string cuda_err;
  // pinned staging buffer for the small host tables of nazb_pack (async upload on the caller's stream)
  void* stage_host = nullptr;
  size_t stage_cap = 0, stage_off = 0;
  void* stage_ev = nullptr;   // cudaEvent_t recorded after the last staged copy
  // tcgen05 engine state (opaque here; defined in flow_tc.cu)
  void* tc = nullptr;
};

cudaError_t nazb_stage_upload(nazb_handle* h, void* dst, const void* src, size_t bytes, cudaStream_t st);

// ---- launchers implemented in the .cu files ----
cudaError_t nazb_simt_launch(const nazb_handle* h, const IoArgs& io, int n_groups, cudaStream_t st);
size_t nazb_simt_smem_bytes(const FlowGeom& g, int P);
int nazb_simt_pick_P(const FlowGeom& g);

cudaError_t nazb_pack_simt(nazb_handle* h, const float* const* W, const float* const* b,
                           const int64_t* wst, const int64_t* bst, const float* const* mask,
                           const float* keep, float p_drop, cudaStream_t st, const DrawMap& dm = DrawMap());

// tcgen05 engine
bool nazb_tc_supported(const FlowGeom& g, std::string* why);
cudaError_t nazb_tc_create(nazb_handle* h);
void nazb_tc_destroy(nazb_handle* h);
cudaError_t nazb_tc_pack(nazb_handle* h, const float* const* W, const float* const* b,
                         const int64_t* wst, const int64_t* bst, const float* const* mask,
                         const float* keep, float p_drop, cudaStream_t st, const DrawMap& dm = DrawMap());
cudaError_t nazb_tc_launch(const nazb_handle* h, const IoArgs& io, int n_groups, cudaStream_t st);
int64_t nazb_tc_packed_bytes(const nazb_handle* h);

// gradient of sum_n lp (masked-affine flows, SIMT image)
cudaError_t nazb_grad_launch(nazb_handle* h, const IoArgs& io, const void* tabs, float* dx, float* dctx, const float* wgt,
                             long long wgt_stride, cudaStream_t st);
bool nazb_grad_fits(const FlowGeom& g);

void nazb_count_launch(int n = 1);
