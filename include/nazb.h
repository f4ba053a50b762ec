/*
 * nazb.h — C ABI of libnazb.so: B200-native draw-batched normalizing-flow evaluation.
 *
 * This is the drop-in boundary for ONE path of AnaryaRay1/naz: `log_prob` / `sample` of its discrete
 * flows (masked-affine-autoregressive "maf", neural-spline-autoregressive "nsa") evaluated for S weight
 * draws x N points.  The reference has no FFI of its own (it is 100 % Python) — the
 * interface it exposes for this path is a set of Python methods.  Each entry point below names the
 * reference code it replaces (paths relative to the reference tree):
 *
 *   nazb_create / nazb_destroy   NormalizingFlow.__init__ + flow_makers        src/naz/flows/flow.py:21-42
 *                                masked_affine_autoregressive                  src/naz/flows/transforms.py:133-160
 *                                neural_spline_autoregressive                  src/naz/flows/transforms.py:165-198
 *   nazb_pack                    set_params (per-draw in-place weight swap)    src/naz/trainers/train_flows.py:47-71
 *                                torch_to_jax (weights/masks/perm export)      src/naz/flows/bflow_jax_maf.py:26-46
 *                                masked_linear's `W * mask` (re-done per call) src/naz/flows/bflow_jax_maf.py:74-77
 *                                dropout conditioners (per-draw keep masks)    src/naz/flows/transforms.py:29-65
 *   nazb_inverse                 NormalizingFlow.log_prob                      src/naz/flows/flow.py:45-79
 *                                make_normalizing_flow(...)["lp"]              src/naz/flows/bflow_jax_maf.py:210-212
 *                                inverse_fn (D-pass autoregressive inverse)    src/naz/flows/bflow_jax_maf.py:181-194
 *                                the per-draw lp loops                         examples/papers/2506.05657/compute_bic_simpler.py:116-120
 *   nazb_forward                 NormalizingFlow.sample                        src/naz/flows/flow.py:94-129
 *                                make_normalizing_flow(...)["sampler"]         src/naz/flows/bflow_jax_maf.py:214-223
 *                                predict (loop over posterior draws)           src/naz/trainers/train_flows.py:384-422
 *                                MCDPNormalizingFlow.sample_uncertain          src/naz/flows/mcdpflow.py:39-56
 *   nazb_lse_reduce/_finish      mean_s exp(lp_s) posterior predictive         examples/papers/2506.05657/plot.py:272-275
 *   nazb_importance              Importance(...).run + posterior.ESS()         src/naz/trainers/train_flows.py:358-380
 *                                compute_bic (max_s sum_n lp)                  src/naz/flows/bflow_jax_maf.py:474-475
 *   nazb_pack_draw_map           theta_0 * (1 + scale * standard_params)        src/naz/flows/bflow_jax_maf.py:239-240
 *   nazb_truncnorm_sample        TruncatedNormalTransform.__call__ / log_prob   src/naz/priors/TruncatedNormal.py:14-60
 *   nazb_inverse_grad            jax.value_and_grad(log_prob) / autograd of    src/naz/flows/bflow_jax_maf.py:233-246,277-287,
 *                                the summed log-likelihood (NUTS, SVI, MLE)    :321-327,:344-348; trainers/train_flows.py:195-213
 *   nazb_histogramdd             per-draw np.histogram2d / jnp.histogramdd     src/naz/flows/bflow_jax_maf.py:436-441
 *                                of the [S][N][D] sample tensor (density=True)
 *   nazb_hpd                     hpd_vectorized across draws                   src/naz/statutils.py:22-46
 *
 * Conventions
 *   - Every data pointer is a DEVICE pointer owned by the caller (e.g. torch.Tensor.data_ptr()),
 *     contiguous row-major fp32 unless stated.  Pointer *tables* (W, b, mask) and `perm`, `hid_deg`
 *     are HOST arrays.  The library owns only the handle (packed weights + workspace).
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation
 *     except inside nazb_create / nazb_destroy.
 *   - Return value: 0 on success, negative nazb_status otherwise; nazb_strerror() explains; the last
 *     CUDA error string is kept per handle (nazb_last_cuda_error).
 *   - Base noise is always an INPUT (RNG stays with the caller; parity needs identical noise).
 *   - One handle per device; a handle is not thread-safe, distinct handles are.
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef NAZB_H_
#define NAZB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAZB_MAX_HIDDEN_LAYERS 8
#define NAZB_MAX_DIM 32

typedef enum {
  NAZB_OK = 0,
  NAZB_ERR_BAD_ARG = -1,       /* null pointer / non-positive size / inconsistent shapes            */
  NAZB_ERR_UNSUPPORTED = -2,   /* configuration outside what any engine supports                      */
  NAZB_ERR_CUDA = -3,          /* a CUDA runtime call failed; see nazb_last_cuda_error()              */
  NAZB_ERR_NOT_PACKED = -4,    /* nazb_forward / nazb_inverse before nazb_pack                        */
  NAZB_ERR_NO_DEVICE = -5      /* no CUDA device / wrong architecture (needs sm_100)                  */
} nazb_status;

typedef enum {
  NAZB_KIND_AFFINE = 0,        /* pyro AffineAutoregressive: y = mu + x*exp(clip(s)), M = 2           */
  NAZB_KIND_RQS = 1,           /* SplineAutoregressive order="quadratic" (naz default), M = 3K-1      */
  NAZB_KIND_RLS = 2            /* SplineAutoregressive order="linear", M = 4K-1                       */
} nazb_kind;

typedef enum {
  NAZB_ENGINE_AUTO = 0,        /* tcgen05 when the shape fits its envelope, else SIMT                 */
  NAZB_ENGINE_SIMT = 1,        /* fp32 CUDA-core kernels, any shape                                   */
  NAZB_ENGINE_TCGEN05 = 2      /* fp16 hi/lo-split tensor-core kernels (error if shape unsupported)   */
} nazb_engine;

typedef enum {
  NAZB_INV_INCREMENTAL = 0,    /* one block-triangular pass per flow layer (needs hid_deg)            */
  NAZB_INV_JACOBI = 1          /* the reference's D full conditioner passes per flow layer            */
} nazb_inverse_mode;

typedef struct nazb_handle nazb_handle;

/* Static description of the flow; mirrors the positional arguments of naz's factories
 * (theta_dim, condition_dim, hidden_dim, num_layers[, count_bins]) — transforms.py:133,165. */
typedef struct nazb_desc {
  int32_t kind;                               /* nazb_kind                                            */
  int32_t D;                                  /* theta_dim, 1..NAZB_MAX_DIM                           */
  int32_t C;                                  /* condition_dim, 0 = unconditional                     */
  int32_t L;                                  /* num_layers (flow layers)                             */
  int32_t n_hidden;                           /* len(hidden_dims), 1..NAZB_MAX_HIDDEN_LAYERS          */
  int32_t hidden[NAZB_MAX_HIDDEN_LAYERS];     /* hidden_dims                                          */
  int32_t count_bins;                         /* K (splines), ignored for affine                      */
  float bound;                                /* spline box half-width B (pyro default 3.0)           */
  float clip_lo, clip_hi;                     /* affine log-scale clamp (pyro default -5, 3)          */
  int32_t S;                                  /* number of weight draws this handle holds (S_local)   */
  int32_t engine;                             /* nazb_engine                                          */
  int32_t inverse_mode;                       /* nazb_inverse_mode                                    */
  int32_t device;                             /* CUDA device ordinal                                  */
} nazb_desc;

int nazb_create(nazb_handle** out, const nazb_desc* desc);
void nazb_destroy(nazb_handle* h);

/* Which engine the handle resolved to (nazb_engine), or a negative status. */
int nazb_engine_in_use(const nazb_handle* h);

/* Engine serving one direction after nazb_pack: dir 0 = nazb_inverse, 1 = nazb_forward.  A direction whose
 * tensor-core program does not fit TMEM is served by the SIMT engine when the handle was created with
 * NAZB_ENGINE_AUTO (nazb_pack fails with NAZB_ERR_UNSUPPORTED when NAZB_ENGINE_TCGEN05 was forced). */
int nazb_engine_for_direction(const nazb_handle* h, int dir);

/* Fold masks (and optional per-draw dropout keep-masks) into the packed, engine-specific weight
 * image for all S draws.  n_lin = n_hidden + 1 linears per flow layer, tables indexed [l*n_lin + j].
 *   W[i]         device, [S][out][in] with draw stride w_draw_stride[i] floats (0 = one shared set)
 *   b[i]         device, [S][out]     with draw stride b_draw_stride[i] floats (0 = shared)
 *   mask[i]      device, [out][in] fp32 0/1 (arn.masks as exported by torch_to_jax)
 *   perm         host int64 [L][D] (arn.permutation)
 *   hid_deg      host int32 [n_hidden][max(hidden)] MADE degrees of the hidden units (non-decreasing
 *                per layer); required for NAZB_INV_INCREMENTAL, may be NULL for NAZB_INV_JACOBI
 *   keep         device [S][L][n_hidden][max(hidden)] 0/1 or NULL; unit j of hidden layer k of flow
 *                layer l in draw s is multiplied by keep/(1-p_drop) (transforms.py:38-43), folded into
 *                the NEXT linear's packed columns
 */
int nazb_pack(nazb_handle* h, const float* const* W, const float* const* b,
              const int64_t* w_draw_stride, const int64_t* b_draw_stride,
              const float* const* mask, const int64_t* perm, const int32_t* hid_deg,
              const float* keep, float p_drop, void* stream);

/* Reference `log_prob` direction (autoregressive inverse) for draws [s_begin, s_begin + s_count).
 *   x        [N][D] points (shared by all draws)
 *   ctx      [ctx_rows][C]; ctx_rows == N (per point) or 1 (broadcast); NULL iff C == 0
 *   lo, hi   [D] bounding box (flow.py:70-73) or both NULL
 *   z        out [s_count][N][D] base-space points, or NULL
 *   lp       out [s_count][N] log p(x_n | theta_s), or NULL
 *   log_w    [s_count] per-draw log-weights added before the cross-draw logsumexp, or NULL (= 0)
 *   lse_max, lse_sum   out [n_groups][N] running (max, sum exp) partials over the draws of each
 *                      group (draw s belongs to group s % n_groups), or both NULL
 *   sum_n    out [s_count] double, sum over points of lp (must be zeroed by the caller), or NULL
 */
int nazb_inverse(nazb_handle* h, int32_t s_begin, int32_t s_count,
                 const float* x, const float* ctx, int32_t ctx_rows, int32_t N,
                 const float* lo, const float* hi,
                 float* z, float* lp, const float* log_w,
                 float* lse_max, float* lse_sum, int32_t n_groups,
                 double* sum_n, void* stream);

/* Reference `sample` direction (one conditioner pass per flow layer).
 *   z        base noise, [s_count][N][D] (z_shared == 0) or [N][D] shared by all draws (z_shared != 0)
 *   x        out [s_count][N][D] samples (inverse bounding applied when lo/hi given)
 *   logdet   out [s_count][N] sum of forward log|det J| over flow layers, or NULL
 */
int nazb_forward(nazb_handle* h, int32_t s_begin, int32_t s_count,
                 const float* z, int32_t z_shared, const float* ctx, int32_t ctx_rows, int32_t N,
                 const float* lo, const float* hi,
                 float* x, float* logdet, void* stream);

/* nazb_pack with the Bayesian-flow draw map applied while packing:
 *   theta_s = theta_0 * (1 + scale * u_s)   src/naz/flows/bflow_jax_maf.py:239-240 (fp32, rounded after every operation)
 * W0 / b0: HOST tables [L * (n_hidden + 1)] of device pointers to the MLE weights (no draw axis);
 * uW / ub + uwst / ubst: the standard parameters u_s in [-1, 1] with the same table / stride convention as nazb_pack
 * (e.g. views into the reference's flat `standard_params` [S, P] matrix: stride P). */
int nazb_pack_draw_map(nazb_handle* h, const float* const* W0, const float* const* b0, const float* const* uW,
                       const float* const* ub, const int64_t* uwst, const int64_t* ubst, float scale,
                       const float* const* mask, const int64_t* perm, const int32_t* hid_deg,
                       const float* keep, float p_drop, void* stream);

/* Truncated-normal guide of the SVI / importance path: src/naz/priors/TruncatedNormal.py:14-60, used by
 * src/naz/flows/bflow.py:35,43 and (numpyro's twin) bflow_jax_maf.py:257.  x: device fp32 [S][P] uniforms in (0, 1)
 * supplied by the caller (RNG stays with the caller); loc / scale / low / high: device fp32 with 1 or P elements each;
 * y: [S][P] samples (feed them to nazb_pack_draw_map as standard parameters); log_q: device double [S] = sum over the P
 * parameters of log pdf(y) - log(cdf(high) - cdf(low))  (the log q(theta_s) of nazb_importance). */
int nazb_truncnorm_sample(const float* x, int32_t S, int64_t P, const float* loc, int32_t loc_n, const float* scale,
                          int32_t scale_n, const float* low, int32_t low_n, const float* high, int32_t high_n,
                          float* y, double* log_q, void* stream);

/* SURVEY §8(f) f1 — value and gradient of the summed log-likelihood of draws [s_begin, s_begin + s_count):
 *   sum_n[s] += sum_n log p(x_n | ctx_n; theta_s)   (device double [s_count], like nazb_inverse; may be NULL)
 *   gW[i][s] += d sum_n / d W_i  (reference layout [out][in]; entries where mask_i == 0 are not touched)
 *   gb[i][s] += d sum_n / d b_i
 *   dx[s][n][:] = d lp[s][n] / d x_n  (device fp32 [s_count][N][D], in data space when lo / hi are given) or NULL
 *   lp           device fp32 [s_count][N] or NULL.
 * i runs over [flow layer][linear] exactly as in nazb_pack; mask / gW / gb are HOST tables of DEVICE pointers, gwst / gbst
 * HOST arrays with the number of floats between consecutive draws of gW[i] / gb[i] (indexed by the GLOBAL draw s).  The
 * gradient arrays are accumulated into (zero them first).  With nazb_pack_draw_map the gradient with respect to the standard
 * parameters is scale * theta_0 * gW (chain rule on bflow_jax_maf.py:239-240; done by the caller).
 * Covered: every flow kind (masked-affine, quadratic and linear-order neural-spline) on a handle created with
 * NAZB_ENGINE_SIMT, no dropout keep-masks; anything else returns NAZB_ERR_UNSUPPORTED.  fp32 atomics: the sum over points is not bit-reproducible run to run. */
int nazb_inverse_grad(nazb_handle* h, int32_t s_begin, int32_t s_count, const float* x, const float* ctx,
                      int32_t ctx_rows, int32_t N, const float* lo, const float* hi, const float* const* mask,
                      float* const* gW, float* const* gb, const int64_t* gwst, const int64_t* gbst, float* dx, float* lp,
                      double* sum_n, void* stream);

/* The same pass with caller-given cotangents of lp — the vector-Jacobian product behind `loss.backward()` of the reference's
 * MLE loop (src/naz/trainers/train_flows.py:195-213: loss = -flow.log_prob(x_batch, condition=y_batch).mean(); loss.backward()):
 *   gW[i][s] += sum_n w[s][n] d lp[s][n] / d W_i,   gb likewise,   dx[s][n][:] = w[s][n] d lp[s][n] / d x_n,
 *   dctx[s][n][:] = w[s][n] d lp[s][n] / d ctx_n  (device fp32 [s_count][N][C] or NULL; for a broadcast context the caller sums
 *   over n) — the cotangent a trainable embedding network in front of the flow needs (src/naz/flows/flow.py:30-36, :75).
 * w: device fp32, draw s at w + (s - s_begin) * w_draw_stride (0 = one [N] vector shared by the draws).  Same coverage as
 * nazb_inverse_grad, which is the case w == 1. */
int nazb_inverse_vjp(nazb_handle* h, int32_t s_begin, int32_t s_count, const float* x, const float* ctx,
                     int32_t ctx_rows, int32_t N, const float* lo, const float* hi,
                     const float* const* mask, float* const* gW, float* const* gb, const int64_t* gwst,
                     const int64_t* gbst, float* dx, float* dctx, float* lp, const float* w, int64_t w_draw_stride,
                     void* stream);

/* Stand-alone cross-draw reduction over a materialised lp[S][N] (kernel group 4):
 * partial (max, sum exp) per point over this rank's draws.  HBM-bound: 4*S*N bytes read. */
int nazb_lse_reduce(const float* lp, int32_t S, int32_t N, const float* log_w,
                    float* lse_max, float* lse_sum, void* stream);

/* Combine G partials (from draw groups and/or ranks): out[n] = log sum_g sum_g[n] exp(max_g[n]) + log_norm. */
int nazb_lse_finish(const float* lse_max, const float* lse_sum, int32_t G, int32_t N,
                    float log_norm, float* out, void* stream);

/* Importance weights: log_w[s] = log_prior[s] + sum_n[s] - log_q[s]; log_evidence = lse(log_w) - log S;
 * ess = exp(2 lse(log_w) - lse(2 log_w)); max_sum = max_s sum_n[s] (BIC).  log_prior / log_q may be NULL.
 * out3 is a device double[3] = {log_evidence, ess, max_sum}; log_w_out device double[S] or NULL. */
int nazb_importance(const double* sum_n, const float* log_prior, const float* log_q, int32_t S,
                    double* log_w_out, double* out3, void* stream);

/* Consumers of the sample tensor.
 * nazb_histogramdd: numpy.histogramdd semantics per draw — bin = searchsorted(edges_d, v, "right") - 1, the right-most
 * edge belongs to the last bin, samples outside any dim's edges (or NaN) are dropped; comparisons in double.
 *   x        device fp32 [S][N][D];   edges  device double, the D edge arrays concatenated (nbins[d] + 1 each);
 *   nbins    HOST int32[D];           counts device uint32 [S][prod nbins] (zeroed by the call, C order);
 *   density  device fp32 [S][prod nbins] or NULL: counts / (sum of the draw's counts * bin volume)  (density=True).
 * nazb_hpd: v device fp32 [S][M] (e.g. the densities above, M = prod nbins) -> per column the narrowest interval
 * holding floor((1 - alpha) * S) + 1 order statistics: lo[M], hi[M]  (first minimum on ties, as numpy.argmin);
 * S <= 32768 draws (the per-column sort runs in shared memory), NAZB_ERR_UNSUPPORTED beyond. */
int nazb_histogramdd(const float* x, int32_t S, int64_t N, int32_t D, const double* edges, const int32_t* nbins,
                     uint32_t* counts, float* density, void* stream);
int nazb_hpd(const float* v, int32_t S, int64_t M, double alpha, float* lo, float* hi, void* stream);

/* Engine options (tuning and A/B switches; this library never reads the environment).  Names, tcgen05 engine:
 *   "inv_kernel"   5 (default: one 128-row chain, 16 epilogue warps) | 6 (24 epilogue warps, split pushes) |
 *                  4 (two 64-row chains) | 3 (round-1 kernel)                                   — needs a new nazb_pack
 *   "inv_merge_n"  pushes with N <= value are issued unsplit; 0 = always critical columns first; -1 (default) = by shape
 *                  (split for flow layers with >= 4 hidden blocks)                              — needs a new nazb_pack
 *   "inv_align"    block-aligned accumulator / operand columns: 1 on, 0 off, -1 (default) by shape — needs a new nazb_pack
 *   "inv_trim"     1 (default) context-folded programs drop the dead degree-0 accumulator columns — needs a new nazb_pack
 *   "inv_gaps"     1: build inverse programs for MADE degree ladders with unpopulated degrees (the single-degree form of
 *                  coupling layers); 0 (default): such ladders are served by the fp32 engine            — needs a new nazb_pack
 *   "inv_a_tmem"   1 (default) A operand of the pushes in tensor memory when the plan has room for it
 *   "inv_fold"     1 (default) fold a broadcast context (ctx_rows == 1) into per-draw constants inside nazb_inverse
 *   "inv_gate"     bound the drift of CTAs across draw groups (keeps the weight images L2-resident): 0 off, 2 = no CTA starts
 *                  a group before all finished issuing the previous one, 3 = one group of slack, 1 (default) = by tile count
 *   "grad_diag" / "grad_tile"  nazb_inverse_grad (any engine): 1 = skip the gradient atomics (timing diagnosis) / 16 = force
 *                  16-point tiles (default 0: by shared-memory fit)
 *   "grad_stash"   1 (default) nazb_inverse_grad parks the conditioner activations of its value pass in a per-CTA scratch area
 *                  (grow-only device allocation, ~1 MB per resident CTA) instead of recomputing them in the adjoint pass
 * nazb_get_option also answers "inv_fold_available", "inv_kernel_in_use", "inv_a_tmem_in_use", "inv_block_width" and "watchdog" (non-zero after a kernel aborted on a barrier time-out:
 * site | warp << 8 | block << 16).  Unknown names return NAZB_ERR_BAD_ARG, the SIMT engine NAZB_ERR_UNSUPPORTED. */
int nazb_set_option(nazb_handle* h, const char* name, int32_t value);

/* Element-wise affine after every flow layer, the eval()-mode form of pyro's T.BatchNorm that naz appends with
 * use_batchnorm=True (src/naz/flows/transforms.py:157-158, :195-196):  sampling direction x <- a[l][d] * x + b[l][d] after
 * flow layer l, log-det sum_d log a[l][d];  nazb_inverse applies (y - b) / a before inverting layer l.  a, b: HOST fp32
 * [L][D] (a > 0), copied on `stream`; a == NULL removes the step.  Flows with the step are served by the fp32 SIMT engine in
 * both directions (call this BEFORE nazb_pack; a handle created with NAZB_ENGINE_TCGEN05 returns NAZB_ERR_UNSUPPORTED);
 * nazb_inverse_grad / nazb_inverse_vjp treat a and b as constants (no gradient with respect to them). */
int nazb_set_layer_affine(nazb_handle* h, const float* a, const float* b, void* stream);
int nazb_get_option(const nazb_handle* h, const char* name, int32_t* value);

/* Test hook, needs no GPU: the device routine behind the neural-spline gradient (naz_b200/csrc/spline_grad.cuh) compiled for
 * the host.  For spline input x and the 3K-1 (quadratic order) or 4K-1 (linear_order != 0) raw conditioner outputs of one
 * (point, dimension):  inv_tx = 1 / (dT/dx),  ldx = d log T'(x) / dx,  ca[m] = -(dT/draw_m) / (dT/dx),
 * cb[m] = -d log T'(x) / draw_m   (spline on [-bound, bound], identity outside). */
int nazb_host_spline_grad(float x, int32_t K, float bound, int32_t linear_order, const float* raw, float* ca, float* cb,
                          float* inv_tx, float* ldx);

const char* nazb_strerror(int status);
const char* nazb_last_cuda_error(const nazb_handle* h);

/* Introspection used by bench.py / tests: bytes of packed weights held, kernels launched so far. */
int64_t nazb_packed_bytes(const nazb_handle* h);
int64_t nazb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NAZB_H_ */
