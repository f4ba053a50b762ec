// tcgen05 engine: the MADE conditioner as fp16 hi/lo-split tensor-core contractions with fp32
// accumulation in TMEM, the affine / rational-spline transform + log-det fused as the epilogue.
//
//   * Points ride the M axis (one CTA = one 128-point tile = the 128 TMEM lanes), weights are the B
//     operand.  Every value is split x = hi + lo (two fp16) and a product is three MMAs
//     (hi*hi + hi*lo + lo*hi) — measured 4e-6 abs error on a K = 160 contraction, i.e. fp32-class;
//     a single fp16/bf16/tf32 MMA misses the 1e-4 / 1e-5 parity bar by 10-100x.
//   * Weights for all draws are pre-packed (masks, dropout keep-masks and biases folded) into the
//     exact shared-memory image each MMA step needs (no-swizzle K-major core matrices), so the
//     producer warp streams them with plain TMA bulk copies (cp.async.bulk) through an mbarrier ring.
//   * The kernel is table-driven: a per-flow-layer list of steps {weights to stream, MMAs to issue,
//     epilogue to run}.  Two programs exist:
//       forward  (reference `sample`):  dense GEMM chain, one pass per flow layer.
//       inverse  (reference `log_prob`): the D-pass autoregressive inverse collapsed into ONE
//                block-triangular pass, "push" style — when the hidden units of MADE degree r of
//                layer j become final they are immediately multiplied into the pre-activation
//                accumulators of layer j+1, which stay resident in TMEM for the whole flow layer.
//   * Warp roles: warps 0-7 epilogue (TMEM -> registers -> tanh / transform -> fp16 hi/lo A operand
//     in shared memory), warp 8 TMA producer, warp 9 TMEM allocator + single-thread MMA issuer.
//
// Reference semantics: src/naz/flows/bflow_jax_maf.py:135-194,210-223; pyro SplineAutoregressive
// (see oracle/flow_oracle.py).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <type_traits>
#include "nazb_internal.h"
#include "tc_ptx.cuh"
#include "transforms.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kEpiWarps = 16;   // 4 per TMEM lane quadrant; the column range of a step is split 4 ways
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kParts = kEpiWarps / 4;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kSlotBytes = 40960;
constexpr int kFwdSlotBytes = 40960;   // pipelined forward kernels (20 KB slots measured 8-12 % slower: more sub-steps, waits and commits)
constexpr int kTmemCols = 512;

enum : uint8_t { EPI_NONE = 0, EPI_TANH = 1, EPI_XINV = 2, EPI_XFWD = 3, EPI_FIRST = 4,
                 EPI_XINV0C = 5 };   // v4 inverse: rank-0 inverse transform from context-folded (per-draw constant) parameters
enum : uint8_t { A_IN = 0, A_H = 1, A_X = 2 };   // A_X: v3 inverse first-layer operand [ctx | x | 1]

struct Step {
  uint32_t w_off;      // byte offset of the weight image inside a flow layer's block
  uint32_t w_bytes;    // hi image + lo image; 0 = no MMA in this step
  uint16_t a_chunk0;   // first 16-byte K-chunk of the A operand
  uint16_t ksteps;     // K / 16
  uint16_t n;          // N extent (multiple of 16)
  uint16_t d_col;      // TMEM column of D
  uint8_t a_buf, nsplit, accumulate, epi;
  uint8_t stage, nranks, flags, pad1;   // flags bit 0: the accumulator holds nothing yet (bias-only output);
                                        // bit 2: TANH block whose accumulator already holds scaled pre-activation + bias
  uint16_t e_col, e_ncols, e_dst_chunk;
  uint16_t e_aux;   // v2 inverse: TANH / XINV -> float offset of the bias of column e_col inside the layer
                    // constants; FIRST -> first hidden unit of the block
  uint16_t n_crit;  // inverse pushes: the first n_crit image rows (the columns of block r) are the critical part
};

struct Image {          // how the pack kernel fills one step's weight image
  uint32_t w_off, w_bytes;
  int lin, n_ext, k_ext;
  int row_mode;         // 0: row -> hidden unit n0 + n; 1: out, rank r0 + n / Mp, slot n % Mp; 2: out, rank r0 + n / M, slot n % M
  int n0, r0, r1;
  int k0, kv0, kv1;     // image column c <-> source column k0 + c, used iff kv0 <= k0 + c < kv1
  int bias_col, bias_only;
  float scale;          // multiplies every packed value (1, or 2 log2 e for tensor-core first-layer blocks)
};

struct TcPlan {
  bool ok[2] = {false, false};   // [0] inverse, [1] forward
  std::vector<Step> steps[2];
  std::vector<Image> images[2];
  size_t layer_bytes[2] = {0, 0};
  int kin_pad = 0, hp_max = 0, mp = 0, nslots = 0;
  size_t smem_bytes = 0;
  uint32_t off_in, off_h, off_x, off_y, off_xo, off_ctx, off_misc, off_scratch, off_ring;
  // v2 inverse: per-(draw, layer) fp32 constants [W0 (Hp0 x kinp) | b_0 .. b_{nh-1} | b_out (D x Mp, rank-major)]
  int kinp = 0, lc_floats = 0, lc_b[NAZB_MAX_HIDDEN_LAYERS] = {0}, lc_bout = 0;
  // v3 inverse kernel: double-buffered A blocks of kr_max columns per chain
  int kr_max = 0;
  // slice-pipelined forward kernel (flow_tc_fwd3.cuh)
  bool fwd3 = false;
  bool fwd4 = false;    // two 128-point tiles per CTA sharing one A ring (flow_tc_fwd4.cuh); same program as fwd3
  uint32_t g_ax = 0, g_a = 0, g_x = 0, g_ctx = 0, g_misc = 0, g_lc = 0, g_ring = 0;
  uint32_t g_slot_off[8] = {0};
  int g_nslots = 0;
  size_t g_smem_bytes = 0;
  int f_lc_floats = 0, f_lc_b[NAZB_MAX_HIDDEN_LAYERS] = {0}, f_lc_bout = 0, f_nslots = 0;
  uint32_t f_ax = 0, f_a = 0, f_x = 0, f_ctx = 0, f_misc = 0, f_lc = 0, f_ring = 0;
  size_t f_smem_bytes = 0;
  bool xf = false;      // first conditioner layer on tensor cores (K = 16 slice [ctx | x | 1])
  uint32_t j_ax = 0, j_xring = 0;
  int xslot_bytes = 0;  // small-ring slot: the widest first-layer image (units x K = 16, hi + lo), rounded up to 1 KB
  uint32_t j_xin = 0, j_lc = 0, j_a = 0, j_y = 0, j_xo = 0, j_misc = 0, j_scratch = 0, j_ring = 0;
  int j_nslots = 0;
  size_t j_smem_bytes = 0;
  // v4 inverse kernel (flow_tc_inv4.cuh): steps[0] is its GENERAL program; steps_fold the context-folded one (same images)
  int inv_ver = 4;
  int mp_inv = 0;                 // output columns per rank in the inverse accumulators / images (v4: ceil8(M))
  bool fold_ok = false;
  std::vector<Step> steps_fold;
  std::vector<Image> fold_images; // stage-0 push images (what inv4_fold_kernel contracts with the degree-0 activations)
  int lc_w0x = 0, lc_w0c = 0, lc_r0c = 0, dp4 = 0, cp4 = 0;
  uint32_t j_xr = 0;
  bool fold_a_tmem = false;       // ... same for the folded program
  int t_a = 0;                    // v5: TMEM column of the A operand (hi image; lo at + kr_max / 2), valid when a_tmem
  bool a_tmem = false;
  int t_a_fold = 0;               // ... of the folded program (its accumulators drop the degree-0 columns)
  bool split = false, split_fold = false;   // the program has pushes whose critical columns are issued first
  bool a_tmem2 = false, fold_a_tmem2 = false;   // room for a second A buffer in TMEM (deferred trailing MMAs, kSplit = 2)
  // Block-aligned column layout of the inverse programs (v5 / v6; 0 = units packed contiguously in degree order): the hidden
  // units of MADE degree r of every hidden layer occupy columns [r bw, r bw + n_r), the rest of the block is zero padding
  // (zero image rows / columns, zero biases), so an A block is exactly bw / 16 K slices and the epilogue of a block touches
  // aw = ceil8(max n_r) columns instead of the 8-aligned hull of an unaligned range (cfg3: 40 instead of 56-64).
  int bw = 0, aw = 0;
  bool allow_gaps = false;        // inverse programs for degree ladders with unpopulated degrees (set from opt_gaps)
  bool trim = true;               // folded programs: accumulators start at the first live block (set from opt_trim before building)
  int hpad[NAZB_MAX_HIDDEN_LAYERS] = {0};   // column count of each hidden layer in the inverse programs (ceil16(H) or D bw)
};

struct TcState {
  TcPlan plan;
  uint8_t* wimg[2] = {nullptr, nullptr};
  Step* steps_dev[2] = {nullptr, nullptr};
  const float** tab_dev = nullptr;   // [3][L * n_lin] W / b / mask pointer tables
  size_t draw_bytes[2] = {0, 0};
  float* lc_dev = nullptr;           // [S][L][lc_floats]
  float* lcf_dev = nullptr;          // forward layer constants [S][L][f_lc_floats]
  float* lcfold_dev = nullptr;       // v4 inverse: context-folded layer constants [S][L][lc_floats] (rewritten per call)
  int* grp_done = nullptr;           // v4 inverse: draw-group gate counters [65536]
  short* pmap_dev = nullptr;         // block-aligned layout: [NAZB_MAX_HIDDEN_LAYERS][256] column -> hidden unit (-1 = padding)
  unsigned int* wd_host = nullptr;   // watchdog word: mapped pinned host memory ...
  unsigned int* wd_dev = nullptr;    // ... and its device alias
  size_t cap_wimg[2] = {0, 0}, cap_lc = 0, cap_lcf = 0, cap_lcfold = 0, cap_tab = 0;
  // options (nazb_set_option); recorded by bench.py
  int opt_inv_kernel = 5;            // 3 = round-1 kernel, 4 = v4 (two 64-row chains), 5 = v5 (one 128-row chain, 16 epilogue warps),
                                     // 6 = v6 (one 128-row chain, 24 epilogue warps, split pushes)
  int opt_fold = 1;                  // context fold when ctx_rows == 1
  int opt_merge_n = -1;              // pushes with N <= merge_n are issued unsplit (critical + deferred columns in one MMA);
                                     // -1 = kernel default (v4: 0 = always split, v5: 256 = never split)
  int opt_gate = 1;                  // draw-group gate for large N: 0 = off; 2 = distance 1 (no CTA starts a group before all have finished
                                     // issuing the previous one); 3 = distance 2 (a CTA may run one group ahead of the slowest); 1 = auto:
                                     // distance 1 when every CTA has >= 32 tiles per group (measured: free at 53 tiles, +2.3 % at 26; the wait at a group boundary is then < 3 % of the
                                     // group), else distance 2.  Full-size cfg3: 76 GB of DRAM reads per launch at distance 2, 5.4 GB at 1.
  int opt_a_tmem = 1;                // v5: A operand in tensor memory when the plan allows it
  int opt_defer = 2;                 // v5 split pushes with room for a second A buffer in TMEM: 1 = hold the trailing MMAs back until the
                                     // accumulator reads are done (ld_done barrier), 2 = double buffer only (no a_free hand-shake), 0 = a_free
  int opt_park = 0;                  // v5: 1 = the issuer / producer warps wait with a suspend-time hint instead of polling (measured: the stragglers move
                                     // to other sub-partitions, the phase spread stays 150-400 cycles, throughput -1 %: off)
  int opt_trim = 1;                  // folded v5 / v6 programs drop the dead degree-0 accumulator columns
  int opt_gaps = 0;                  // 1: build inverse programs for ladders with unpopulated degrees (coupling layers)
  int opt_align = -1;                // v5 / v6: block-aligned column layout when it fits tensor memory: 1 = on, 0 = off, -1 = auto (on for
                                     // flow layers with >= 4 hidden blocks: cfg2 +13 %; no effect on cfg3 / cfg4 whose blocks are wide)
};

constexpr int kMaxSteps = 80;

struct KParams {
  long long* dbg;          // optional per-step clock stamps of CTA 0 (dev tool), layout [step][8]
  Step steps[kMaxSteps];   // the program, read through the constant bank (uniform loads)
  int nsteps;
  const uint8_t* wimg;
  unsigned long long draw_bytes, layer_bytes;
  const int* perm;
  int D, C, L, M, Mp, K, kind, kin, kin_pad, hp_max, nslots;
  float bound, clip_lo, clip_hi;
  uint32_t off_in, off_h, off_x, off_y, off_xo, off_ctx, off_misc, off_scratch, off_ring;
};

struct KParamsInv {
  long long* dbg;
  Step steps[kMaxSteps];
  int nsteps;
  const uint8_t* wimg;
  unsigned long long draw_bytes, layer_bytes;
  const float* lc;                 // [S][L][lc_floats]
  int lc_floats, lc_b0;
  int phase_delay;
  const int* perm;
  int D, C, L, M, Mp, K, kind, kin, kinp, hp_max, nslots;
  int kr_max;                      // columns of one A block buffer
  int xf;                          // first conditioner layer on tensor cores
  uint32_t off_ax, off_xring;
  int xslot_bytes;
  float bound, clip_lo, clip_hi;
  uint32_t off_xin, off_lc, off_h, off_y, off_xo, off_misc, off_scratch, off_ring;
};

inline int ceil_to(int v, int m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------
// Program construction (host)
// ------------------------------------------------------------------------------------------------
struct Builder {
  std::vector<Step>& steps;
  std::vector<Image>& images;
  uint32_t w_off = 0;
  int slot_bytes = kSlotBytes;
  bool emit = true;   // false: only advance w_off / record the images (steps of a program variant that skips this gemm)
  int k_align = 16;   // K-split sub-steps start at a multiple of this (v5 waits on PAIRS of K slices: 32)
  void gemm(uint8_t a_buf, int a_chunk0, int k_ext, int n_ext, int d_col, int nsplit, int accumulate, Image im,
            Step epi, int n_crit = 0) {
    int k_sub_max = std::min(96, (slot_bytes / (n_ext * 4)) / k_align * k_align);
    for (int k_off = 0; k_off < k_ext; k_off += k_sub_max) {
      int ks = std::min(k_sub_max, k_ext - k_off);
      bool last = (k_off + ks >= k_ext);
      Step s{};
      s.w_off = w_off;
      s.w_bytes = (uint32_t)n_ext * ks * 4;
      s.a_buf = a_buf;
      s.a_chunk0 = (uint16_t)(a_chunk0 + k_off / 8);
      s.ksteps = (uint16_t)(ks / 16);
      s.n = (uint16_t)n_ext;
      s.d_col = (uint16_t)d_col;
      s.nsplit = (uint8_t)nsplit;
      s.accumulate = (uint8_t)((k_off > 0) ? 1 : accumulate);
      s.n_crit = (uint16_t)(n_crit > 0 ? n_crit : n_ext);
      if (last) {
        s.epi = epi.epi; s.stage = epi.stage; s.nranks = epi.nranks; s.flags = epi.flags; s.e_aux = epi.e_aux;
        s.e_col = epi.e_col; s.e_ncols = epi.e_ncols; s.e_dst_chunk = epi.e_dst_chunk;
      }
      Image sub = im;
      sub.w_off = s.w_off; sub.w_bytes = s.w_bytes;
      sub.k_ext = ks; sub.k0 = im.k0 + k_off;
      sub.bias_col = (im.bias_col >= k_off && im.bias_col < k_off + ks) ? im.bias_col - k_off : -1;
      if (emit) steps.push_back(s);
      images.push_back(sub);
      w_off += s.w_bytes;
    }
  }
  void epi_only(Step epi) {
    Step s{};
    s.w_bytes = 0;
    s.epi = epi.epi; s.stage = epi.stage; s.nranks = epi.nranks; s.flags = epi.flags; s.e_aux = epi.e_aux;
    s.e_col = epi.e_col; s.e_ncols = epi.e_ncols; s.e_dst_chunk = epi.e_dst_chunk;
    steps.push_back(s);
  }
};

Step mk_epi(uint8_t kind, int e_col, int e_ncols = 0, int dst_chunk = 0, int stage = 0, int nranks = 0) {
  Step s{};
  s.epi = kind; s.e_col = (uint16_t)e_col; s.e_ncols = (uint16_t)e_ncols; s.e_dst_chunk = (uint16_t)dst_chunk;
  s.stage = (uint8_t)stage; s.nranks = (uint8_t)nranks;
  return s;
}

Image mk_img(int lin, int n_ext, int k_ext, int row_mode, int n0, int r0, int r1, int k0, int kv0, int kv1, int bias_col,
             int bias_only) {
  Image im{};
  im.lin = lin; im.n_ext = n_ext; im.k_ext = k_ext; im.row_mode = row_mode; im.n0 = n0; im.r0 = r0; im.r1 = r1;
  im.k0 = k0; im.kv0 = kv0; im.kv1 = kv1; im.bias_col = bias_col; im.bias_only = bias_only;
  im.scale = 1.f;
  return im;
}

bool base_dims(const FlowGeom& g, TcPlan& P) {
  P.kin_pad = ceil_to(g.kin + 1, 16);
  if (P.kin_pad > 48) return false;
  P.hp_max = 0;
  for (int j = 0; j < g.n_hidden; ++j) {
    int hp = ceil_to(g.hidden[j], 16);
    if (hp > 256) return false;
    P.hp_max = std::max(P.hp_max, hp);
  }
  P.mp = ceil_to(g.M, 16);
  if (g.M > 32) return false;
  return true;
}

bool build_forward(const FlowGeom& g, TcPlan& P) {
  const int nh = g.n_hidden, D = g.D, M = g.M;
  const int T_A = 0, T_OUT = P.hp_max;
  int dims_per_chunk = std::min(D, 256 / M);
  int max_chunk_n = ceil_to(dims_per_chunk * M, 16);
  if (T_OUT + max_chunk_n + 8 > kTmemCols) return false;
  Builder b{P.steps[1], P.images[1]};
  const int kin = g.kin, kp = P.kin_pad;
  auto hp = [&](int j) { return ceil_to(g.hidden[j], 16); };
  b.gemm(A_IN, 0, kp, hp(0), T_A, 3, 0, mk_img(0, hp(0), kp, 0, 0, 0, 0, 0, 0, kin, kin, 0),
         mk_epi(EPI_TANH, T_A, ceil_to(g.hidden[0], 8), 0));
  for (int j = 1; j < nh; ++j) {
    b.gemm(A_IN, 0, kp, hp(j), T_A, 2, 0, mk_img(j, hp(j), kp, 0, 0, 0, 0, 0, 0, 0, kin, 1), mk_epi(EPI_NONE, 0));
    b.gemm(A_H, 0, hp(j - 1), hp(j), T_A, 3, 1, mk_img(j, hp(j), hp(j - 1), 0, 0, 0, 0, 0, 0, g.hidden[j - 1], -1, 0),
           mk_epi(EPI_TANH, T_A, ceil_to(g.hidden[j], 8), 0));
  }
  for (int r0 = 0; r0 < D; r0 += dims_per_chunk) {
    int r1 = std::min(D, r0 + dims_per_chunk);
    int n = ceil_to((r1 - r0) * M, 16);
    b.gemm(A_IN, 0, kp, n, T_OUT, 2, 0, mk_img(nh, n, kp, 2, 0, r0, r1, 0, 0, 0, kin, 1), mk_epi(EPI_NONE, 0));
    b.gemm(A_H, 0, hp(nh - 1), n, T_OUT, 3, 1, mk_img(nh, n, hp(nh - 1), 2, 0, r0, r1, 0, 0, g.hidden[nh - 1], -1, 0),
           mk_epi(EPI_XFWD, T_OUT, 0, 0, r0, r1 - r0));
  }
  P.layer_bytes[1] = b.w_off;
  return (int)P.steps[1].size() <= kMaxSteps;
}

// Forward program for the slice-pipelined kernel (flow_tc_fwd3.cuh): per flow layer
//   XF   [ctx | x | 1] (K = 16) -> pre, tanh;   L_j  h_{j-1} -> pre (other buffer), + bias, tanh;   OUT chunks -> transform.
// Biases of L_j and OUT come from the layer constants; d_col = 0xFFFF means "the pre-activation buffer of this phase".
bool build_forward3(const FlowGeom& g, TcPlan& P) {
  const int nh = g.n_hidden, D = g.D, M = g.M;
  if (g.kin + 1 > 16) return false;
  if (!(g.kind == NAZB_KIND_AFFINE || (g.kind == NAZB_KIND_RQS && g.K == 8))) return false;
  if (P.hp_max > 256) return false;
  const int T_OUT = 2 * P.hp_max;
  if (T_OUT + 16 > kTmemCols) return false;
  int dims_per_chunk = std::min(D, std::min(256, kTmemCols - T_OUT) / M);
  // single output chunk only: the output accumulator is not double-buffered, so a second chunk's MMAs would have to wait
  // for the transform epilogue of the first (the old dense kernel handles those shapes)
  if (dims_per_chunk < D) return false;
  auto hp = [&](int j) { return ceil_to(g.hidden[j], 16); };
  // layer constants: [b_1 * c | ... | b_{nh-1} * c | b_out (rank-major, stride M)]
  int off = 0;
  P.f_lc_b[0] = 0;
  for (int j = 1; j < nh; ++j) { P.f_lc_b[j] = off; off += hp(j); }
  P.f_lc_bout = off; off += ceil_to(D * M, 4);
  P.f_lc_floats = ceil_to(off, 4);
  P.steps[1].clear(); P.images[1].clear();
  Builder b{P.steps[1], P.images[1]};
  b.slot_bytes = kFwdSlotBytes;
  if (kFwdSlotBytes / (P.hp_max * 4) < 16) return false;
  {
    Step t = mk_epi(EPI_TANH, 0, hp(0), 0);
    t.flags = 4;
    Image im = mk_img(0, hp(0), 16, 0, 0, 0, 0, 0, 0, g.kin, g.kin, 0);
    im.scale = 2.885390081777927f;
    b.gemm(A_X, 0, 16, hp(0), 0xFFFF, 3, 0, im, t);
  }
  for (int j = 1; j < nh; ++j) {
    Step t = mk_epi(EPI_TANH, 0, hp(j), 0);
    t.e_aux = (uint16_t)P.f_lc_b[j];
    b.gemm(A_H, 0, hp(j - 1), hp(j), 0xFFFF, 3, 0, mk_img(j, hp(j), hp(j - 1), 0, 0, 0, 0, 0, 0, g.hidden[j - 1], -1, 0), t);
  }
  bool first_chunk = true;
  for (int r0 = 0; r0 < D; r0 += dims_per_chunk) {
    int r1 = std::min(D, r0 + dims_per_chunk);
    int n = ceil_to((r1 - r0) * M, 16);
    Step t = mk_epi(EPI_XFWD, T_OUT, 0, 0, r0, r1 - r0);
    t.e_aux = (uint16_t)(P.f_lc_bout + r0 * M);
    size_t s0 = P.steps[1].size();
    b.gemm(A_H, 0, hp(nh - 1), n, T_OUT, 3, 0, mk_img(nh, n, hp(nh - 1), 2, 0, r0, r1, 0, 0, g.hidden[nh - 1], -1, 0), t);
    if (!first_chunk)
      for (size_t i = s0; i < P.steps[1].size(); ++i) P.steps[1][i].flags |= 8;   // A slices already waited for
    first_chunk = false;
  }
  P.layer_bytes[1] = b.w_off;
  if ((int)P.steps[1].size() > kMaxSteps) return false;
  // shared memory
  uint32_t o = 1024;
  P.f_ax = o;   o += 2u * 16u * kTileM * 2;
  P.f_a = o;    o += 2u * (uint32_t)P.hp_max * kTileM * 2;
  P.f_x = o;    o += (uint32_t)D * kTileM * 4;
  P.f_ctx = o;  o += (uint32_t)std::max(1, g.C) * kTileM * 4;
  P.f_misc = o; o += 4u * kTileM * 4;      // kF3Parts log-det partials
  P.f_lc = o;   o += 2u * (uint32_t)P.f_lc_floats * 4;
  o = (o + 127) & ~127u;
  P.f_ring = o;
  const uint32_t cap = 227 * 1024;
  if (o + 2 * kFwdSlotBytes > cap) return false;
  P.f_nslots = std::min(8u, (cap - o) / kFwdSlotBytes);
  P.f_smem_bytes = o + (size_t)P.f_nslots * kFwdSlotBytes;
  // two-tile variant: each tile owns half of TMEM (pre + transform parameters <= 256 columns)
  P.fwd4 = false;
  if (P.hp_max + ceil_to(D * M, 16) <= kTmemCols / 2) {
    uint32_t q = 1024;
    P.g_ax = q;   q += 2u * 2u * 16u * kTileM * 2;
    P.g_a = q;    q += 2u * (uint32_t)P.hp_max * kTileM * 2;
    P.g_x = q;    q += 2u * (uint32_t)D * kTileM * 4;
    P.g_ctx = q;  q += 2u * (uint32_t)std::max(1, g.C) * kTileM * 4;
    P.g_misc = q; q += 2u * 4u * kTileM * 4;   // [tile][part <= 4] log-det partials
    P.g_lc = q;   q += 2u * (uint32_t)P.f_lc_floats * 4;
    q = (q + 127) & ~127u;
    P.g_ring = q;
    // weight slots: slot j holds sub-step j of the current gemm for BOTH tiles; its size is the largest j-th sub-step
    uint32_t slot_sz[8] = {0};
    int nsub_max = 0, j = 0;
    for (const Step& stp : P.steps[1]) {
      if (j < 8) slot_sz[j] = std::max(slot_sz[j], (stp.w_bytes + 127u) & ~127u);
      nsub_max = std::max(nsub_max, j + 1);
      j = (stp.epi == EPI_NONE) ? j + 1 : 0;
    }
    uint32_t ring_bytes = 0;
    for (int i = 0; i < std::min(nsub_max, 8); ++i) { P.g_slot_off[i] = ring_bytes; ring_bytes += slot_sz[i]; }
    if (nsub_max <= 8 && q + ring_bytes <= cap) {
      P.g_nslots = nsub_max;
      P.g_smem_bytes = q + ring_bytes;
      P.fwd4 = true;
    }
  }
  return true;
}

// Inverse program.  Per flow layer, stage r = 0..D-1 (finalises the dimension of rank r):
//   FIRST(r)   first conditioner layer for the hidden units of degree r -> A operand.  v3: K = 16 MMA over [ctx | x | 1]
//              (or CUDA cores when that slice does not fit); v4: always CUDA cores, by every epilogue warp of the chain
//   PUSH j->j+1 for j = 0..nh-2: pre_{j+1}[cols >= block r] += h_j[block r] . W^T  (tcgen05), epilogue tanh(block r of j+1)
//   PUSH nh-1 -> out: out[ranks >= r] += h_last[block r] . Wout^T, epilogue = inverse transform of rank r
// Biases are added in the epilogues from the layer-constants block, so accumulators need no init pass: the first
// push into each accumulator (stage 0, or stage 1 for unconditional flows) covers its full width with accumulate = 0.
// variant 0: v3 program;  1: v4 general program;  2: v4 context-folded program (stage 0 is constant per draw: its steps
// are replaced by one XINV0C epilogue, its images stay where they are and feed inv4_fold_kernel).
bool build_inverse(const FlowGeom& g, TcPlan& P, int variant, int merge_n, std::vector<Step>& steps_out,
                   std::vector<Image>& images_out, std::vector<Image>* fold_images) {
  if (g.inv_mode != NAZB_INV_INCREMENTAL) return false;
  // variants 3 / 4: the v5 kernel (one 128-row chain, M = 128 MMAs: every N and every accumulator offset a multiple of 16)
  // variants 5 / 6: the v6 kernel (same program as v5; split pushes keep the A operand in tensor memory behind a_free)
  const bool v5 = variant >= 3, v6 = variant >= 5;
  const bool v4 = variant != 0, folded = (variant == 2 || variant == 4 || variant == 6);
  const int nh = g.n_hidden, D = g.D;
  // output columns per rank: v4 / v5 ceil8(M) (v5: a push that does not start at a multiple of 16 columns starts at the
  // 16-aligned column below and carries zero rows for the columns of the already finished rank it overlaps), v3 ceil16(M)
  const int Mp = v4 ? ceil_to(g.M, 8) : P.mp;
  const int out_w = v5 ? ceil_to(D * Mp, 16) : D * Mp;
  P.mp_inv = Mp;
  // v3 feeds [ctx | x | 1] to a K = 16 MMA slice; v4 / v5 compute the first layer on CUDA cores from rank-ordered columns
  if (D * Mp > 256 || (!v4 && g.kin > 16) || D > 16) return false;
  auto hp = [&](int j) { return ceil_to(g.hidden[j], 16); };
  auto hp8 = [&](int j) { return (v4 && !v5) ? ceil_to(g.hidden[j], 8) : hp(j); };   // last column a push has to reach
  // Accumulator columns.  The folded program never touches the columns of degree 0 (constant per draw), so its
  // accumulators start at the first live block (16-aligned) and rank 1: T_PRE / T_OUT are the (possibly negative) column
  // of unit 0 / rank 0, only sums with live offsets are ever used.
  int col = 0;
  int T_PRE[NAZB_MAX_HIDDEN_LAYERS] = {0};
  const bool trim = folded && v5 && P.trim;
  for (int j = 1; j < nh; ++j) {
    const int dead = trim ? (g.blk[j][1] & ~15) : 0;
    T_PRE[j] = col - dead; col += hp(j) - dead;
  }
  const int dead_out = trim ? (v5 ? (Mp & ~15) : Mp) : 0;   // rank 0 is never read by the folded program
  const int T_OUT = col - dead_out; col += out_w - dead_out;
  if (col > kTmemCols) return false;
  const int t_end = col;
  // width of a block as the epilogues see it: the aligned layout pads every block to bw columns of which only aw are live
  auto ep_w = [&](int c0, int c1) { return (P.bw > 0) ? std::min(c1 - c0, P.aw) : (c1 - c0); };
  int xw = 0;
  for (int r = 0; r < D; ++r) xw = std::max(xw, ceil_to(g.blk[0][r + 1], 8) - (g.blk[0][r] & ~7));
  const int T_PRE1 = col;
  if (!v4) {
    // first conditioner layer on tensor cores when TMEM has room for the transient block of pre-activations and
    // [ctx | x | 1] fits one K = 16 slice; otherwise it runs on CUDA cores from the layer constants
    P.xf = (g.kin + 1 <= 16) && (xw <= 128) && (col + ceil_to(xw, 16) <= kTmemCols);
    P.xslot_bytes = ceil_to(xw * 16 * 4, 1024);
    // layer constants: [W0 * c (Hp0 x kinp; CUDA-core first layer only) | b_0 * c .. b_{nh-1} * c | b_out], c = 2 log2 e
    P.kinp = ceil_to(g.kin, 4);
    int off = P.xf ? 0 : hp(0) * P.kinp;
    for (int j = 0; j < nh; ++j) { P.lc_b[j] = off; off += hp(j); }
    P.lc_bout = off; off += D * Mp;
    P.lc_floats = ceil_to(off, 4);
  } else {
    // v4 layer constants: [W0x * c (Hp0 x dp4, x columns BY RANK) | W0c * c (Hp0 x cp4, context columns) |
    //                      b_0 * c .. b_{nh-1} * c | b_out (D x Mp rank-major) | rank-0 constants (32 floats; folded copy only)]
    P.xf = false;
    P.dp4 = ceil_to(D, 4);
    P.cp4 = g.C > 0 ? ceil_to(g.C, 4) : 0;
    int off = 0;
    // v4: unit-major tables W0x[n][dp4], W0c[n][cp4];  v5: input-major tables W0x[q][hp0], W0c[c][hp0] (a thread's 8 units
    // of one input are two 16-byte loads feeding packed FFMA2)
    P.lc_w0x = off; off += v5 ? D * hp(0) : hp(0) * P.dp4;
    P.lc_w0c = off; off += v5 ? g.C * hp(0) : hp(0) * P.cp4;
    for (int j = 0; j < nh; ++j) { P.lc_b[j] = off; off += hp(j); }
    P.lc_bout = off; off += ceil_to(D * Mp, 4);
    P.lc_r0c = off; off += 32;
    P.lc_floats = ceil_to(off, 4);
  }
  for (int r = 0; r < D; ++r) {
    bool empty0 = (g.blk[0][r + 1] == g.blk[0][r]);
    for (int j = 1; j < nh; ++j)
      if ((g.blk[j][r + 1] == g.blk[j][r]) != empty0) return false;   // blocks must be (non)empty together
  }
  // Ladders with unpopulated degrees (beyond degree 0 of a context-free flow) — the single-degree form of a coupling layer,
  // flows/transforms.py::SplineCoupling — need two things this builder used to get wrong: a stage without hidden units
  // reads the output accumulators once ANY earlier stage has pushed into them (flags bit 0 only before the first push), and
  // the last push before such a stage must retire completely before it (unsplit: no later accumulator barrier covers its
  // trailing columns).  `allow_gaps` (engine option "inv_gaps", default off) enables them; off = declined, served by the fp32 engine.
  bool gaps = false;
  for (int r = 0; r < D; ++r) gaps = gaps || ((g.blk[0][r + 1] == g.blk[0][r]) && (r > 0 || g.C > 0));
  if (gaps && !P.allow_gaps) return false;
  if (folded) {
    // needs a context, a non-empty degree-0 block and a transform whose rank-0 parameters fold into a small table
    if (g.C == 0 || g.blk[0][1] == g.blk[0][0]) return false;
    if (!(g.kind == NAZB_KIND_AFFINE || (g.kind == NAZB_KIND_RQS && g.K == 8))) return false;
  }
  Builder b{steps_out, images_out};
  if (v5) b.k_align = 32;
  bool first_push[NAZB_MAX_LIN];
  for (int j = 0; j <= nh; ++j) first_push[j] = true;
  for (int r = 0; r < D; ++r) {
    int b0 = g.blk[0][r], b1 = g.blk[0][r + 1];
    const bool skip = folded && r == 0;   // stage 0 of the folded program: constant, evaluated by inv4_fold_kernel
    b.emit = !skip;
    if (b1 == b0) {
      Step e = mk_epi(EPI_XINV, T_OUT + r * Mp, 0, 0, r);
      e.flags = first_push[nh] ? 1 : 0; e.e_aux = (uint16_t)(P.lc_bout + r * Mp);
      b.epi_only(e);
      continue;
    }
    {
      int ec0 = b0 & ~7, ec1 = ec0 + ep_w(b0 & ~7, ceil_to(b1, 8));
      Step e = mk_epi(EPI_FIRST, 0, ec1 - ec0, 0, r);
      e.e_aux = (uint16_t)ec0;
      if (!skip) b.epi_only(e);
      if (P.xf) {
        Step t = mk_epi(EPI_TANH, T_PRE1, ec1 - ec0, 0);
        t.flags = 4;
        Image im = mk_img(0, ec1 - ec0, 16, 0, ec0, 0, 0, 0, 0, g.kin, g.kin, 0);
        im.scale = 2.885390081777927f;
        b.gemm(A_X, 0, 16, ec1 - ec0, T_PRE1, 3, 0, im, t, ec1 - ec0);
      }
    }
    for (int j = 0; j < nh; ++j) {
      int sb0 = g.blk[j][r], sb1 = g.blk[j][r + 1];
      int sc0 = sb0 & ~7, sc1 = ceil_to(sb1, 8);
      int kr = ceil_to(sc1 - sc0, 16);
      if (kr > P.hp_max) return false;
      P.kr_max = std::max(P.kr_max, kr);
      const size_t img0 = images_out.size();
      if (j + 1 < nh) {
        int tb0 = g.blk[j + 1][r], tb1 = g.blk[j + 1][r + 1];
        int tc0 = tb0 & ~7, tc1 = ceil_to(tb1, 8);
        int tn0 = v5 ? (tc0 & ~15) : tc0;      // M = 64 MMAs take any N % 8 == 0, M = 128 ones N % 16 == 0
        int n = hp8(j + 1) - tn0;
        if (first_push[j + 1] && tn0 != 0 && !(folded && r == 1)) return false;
        Step e = mk_epi(EPI_TANH, T_PRE[j + 1] + tc0, ep_w(tc0, tc1), 0);
        e.e_aux = (uint16_t)(P.lc_b[j + 1] + tc0);
        int n_crit = v5 ? ceil_to(tc1, 16) - tn0 : tc1 - tc0;
        if (v4 && n <= merge_n) n_crit = n;
        b.gemm(A_H, 0, kr, n, T_PRE[j + 1] + tn0, 3, first_push[j + 1] ? 0 : 1,
               mk_img(j + 1, n, kr, 0, tn0, 0, 0, sc0, sb0, sb1, -1, 0), e, n_crit);
        if (!skip) first_push[j + 1] = false;
      } else {
        const int s0 = v5 ? ((r * Mp) & ~15) : r * Mp;          // first accumulator column of this push (M = 128: multiple of 16)
        int n = out_w - s0;
        Step e = mk_epi(EPI_XINV, T_OUT + r * Mp, 0, 0, r);
        e.e_aux = (uint16_t)(P.lc_bout + r * Mp);
        int n_crit = v5 ? ceil_to((r + 1) * Mp, 16) - s0 : Mp;
        if (v4 && n <= merge_n) n_crit = n;
        if (gaps) n_crit = n;   // a stage without hidden units may follow: it reads columns no later barrier would cover
        // image row n <-> output column s0 + n (rank = column / Mp, slot = column % Mp); rows of ranks < r and of the padding are zero
        b.gemm(A_H, 0, kr, n, T_OUT + s0, 3, first_push[nh] ? 0 : 1, mk_img(nh, n, kr, 1, s0, r, D, sc0, sb0, sb1, -1, 0), e, n_crit);
        if (!skip) first_push[nh] = false;
      }
      if (fold_images && r == 0)
        for (size_t i = img0; i < images_out.size(); ++i) fold_images->push_back(images_out[i]);
    }
    if (skip) {
      b.emit = true;
      b.epi_only(mk_epi(EPI_XINV0C, 0, 0, 0, 0));
    }
  }
  P.layer_bytes[0] = b.w_off;
  if (v5) {
    // A operand in tensor memory: one buffer of kr_max / 2 columns each for the hi and lo images behind the accumulators;
    // needs unsplit pushes (every MMA of a push retires before the accumulator barrier that releases the next writer)
    P.t_a = ceil_to(t_end, 16);
    bool unsplit = true;
    for (const Step& st : steps_out) if (st.w_bytes && st.n_crit != st.n) unsplit = false;
    (void)v6;   // split pushes keep the A operand in tensor memory too (a_free barrier in the kernels)
    P.split = !unsplit;
    P.a_tmem2 = (P.t_a + 2 * P.kr_max <= kTmemCols);
    P.a_tmem = (P.t_a + P.kr_max <= kTmemCols);
  }
  return (int)steps_out.size() <= kMaxSteps;
}

bool plan_smem(const FlowGeom& g, TcPlan& P) {
  uint32_t off = 1024;                              // barriers + misc
  P.off_in = off;      off += (uint32_t)P.kin_pad * kTileM * 2 * 2;          // hi then lo
  P.off_h = off;       off += (uint32_t)P.hp_max * kTileM * 2 * 2;
  P.off_x = off;       off += (uint32_t)g.D * kTileM * 4;
  P.off_y = off;       off += (uint32_t)g.D * kTileM * 4;
  P.off_xo = off;      off += (uint32_t)g.D * kTileM * 4;
  P.off_ctx = off;     off += (uint32_t)std::max(1, g.C) * kTileM * 4;
  P.off_misc = off;    off += (1 + kParts) * kTileM * 4;                       // ljac, ld partials
  P.off_scratch = off; off += (g.kind == NAZB_KIND_AFFINE || (g.kind == NAZB_KIND_RQS && g.K == 8)) ? 0 : (uint32_t)kParts * 32 * kTileM * 4;
  off = (off + 127) & ~127u;
  P.off_ring = off;
  const uint32_t cap = 227 * 1024;
  if (off + 2 * kSlotBytes > cap) return false;
  P.nslots = std::min(6u, (cap - off) / kSlotBytes);
  P.smem_bytes = off + (size_t)P.nslots * kSlotBytes;
  return true;
}

bool plan_smem_inv3(const FlowGeom& g, TcPlan& P) {
  if (P.kr_max <= 0 || P.kr_max > 128) return false;      // <= kV3MaxSlices K slices per A block
  uint32_t off = 1024;
  P.j_xin = off;     off += (uint32_t)ceil_to(g.kin, 4) * kTileM * 4;
  P.j_lc = off;      off += 2u * (uint32_t)P.lc_floats * 4;
  off = (off + 127) & ~127u;
  P.j_a = off;       off += 2u * 2u * (uint32_t)P.kr_max * (kTileM / 2) * 2 * 2;   // [chain][buffer][hi | lo]
  P.j_ax = off;      off += 2u * 2u * 16u * (kTileM / 2) * 2;                      // [chain][hi | lo] one K = 16 slice
  P.j_xring = off;   off += P.xf ? 4u * (uint32_t)P.xslot_bytes : 0u;                  // kXSlots small-ring slots
  P.j_y = off;       off += (uint32_t)g.D * kTileM * 4;
  P.j_xo = off;      off += (uint32_t)g.D * kTileM * 4;
  P.j_misc = off;    off += kTileM * 4;
  P.j_scratch = off; off += (g.kind == NAZB_KIND_AFFINE || (g.kind == NAZB_KIND_RQS && g.K == 8)) ? 0 : 32u * kTileM * 4;
  off = (off + 127) & ~127u;
  P.j_ring = off;
  const uint32_t cap = 227 * 1024;
  if (off + 2 * kSlotBytes > cap) return false;
  P.j_nslots = std::min(6u, (cap - off) / kSlotBytes);
  P.j_smem_bytes = off + (size_t)P.j_nslots * kSlotBytes;
  return true;
}

bool plan_smem_inv4(const FlowGeom& g, TcPlan& P) {
  if (P.kr_max <= 0 || P.kr_max > 128) return false;      // <= kV4MaxSlices K slices per A block
  uint32_t off = 1024;
  P.j_xin = off;     off += (uint32_t)std::max(1, g.C) * kTileM * 4;
  P.j_lc = off;      off += 2u * (uint32_t)P.lc_floats * 4;
  off = (off + 127) & ~127u;
  P.j_a = off;       off += 2u * 2u * (uint32_t)P.kr_max * (kTileM / 2) * 2 * 2;   // [chain][buffer][hi | lo]
  P.j_y = off;       off += (uint32_t)g.D * kTileM * 4;
  P.j_xo = off;      off += (uint32_t)g.D * kTileM * 4;
  P.j_xr = off;      off += (uint32_t)g.D * kTileM * 4;
  P.j_misc = off;    off += kTileM * 4;
  P.j_scratch = off; off += (g.kind == NAZB_KIND_AFFINE || (g.kind == NAZB_KIND_RQS && g.K == 8)) ? 0 : 32u * kTileM * 4;
  off = (off + 127) & ~127u;
  P.j_ring = off;
  const uint32_t cap = 227 * 1024;
  if (off + 2 * kSlotBytes > cap) return false;
  P.j_nslots = std::min(6u, (cap - off) / kSlotBytes);
  P.j_smem_bytes = off + (size_t)P.j_nslots * kSlotBytes;
  return true;
}

// ------------------------------------------------------------------------------------------------
// Pack kernel: one 16-byte K-chunk (8 fp16) of the hi image and of the lo image per thread.
// ------------------------------------------------------------------------------------------------
struct PackGeom {
  int S, L, n_lin, D, M, Mp;
  int kdim[NAZB_MAX_LIN], ndim[NAZB_MAX_LIN];
  long long wst[NAZB_MAX_LIN * 1], bst[NAZB_MAX_LIN * 1];   // unused placeholders (strides come per pointer table)
};

__global__ void tc_pack_kernel(Image im, int S, int L, int n_lin, int D, int M, int Mp, int kdim, int ndim,
                               const float* const* __restrict__ Wtab, const float* const* __restrict__ btab,
                               const float* const* __restrict__ mtab, const long long* __restrict__ wst,
                               const long long* __restrict__ bst, const int* __restrict__ perm,
                               const float* __restrict__ keep, long long keep_draw_stride, long long keep_layer_stride,
                               int keep_hk, float inv_keep, uint8_t* __restrict__ dst, unsigned long long draw_bytes,
                               unsigned long long layer_bytes, const float* const* __restrict__ bWtab,
                               const float* const* __restrict__ bbtab, float dm_scale, const short* __restrict__ pmap) {
  // pmap (block-aligned inverse layout): [hidden layer][256] column -> hidden unit, -1 = padding column; null = identity
  const short* map_out = (pmap && im.row_mode == 0) ? pmap + (size_t)im.lin * 256 : nullptr;
  const short* map_in = (pmap && im.lin > 0) ? pmap + (size_t)(im.lin - 1) * 256 : nullptr;
  const int KC = im.k_ext >> 3;
  const long long per_layer = (long long)KC * im.n_ext;
  const long long total = per_layer * L * S;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int n = (int)(idx % im.n_ext);
    int kc = (int)((idx / im.n_ext) % KC);
    int l = (int)((idx / per_layer) % L);
    int s = (int)(idx / (per_layer * L));
    // image row -> source output unit
    int o = -1;
    if (im.row_mode == 0) {
      int u = im.n0 + n;
      if (map_out) u = (u < 256) ? map_out[u] : -1;
      if (u >= 0 && u < ndim) o = u;
    } else if (im.row_mode == 1) {
      // n0 = output column of image row 0; ranks below r0 (columns a 16-aligned start overlaps) and padding rows stay zero
      const int colo = im.n0 + n;
      int rank = colo / Mp, m = colo % Mp;
      if (rank >= im.r0 && rank < im.r1 && m < M) o = m * D + perm[l * D + rank];
    } else {
      int rank = im.r0 + n / M, m = n % M;
      if (rank < im.r1) o = m * D + perm[l * D + rank];
    }
    const int ti = l * n_lin + im.lin;
    const float* W = Wtab[ti] + (size_t)s * wst[ti];
    const float* bb = btab[ti] + (size_t)s * bst[ti];
    const float* mk = mtab[ti];
    const float* bW = bWtab ? bWtab[ti] : nullptr;     // draw map: W / bb hold the standard parameters u_s
    const float* bB = bbtab ? bbtab[ti] : nullptr;
    const float* kp = (keep && im.lin > 0) ? keep + (size_t)s * keep_draw_stride + (size_t)l * keep_layer_stride +
                                                 (size_t)(im.lin - 1) * keep_hk
                                           : nullptr;
    __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int c = kc * 8 + e;
      float v = 0.f;
      if (o >= 0) {
        if (c == im.bias_col) { v = bb[o]; if (bB) v = nazb_draw_map(bB[o], v, dm_scale); }
        else if (!im.bias_only) {
          int ks = im.k0 + c;
          const bool in_range = (ks >= im.kv0 && ks < im.kv1);
          if (map_in) ks = (ks >= 0 && ks < 256) ? map_in[ks] : -1;
          if (in_range && ks >= 0 && ks < kdim) {
            v = W[(size_t)o * kdim + ks];
            if (bW) v = nazb_draw_map(bW[(size_t)o * kdim + ks], v, dm_scale);
            v *= mk[(size_t)o * kdim + ks];
            if (kp) v *= kp[ks] * inv_keep;
          }
        }
      }
      v = fminf(fmaxf(v * im.scale, -65504.f), 65504.f);
      hi[e] = __float2half_rn(v);
      lo[e] = __float2half_rn(v - __half2float(hi[e]));
    }
    uint8_t* base = dst + (size_t)s * draw_bytes + (size_t)l * layer_bytes + im.w_off;
    size_t eo = ((size_t)kc * im.n_ext + n) * 16;
    *reinterpret_cast<uint4*>(base + eo) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(base + (im.w_bytes >> 1) + eo) = *reinterpret_cast<uint4*>(lo);
  }
}

// ------------------------------------------------------------------------------------------------
// Main kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }
#define DBG(slot)                                                                       \
  if (p.dbg && blockIdx.x == 0 && dbg_i < 256) p.dbg[dbg_i * 8 + (slot)] = clk();

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory"); }

// 8 consecutive accumulator columns -> tanh -> fp16 hi / lo chunks
__device__ __forceinline__ void tanh_chunk(const uint32_t* r, uint4& hi4, uint4& lo4) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = tcx::tanh_fast(__uint_as_float(r[i]));
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = tcx::pack_hi2(v[2 * i], v[2 * i + 1]);
    float a, b;
    tcx::unpack2(h[i], a, b);
    l[i] = tcx::pack_hi2(v[2 * i] - a, v[2 * i + 1] - b);
  }
  hi4 = make_uint4(h[0], h[1], h[2], h[3]);
  lo4 = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void store_split(__half* hi_base, __half* lo_base, int col, int row, float v) {
  // element (row, col) of an A operand buffer laid out [col / 8][row][col % 8]
  float c = fminf(fmaxf(v, -65504.f), 65504.f);
  __half h = __float2half_rn(c);
  __half l = __float2half_rn(c - __half2float(h));
  size_t o = ((size_t)(col >> 3) * kTileM + row) * 8 + (col & 7);
  hi_base[o] = h;
  lo_base[o] = l;
}

__global__ void __launch_bounds__(kThreads, 1) flow_tc_kernel(const __grid_constant__ KParams p, const __grid_constant__ IoArgs io,
                                                               int n_groups) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);          // [nslots]
  uint64_t* bar_empty = bar_full + 8;                               // [nslots]
  uint64_t* bar_acc = bar_empty + 8;                                // MMA -> epilogue
  uint64_t* bar_a = bar_acc + 1;                                    // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a + 1);
  __half* in_hi = reinterpret_cast<__half*>(smem + p.off_in);
  __half* in_lo = in_hi + (size_t)p.kin_pad * kTileM;
  __half* h_hi = reinterpret_cast<__half*>(smem + p.off_h);
  __half* h_lo = h_hi + (size_t)p.hp_max * kTileM;
  float* xcur = reinterpret_cast<float*>(smem + p.off_x);           // [D][128]
  float* ycur = reinterpret_cast<float*>(smem + p.off_y);           // [D][128]
  float* xorig = reinterpret_cast<float*>(smem + p.off_xo);         // [D][128] tile input (after bounding)
  float* ctxs = reinterpret_cast<float*>(smem + p.off_ctx);         // [C][128]
  float* ljac = reinterpret_cast<float*>(smem + p.off_misc);        // [128]
  float* ldpart = ljac + kTileM;                                    // [kParts][128]
  float* scratch = reinterpret_cast<float*>(smem + p.off_scratch);  // [kParts][32][128]
  uint8_t* ring = smem + p.off_ring;

  const int tid = threadIdx.x, lane = tid & 31;
  // broadcast so the compiler can prove the role index warp-uniform (role branches stay convergent and
  // warp-uniform values are eligible for uniform registers)
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int D = p.D, C = p.C, M = p.M;
  const bool inverse = io.dir == 0;

  if (tid == 0) {
    for (int i = 0; i < p.nslots; ++i) { tcx::mbar_init(bar_full + i, 1); tcx::mbar_init(bar_empty + i, 1); }
    tcx::mbar_init(bar_acc, 1);
    tcx::mbar_init(bar_a, kEpiWarps);
    tcx::mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tcx::tmem_alloc(tmem_slot, kTmemCols);
  // zero the A operand buffers once
  for (uint32_t i = tid; i < ((uint32_t)(p.kin_pad + p.hp_max) * kTileM * 4) / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem + p.off_in)[i] = make_uint4(0, 0, 0, 0);
  tcx::fence_async_smem();
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n_tiles = (io.N + kTileM - 1) / kTileM;
  const long long n_items = (long long)n_tiles * n_groups;

  if (warp == kEpiWarps) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = (int)(item / n_tiles);
        for (int si = grp; si < io.s_count; si += n_groups) {
          const uint8_t* wdraw = p.wimg + (size_t)(io.s_begin + si) * p.draw_bytes;
          for (int li = 0; li < p.L; ++li) {
            const int l = inverse ? (p.L - 1 - li) : li;
            const uint8_t* wl = wdraw + (size_t)l * p.layer_bytes;
            for (int st = 0; st < p.nsteps; ++st) {
              const uint32_t wb = p.steps[st].w_bytes;
              if (wb == 0) continue;
              const uint32_t slot = cnt % p.nslots, use = cnt / p.nslots;
              tcx::mbar_wait(bar_empty + slot, (use & 1) ^ 1);
              tcx::mbar_expect_tx(bar_full + slot, wb);
              tcx::bulk_g2s(ring + (size_t)slot * kSlotBytes, wl + p.steps[st].w_off, wb, bar_full + slot);
              ++cnt;
            }
          }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop convergently on warp-uniform values (program in constant space), so the
    // descriptors live in uniform registers; only the tcgen05 instructions are predicated on the elected lane.
    const uint32_t elected = tcx::elect_one();
    uint32_t slot = 0, use = 0, par_a = 0;
    int dbg_i = 0;
    const uint32_t in_hi_a = tcx::smem_u32(in_hi), in_lo_a = tcx::smem_u32(in_lo);
    const uint32_t h_hi_a = tcx::smem_u32(h_hi), h_lo_a = tcx::smem_u32(h_lo);
    const uint32_t ring_a = tcx::smem_u32(ring);
    constexpr uint32_t lbo_a = kTileM * 16;
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);            // SBO = 128 B, descriptor version 1
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int grp = (int)(item / n_tiles);
      for (int si = grp; si < io.s_count; si += n_groups) {
        bool need_a = true;   // draw start: wait for the epilogue warps to stage `in`
        for (int li = 0; li < p.L; ++li) {
          for (int st = 0; st < p.nsteps; ++st) {
            // hoist every field of the step out of constant space BEFORE the waits (the asm "memory" clobbers
            // would otherwise force indexed constant re-loads between MMAs)
            const uint32_t s_wbytes = p.steps[st].w_bytes, s_n = p.steps[st].n, s_dcol = p.steps[st].d_col;
            const uint32_t s_chunk0 = p.steps[st].a_chunk0, s_abuf = p.steps[st].a_buf, s_nsplit = p.steps[st].nsplit;
            const int ksteps = p.steps[st].ksteps;
            const uint32_t s_acc = p.steps[st].accumulate, s_epi = p.steps[st].epi;
            const uint32_t idesc = tcx::make_idesc_f16(s_n);
            const uint32_t lbo_b = s_n * 16;
            const uint32_t a_off = s_chunk0 * lbo_a;
            const uint32_t a_hi = (s_abuf == A_IN ? in_hi_a : h_hi_a) + a_off;
            const uint32_t a_lo = (s_abuf == A_IN ? in_lo_a : h_lo_a) + a_off;
            const uint32_t b_hi = ring_a + slot * kSlotBytes, b_lo = b_hi + (s_wbytes >> 1);
            // low descriptor words: (addr >> 4) | (LBO >> 4) << 16; one k-step advances the address by 2 LBO
            const uint32_t da_hi0 = (a_hi >> 4) | ((lbo_a >> 4) << 16), da_lo0 = (a_lo >> 4) | ((lbo_a >> 4) << 16);
            const uint32_t db_hi0 = (b_hi >> 4) | ((lbo_b >> 4) << 16), db_lo0 = (b_lo >> 4) | ((lbo_b >> 4) << 16);
            const uint32_t da_step = (2 * lbo_a) >> 4, db_step = (2 * lbo_b) >> 4;
            const uint32_t d_addr = tmem + s_dcol;
            constexpr uint64_t dhi = (uint64_t)desc_hi << 32;
            if (elected) { DBG(0) }
            if (need_a) {
              tcx::mbar_wait(bar_a, par_a);
              par_a ^= 1;
              need_a = false;
            }
            if (s_wbytes) {
              tcx::mbar_wait(bar_full + slot, use & 1);
              tcx::tc_fence_after();
              if (elected) { DBG(2) }
              uint32_t da = da_hi0, db = db_hi0;
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k) {            // a_hi * w_hi
                tcx::mma_f16_ss_elect(d_addr, dhi | da, dhi | db, idesc, (k == 0) ? s_acc : 1u, elected);
                da += da_step; db += db_step;
              }
              if (elected) { DBG(1) }
              da = da_hi0; db = db_lo0;
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k) {            // a_hi * w_lo
                tcx::mma_f16_ss_elect(d_addr, dhi | da, dhi | db, idesc, 1u, elected);
                da += da_step; db += db_step;
              }
              if (s_nsplit == 3) {
                da = da_lo0; db = db_hi0;
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k) {          // a_lo * w_hi
                  tcx::mma_f16_ss_elect(d_addr, dhi | da, dhi | db, idesc, 1u, elected);
                  da += da_step; db += db_step;
                }
              }
              tcx::mma_commit_elect(bar_empty + slot, elected);   // weights slot is free once these MMAs retire
              if (++slot == (uint32_t)p.nslots) { slot = 0; ++use; }
            }
            if (elected) { DBG(3) }
            ++dbg_i;
            if (s_epi != EPI_NONE) {
              tcx::mma_commit_elect(bar_acc, elected);
              // the next step reads what this epilogue writes, except after the last step of a draw
              need_a = !(li == p.L - 1 && st == p.nsteps - 1);
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, half = warp >> 2;   // `half` = column part 0..kParts-1
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t par_acc = 0;
    int dbg_i = 0;
    const bool dbg_me = (tid == 0);
    const bool spline = p.kind != NAZB_KIND_AFFINE;
    const bool fast_rqs = (p.kind == NAZB_KIND_RQS && p.K == 8);
    float* scr = scratch + (size_t)half * 32 * kTileM + row;   // [m * 128]
    auto raw = [&](int m) { return scr[m * kTileM]; };
    auto setw = [&](int m, float v) { scr[m * kTileM] = v; };

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = (int)(item % n_tiles), grp = (int)(item / n_tiles);
      const int n0 = tile * kTileM;
      const int npts = min(kTileM, io.N - n0);
      float run_m = -INFINITY, run_s = 0.f;
      // ---- tile load (shared across this item's draws) ----
      epi_bar_sync();   // previous item's readers are done with xorig / ctxs
      for (int i = tid; i < kTileM * C; i += kEpiThreads) {
        int pt = i / C, c = i % C;
        float v = 0.f;
        if (pt < npts) v = io.ctx[((io.ctx_rows == 1) ? 0 : (size_t)(n0 + pt)) * C + c];
        ctxs[c * kTileM + pt] = v;
      }
      if (inverse || io.x_draw_stride == 0) {
        for (int i = tid; i < kTileM * D; i += kEpiThreads) {
          int pt = i / D, d = i % D;
          xorig[d * kTileM + pt] = (pt < npts) ? io.x[(size_t)(n0 + pt) * D + d] : 0.f;
        }
      }
      epi_bar_sync();
      if (inverse && half == 0) {
        float lj = 0.f;
        if (io.lo != nullptr && row < npts)
          for (int d = 0; d < D; ++d) xorig[d * kTileM + row] = nazb::bound_fwd(xorig[d * kTileM + row], io.lo[d], io.hi[d], lj);
        ljac[row] = lj;
      }

      for (int si = grp; si < io.s_count; si += n_groups) {
        // ---- draw start: stage `in`, reset state ----
        if (!inverse && io.x_draw_stride != 0) {
          epi_bar_sync();
          const float* zs = io.x + (size_t)si * io.x_draw_stride;
          for (int i = tid; i < kTileM * D; i += kEpiThreads) {
            int pt = i / D, d = i % D;
            xorig[d * kTileM + pt] = (pt < npts) ? zs[(size_t)(n0 + pt) * D + d] : 0.f;
          }
          epi_bar_sync();
        }
        for (uint32_t i = tid; i < ((uint32_t)p.hp_max * kTileM * 4) / 16; i += kEpiThreads)
          reinterpret_cast<uint4*>(smem + p.off_h)[i] = make_uint4(0, 0, 0, 0);
        float ld_acc = 0.f;
        if (half == 0) {
          for (int c = 0; c < p.kin_pad; ++c) {
            float v = 0.f;
            if (c < C) v = ctxs[c * kTileM + row];
            else if (c < C + D) v = inverse ? 0.f : xorig[(c - C) * kTileM + row];
            else if (c == p.kin) v = 1.f;
            store_split(in_hi, in_lo, c, row, v);
          }
          for (int d = 0; d < D; ++d) {
            float v = xorig[d * kTileM + row];
            if (inverse) ycur[d * kTileM + row] = v; else xcur[d * kTileM + row] = v;
          }
        }
        tcx::fence_async_smem();
        __syncwarp();
        if (lane == 0) tcx::mbar_arrive(bar_a);

        for (int li = 0; li < p.L; ++li) {
          const int l = inverse ? (p.L - 1 - li) : li;
          const int* perm = p.perm + l * D;
          for (int st = 0; st < p.nsteps; ++st) {
            // copy the step's epilogue fields out of constant space before parking on the barrier
            struct { uint32_t epi, e_col, e_ncols, e_dst_chunk, stage, nranks; } s;
            s.epi = p.steps[st].epi; s.e_col = p.steps[st].e_col; s.e_ncols = p.steps[st].e_ncols;
            s.e_dst_chunk = p.steps[st].e_dst_chunk; s.stage = p.steps[st].stage; s.nranks = p.steps[st].nranks;
            if (s.epi == EPI_NONE) { ++dbg_i; continue; }
            tcx::mbar_wait(bar_acc, par_acc);
            par_acc ^= 1;
            tcx::tc_fence_after();
            if (dbg_me) { DBG(5) }
            if (s.epi == EPI_TANH) {
              const int nchunks = s.e_ncols >> 3;
              const int per = (nchunks + kParts - 1) / kParts;
              const int cb = min(nchunks, half * per), ce = min(nchunks, cb + per);
              // up to 4 chunks (32 columns) per batch: all TMEM loads in flight before the single wait
              for (int c = cb; c < ce; c += 4) {
                uint32_t r[32];
                const int nb = min(4, ce - c);
                const uint32_t ta = lane_base + s.e_col + c * 8;
                if (nb == 4) tcx::tmem_ld32(ta, r);
                else if (nb == 3) { tcx::tmem_ld16(ta, r); tcx::tmem_ld8(ta + 16, r + 16); }
                else if (nb == 2) tcx::tmem_ld16(ta, r);
                else tcx::tmem_ld8(ta, r);
                tcx::tmem_ld_wait();
                if (dbg_me) { DBG(4) }
                size_t o = ((size_t)(s.e_dst_chunk + c) * kTileM + row) * 8;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (u < nb) {
                    uint4 hi4, lo4;
                    tanh_chunk(r + 8 * u, hi4, lo4);
                    *reinterpret_cast<uint4*>(h_hi + o) = hi4;
                    *reinterpret_cast<uint4*>(h_lo + o) = lo4;
                    o += (size_t)kTileM * 8;
                  }
                }
              }
            } else if (s.epi == EPI_XINV) {
              if (half == 0) {
                const int r = s.stage, d = perm[r];
                const float yv = ycur[d * kTileM + row];
                float xv, ld;
                if (!spline) {
                  uint32_t rr[2];
                  tcx::tmem_ld2(lane_base + s.e_col, rr);
                  tcx::tmem_ld_wait();
                  float mu = __uint_as_float(rr[0]);
                  float sc = fminf(fmaxf(__uint_as_float(rr[1]), p.clip_lo), p.clip_hi);
                  xv = (yv - mu) * expf(-sc);
                  ld = sc;
                } else if (fast_rqs) {
                  uint32_t rr[24];
                  tcx::tmem_ld16(lane_base + s.e_col, rr);
                  tcx::tmem_ld8(lane_base + s.e_col + 16, rr + 16);
                  tcx::tmem_ld_wait();
                  float rf[24];
#pragma unroll
                  for (int e = 0; e < 24; ++e) rf[e] = __uint_as_float(rr[e]);
                  nazb::rqs_fast<8>(yv, p.bound, true, rf, xv, ld);
                } else {
                  for (int m0 = 0; m0 < p.Mp; m0 += 16) {
                    uint32_t rr[16];
                    tcx::tmem_ld16(lane_base + s.e_col + m0, rr);
                    tcx::tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                      if (m0 + e < M) scr[(m0 + e) * kTileM] = __uint_as_float(rr[e]);
                  }
                  if (p.kind == NAZB_KIND_RQS) nazb::rational_spline<false>(yv, p.K, p.bound, true, raw, setw, xv, ld);
                  else nazb::rational_spline<true>(yv, p.K, p.bound, true, raw, setw, xv, ld);
                }
                ld_acc += ld;
                xcur[d * kTileM + row] = xv;
                if (r == D - 1) {
                  // end of this flow layer: x becomes the y of the next (earlier) layer, x restarts at 0
                  for (int dd = 0; dd < D; ++dd) {
                    ycur[dd * kTileM + row] = xcur[dd * kTileM + row];
                    store_split(in_hi, in_lo, C + dd, row, 0.f);
                  }
                } else {
                  store_split(in_hi, in_lo, C + d, row, xv);
                }
              }
            } else {   // EPI_XFWD: transform the dims of ranks [stage, stage + nranks)
              for (int i = half; i < (int)s.nranks; i += kParts) {
                const int rr_ = s.stage + i, d = perm[rr_];
                const float xv = xcur[d * kTileM + row];
                float yv, ld;
                if (!spline) {
                  uint32_t rr[2];
                  tcx::tmem_ld2(lane_base + s.e_col + i * 2, rr);
                  tcx::tmem_ld_wait();
                  float mu = __uint_as_float(rr[0]);
                  float sc = fminf(fmaxf(__uint_as_float(rr[1]), p.clip_lo), p.clip_hi);
                  yv = mu + xv * expf(sc);
                  ld = sc;
                } else if (fast_rqs) {
                  uint32_t rr[24];
                  const uint32_t ta = lane_base + s.e_col + i * 23;
                  tcx::tmem_ld8(ta, rr);
                  tcx::tmem_ld8(ta + 8, rr + 8);
                  tcx::tmem_ld8(ta + 16, rr + 16);
                  tcx::tmem_ld_wait();
                  float rf[24];
#pragma unroll
                  for (int e = 0; e < 24; ++e) rf[e] = __uint_as_float(rr[e]);
                  nazb::rqs_fast<8>(xv, p.bound, false, rf, yv, ld);
                } else {
                  for (int m0 = 0; m0 < M; m0 += 8) {
                    uint32_t rr[8];
                    tcx::tmem_ld8(lane_base + s.e_col + i * M + m0, rr);
                    tcx::tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                      if (m0 + e < M) scr[(m0 + e) * kTileM] = __uint_as_float(rr[e]);
                  }
                  if (p.kind == NAZB_KIND_RQS) nazb::rational_spline<false>(xv, p.K, p.bound, false, raw, setw, yv, ld);
                  else nazb::rational_spline<true>(xv, p.K, p.bound, false, raw, setw, yv, ld);
                }
                ld_acc += ld;
                xcur[d * kTileM + row] = yv;
                store_split(in_hi, in_lo, C + d, row, yv);
              }
            }
            const bool last_of_draw = (li == p.L - 1 && st == p.nsteps - 1);
            if (dbg_me) { DBG(6) }
            if (!last_of_draw) {
              tcx::tc_fence_before();
              tcx::fence_async_smem();
              __syncwarp();
              if (lane == 0) tcx::mbar_arrive(bar_a);
            }
            if (dbg_me) { DBG(7) }
            ++dbg_i;
          }
        }

        // ---- draw end ----
        if (inverse) {
          if (half == 0) {
            float qd = 0.f;
            for (int d = 0; d < D; ++d) { float z = ycur[d * kTileM + row]; qd += 0.5f * z * z; }
            float lp = -qd - 0.5f * D * NAZB_LOG_2PI - ld_acc + ljac[row];
            if (row < npts) {
              if (io.out_l) io.out_l[(size_t)si * io.N + n0 + row] = lp;
              if (io.lse_max) {
                float v = lp + (io.log_w ? io.log_w[si] : 0.f);
                if (!(v <= run_m)) { run_s = run_s * expf(run_m - v) + 1.f; run_m = v; }
                else if (v > -INFINITY) run_s += expf(v - run_m);
              }
              if (io.out_x) {
                float* dst = io.out_x + ((size_t)si * io.N + n0 + row) * D;
                for (int d = 0; d < D; ++d) dst[d] = ycur[d * kTileM + row];
              }
            }
            if (io.sum_n) {
              double v = (row < npts) ? (double)lp : 0.0;
              for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
              if (lane == 0) atomicAdd(io.sum_n + si, v);
            }
          }
        } else {
          ldpart[half * kTileM + row] = ld_acc;
          epi_bar_sync();
          if (half == 0 && row < npts) {
            if (io.out_l) {
              float a = 0.f;
              for (int pp = 0; pp < kParts; ++pp) a += ldpart[pp * kTileM + row];
              io.out_l[(size_t)si * io.N + n0 + row] = a;
            }
            float* dst = io.out_x + ((size_t)si * io.N + n0 + row) * D;
            for (int d = 0; d < D; ++d) {
              float v = xcur[d * kTileM + row];
              if (io.lo != nullptr) v = nazb::bound_inv(v, io.lo[d], io.hi[d]);
              dst[d] = v;
            }
          }
        }
      }
      if (inverse && io.lse_max && half == 0 && row < npts) {
        io.lse_max[(size_t)grp * io.N + n0 + row] = run_m;
        io.lse_sum[(size_t)grp * io.N + n0 + row] = run_s;
      }
    }
  }
  tcx::tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) tcx::tmem_dealloc(tmem, kTmemCols);
}



// ------------------------------------------------------------------------------------------------
// Layer-constants pack kernel (inverse): fp32 [W0 masked * c (Hp0 x kinp; absent when the first layer runs on tensor
// cores) | hidden biases * c | output bias rank-major], c = 2 log2 e (the tanh epilogues take scaled pre-activations)
// ------------------------------------------------------------------------------------------------
struct LcGeom { int lc_b[NAZB_MAX_HIDDEN_LAYERS]; int hdim[NAZB_MAX_HIDDEN_LAYERS]; };
__global__ void tc_pack_lc_kernel(int S, int L, int n_lin, int D, int M, int Mp, int kin, int kinp, int w0_floats, int h0,
                                  float hscale, int lc_floats, LcGeom lg, int lc_bout,
                                  const float* const* __restrict__ Wtab, const float* const* __restrict__ btab,
                                  const float* const* __restrict__ mtab, const long long* __restrict__ wst,
                                  const long long* __restrict__ bst, const int* __restrict__ perm, float* __restrict__ dst,
                                  const float* const* __restrict__ bWtab, const float* const* __restrict__ bbtab, float dm_scale) {
  const long long total = (long long)S * L * lc_floats;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int f = (int)(idx % lc_floats);
    int l = (int)((idx / lc_floats) % L);
    int s = (int)(idx / ((long long)lc_floats * L));
    float v = 0.f;
    if (f < w0_floats) {
      int n = f / kinp, k = f % kinp;
      if (n < h0 && k < kin) {
        int ti = l * n_lin;
        float w = Wtab[ti][(size_t)s * wst[ti] + (size_t)n * kin + k];
        if (bWtab) w = nazb_draw_map(bWtab[ti][(size_t)n * kin + k], w, dm_scale);
        v = hscale * w * mtab[ti][(size_t)n * kin + k];
      }
    } else if (f < lc_bout) {
      int j = 0;
      while (j + 1 < n_lin - 1 && f >= lg.lc_b[j + 1]) ++j;
      int n = f - lg.lc_b[j];
      if (n < lg.hdim[j]) {
        int ti = l * n_lin + j;
        float bv = btab[ti][(size_t)s * bst[ti] + n];
        if (bbtab) bv = nazb_draw_map(bbtab[ti][n], bv, dm_scale);
        v = hscale * bv;
      }
    } else if (f < lc_bout + D * Mp) {
      int n = f - lc_bout, rank = n / Mp, m = n % Mp;
      if (m < M) {
        int ti = l * n_lin + (n_lin - 1);
        const int o = m * D + perm[l * D + rank];
        v = btab[ti][(size_t)s * bst[ti] + o];
        if (bbtab) v = nazb_draw_map(bbtab[ti][o], v, dm_scale);
      }
    }
    dst[idx] = v;
  }
}

// v4 inverse layer constants: [W0x * c (Hp0 x dp4, x columns by RANK) | W0c * c (Hp0 x cp4) | hidden biases * c |
// output bias rank-major (stride Mp) | 32 spare floats (rank-0 constants, written by inv4_fold_kernel)]
struct Lc4Geom {
  int lc_w0x, lc_w0c, lc_b[NAZB_MAX_HIDDEN_LAYERS], lc_bout, lc_r0c, lc_floats;
  int hdim[NAZB_MAX_HIDDEN_LAYERS], hp[NAZB_MAX_HIDDEN_LAYERS];
  int dp4, cp4;
  int qmajor;   // v5: first-layer tables stored input-major ([q][hp0], [c][hp0])
};
__global__ void tc_pack_lc4_kernel(int S, int L, int n_lin, int D, int C, int M, int Mp, float hscale, Lc4Geom lg,
                                   const float* const* __restrict__ Wtab, const float* const* __restrict__ btab,
                                   const float* const* __restrict__ mtab, const long long* __restrict__ wst,
                                   const long long* __restrict__ bst, const int* __restrict__ perm, float* __restrict__ dst,
                                   const float* const* __restrict__ bWtab, const float* const* __restrict__ bbtab, float dm_scale,
                                   const short* __restrict__ pmap) {
  const int kin = C + D, nh = n_lin - 1;
  const long long total = (long long)S * L * lg.lc_floats;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int f = (int)(idx % lg.lc_floats);
    int l = (int)((idx / lg.lc_floats) % L);
    int s = (int)(idx / ((long long)lg.lc_floats * L));
    float v = 0.f;
    if (f < lg.lc_b[0]) {
      // first-layer weights: x part (by rank) or context part
      int n, k = -1;
      if (lg.qmajor) {
        if (f < lg.lc_w0c) { int q = f / lg.hp[0]; n = f % lg.hp[0]; if (q < D) k = C + perm[l * D + q]; }
        else { int g2 = f - lg.lc_w0c; int c = g2 / lg.hp[0]; n = g2 % lg.hp[0]; if (c < C) k = c; }
      } else if (f < lg.lc_w0c) { n = f / lg.dp4; int q = f % lg.dp4; if (q < D) k = C + perm[l * D + q]; }
      else { int g2 = f - lg.lc_w0c; n = g2 / lg.cp4; int c = g2 % lg.cp4; if (c < C) k = c; }
      if (pmap) n = (n < 256) ? pmap[n] : -1;
      if (n >= 0 && n < lg.hdim[0] && k >= 0) {
        int ti = l * n_lin;
        float w = Wtab[ti][(size_t)s * wst[ti] + (size_t)n * kin + k];
        if (bWtab) w = nazb_draw_map(bWtab[ti][(size_t)n * kin + k], w, dm_scale);
        v = hscale * w * mtab[ti][(size_t)n * kin + k];
      }
    } else if (f < lg.lc_bout) {
      int j = 0;
      while (j + 1 < nh && f >= lg.lc_b[j + 1]) ++j;
      int n = f - lg.lc_b[j];
      if (pmap) n = (n < 256) ? pmap[(size_t)j * 256 + n] : -1;
      if (n >= 0 && n < lg.hdim[j]) {
        int ti = l * n_lin + j;
        float bv = btab[ti][(size_t)s * bst[ti] + n];
        if (bbtab) bv = nazb_draw_map(bbtab[ti][n], bv, dm_scale);
        v = hscale * bv;
      }
    } else if (f < lg.lc_bout + D * Mp) {
      int n = f - lg.lc_bout, rank = n / Mp, m = n % Mp;
      if (m < M) {
        int ti = l * n_lin + (n_lin - 1);
        const int o = m * D + perm[l * D + rank];
        v = btab[ti][(size_t)s * bst[ti] + o];
        if (bbtab) v = nazb_draw_map(bbtab[ti][o], v, dm_scale);
      }
    }
    dst[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Inverse kernel (flow_tc_inv3.cuh): two independent 64-row chains per 128-point tile.
//
// An M = 64 tcgen05.mma writes its rows to lanes 0-15 of each 32-lane TMEM quadrant, and a lane offset of
// 16 in the D address selects lanes 16-31 (probed: tools/tc_probe_m64.cu).  The 128-point tile is therefore
// run as TWO independent 64-row sub-tiles ("chains") that share every TMEM column, the weight ring and the
// layer constants, but have their own A-operand buffers, barriers and dependency chain: while one chain
// sits in its MMA / barrier latency the other one owns the MUFU and issue slots.
// ------------------------------------------------------------------------------------------------
constexpr int kChains = 2;
constexpr int kChainRows = kTileM / kChains;          // 64

#include "flow_tc_inv3.cuh"
#include "flow_tc_inv4.cuh"
#include "flow_tc_inv5.cuh"
#include "flow_tc_inv6.cuh"
#include "flow_tc_fwd3.cuh"
#include "flow_tc_fwd4.cuh"

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host glue
// ------------------------------------------------------------------------------------------------
static long long* g_tc_dbg = nullptr;
// dev tool (not part of include/nazb.h): device buffer [256][8] receiving CTA 0's per-step clock stamps
extern "C" void nazb_debug_set_clock_buffer(long long* dev_buf) { g_tc_dbg = dev_buf; }
static int g_tc_dbg_all = 0;
extern "C" void nazb_debug_set_all_warps(int on) { g_tc_dbg_all = on; }
extern "C" int nazb_debug_program(const nazb_handle* h, int dir, int* out, int cap);
bool nazb_tc_supported(const FlowGeom& g, std::string* why) {
  TcPlan P;
  if (!base_dims(g, P)) { if (why) *why = "hidden width > 256, D + C + 1 > 48 or M > 32"; return false; }
  if (!plan_smem(g, P)) { if (why) *why = "shared memory plan does not fit"; return false; }
  if (!build_forward(g, P)) { if (why) *why = "forward TMEM plan does not fit"; return false; }
  return true;
}

cudaError_t nazb_tc_create(nazb_handle* h) {
  TcState* t = new TcState();
  h->tc = t;
  cudaError_t e = cudaMalloc(&t->grp_done, sizeof(int) * 65536);
  if (e != cudaSuccess) return e;
  // watchdog word in mapped pinned host memory: readable by the host even after a device-side trap
  e = cudaHostAlloc(&t->wd_host, sizeof(unsigned int) * 4, cudaHostAllocMapped);
  if (e != cudaSuccess) return e;
  t->wd_host[0] = 0;
  return cudaHostGetDevicePointer(&t->wd_dev, t->wd_host, 0);
}

void nazb_tc_destroy(nazb_handle* h) {
  TcState* t = static_cast<TcState*>(h->tc);
  if (!t) return;
  for (int d = 0; d < 2; ++d)
    if (t->wimg[d]) cudaFree(t->wimg[d]);
  if (t->tab_dev) cudaFree(t->tab_dev);
  if (t->lc_dev) cudaFree(t->lc_dev);
  if (t->lcf_dev) cudaFree(t->lcf_dev);
  if (t->lcfold_dev) cudaFree(t->lcfold_dev);
  if (t->grp_done) cudaFree(t->grp_done);
  if (t->pmap_dev) cudaFree(t->pmap_dev);
  if (t->wd_host) cudaFreeHost(t->wd_host);
  delete t;
  h->tc = nullptr;
}

// Options of the tcgen05 engine (nazb_set_option).  They replace the round-1 environment overrides: whatever is set is
// visible through nazb_get_option and recorded by bench.py.
int nazb_tc_set_option(nazb_handle* h, const char* name, int value) {
  TcState* t = static_cast<TcState*>(h->tc);
  if (!t) return NAZB_ERR_UNSUPPORTED;
  if (!strcmp(name, "inv_kernel")) { if (value < 3 || value > 6) return NAZB_ERR_BAD_ARG; t->opt_inv_kernel = value; h->is_packed = false; return NAZB_OK; }
  if (!strcmp(name, "inv_fold")) { t->opt_fold = value ? 1 : 0; return NAZB_OK; }
  if (!strcmp(name, "inv_merge_n")) { if (value < -1 || value > 256) return NAZB_ERR_BAD_ARG; t->opt_merge_n = value; h->is_packed = false; return NAZB_OK; }
  if (!strcmp(name, "inv_gate")) { if (value < 0 || value > 3) return NAZB_ERR_BAD_ARG; t->opt_gate = value; return NAZB_OK; }
  if (!strcmp(name, "inv_a_tmem")) { t->opt_a_tmem = value ? 1 : 0; return NAZB_OK; }
  if (!strcmp(name, "inv_park")) { if (value < 0 || value > 2) return NAZB_ERR_BAD_ARG; t->opt_park = value; return NAZB_OK; }
  if (!strcmp(name, "inv_trim")) { t->opt_trim = value ? 1 : 0; h->is_packed = false; return NAZB_OK; }
  if (!strcmp(name, "inv_gaps")) { t->opt_gaps = value ? 1 : 0; h->is_packed = false; return NAZB_OK; }
  if (!strcmp(name, "inv_defer")) { if (value < 0 || value > 2) return NAZB_ERR_BAD_ARG; t->opt_defer = value; h->is_packed = false; return NAZB_OK; }
  if (!strcmp(name, "inv_align")) { if (value < -1 || value > 1) return NAZB_ERR_BAD_ARG; t->opt_align = value; h->is_packed = false; return NAZB_OK; }
  return NAZB_ERR_BAD_ARG;
}
int nazb_tc_get_option(const nazb_handle* h, const char* name, int* value) {
  const TcState* t = static_cast<const TcState*>(h->tc);
  if (!t) return NAZB_ERR_UNSUPPORTED;
  if (!strcmp(name, "inv_kernel")) *value = t->opt_inv_kernel;
  else if (!strcmp(name, "inv_fold")) *value = t->opt_fold;
  else if (!strcmp(name, "inv_merge_n")) *value = t->opt_merge_n;
  else if (!strcmp(name, "inv_gate")) *value = t->opt_gate;
  else if (!strcmp(name, "inv_a_tmem")) *value = t->opt_a_tmem;
  else if (!strcmp(name, "inv_trim")) *value = t->opt_trim;
  else if (!strcmp(name, "inv_gaps")) *value = t->opt_gaps;
  else if (!strcmp(name, "inv_park")) *value = t->opt_park;
  else if (!strcmp(name, "inv_defer")) *value = t->opt_defer;
  else if (!strcmp(name, "inv_align")) *value = t->opt_align;
  else if (!strcmp(name, "inv_block_width")) *value = t->plan.ok[0] ? t->plan.bw : 0;
  else if (!strcmp(name, "inv_kernel_in_use")) *value = t->plan.ok[0] ? t->plan.inv_ver : 0;
  else if (!strcmp(name, "inv_a_tmem_in_use")) *value = (t->plan.ok[0] && t->plan.inv_ver >= 5 && t->opt_a_tmem) ? (t->plan.a_tmem ? 1 : 0) + (t->plan.fold_a_tmem ? 2 : 0) : 0;   // bit 0: general program, bit 1: folded
  else if (!strcmp(name, "inv_fold_available")) *value = (t->plan.ok[0] && t->plan.inv_ver >= 4 && t->plan.fold_ok) ? 1 : 0;
  else return NAZB_ERR_BAD_ARG;
  return NAZB_OK;
}
unsigned int nazb_tc_watchdog(const nazb_handle* h) {
  const TcState* t = static_cast<const TcState*>(h->tc);
  return (t && t->wd_host) ? t->wd_host[0] : 0u;
}

int64_t nazb_tc_packed_bytes(const nazb_handle* h) {
  const TcState* t = static_cast<const TcState*>(h->tc);
  if (!t) return 0;
  return (int64_t)h->desc.S * (int64_t)(t->draw_bytes[0] + t->draw_bytes[1] + (t->lc_dev ? sizeof(float) * (size_t)h->geom.L * t->plan.lc_floats : 0) +
                                     (t->lcf_dev ? sizeof(float) * (size_t)h->geom.L * t->plan.f_lc_floats : 0));
}

// points per kernel work item (the two-tile forward kernel takes 256-point pairs)
int nazb_tc_rows_per_item(const nazb_handle* h, int dir) {
  const TcState* t = static_cast<const TcState*>(h->tc);
  return (t && dir == 1 && t->plan.fwd3 && t->plan.fwd4) ? 2 * kTileM : kTileM;
}

bool nazb_tc_direction_ok(const nazb_handle* h, int dir) {
  const TcState* t = static_cast<const TcState*>(h->tc);
  return t && t->plan.ok[dir];
}

// grow-only device buffer: repacking with unchanged sizes (every MCMC / SVI step) allocates nothing, so it never
// triggers the implicit device synchronisation of cudaFree / cudaMalloc
template <class T>
static cudaError_t ensure_cap(T** ptr, size_t* cap, size_t bytes) {
  if (*ptr && *cap >= bytes) return cudaSuccess;
  if (*ptr) { cudaFree(*ptr); *ptr = nullptr; *cap = 0; }
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) *cap = bytes;
  return e;
}

cudaError_t nazb_tc_pack(nazb_handle* h, const float* const* W, const float* const* b, const int64_t* wst,
                         const int64_t* bst, const float* const* mask, const float* keep, float p_drop,
                         cudaStream_t st, const DrawMap& dm) {
  TcState* t = static_cast<TcState*>(h->tc);
  const FlowGeom& g = h->geom;
  const int S = h->desc.S, L = g.L, n_lin = g.n_hidden + 1, ntab = L * n_lin;
  t->plan.ok[0] = t->plan.ok[1] = false;   // stays false if anything below fails
  // (re)build the programs now that the MADE block structure is known
  TcPlan P;
  if (!base_dims(g, P) || !plan_smem(g, P)) return cudaErrorInvalidConfiguration;
  P.fwd3 = build_forward3(g, P);
  if (!P.fwd3) { P.steps[1].clear(); P.images[1].clear(); }
  P.ok[1] = P.fwd3 || build_forward(g, P);
  P.inv_ver = t->opt_inv_kernel;
  P.trim = t->opt_trim != 0;
  P.allow_gaps = t->opt_gaps != 0;
  // geometry the inverse programs are built on: the real one (units contiguous in degree order) or the block-aligned one
  FlowGeom gi = g;
  std::vector<short> pmap;
  int n_live_blocks = 0;
  for (int r = 0; r < g.D; ++r) n_live_blocks += (g.blk[0][r + 1] > g.blk[0][r]) ? 1 : 0;
  if (g.C > 0 && n_live_blocks > 0) --n_live_blocks;   // the degree-0 block is folded away for a broadcast context
  auto plan_inverse = [&](const FlowGeom& gg) -> bool {
    P.steps[0].clear(); P.images[0].clear(); P.fold_images.clear(); P.steps_fold.clear(); P.kr_max = 0; P.fold_ok = false;
    const int vgen = (P.inv_ver == 6) ? 5 : (P.inv_ver == 5) ? 3 : 1;
    // Split pushes (critical columns first), v5 defaults by shape, measured at the dev sizes:
    //  * a program whose plan has room for a SECOND A buffer in tensor memory is split and needs no hand-shake at all
    //    (kSplit = 3): cfg3 folded 81.2 -> 84.0 M evals/s, cfg2 +8 %, cfg5 8|4 +6.5 %, cfg4 +0 %;
    //  * otherwise (single A buffer + a_free hand-shake) splitting pays only for flow layers with >= 4 hidden blocks, where the
    //    trailing columns are most of a push (cfg2, cfg5 8|4: +1-2 %); with 2-3 wide blocks it costs 3 % (cfg3: 81.4 -> 79.0).
    // The general and the folded program share the images but not the step lists, so each takes its own decision.
    const int merge_fallback = (P.inv_ver == 5 ? (n_live_blocks >= 4 ? 0 : 256) : 0);
    auto build_prog = [&](TcPlan& T, int variant, std::vector<Step>& steps, std::vector<Image>& imgs, std::vector<Image>* fold) -> bool {
      // built into locals: `steps` / `imgs` / `fold` may be members of T itself, which the plan copy below would clobber
      auto attempt = [&](int merge_n, bool need_room) -> bool {
        TcPlan T2 = T;
        std::vector<Step> st;
        std::vector<Image> im, fo;
        if (!build_inverse(gg, T2, variant, merge_n, st, im, fold ? &fo : nullptr)) return false;
        if (need_room && !(T2.split && T2.a_tmem2)) return false;
        T = T2;
        steps = std::move(st); imgs = std::move(im);
        if (fold) *fold = std::move(fo);
        return true;
      };
      if (t->opt_merge_n >= 0) return attempt(t->opt_merge_n, false);
      if (P.inv_ver == 5 && t->opt_defer == 2 && t->opt_a_tmem && attempt(0, true)) return true;
      return attempt(merge_fallback, false);
    };
    if (!(build_prog(P, vgen, P.steps[0], P.images[0], &P.fold_images) && plan_smem_inv4(gg, P))) return false;
    std::vector<Image> scratch_images;
    TcPlan Q = P;   // the folded variant must not disturb kr_max / layer_bytes of the general plan
    Q.split = false; Q.a_tmem2 = false;
    P.fold_ok = build_prog(Q, vgen + 1, P.steps_fold, scratch_images, nullptr) &&
                Q.layer_bytes[0] == P.layer_bytes[0] && (int)P.fold_images.size() <= kMaxFoldImgs;
    if (!P.fold_ok) P.steps_fold.clear();
    P.fold_a_tmem = P.fold_ok && Q.a_tmem;
    P.t_a_fold = Q.t_a;
    P.split_fold = Q.split;
    P.fold_a_tmem2 = P.fold_ok && Q.a_tmem2;
    return true;
  };
  if (P.inv_ver >= 4) {
    const bool want_align = t->opt_align > 0 || (t->opt_align < 0 && n_live_blocks >= 4);   // auto: measured gain only for many narrow blocks
    if (P.inv_ver >= 5 && want_align && g.inv_mode == NAZB_INV_INCREMENTAL) {
      int maxblk = 0, minblk = 1 << 20;
      for (int j = 0; j < g.n_hidden; ++j)
        for (int r = 0; r < g.D; ++r) {
          maxblk = std::max(maxblk, g.blk[j][r + 1] - g.blk[j][r]);
          minblk = std::min(minblk, g.blk[j][r + 1] - g.blk[j][r]);
        }
      const int bw = ceil_to(std::max(maxblk, 1), 16);
      if (minblk > 0 && g.D * bw <= 256) {   // (an empty degree block keeps the contiguous layout and its shortcuts)
        P.bw = bw; P.aw = ceil_to(std::max(maxblk, 1), 8);
        pmap.assign((size_t)NAZB_MAX_HIDDEN_LAYERS * 256, (short)-1);
        for (int j = 0; j < g.n_hidden; ++j) {
          gi.hidden[j] = g.D * bw;
          for (int r = 0; r <= g.D; ++r) gi.blk[j][r] = (short)(r * bw);
          for (int r = 0; r < g.D; ++r)
            for (int u = g.blk[j][r]; u < g.blk[j][r + 1]; ++u) pmap[(size_t)j * 256 + r * bw + (u - g.blk[j][r])] = (short)u;
        }
        P.ok[0] = plan_inverse(gi);
        if (!P.ok[0]) { P.bw = P.aw = 0; gi = g; pmap.clear(); }
      }
    }
    if (!P.ok[0]) P.ok[0] = plan_inverse(g);
    if (!P.ok[0] && P.inv_ver >= 5) {
      // shapes the 128-row kernels cannot hold (output accumulators at stride ceil16(M)) fall back to the two-chain kernel
      P.inv_ver = 4;
      P.ok[0] = plan_inverse(g);
    }
    for (int j = 0; j < g.n_hidden; ++j) P.hpad[j] = ceil_to(gi.hidden[j], 16);
  } else {
    P.ok[0] = build_inverse(g, P, 0, 0, P.steps[0], P.images[0], nullptr) && plan_smem_inv3(g, P);
  }
  if (!P.ok[0]) { P.steps[0].clear(); P.images[0].clear(); P.layer_bytes[0] = 0; P.fold_ok = false; }
  if (!P.ok[1]) return cudaErrorInvalidConfiguration;
  cudaError_t e;
  for (int d = 0; d < 2; ++d) {
    t->draw_bytes[d] = 0;
    if (!P.ok[d]) continue;
    t->draw_bytes[d] = P.layer_bytes[d] * (size_t)L;
    if ((e = ensure_cap(&t->wimg[d], &t->cap_wimg[d], t->draw_bytes[d] * (size_t)S)) != cudaSuccess) return e;
  }
  // pointer / stride tables on the device (pinned staging + async copy on the caller's stream: nazb_stage_upload)
  std::vector<const float*> tabs(5 * (size_t)ntab);
  std::vector<long long> strides(2 * (size_t)ntab);
  for (int i = 0; i < ntab; ++i) {
    tabs[i] = W[i]; tabs[ntab + i] = b[i]; tabs[2 * ntab + i] = mask[i];
    tabs[3 * ntab + i] = dm.baseW ? dm.baseW[i] : nullptr; tabs[4 * ntab + i] = dm.baseB ? dm.baseB[i] : nullptr;
    strides[i] = wst[i]; strides[ntab + i] = bst[i];
  }
  const bool has_dm = dm.baseW != nullptr && dm.baseB != nullptr;
  size_t tab_bytes = tabs.size() * sizeof(float*) + strides.size() * sizeof(long long);
  if ((e = ensure_cap(&t->tab_dev, &t->cap_tab, tab_bytes)) != cudaSuccess) return e;
  if ((e = nazb_stage_upload(h, t->tab_dev, tabs.data(), tabs.size() * sizeof(float*), st)) != cudaSuccess) return e;
  long long* strides_dev = reinterpret_cast<long long*>(t->tab_dev + tabs.size());
  if ((e = nazb_stage_upload(h, strides_dev, strides.data(), strides.size() * sizeof(long long), st)) != cudaSuccess) return e;
  if (P.bw > 0) {
    if (!t->pmap_dev && (e = cudaMalloc(&t->pmap_dev, sizeof(short) * NAZB_MAX_HIDDEN_LAYERS * 256)) != cudaSuccess) return e;
    if ((e = nazb_stage_upload(h, t->pmap_dev, pmap.data(), sizeof(short) * pmap.size(), st)) != cudaSuccess) return e;
  }
  int hk = 0;
  for (int j = 0; j < g.n_hidden; ++j) hk = std::max(hk, g.hidden[j]);
  for (int d = 0; d < 2; ++d) {
    if (!P.ok[d]) continue;
    const int mp_d = (d == 0) ? P.mp_inv : P.mp;
    for (const Image& im : P.images[d]) {
      long long total = (long long)(im.k_ext / 8) * im.n_ext * L * S;
      int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
      tc_pack_kernel<<<blocks, 256, 0, st>>>(im, S, L, n_lin, g.D, g.M, mp_d, g.kdim[im.lin], g.ndim[im.lin], t->tab_dev,
                                             t->tab_dev + ntab, t->tab_dev + 2 * ntab, strides_dev, strides_dev + ntab,
                                             h->perm_dev, keep, (long long)L * g.n_hidden * hk, (long long)g.n_hidden * hk,
                                             hk, 1.f / (1.f - p_drop), t->wimg[d], (unsigned long long)t->draw_bytes[d],
                                             (unsigned long long)P.layer_bytes[d], has_dm ? t->tab_dev + 3 * ntab : nullptr,
                                             has_dm ? t->tab_dev + 4 * ntab : nullptr, dm.scale,
                                             (d == 0 && P.bw > 0) ? t->pmap_dev : nullptr);
      nazb_count_launch();
    }
  }
  if (P.ok[0]) {
    const size_t lc_bytes = sizeof(float) * (size_t)S * L * P.lc_floats;
    if ((e = ensure_cap(&t->lc_dev, &t->cap_lc, lc_bytes)) != cudaSuccess) return e;
    long long total = (long long)S * L * P.lc_floats;
    int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    if (P.inv_ver >= 4) {
      Lc4Geom lg{};
      lg.lc_w0x = P.lc_w0x; lg.lc_w0c = P.lc_w0c; lg.lc_bout = P.lc_bout; lg.lc_r0c = P.lc_r0c; lg.lc_floats = P.lc_floats;
      lg.dp4 = P.dp4; lg.cp4 = std::max(P.cp4, 1); lg.qmajor = (P.inv_ver >= 5) ? 1 : 0;
      for (int j = 0; j < g.n_hidden; ++j) { lg.lc_b[j] = P.lc_b[j]; lg.hdim[j] = g.hidden[j]; lg.hp[j] = P.hpad[j]; }
      tc_pack_lc4_kernel<<<blocks, 256, 0, st>>>(S, L, n_lin, g.D, g.C, g.M, P.mp_inv, 2.885390081777927f, lg, t->tab_dev,
                                                 t->tab_dev + ntab, t->tab_dev + 2 * ntab, strides_dev, strides_dev + ntab,
                                                 h->perm_dev, t->lc_dev, has_dm ? t->tab_dev + 3 * ntab : nullptr,
                                                 has_dm ? t->tab_dev + 4 * ntab : nullptr, dm.scale, P.bw > 0 ? t->pmap_dev : nullptr);
      nazb_count_launch();
      if (P.fold_ok && (e = ensure_cap(&t->lcfold_dev, &t->cap_lcfold, lc_bytes)) != cudaSuccess) return e;
    } else {
      LcGeom lg{};
      for (int j = 0; j < g.n_hidden; ++j) { lg.lc_b[j] = P.lc_b[j]; lg.hdim[j] = g.hidden[j]; }
      tc_pack_lc_kernel<<<blocks, 256, 0, st>>>(S, L, n_lin, g.D, g.M, P.mp, g.kin, P.kinp, P.lc_b[0], g.hidden[0],
                                                2.885390081777927f, P.lc_floats, lg, P.lc_bout, t->tab_dev, t->tab_dev + ntab, t->tab_dev + 2 * ntab,
                                                strides_dev, strides_dev + ntab, h->perm_dev, t->lc_dev,
                                                has_dm ? t->tab_dev + 3 * ntab : nullptr, has_dm ? t->tab_dev + 4 * ntab : nullptr, dm.scale);
      nazb_count_launch();
    }
  }
  if (P.fwd3) {
    if ((e = ensure_cap(&t->lcf_dev, &t->cap_lcf, sizeof(float) * (size_t)S * L * P.f_lc_floats)) != cudaSuccess) return e;
    LcGeom lg{};
    for (int j = 0; j < g.n_hidden; ++j) { lg.lc_b[j] = P.f_lc_b[j]; lg.hdim[j] = g.hidden[j]; }
    long long total = (long long)S * L * P.f_lc_floats;
    int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    tc_pack_lc_kernel<<<blocks, 256, 0, st>>>(S, L, n_lin, g.D, g.M, g.M, g.kin, P.kinp > 0 ? P.kinp : 4, 0, g.hidden[0],
                                              2.885390081777927f, P.f_lc_floats, lg, P.f_lc_bout, t->tab_dev, t->tab_dev + ntab,
                                              t->tab_dev + 2 * ntab, strides_dev, strides_dev + ntab, h->perm_dev, t->lcf_dev,
                                              has_dm ? t->tab_dev + 3 * ntab : nullptr, has_dm ? t->tab_dev + 4 * ntab : nullptr, dm.scale);
    nazb_count_launch();
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  t->plan = P;
  return cudaSuccess;
}

static cudaError_t launch_inv4(const nazb_handle* h, const TcState* t, const IoArgs& io, int n_groups, cudaStream_t st) {
  const FlowGeom& g = h->geom;
  const TcPlan& P = t->plan;
  const bool fold = P.fold_ok && t->opt_fold && g.C > 0 && io.ctx_rows == 1;
  const std::vector<Step>& prog = fold ? P.steps_fold : P.steps[0];
  KParamsInv4 kp{};
  kp.dbg = g_tc_dbg;
  kp.dbg_all = g_tc_dbg_all;
  kp.park = t->opt_park;
  kp.nsteps = (int)prog.size();
  for (int i = 0; i < kp.nsteps; ++i) kp.steps[i] = prog[i];
  kp.wimg = t->wimg[0];
  kp.draw_bytes = t->draw_bytes[0];
  kp.layer_bytes = P.layer_bytes[0];
  kp.lc = fold ? t->lcfold_dev : t->lc_dev;
  kp.lc_floats = P.lc_floats; kp.lc_s0 = fold ? 0 : io.s_begin;
  kp.lc_w0x = P.lc_w0x; kp.lc_w0c = P.lc_w0c; kp.lc_b0 = P.lc_b[0]; kp.lc_r0c = P.lc_r0c;
  kp.dp4 = (P.inv_ver >= 5) ? P.hpad[0] : P.dp4;   // v5: row stride of the input-major first-layer tables
  kp.cp4 = P.cp4;
  kp.perm = h->perm_dev;
  kp.D = g.D; kp.C = g.C; kp.L = g.L; kp.M = g.M; kp.Mp = P.mp_inv; kp.K = g.K; kp.kind = g.kind;
  kp.nslots = P.j_nslots; kp.kr_max = P.kr_max;
  kp.folded = fold ? 1 : 0;
  kp.bound = g.bound; kp.clip_lo = g.clip_lo; kp.clip_hi = g.clip_hi;
  kp.off_xin = P.j_xin; kp.off_lc = P.j_lc; kp.off_h = P.j_a; kp.off_y = P.j_y; kp.off_xo = P.j_xo; kp.off_xr = P.j_xr;
  kp.off_misc = P.j_misc; kp.off_scratch = P.j_scratch; kp.off_ring = P.j_ring;
  kp.wd = t->wd_dev;
  const int n_tiles = (io.N + kTileM - 1) / kTileM;
  const int grid = (int)std::min<long long>((long long)n_tiles * n_groups, h->sm_count);
  // draw-group gate: only when every CTA works in every group (n_tiles >= 2 grid) and there are groups to drift across
  kp.grp_done = nullptr;
  cudaError_t e;
  if (t->opt_gate && n_groups >= 3 && n_tiles >= 2 * grid && n_groups <= 65536) {
    if ((e = cudaMemsetAsync(t->grp_done, 0, sizeof(int) * (size_t)n_groups, st)) != cudaSuccess) return e;
    kp.grp_done = t->grp_done;
  }
  kp.gate_dist = (t->opt_gate == 2 || (t->opt_gate == 1 && n_tiles >= 32 * grid)) ? 1 : 2;
  if (fold) {
    FoldParams fp{};
    fp.n_img = (int)P.fold_images.size();
    for (int i = 0; i < fp.n_img; ++i) {
      const Image& im = P.fold_images[i];
      FoldImg& fi = fp.img[i];
      fi.w_off = im.w_off; fi.w_bytes = im.w_bytes; fi.lin = im.lin; fi.n_ext = im.n_ext; fi.k_ext = im.k_ext;
      fi.n0 = im.n0; fi.k0 = im.k0;   // hidden unit / output column of image row 0
    }
    fp.wimg = t->wimg[0]; fp.draw_bytes = t->draw_bytes[0]; fp.layer_bytes = P.layer_bytes[0];
    fp.lc = t->lc_dev; fp.lcf = t->lcfold_dev; fp.lc_floats = P.lc_floats; fp.lc_w0c = P.lc_w0c; fp.lc_r0c = P.lc_r0c;
    fp.cp4 = std::max(P.cp4, 1); fp.lc_bout = P.lc_bout;
    fp.w0c_sn = (P.inv_ver >= 5) ? 1 : fp.cp4;                           // W0c[n][c] at n * w0c_sn + c * w0c_sc
    fp.w0c_sc = (P.inv_ver >= 5) ? P.hpad[0] : 1;
    // blk1 = number of degree-0 units: they occupy the first columns in both layouts
    for (int j = 0; j < g.n_hidden; ++j) { fp.lc_b[j] = P.lc_b[j]; fp.hp[j] = P.hpad[j]; fp.blk1[j] = g.blk[j][1]; }
    fp.n_hidden = g.n_hidden; fp.L = g.L; fp.C = g.C; fp.D = g.D; fp.Mp = P.mp_inv; fp.kind = g.kind;
    fp.bound = g.bound; fp.clip_lo = g.clip_lo; fp.clip_hi = g.clip_hi;
    fp.ctx = io.ctx; fp.s_begin = io.s_begin;
    inv4_fold_kernel<<<io.s_count * g.L, 256, 0, st>>>(fp);
    nazb_count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  const int mode = (g.kind == NAZB_KIND_AFFINE) ? 0 : (g.kind == NAZB_KIND_RQS && g.K == 8) ? 1 : 2;
  auto kern = flow_tc_inv4_kernel<false, 2>;
  int threads = kV4Threads;
  if (P.inv_ver == 6) {
    threads = kV6Threads;
    const bool atm = t->opt_a_tmem && (fold ? P.fold_a_tmem : P.a_tmem);
    kp.t_a = (uint32_t)(fold ? P.t_a_fold : P.t_a);
    if (atm) {
      kern = flow_tc_inv6_kernel<false, 2, true>;
      if (mode == 0) kern = flow_tc_inv6_kernel<false, 0, true>;
      else if (mode == 1) kern = g_tc_dbg ? flow_tc_inv6_kernel<true, 1, true> : flow_tc_inv6_kernel<false, 1, true>;
    } else {
      kern = flow_tc_inv6_kernel<false, 2, false>;
      if (mode == 0) kern = flow_tc_inv6_kernel<false, 0, false>;
      else if (mode == 1) kern = g_tc_dbg ? flow_tc_inv6_kernel<true, 1, false> : flow_tc_inv6_kernel<false, 1, false>;
    }
  } else if (P.inv_ver == 5) {
    threads = kV5Threads;
    const bool atm = t->opt_a_tmem && (fold ? P.fold_a_tmem : P.a_tmem);
    kp.t_a = (uint32_t)(fold ? P.t_a_fold : P.t_a);
    const bool afree = atm && (fold ? P.split_fold : P.split);
    kp.a_free = afree ? 1 : 0;
    const bool dbl = afree && (fold ? P.fold_a_tmem2 : P.a_tmem2);
    const bool defer = dbl && t->opt_defer == 1;
    kp.a_free = (afree && !(dbl && t->opt_defer >= 1)) ? 1 : 0;
    // <debug, transform kind, A in TMEM, kSplit, per-point context code>: context-folded programs get the instantiation without it
#define NAZB_V5K(dbg, md, at, sp) (fold ? flow_tc_inv5_kernel<dbg, md, at, sp, false> : flow_tc_inv5_kernel<dbg, md, at, sp, true>)
    if (atm && dbl && t->opt_defer == 2) {
      kern = NAZB_V5K(false, 2, true, 3);
      if (mode == 0) kern = NAZB_V5K(false, 0, true, 3);
      else if (mode == 1) kern = g_tc_dbg ? NAZB_V5K(true, 1, true, 3) : NAZB_V5K(false, 1, true, 3);
    } else if (atm && defer) {
      kern = NAZB_V5K(false, 2, true, 2);
      if (mode == 0) kern = NAZB_V5K(false, 0, true, 2);
      else if (mode == 1) kern = NAZB_V5K(false, 1, true, 2);
    } else if (atm && afree) {
      kern = NAZB_V5K(false, 2, true, 1);
      if (mode == 0) kern = NAZB_V5K(false, 0, true, 1);
      else if (mode == 1) kern = g_tc_dbg ? NAZB_V5K(true, 1, true, 1) : NAZB_V5K(false, 1, true, 1);
    } else if (atm) {
      kern = NAZB_V5K(false, 2, true, 0);
      if (mode == 0) kern = NAZB_V5K(false, 0, true, 0);
      else if (mode == 1) kern = g_tc_dbg ? NAZB_V5K(true, 1, true, 0) : NAZB_V5K(false, 1, true, 0);
    } else {
      kern = NAZB_V5K(false, 2, false, 0);
      if (mode == 0) kern = NAZB_V5K(false, 0, false, 0);
      else if (mode == 1) kern = g_tc_dbg ? NAZB_V5K(true, 1, false, 0) : NAZB_V5K(false, 1, false, 0);
    }
#undef NAZB_V5K
  } else {
    if (mode == 0) kern = flow_tc_inv4_kernel<false, 0>;
    else if (mode == 1) kern = flow_tc_inv4_kernel<false, 1>;
  }
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.j_smem_bytes);
  if (e != cudaSuccess) return e;
  kern<<<grid, threads, P.j_smem_bytes, st>>>(kp, io, n_groups);
  nazb_count_launch();
  return cudaGetLastError();
}

cudaError_t nazb_tc_launch(const nazb_handle* h, const IoArgs& io, int n_groups, cudaStream_t st) {
  const TcState* t = static_cast<const TcState*>(h->tc);
  const FlowGeom& g = h->geom;
  const TcPlan& P = t->plan;
  const int d = io.dir;
  if (!P.ok[d]) return cudaErrorNotSupported;
  if (d == 0 && P.inv_ver >= 4) return launch_inv4(h, t, io, n_groups, st);
  if (d == 0) {
    KParamsInv kp{};
    kp.dbg = g_tc_dbg;
    kp.nsteps = (int)P.steps[0].size();
    for (int i = 0; i < kp.nsteps; ++i) kp.steps[i] = P.steps[0][i];
    kp.wimg = t->wimg[0];
    kp.draw_bytes = t->draw_bytes[0];
    kp.layer_bytes = P.layer_bytes[0];
    kp.lc = t->lc_dev; kp.lc_floats = P.lc_floats; kp.lc_b0 = P.lc_b[0];
    kp.phase_delay = 0;
    kp.perm = h->perm_dev;
    kp.D = g.D; kp.C = g.C; kp.L = g.L; kp.M = g.M; kp.Mp = P.mp; kp.K = g.K; kp.kind = g.kind; kp.kin = g.kin;
    kp.kinp = P.kinp; kp.hp_max = P.hp_max; kp.nslots = P.j_nslots;
    kp.kr_max = P.kr_max; kp.xf = P.xf ? 1 : 0;
    kp.bound = g.bound; kp.clip_lo = g.clip_lo; kp.clip_hi = g.clip_hi;
    kp.off_xin = P.j_xin; kp.off_lc = P.j_lc; kp.off_h = P.j_a; kp.off_y = P.j_y; kp.off_xo = P.j_xo;
    kp.off_misc = P.j_misc; kp.off_scratch = P.j_scratch; kp.off_ring = P.j_ring;
    kp.off_ax = P.j_ax; kp.off_xring = P.j_xring; kp.xslot_bytes = P.xslot_bytes;
    const int n_tiles = (io.N + kTileM - 1) / kTileM;
    const int grid = (int)std::min<long long>((long long)n_tiles * n_groups, h->sm_count);
    const int mode = (g.kind == NAZB_KIND_AFFINE) ? 0 : (g.kind == NAZB_KIND_RQS && g.K == 8) ? 1 : 2;
    auto kern = flow_tc_inv3_kernel<false, 2>;
    if (mode == 0) kern = g_tc_dbg ? flow_tc_inv3_kernel<true, 0> : flow_tc_inv3_kernel<false, 0>;
    else if (mode == 1) kern = g_tc_dbg ? flow_tc_inv3_kernel<true, 1> : flow_tc_inv3_kernel<false, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.j_smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, kV3Threads, P.j_smem_bytes, st>>>(kp, io, n_groups);
    nazb_count_launch();
    return cudaGetLastError();
  }
  if (P.fwd3) {
    KParamsFwd3 kp{};
    kp.nsteps = (int)P.steps[1].size();
    for (int i = 0; i < kp.nsteps; ++i) kp.steps[i] = P.steps[1][i];
    kp.wimg = t->wimg[1];
    kp.draw_bytes = t->draw_bytes[1];
    kp.layer_bytes = P.layer_bytes[1];
    kp.lc = t->lcf_dev; kp.lc_floats = P.f_lc_floats;
    kp.perm = h->perm_dev;
    kp.D = g.D; kp.C = g.C; kp.L = g.L; kp.M = g.M; kp.K = g.K; kp.kind = g.kind; kp.kin = g.kin;
    kp.hp_max = P.hp_max; kp.nslots = P.f_nslots; kp.t_pre1 = P.hp_max;
    kp.bound = g.bound; kp.clip_lo = g.clip_lo; kp.clip_hi = g.clip_hi;
    kp.off_ax = P.f_ax; kp.off_a = P.f_a; kp.off_x = P.f_x; kp.off_ctx = P.f_ctx; kp.off_misc = P.f_misc;
    kp.off_lc = P.f_lc; kp.off_ring = P.f_ring;
    if (P.fwd4) {
      kp.nslots = P.g_nslots;
      kp.off_ax = P.g_ax; kp.off_a = P.g_a; kp.off_x = P.g_x; kp.off_ctx = P.g_ctx; kp.off_misc = P.g_misc;
      kp.off_lc = P.g_lc; kp.off_ring = P.g_ring;
      for (int i = 0; i < 8; ++i) kp.slot_off[i] = P.g_slot_off[i];
      auto kern4 = g_tc_dbg ? flow_tc_fwd4_kernel<true> : flow_tc_fwd4_kernel<false>;
      cudaError_t e4 = cudaFuncSetAttribute(kern4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.g_smem_bytes);
      if (e4 != cudaSuccess) return e4;
      const int n_pairs = (io.N + 2 * kTileM - 1) / (2 * kTileM);
      const int grid4 = (int)std::min<long long>((long long)n_pairs * n_groups, h->sm_count);
      kern4<<<grid4, kF4Threads, P.g_smem_bytes, st>>>(g_tc_dbg, kp, io, n_groups);
      nazb_count_launch();
      return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(flow_tc_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.f_smem_bytes);
    if (e != cudaSuccess) return e;
    const int n_tiles = (io.N + kTileM - 1) / kTileM;
    const int grid = (int)std::min<long long>((long long)n_tiles * n_groups, h->sm_count);
    flow_tc_fwd3_kernel<<<grid, kF3Threads, P.f_smem_bytes, st>>>(kp, io, n_groups);
    nazb_count_launch();
    return cudaGetLastError();
  }
  KParams kp{};
  kp.dbg = g_tc_dbg;
  kp.nsteps = (int)P.steps[d].size();
  for (int i = 0; i < kp.nsteps; ++i) kp.steps[i] = P.steps[d][i];
  kp.wimg = t->wimg[d];
  kp.draw_bytes = t->draw_bytes[d];
  kp.layer_bytes = P.layer_bytes[d];
  kp.perm = h->perm_dev;
  kp.D = g.D; kp.C = g.C; kp.L = g.L; kp.M = g.M; kp.Mp = P.mp; kp.K = g.K; kp.kind = g.kind; kp.kin = g.kin;
  kp.kin_pad = P.kin_pad; kp.hp_max = P.hp_max; kp.nslots = P.nslots;
  kp.bound = g.bound; kp.clip_lo = g.clip_lo; kp.clip_hi = g.clip_hi;
  kp.off_in = P.off_in; kp.off_h = P.off_h; kp.off_x = P.off_x; kp.off_y = P.off_y; kp.off_xo = P.off_xo;
  kp.off_ctx = P.off_ctx; kp.off_misc = P.off_misc; kp.off_scratch = P.off_scratch; kp.off_ring = P.off_ring;
  cudaError_t e = cudaFuncSetAttribute(flow_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes);
  if (e != cudaSuccess) return e;
  const int n_tiles = (io.N + kTileM - 1) / kTileM;
  long long items = (long long)n_tiles * n_groups;
  int grid = (int)std::min<long long>(items, h->sm_count);
  flow_tc_kernel<<<grid, kThreads, P.smem_bytes, st>>>(kp, io, n_groups);
  nazb_count_launch();
  return cudaGetLastError();
}

// dev tool: dump the step program (n, ksteps, nsplit, epi, e_ncols, w_bytes) of one direction
extern "C" int nazb_debug_program(const nazb_handle* h, int dir, int* out, int cap) {
  const TcState* t = static_cast<const TcState*>(h->tc);
  if (!t || !t->plan.ok[dir]) return 0;
  int n = 0;
  for (const Step& s : t->plan.steps[dir]) {
    if ((n + 1) * 8 > cap) break;
    int* o = out + n * 8;
    o[0] = s.n; o[1] = s.ksteps; o[2] = s.nsplit; o[3] = s.epi; o[4] = s.e_ncols; o[5] = (int)s.w_bytes; o[6] = s.d_col; o[7] = s.a_buf;
    ++n;
  }
  return n;
}
