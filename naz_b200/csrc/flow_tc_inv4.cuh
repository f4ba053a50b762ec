// v4 inverse (`log_prob` direction) kernel of the tcgen05 engine.  Included by flow_tc.cu inside its anonymous
// namespace (shares Step / IoArgs and the helpers defined there).
//
// Same mathematics as v3 (push-style incremental inverse of the reference's D-pass autoregressive inverse,
// bflow_jax_maf.py:181-194; two 64-row chains per 128-point tile; accumulators resident in TMEM for a whole flow
// layer; fp16 hi/lo 3-MMA products).  What changed, and why (round-1 profile: one chain's 20-phase dependency chain
// per flow layer bounds the kernel, a third of the issued instructions were mbarrier spins):
//
//   * CONTEXT FOLD.  Hidden units of MADE degree 0 see only the context.  With a broadcast context vector
//     (ctx_rows == 1: calibrate.py:85,126 `test_lambda`) everything stage 0 of a flow layer computes — the degree-0
//     blocks of all hidden layers, their pushes into later blocks, the rank-0 transform parameters — is the same for
//     every point.  `inv4_fold_kernel` evaluates it once per (draw, flow layer) inside the same nazb_inverse call and
//     writes it into the layer constants (effective biases + the rank-0 knot table); the main kernel then runs D-1
//     stages instead of D: 5 of 20 phases and ~40 % of the MMAs disappear.  Per-point contexts run the general program.
//   * FIRST CONDITIONER LAYER WITHOUT AN MMA ROUND TRIP.  pre_1[block r] = b' + W0[:, x_{<r}] x_{<r} needs r <= D-1
//     multiply-adds per unit, so every epilogue warp of the chain computes its K slices of tanh(pre_1) straight from
//     the layer constants as soon as x_{r-1} is known (64-thread named barrier between the two warps that share a
//     TMEM quadrant); the K = 16 MMA + commit + tcgen05.ld hop of v3 is gone (one of four round trips per stage).
//   * mbarrier waits carry a clock watchdog that writes a tag to mapped host memory and traps instead of hanging the GPU.
//   * x_r is published before its log-det is computed; pushes are trimmed to the last real column.
//   * Optional draw-group gate: a CTA's producer does not start draw group g before every CTA has finished issuing
//     group g-2, so the weight images in flight stay L2-resident (DRAM traffic of round 1: 140x algorithmic).
#pragma once

constexpr int kV4Parts = 2;                          // epilogue warps per (chain, quadrant)
constexpr int kV4EpiWarps = kChains * 4 * kV4Parts;
constexpr int kV4Issuer0 = kV4EpiWarps;              // next kChains warps: MMA issuers
constexpr int kV4Producer = kV4EpiWarps + kChains;   // last warp: TMA producer
constexpr int kV4Threads = (kV4EpiWarps + kChains + 1) * 32;
constexpr int kV4MaxSlices = 8;                      // A block <= 128 columns


struct KParamsInv4 {
  Step steps[kMaxSteps];
  int nsteps;
  const uint8_t* wimg;
  unsigned long long draw_bytes, layer_bytes;
  const float* lc;                 // [rows][L][lc_floats]; row of local draw si = lc_s0 + si
  int lc_floats, lc_s0;
  int lc_w0x, lc_w0c, lc_b0, lc_r0c, dp4, cp4;
  const int* perm;
  int D, C, L, M, Mp, K, kind, nslots, kr_max;
  int folded;                      // context folded into the layer constants: stage 0 is constant, ctx adds nothing
  uint32_t t_a;                    // v5, A operand in tensor memory: first TMEM column of the hi image (lo at + kr_max / 2)
  int a_free;                      // v5: the program has split pushes and A lives in TMEM: writers of an A block wait for the
                                   // a_free barrier (measured: the extra commit + wait cost 7 % when no push is split, so it is off then)
  float bound, clip_lo, clip_hi;
  uint32_t off_xin, off_lc, off_h, off_y, off_xo, off_xr, off_misc, off_scratch, off_ring;
  int* grp_done;                   // [n_groups] producers that finished issuing a draw group (gate), or null
  int gate_dist;                   // a producer starts group g once every CTA has finished issuing group g - gate_dist (1 or 2)
  unsigned int* wd;                // watchdog word (mapped host memory) or null
  long long* dbg;
  int park;                        // issuer / producer waits use the suspend-time hint (mbar_wait4_parked)
  int dbg_all;                     // dev tool: log every epilogue warp (slot = warp id, issuer = slot 16) instead of warps 0 / 8 / issuer
};

// mbarrier wait (a suspend-time hint was measured: 2 % slower) with a poll-count watchdog: on timeout the tag goes to mapped host
// memory and the kernel traps, so a protocol bug surfaces as a launch failure with a location instead of a hung GPU.
__device__ __forceinline__ void mbar_wait4(uint64_t* bar, uint32_t parity, unsigned int* wd, uint32_t tag) {
  // the watchdog counts polls instead of reading the clock (5 instead of 8 instructions per poll: the pollers share their
  // sub-partition's issue slots with working warps); 2^27 polls of >= 20 cycles each are seconds, far beyond any legal wait
  asm volatile(
      "{\n\t.reg .pred P1, P2;\n\t.reg .u32 n;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE4;\n\t"
      "mov.u32 n, 0;\n\t"
      "LAB_WAIT4:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE4;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.gt.u32 P2, n, 0x8000000;\n\t"
      "@!P2 bra LAB_WAIT4;\n\t"
      "setp.ne.u64 P2, %2, 0;\n\t"
      "@P2 st.volatile.global.u32 [%2], %3;\n\t"
      "fence.acq_rel.sys;\n\t"
      "trap;\n\t"
      "DONE4:\n\t}\n" ::"r"(tcx::smem_u32(bar)),
      "r"(parity), "l"(wd), "r"(tag)
      : "memory");
}
// Same wait for the single-purpose warps (MMA issuer, TMA producer): try_wait with a suspend-time hint, so that the warp is
// parked by the hardware until the phase completes instead of polling.  Measured with the all-warp event log
// (tools/inv5_spread.py): the polling issuer / producer took 15-25 % of the issue slots of the sub-partition they live on, and the four
// epilogue warps of that sub-partition finished every phase 150-500 cycles after the others — the whole tile waits for them.
__device__ __forceinline__ void mbar_wait4_parked(uint64_t* bar, uint32_t parity, unsigned int* wd, uint32_t tag) {
  asm volatile(
      "{\n\t.reg .pred P1, P2;\n\t.reg .u64 t0, t1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x10000;\n\t"
      "@P1 bra DONE4P;\n\t"
      "mov.u64 t0, %%clock64;\n\t"
      "LAB_WAIT4P:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x10000;\n\t"
      "@P1 bra DONE4P;\n\t"
      "mov.u64 t1, %%clock64;\n\t"
      "sub.u64 t1, t1, t0;\n\t"
      "setp.gt.u64 P2, t1, 0x100000000;\n\t"
      "@!P2 bra LAB_WAIT4P;\n\t"
      "setp.ne.u64 P2, %2, 0;\n\t"
      "@P2 st.volatile.global.u32 [%2], %3;\n\t"
      "fence.acq_rel.sys;\n\t"
      "trap;\n\t"
      "DONE4P:\n\t}\n" ::"r"(tcx::smem_u32(bar)),
      "r"(parity), "l"(wd), "r"(tag)
      : "memory");
}
__device__ __forceinline__ void mbar_wait4_sleepy(uint64_t* bar, uint32_t parity, unsigned int* wd, uint32_t tag) {
  asm volatile(
      "{\n\t.reg .pred P1, P2;\n\t.reg .u32 n;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE4S;\n\t"
      "mov.u32 n, 0;\n\t"
      "LAB_WAIT4S:\n\t"
      "nanosleep.u32 64;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE4S;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.gt.u32 P2, n, 0x8000000;\n\t"
      "@!P2 bra LAB_WAIT4S;\n\t"
      "setp.ne.u64 P2, %2, 0;\n\t"
      "@P2 st.volatile.global.u32 [%2], %3;\n\t"
      "fence.acq_rel.sys;\n\t"
      "trap;\n\t"
      "DONE4S:\n\t}\n" ::"r"(tcx::smem_u32(bar)),
      "r"(parity), "l"(wd), "r"(tag)
      : "memory");
}
#define WD_TAG(site) ((uint32_t)(site) | ((uint32_t)warp << 8) | ((uint32_t)blockIdx.x << 16))

__device__ __forceinline__ void pair_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// Inverse rational-quadratic spline (K = 8) from a precomputed knot table t = [x knots 0..8 | y knots 0..8 | derivatives 0..8]
// (shared memory, the same for every row: context-folded rank 0).  Formulas of rqs8_inv_pair / rqs_fast<8>.
__device__ __forceinline__ void rqs8_inv_knots(float in, float B, const float* __restrict__ t, float& out, float& ld_fwd) {
  const float eps = 1e-6f, LN2 = 0.6931471805599453f;
  int k = 0;
#pragma unroll
  for (int j = 1; j < 8; ++j) k += (in >= t[9 + j] + eps) ? 1 : 0;
  const float sel_x = t[k], sel_w = t[k + 1] - t[k];
  const float sel_y = t[9 + k], sel_h = t[10 + k] - t[9 + k];
  const float d0 = t[18 + k], d1 = t[19 + k];
  const float delta = sel_h * tcx::rcp_approx(sel_w);
  const float t2 = d0 + d1 - 2.f * delta;
  const float dy = in - sel_y;
  const float a = fmaf(dy, t2, sel_h * (delta - d0));
  const float b = fmaf(-dy, t2, sel_h * d0);
  const float c = -delta * dy;
  const float disc = fmaf(b, b, -4.f * a * c);
  const float th = (2.f * c) * tcx::rcp_approx(-b - tcx::sqrt_approx(fmaxf(disc, 0.f)));
  const bool inside = (in >= -B && in <= B);
  out = inside ? fmaf(th, sel_w, sel_x) : in;
  const float tomt = th * (1.f - th), omt = 1.f - th;
  const float den = fmaf(t2, tomt, delta);
  const float dnum = delta * delta * fmaf(d1, th * th, fmaf(2.f * delta, tomt, d0 * omt * omt));
  ld_fwd = inside ? (tcx::lg2_approx(dnum) - 2.f * tcx::lg2_approx(den)) * LN2 : 0.f;
}

// kMode: 0 = affine, 1 = rational-quadratic spline with K = 8 (two lanes per row), 2 = any other spline.
template <bool kDbg, int kMode>
__global__ void __launch_bounds__(kV4Threads, 1) flow_tc_inv4_kernel(const __grid_constant__ KParamsInv4 p,
                                                                      const __grid_constant__ IoArgs io, int n_groups) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);            // [nslots] TMA -> issuers
  uint64_t* w_empty = w_full + 8;                                   // [nslots] count = kChains (one commit per issuer)
  uint64_t* bar_acc = w_empty + 8;                                  // [kChains] issuer -> epilogue warps
  uint64_t* lc_full = bar_acc + kChains;                            // [2]
  uint64_t* lc_empty = lc_full + 2;                                 // [2], count = kV4EpiWarps
  uint64_t* a_ready = lc_empty + 2;                                 // [kChains][2 buffers][kV4MaxSlices]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + kChains * 2 * kV4MaxSlices);
  float* xin = reinterpret_cast<float*>(smem + p.off_xin);          // [C][128] context rows (per-point contexts)
  float* lcs = reinterpret_cast<float*>(smem + p.off_lc);           // [2][lc_floats]
  float* ycur = reinterpret_cast<float*>(smem + p.off_y);           // [D][128] by dimension
  float* xorig = reinterpret_cast<float*>(smem + p.off_xo);         // [D][128]
  float* xr = reinterpret_cast<float*>(smem + p.off_xr);            // [D][128] by RANK: x of the current flow layer
  float* ljac = reinterpret_cast<float*>(smem + p.off_misc);        // [128]
  float* scratch = reinterpret_cast<float*>(smem + p.off_scratch);  // [32][128] (generic spline only)
  uint8_t* ring = smem + p.off_ring;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int D = p.D, C = p.C, M = p.M;
  const uint32_t a_img_bytes = (uint32_t)p.kr_max * kChainRows * 2;   // one fp16 image (hi or lo) of an A block
  const uint32_t a_buf_bytes = 2 * a_img_bytes;

  if (tid == 0) {
    for (int i = 0; i < p.nslots; ++i) { tcx::mbar_init(w_full + i, 1); tcx::mbar_init(w_empty + i, kChains); }
    for (int i = 0; i < kChains; ++i) tcx::mbar_init(bar_acc + i, 1);
    for (int i = 0; i < 2; ++i) { tcx::mbar_init(lc_full + i, 1); tcx::mbar_init(lc_empty + i, kV4EpiWarps); }
    // slice 0 of every A block also collects one arrival from each NON-producing warp of the chain: a warp that waits
    // on the accumulator barrier must be needed for the next MMA, otherwise the issuer could complete two accumulator
    // phases before a late warp has observed the first one and its parity wait would never return
    for (int i = 0; i < kChains * 2 * kV4MaxSlices; ++i)
      tcx::mbar_init(a_ready + i, (i % kV4MaxSlices == 0) ? 4 * kV4Parts : 4);
    tcx::mbar_fence_init();
  }
  if (warp == 0) tcx::tmem_alloc(tmem_slot, kTmemCols);
  for (uint32_t i = tid; i < (kChains * 2 * a_buf_bytes) / 16; i += kV4Threads)
    reinterpret_cast<uint4*>(smem + p.off_h)[i] = make_uint4(0, 0, 0, 0);
  tcx::fence_async_smem();
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n_tiles = (io.N + kTileM - 1) / kTileM;
  const long long n_items = (long long)n_tiles * n_groups;

  if (warp == kV4Producer) {
    // ===================== TMA producer: weight images + layer constants =====================
    if (lane == 0) {
      uint32_t cnt = 0, lcnt = 0;
      int prev_grp = -1;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = (int)(item / n_tiles);
        if (p.grp_done != nullptr && grp != prev_grp) {
          // draw-group gate: finish group prev_grp, start group grp only once every CTA is done issuing grp - gate_dist
          if (prev_grp >= 0) {
            __threadfence();
            atomicAdd(p.grp_done + prev_grp, 1);
          }
          if (grp >= p.gate_dist) {
            const long long t0 = clk();
            const int* flag = p.grp_done + (grp - p.gate_dist);
            while (*reinterpret_cast<const volatile int*>(flag) < (int)gridDim.x) {
              __nanosleep(200);
              if (clk() - t0 > (1ll << 26)) break;   // the gate is an optimisation only: never wait more than ~30 ms
            }
          }
          prev_grp = grp;
        }
        for (int si = grp; si < io.s_count; si += n_groups) {
          const uint8_t* wdraw = p.wimg + (size_t)(io.s_begin + si) * p.draw_bytes;
          const float* lcdraw = p.lc + (size_t)(p.lc_s0 + si) * p.L * p.lc_floats;
          for (int li = 0; li < p.L; ++li) {
            const int l = p.L - 1 - li;
            {
              const uint32_t b = lcnt & 1, use = lcnt >> 1;
              mbar_wait4(lc_empty + b, (use & 1) ^ 1, p.wd, WD_TAG(1));
              tcx::mbar_expect_tx(lc_full + b, (uint32_t)p.lc_floats * 4);
              tcx::bulk_g2s(lcs + (size_t)b * p.lc_floats, lcdraw + (size_t)l * p.lc_floats, (uint32_t)p.lc_floats * 4, lc_full + b);
              ++lcnt;
            }
            const uint8_t* wl = wdraw + (size_t)l * p.layer_bytes;
            for (int st = 0; st < p.nsteps; ++st) {
              const uint32_t wb = p.steps[st].w_bytes;
              if (wb == 0) continue;
              const uint32_t slot = cnt % p.nslots, use = cnt / p.nslots;
              mbar_wait4(w_empty + slot, (use & 1) ^ 1, p.wd, WD_TAG(2));
              tcx::mbar_expect_tx(w_full + slot, wb);
              tcx::bulk_g2s(ring + (size_t)slot * kSlotBytes, wl + p.steps[st].w_off, wb, w_full + slot);
              ++cnt;
            }
          }
        }
      }
      if (p.grp_done != nullptr && prev_grp >= 0) {
        __threadfence();
        atomicAdd(p.grp_done + prev_grp, 1);
      }
    }
  } else if (warp >= kV4Issuer0) {
    // ===================== MMA issuer of chain `ch` =====================
    // Whole warp convergent, tcgen05 instructions predicated on the elected lane (descriptors stay uniform).
    const int ch = warp - kV4Issuer0;
    const uint32_t elected = tcx::elect_one();
    const uint32_t ring_a = tcx::smem_u32(ring);
    const uint32_t a_base = tcx::smem_u32(smem + p.off_h) + (uint32_t)ch * 2 * a_buf_bytes;
    constexpr uint32_t lbo_a = kChainRows * 16;
    constexpr uint64_t dhi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
    const uint32_t d_lane = tmem + ((uint32_t)(ch * 16) << 16);
    uint64_t* my_acc = bar_acc + ch;
    uint64_t* my_ready = a_ready + ch * 2 * kV4MaxSlices;
    uint32_t slot = 0, use = 0, buf = 0, apar = 0;   // apar: one parity bit per (buffer, slice) barrier
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int grp = (int)(item / n_tiles);
      for (int si = grp; si < io.s_count; si += n_groups) {
        for (int li = 0; li < p.L; ++li) {
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_wbytes = p.steps[st].w_bytes;
            if (s_wbytes == 0) continue;
            const uint32_t s_n = p.steps[st].n, s_ncrit = p.steps[st].n_crit, s_dcol = p.steps[st].d_col;
            const int ksteps = p.steps[st].ksteps;
            const uint32_t s_acc = p.steps[st].accumulate, s_last = (p.steps[st].epi != EPI_NONE);
            const uint32_t slice0 = p.steps[st].a_chunk0 >> 1;
            const uint32_t n_rest = s_n - s_ncrit;
            const uint32_t idesc_c = tcx::make_idesc_f16_m64(s_ncrit);
            const uint32_t idesc_r = tcx::make_idesc_f16_m64(n_rest);
            const uint32_t lbo_b = s_n * 16;
            const uint32_t b_hi = ring_a + slot * kSlotBytes;
            const uint32_t b_lo = b_hi + (s_wbytes >> 1);
            const uint32_t lbo_b_hi16 = (lbo_b >> 4) << 16, lbo_a_hi16 = (lbo_a >> 4) << 16;
            const uint32_t a_hi = a_base + buf * a_buf_bytes;
            const uint32_t a_lo = a_hi + a_img_bytes;
            const uint32_t d_c = d_lane + s_dcol, d_r = d_c + s_ncrit;
            const uint32_t ro = s_ncrit;              // image row n_crit = byte offset n_crit * 16, >> 4
            mbar_wait4(w_full + slot, use & 1, p.wd, WD_TAG(3));
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t sl = slice0 + k, bit = 1u << (buf * kV4MaxSlices + sl);
              mbar_wait4(my_ready + buf * kV4MaxSlices + sl, (apar & bit) ? 1u : 0u, p.wd, WD_TAG(4));
              apar ^= bit;
              tcx::tc_fence_after();
              const uint32_t ao = sl * 2 * lbo_a, bo = (uint32_t)k * 2 * lbo_b;
              const uint64_t da_h = dhi | (((a_hi + ao) >> 4) | lbo_a_hi16), da_l = dhi | (((a_lo + ao) >> 4) | lbo_a_hi16);
              const uint64_t db_h = dhi | (((b_hi + bo) >> 4) | lbo_b_hi16), db_l = dhi | (((b_lo + bo) >> 4) | lbo_b_hi16);
              const uint32_t acc0 = (k == 0) ? s_acc : 1u;
              // critical columns (block r): hi*hi + hi*lo + lo*hi
              tcx::mma_f16_ss_elect(d_c, da_h, db_h, idesc_c, acc0, elected);
              tcx::mma_f16_ss_elect(d_c, da_h, db_l, idesc_c, 1u, elected);
              tcx::mma_f16_ss_elect(d_c, da_l, db_h, idesc_c, 1u, elected);
              if (s_last && k == ksteps - 1) tcx::mma_commit_elect(my_acc, elected);
              if (n_rest) {
                tcx::mma_f16_ss_elect(d_r, da_h, db_h + ro, idesc_r, acc0, elected);
                tcx::mma_f16_ss_elect(d_r, da_h, db_l + ro, idesc_r, 1u, elected);
                tcx::mma_f16_ss_elect(d_r, da_l, db_h + ro, idesc_r, 1u, elected);
              }
            }
            tcx::mma_commit_elect(w_empty + slot, elected);   // the slot is free once these MMAs retire
            if (++slot == (uint32_t)p.nslots) { slot = 0; ++use; }
            if (s_last) buf ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps: kV4Parts per (chain, TMEM lane quadrant) =====================
    const int ch = warp / (4 * kV4Parts), part = (warp >> 2) % kV4Parts, q = warp & 3;
    const int hw = lane >> 4, lr = lane & 15;
    const int crow = q * 16 + lr;                  // row inside the chain's 64-row sub-tile
    const int trow = ch * kChainRows + crow;       // row inside the 128-point tile
    const int wrow0 = ch * kChainRows + q * 16;    // first tile row owned by this quadrant
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32 + ch * 16) << 16);
    constexpr bool spline = kMode != 0;
    constexpr bool fast_rqs = kMode == 1;
    const bool rows_mine = (part == 0);            // this warp owns the per-row state of its 16 rows
    const bool owner = rows_mine && (hw == 0);
    const bool add_ctx = (C > 0) && !p.folded;
    const int pair_id = 1 + ch * 4 + q;            // named barrier of the kV4Parts warps sharing these 16 rows
    float* scr = scratch + trow;
    auto raw = [&](int m) { return scr[m * kTileM]; };
    auto setw = [&](int m, float v) { scr[m * kTileM] = v; };
    uint8_t* a_chain = smem + p.off_h + (size_t)ch * 2 * a_buf_bytes;
    uint64_t* my_acc = bar_acc + ch;
    uint64_t* my_ready = a_ready + ch * 2 * kV4MaxSlices;
    uint32_t par_acc = 0, lcnt = 0, buf = 0;
    const uint64_t scale2 = tcx::pk2(kTanhScale, kTanhScale);

    // write this thread's 8-column chunk `c` (hi / lo fp16) of the A block: layout [chunk][64 rows][8 halves]
    auto store_chunk = [&](int c, const uint4& hi4, const uint4& lo4) {
      uint8_t* dst = a_chain + (size_t)buf * a_buf_bytes + ((size_t)c * kChainRows + crow) * 16;
      *reinterpret_cast<uint4*>(dst) = hi4;
      *reinterpret_cast<uint4*>(dst + a_img_bytes) = lo4;
    };
    // publish K slice `sl` of the A block under construction
    auto publish = [&](int sl) {
      tcx::fence_async_smem();
      __syncwarp();
      if (lane == 0) tcx::mbar_arrive(my_ready + buf * kV4MaxSlices + sl);
    };

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = (int)(item % n_tiles), grp = (int)(item / n_tiles);
      const int n0 = tile * kTileM;
      const int npts = min(kTileM, io.N - n0);
      float run_m = -INFINITY, run_s = 0.f;
      // ---- tile load: the 16 rows of this quadrant (part-0 warp), then visible to the other parts ----
      pair_bar_sync(pair_id, 32 * kV4Parts);       // every part is done with the previous tile's rows
      if (rows_mine) {
        if (add_ctx)
          for (int i = lane; i < 16 * C; i += 32) {
            const int pt = wrow0 + i / C, c = i % C;
            float v = 0.f;
            if (pt < npts) v = io.ctx[((io.ctx_rows == 1) ? 0 : (size_t)(n0 + pt)) * C + c];
            xin[c * kTileM + pt] = v;
          }
        for (int i = lane; i < 16 * D; i += 32) {
          const int pt = wrow0 + i / D, d = i % D;
          xorig[d * kTileM + pt] = (pt < npts) ? io.x[(size_t)(n0 + pt) * D + d] : 0.f;
        }
        __syncwarp();
        if (owner) {
          float lj = 0.f;
          if (io.lo != nullptr && trow < npts)
            for (int d = 0; d < D; ++d) xorig[d * kTileM + trow] = nazb::bound_fwd(xorig[d * kTileM + trow], io.lo[d], io.hi[d], lj);
          ljac[trow] = lj;
        }
      }
      pair_bar_sync(pair_id, 32 * kV4Parts);       // context rows are in shared memory

      for (int si = grp; si < io.s_count; si += n_groups) {
        // ---- draw start ----
        float ld_acc = 0.f;
        if (owner)
          for (int d = 0; d < D; ++d) ycur[d * kTileM + trow] = xorig[d * kTileM + trow];

        for (int li = 0; li < p.L; ++li) {
          const int l = p.L - 1 - li;
          const int* perm = p.perm + l * D;
          const float* lc = lcs + (size_t)(lcnt & 1) * p.lc_floats;
          mbar_wait4(lc_full + (lcnt & 1), (lcnt >> 1) & 1, p.wd, WD_TAG(5));
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_epi = p.steps[st].epi;
            if (s_epi == EPI_NONE) continue;   // K-split sub-step: nothing to do on this side
            const uint32_t s_ecol = p.steps[st].e_col, s_encols = p.steps[st].e_ncols, s_eaux = p.steps[st].e_aux;
            const uint32_t s_stage = p.steps[st].stage, s_flags = p.steps[st].flags;
            if (p.steps[st].w_bytes) {
              mbar_wait4(my_acc, par_acc, p.wd, WD_TAG(6));
              par_acc ^= 1;
              tcx::tc_fence_after();
            }
            if (s_epi == EPI_TANH) {
              // block of nch 8-column chunks = nsl K slices; slice s = chunks {2s (half-warp 0), 2s+1 (half-warp 1)}
              const int nch = s_encols >> 3, nsl = (nch + 1) >> 1;
              const int nmine = (nsl - part + kV4Parts - 1) / kV4Parts;   // slices part, part + kV4Parts, ...
              if (part != 0 && lane == 0) tcx::mbar_arrive(my_ready + buf * kV4MaxSlices);   // observer arrival on slice 0
              for (int j0 = 0; j0 < nmine; j0 += 2) {
                uint32_t r[16];
                const int nj = min(2, nmine - j0);
                const int sl0 = part + j0 * kV4Parts;
                const uint32_t ta = lane_base + s_ecol + sl0 * 16;
                tcx::tmem_ld16x2_8<8>(ta, r);
                if (nj > 1) tcx::tmem_ld16x2_8<8>(ta + 16 * kV4Parts, r + 8);
                tcx::tmem_ld_wait();
                tcx::tc_fence_before();
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  if (u < nj) {
                    const int sl = sl0 + u * kV4Parts;
                    const int c = sl * 2 + hw;
                    const int cl = min(c, nch - 1);
                    const uint32_t* ru = r + 8 * u;
                    uint64_t s2[4];
                    const ulonglong2* bv = reinterpret_cast<const ulonglong2*>(lc + s_eaux + cl * 8);
                    const ulonglong2 b0 = bv[0], b1 = bv[1];   // biases already multiplied by 2 log2 e
                    s2[0] = tcx::fma2(tcx::pk2(__uint_as_float(ru[0]), __uint_as_float(ru[1])), scale2, b0.x);
                    s2[1] = tcx::fma2(tcx::pk2(__uint_as_float(ru[2]), __uint_as_float(ru[3])), scale2, b0.y);
                    s2[2] = tcx::fma2(tcx::pk2(__uint_as_float(ru[4]), __uint_as_float(ru[5])), scale2, b1.x);
                    s2[3] = tcx::fma2(tcx::pk2(__uint_as_float(ru[6]), __uint_as_float(ru[7])), scale2, b1.y);
                    uint4 hi4, lo4;
                    tcx::tanh8_scaled(s2, hi4, lo4);
                    if (c >= nch) { hi4 = make_uint4(0, 0, 0, 0); lo4 = hi4; }   // K padding chunk
                    store_chunk(c, hi4, lo4);
                    publish(sl);
                  }
                }
              }
              buf ^= 1;
            } else if (s_epi == EPI_FIRST) {
              // first conditioner layer of block `stage` on CUDA cores: s = b'[n] + sum_c W0c[n][c] ctx_c + sum_{q < stage} W0x[n][q] x_q
              // (everything pre-multiplied by 2 log2 e), tanh, fp16 hi/lo A block.  x_q (by rank) were written by the row owners.
              const int r = (int)s_stage;
              if (r > 0) pair_bar_sync(pair_id, 32 * kV4Parts);   // x_{r-1} of these rows is visible
              const int nch = s_encols >> 3, nsl = (nch + 1) >> 1;
              const int u0 = s_eaux;
              const int nmine = (nsl - part + kV4Parts - 1) / kV4Parts;
              if (part != 0 && lane == 0) tcx::mbar_arrive(my_ready + buf * kV4MaxSlices);   // observer arrival on slice 0
              float xv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < r) xv[k] = xr[k * kTileM + trow];
              float cv[4] = {0.f, 0.f, 0.f, 0.f};
              if (add_ctx) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < C) cv[k] = xin[k * kTileM + trow];
              }
              for (int j = 0; j < nmine; ++j) {
                const int sl = part + j * kV4Parts;
                const int c = sl * 2 + hw;
                uint4 hi4 = make_uint4(0, 0, 0, 0), lo4 = hi4;
                if (c < nch) {
                  const int nb = u0 + c * 8;
                  float acc[8];
                  const float4* bb = reinterpret_cast<const float4*>(lc + p.lc_b0 + nb);
                  const float4 ba = bb[0], bc = bb[1];
                  acc[0] = ba.x; acc[1] = ba.y; acc[2] = ba.z; acc[3] = ba.w;
                  acc[4] = bc.x; acc[5] = bc.y; acc[6] = bc.z; acc[7] = bc.w;
                  if (add_ctx) {
                    for (int k4 = 0; k4 < p.cp4; k4 += 4) {
#pragma unroll
                      for (int e = 0; e < 8; ++e) {
                        const float4 w = *reinterpret_cast<const float4*>(lc + p.lc_w0c + (size_t)(nb + e) * p.cp4 + k4);
                        if (k4 == 0) {
                          acc[e] = fmaf(w.x, cv[0], acc[e]); acc[e] = fmaf(w.y, cv[1], acc[e]);
                          acc[e] = fmaf(w.z, cv[2], acc[e]); acc[e] = fmaf(w.w, cv[3], acc[e]);
                        } else {
                          acc[e] = fmaf(w.x, xin[(k4 + 0) * kTileM + trow], acc[e]);
                          acc[e] = fmaf(w.y, (k4 + 1 < C) ? xin[(k4 + 1) * kTileM + trow] : 0.f, acc[e]);
                          acc[e] = fmaf(w.z, (k4 + 2 < C) ? xin[(k4 + 2) * kTileM + trow] : 0.f, acc[e]);
                          acc[e] = fmaf(w.w, (k4 + 3 < C) ? xin[(k4 + 3) * kTileM + trow] : 0.f, acc[e]);
                        }
                      }
                    }
                  }
                  if (r > 0) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                      const float4 w = *reinterpret_cast<const float4*>(lc + p.lc_w0x + (size_t)(nb + e) * p.dp4);
                      acc[e] = fmaf(w.x, xv[0], acc[e]); acc[e] = fmaf(w.y, xv[1], acc[e]);
                      acc[e] = fmaf(w.z, xv[2], acc[e]); acc[e] = fmaf(w.w, xv[3], acc[e]);
                    }
                  }
                  for (int k4 = 4; k4 < r; k4 += 4) {   // D > 5 only
                    const float x0 = xr[k4 * kTileM + trow];
                    const float x1 = (k4 + 1 < r) ? xr[(k4 + 1) * kTileM + trow] : 0.f;
                    const float x2 = (k4 + 2 < r) ? xr[(k4 + 2) * kTileM + trow] : 0.f;
                    const float x3 = (k4 + 3 < r) ? xr[(k4 + 3) * kTileM + trow] : 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                      const float4 w = *reinterpret_cast<const float4*>(lc + p.lc_w0x + (size_t)(nb + e) * p.dp4 + k4);
                      acc[e] = fmaf(w.x, x0, acc[e]); acc[e] = fmaf(w.y, x1, acc[e]);
                      acc[e] = fmaf(w.z, x2, acc[e]); acc[e] = fmaf(w.w, x3, acc[e]);
                    }
                  }
                  uint64_t s2[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) s2[i] = tcx::pk2(acc[2 * i], acc[2 * i + 1]);
                  tcx::tanh8_scaled(s2, hi4, lo4);
                }
                store_chunk(c, hi4, lo4);
                publish(sl);
              }
              buf ^= 1;
            } else if (s_epi == EPI_XINV0C) {
              // context-folded rank 0: the transform parameters are per-draw constants (lc_r0c), no accumulator involved
              if (owner) {
                const int d = perm[0];
                const float yv = ycur[d * kTileM + trow];
                const float* t = lc + p.lc_r0c;
                float xv, ld;
                if (!spline) { xv = (yv - t[0]) * t[2]; ld = t[1]; }
                else rqs8_inv_knots(yv, p.bound, t, xv, ld);
                xr[trow] = xv;
                ld_acc += ld;
                if (D == 1) ycur[d * kTileM + trow] = xv;
              }
            } else if (s_epi == EPI_XINV && rows_mine) {
              const int r = s_stage, d = perm[r];
              const float yv = ycur[d * kTileM + trow];
              const float* bo = lc + s_eaux;
              const bool has_acc = !(s_flags & 1);
              float xv = 0.f, ld = 0.f;
              if (!spline) {
                uint32_t rr[2] = {0u, 0u};
                if (has_acc) { tcx::tmem_ld16x2_2<0>(lane_base + s_ecol, rr); tcx::tmem_ld_wait(); }
                float mu = __uint_as_float(rr[0]) + bo[0];
                float sc = fminf(fmaxf(__uint_as_float(rr[1]) + bo[1], p.clip_lo), p.clip_hi);
                xv = (yv - mu) * expf(-sc);
                ld = sc;
              } else if (fast_rqs) {
                // the two lanes of a row split the spline: half-warp 0 takes the widths, half-warp 1 the heights
                uint32_t ro[8], rd[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { ro[e] = 0u; rd[e] = 0u; }
                if (has_acc) {
                  tcx::tmem_ld16x2_8<8>(lane_base + s_ecol, ro);        // cols 0-7 (w) | 8-15 (h)
                  tcx::tmem_ld16x2_8<0>(lane_base + s_ecol + 16, rd);   // cols 16-23 (derivatives) to both
                  tcx::tmem_ld_wait();
                }
                const float4* bw = reinterpret_cast<const float4*>(bo + hw * 8);
                const float4* bd = reinterpret_cast<const float4*>(bo + 16);
                const float4 w0 = bw[0], w1 = bw[1], e0 = bd[0], e1 = bd[1];
                const float own[8] = {__uint_as_float(ro[0]) + w0.x, __uint_as_float(ro[1]) + w0.y, __uint_as_float(ro[2]) + w0.z,
                                      __uint_as_float(ro[3]) + w0.w, __uint_as_float(ro[4]) + w1.x, __uint_as_float(ro[5]) + w1.y,
                                      __uint_as_float(ro[6]) + w1.z, __uint_as_float(ro[7]) + w1.w};
                const float dr[8] = {__uint_as_float(rd[0]) + e0.x, __uint_as_float(rd[1]) + e0.y, __uint_as_float(rd[2]) + e0.z,
                                     __uint_as_float(rd[3]) + e0.w, __uint_as_float(rd[4]) + e1.x, __uint_as_float(rd[5]) + e1.y,
                                     __uint_as_float(rd[6]) + e1.z, 0.f};
                nazb::rqs8_inv_pair(yv, p.bound, own, dr, hw, xv, ld);
              } else {
                for (int m0 = 0; m0 < p.Mp; m0 += 8) {
                  uint32_t rr[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) rr[e] = 0u;
                  if (has_acc) { tcx::tmem_ld16x2_8<0>(lane_base + s_ecol + m0, rr); tcx::tmem_ld_wait(); }
                  if (owner) {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                      if (m0 + e < M) scr[(m0 + e) * kTileM] = __uint_as_float(rr[e]) + bo[m0 + e];
                  }
                }
                if (owner) {
                  if (p.kind == NAZB_KIND_RQS) nazb::rational_spline<false>(yv, p.K, p.bound, true, raw, setw, xv, ld);
                  else nazb::rational_spline<true>(yv, p.K, p.bound, true, raw, setw, xv, ld);
                }
              }
              tcx::tc_fence_before();
              if (owner) {
                xr[r * kTileM + trow] = xv;   // published first: the next stage's first layer waits for it
                ld_acc += ld;
                if (r == D - 1) {
                  // end of this flow layer: x becomes the y of the next (earlier) layer
                  for (int rr2 = 0; rr2 < D; ++rr2) ycur[perm[rr2] * kTileM + trow] = xr[rr2 * kTileM + trow];
                }
              }
            }
          }
          // this layer's constants are no longer needed by this warp
          __syncwarp();
          if (lane == 0) tcx::mbar_arrive(lc_empty + (lcnt & 1));
          ++lcnt;
        }

        // ---- draw end ----
        if (rows_mine) {
          float lp = 0.f;
          const bool mine = owner && trow < npts;
          if (owner) {
            float qd = 0.f;
            for (int d = 0; d < D; ++d) { float z = ycur[d * kTileM + trow]; qd += 0.5f * z * z; }
            lp = -qd - 0.5f * D * NAZB_LOG_2PI - ld_acc + ljac[trow];
          }
          if (mine) {
            if (io.out_l) io.out_l[(size_t)si * io.N + n0 + trow] = lp;
            if (io.lse_max) {
              float v = lp + (io.log_w ? io.log_w[si] : 0.f);
              if (!(v <= run_m)) { run_s = run_s * expf(run_m - v) + 1.f; run_m = v; }
              else if (v > -INFINITY) run_s += expf(v - run_m);
            }
            if (io.out_x) {
              float* dst = io.out_x + ((size_t)si * io.N + n0 + trow) * D;
              for (int d = 0; d < D; ++d) dst[d] = ycur[d * kTileM + trow];
            }
          }
          if (io.sum_n) {
            double v = mine ? (double)lp : 0.0;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(io.sum_n + si, v);
          }
        }
      }
      if (io.lse_max && owner && trow < npts) {
        io.lse_max[(size_t)grp * io.N + n0 + trow] = run_m;
        io.lse_sum[(size_t)grp * io.N + n0 + trow] = run_s;
      }
    }
  }
  tcx::tc_fence_before();
  __syncthreads();
  if (warp == 0) tcx::tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// Context fold (see the header comment): one CTA per (draw, flow layer).
// ------------------------------------------------------------------------------------------------
struct FoldImg {
  uint32_t w_off, w_bytes;
  int lin;        // linear layer the image belongs to (1 .. n_hidden); its K axis indexes hidden layer lin - 1
  int n_ext, k_ext, n0, k0;
};
constexpr int kMaxFoldImgs = 24;
struct FoldParams {
  FoldImg img[kMaxFoldImgs];
  int n_img;
  const uint8_t* wimg;
  unsigned long long draw_bytes, layer_bytes;
  const float* lc;       // general layer constants [S][L][lc_floats]
  float* lcf;            // folded layer constants [s_count][L][lc_floats]
  int lc_floats, lc_w0c, lc_r0c, cp4;
  int w0c_sn, w0c_sc;    // W0c[n][c] lives at lc_w0c + n * w0c_sn + c * w0c_sc (v4: unit-major, v5: input-major)
  int lc_b[NAZB_MAX_HIDDEN_LAYERS], lc_bout;
  int hp[NAZB_MAX_HIDDEN_LAYERS], blk1[NAZB_MAX_HIDDEN_LAYERS];   // padded widths; number of degree-0 units per hidden layer
  int n_hidden, L, C, D, Mp, kind;
  float bound, clip_lo, clip_hi;
  const float* ctx;      // [C]
  int s_begin;
};

__device__ __forceinline__ float tanh_from_scaled(float s) {   // s = 2 x log2 e
  const float e = exp2f(fminf(s, 30.f));
  return 1.f - 2.f / (e + 1.f);
}

__global__ void __launch_bounds__(256) inv4_fold_kernel(const __grid_constant__ FoldParams f) {
  __shared__ float hvec[256];       // degree-0 activations of the current hidden layer (0 elsewhere)
  __shared__ float accv[512];       // folded contribution to the next linear layer's pre-activations / outputs
  const int s = blockIdx.x / f.L, l = blockIdx.x % f.L;
  const int tid = threadIdx.x;
  const float* lc = f.lc + ((size_t)(f.s_begin + s) * f.L + l) * f.lc_floats;
  float* out = f.lcf + ((size_t)s * f.L + l) * f.lc_floats;
  const uint8_t* wl = f.wimg + (size_t)(f.s_begin + s) * f.draw_bytes + (size_t)l * f.layer_bytes;
  for (int i = tid; i < f.lc_floats; i += blockDim.x) out[i] = lc[i];
  __syncthreads();
  // first layer: b0' = b0 + W0[:, ctx] ctx (scaled domain) for every unit
  for (int n = tid; n < f.hp[0]; n += blockDim.x) {
    float a = lc[f.lc_b[0] + n];
    for (int c = 0; c < f.C; ++c) a = fmaf(lc[f.lc_w0c + (size_t)n * f.w0c_sn + (size_t)c * f.w0c_sc], f.ctx[c], a);
    out[f.lc_b[0] + n] = a;
    hvec[n] = (n < f.blk1[0]) ? tanh_from_scaled(a) : 0.f;
  }
  for (int n = f.hp[0] + tid; n < 256; n += blockDim.x) hvec[n] = 0.f;
  __syncthreads();
  for (int lin = 1; lin <= f.n_hidden; ++lin) {
    for (int i = tid; i < 512; i += blockDim.x) accv[i] = 0.f;
    __syncthreads();
    for (int im = 0; im < f.n_img; ++im) {
      const FoldImg& g = f.img[im];
      if (g.lin != lin) continue;
      const uint8_t* hi = wl + g.w_off;
      const uint8_t* lo = hi + (g.w_bytes >> 1);
      for (int n = tid; n < g.n_ext; n += blockDim.x) {
        float a = 0.f;
        for (int kc = 0; kc < (g.k_ext >> 3); ++kc) {
          const uint4 vh = *reinterpret_cast<const uint4*>(hi + ((size_t)kc * g.n_ext + n) * 16);
          const uint4 vl = *reinterpret_cast<const uint4*>(lo + ((size_t)kc * g.n_ext + n) * 16);
          const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w}, lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float h0, h1, l0, l1;
            tcx::unpack2(hw[e], h0, h1);
            tcx::unpack2(lw[e], l0, l1);
            const int k = g.k0 + kc * 8 + 2 * e;
            a = fmaf(hvec[k], h0 + l0, a);
            a = fmaf(hvec[k + 1], h1 + l1, a);
          }
        }
        accv[g.n0 + n] += a;   // images of one linear layer cover disjoint (n, k) ranges per thread n: no race
      }
      __syncthreads();
    }
    if (lin < f.n_hidden) {
      for (int n = tid; n < f.hp[lin]; n += blockDim.x) {
        const float b = lc[f.lc_b[lin] + n] + kTanhScale * accv[n];
        out[f.lc_b[lin] + n] = b;
        hvec[n] = (n < f.blk1[lin]) ? tanh_from_scaled(b) : 0.f;
      }
      for (int n = f.hp[lin] + tid; n < 256; n += blockDim.x) hvec[n] = 0.f;
    } else {
      for (int n = tid; n < f.D * f.Mp; n += blockDim.x) out[f.lc_bout + n] = lc[f.lc_bout + n] + accv[n];
    }
    __syncthreads();
  }
  // rank-0 transform parameters -> constants
  if (tid == 0) {
    const float* ro = out + f.lc_bout;
    float* t = out + f.lc_r0c;
    if (f.kind == NAZB_KIND_AFFINE) {
      const float sc = fminf(fmaxf(ro[1], f.clip_lo), f.clip_hi);
      t[0] = ro[0]; t[1] = sc; t[2] = expf(-sc);
    } else {
      const int K = 8;
      const float B = f.bound, min_bin = 1e-3f, min_d = 1e-3f;
      for (int ax = 0; ax < 2; ++ax) {
        const float* v = ro + ax * K;
        float m = v[0];
        for (int j = 1; j < K; ++j) m = fmaxf(m, v[j]);
        float e[8], sum = 0.f;
        for (int j = 0; j < K; ++j) { e[j] = expf(v[j] - m); sum += e[j]; }
        const float inv = (1.f - min_bin * K) / sum;
        float cum = 0.f;
        t[ax * 9] = -B;
        for (int j = 0; j < K; ++j) {
          cum += fmaf(e[j], inv, min_bin);
          t[ax * 9 + j + 1] = (j == K - 1) ? B : fmaf(2.f * B, cum, -B);
        }
      }
      t[18] = 1.f - min_d; t[26] = 1.f - min_d;
      for (int j = 0; j < K - 1; ++j) {
        const float a = ro[2 * K + j];
        t[19 + j] = min_d + (a > 20.f ? a : log1pf(expf(a)));
      }
    }
  }
}
