// v3 inverse (`log_prob` direction) kernel of the tcgen05 engine.  Included by flow_tc.cu inside its
// anonymous namespace (it shares Step / KParamsInv / IoArgs and the helpers defined there).
//
// Same mathematics and the same weight images as the v2 kernel (push-style incremental inverse, two 64-row
// chains per 128-point tile, accumulators resident in TMEM for a whole flow layer); what changes is the
// hand-off structure, which is what bounded v2 (16 bar.sync-separated phases per flow layer, each 2.5-3 k cycles):
//
//   * A chain is 4 epilogue warps (one per TMEM lane quadrant); a warp owns its 16 rows of the chain for
//     EVERY phase, two threads per row (the two half-warps take alternate 8-column chunks via
//     tcgen05.ld.16x32bx2).  Spline -> first conditioner layer of the next stage is therefore warp-local
//     (__syncwarp), and no block-level barrier is left anywhere in the steady state.
//   * Each chain has its own MMA-issuer warp.  The A operand (fp16 hi/lo of the activations of one MADE
//     block) is produced in 16-column K slices; every slice has its own mbarrier, and the issuer fires the
//     slice's MMAs as soon as the slice lands, so MMA issue + execution of all but the last slice hides
//     under the tanh of the following slices.
//   * A push h_j[block r] -> pre_{j+1}[columns >= block r] is split into the CRITICAL part (the columns of
//     block r, which the next epilogue needs: committed to the accumulator barrier) and the DEFERRED part
//     (all later columns), issued after the commit so it overlaps the next epilogue.
//   * A operand blocks are double-buffered; the in-order tensor pipe makes two buffers sufficient
//     (the writer of block p+2 starts after the commit of push p+1, which retires after every MMA of push p).
#pragma once

constexpr int kV3Parts = 2;                          // epilogue warps per (chain, quadrant): part p takes K slices p, p + kV3Parts, ...
constexpr int kV3EpiWarps = kChains * 4 * kV3Parts;  // part 0 warps also own the rows (spline, first layer, outputs)
constexpr int kV3Issuer0 = kV3EpiWarps;              // next kChains warps: MMA issuers of chain 0, 1
constexpr int kV3Producer = kV3EpiWarps + kChains;   // last warp: TMA producer
constexpr int kV3Threads = (kV3EpiWarps + kChains + 1) * 32;
constexpr int kV3MaxSlices = 8;                      // A block <= 128 columns
constexpr int kXSlots = 4;
constexpr float kTanhScale = 2.885390081777927f;     // 2 log2(e): pre-activations are scaled so tanh needs no multiply

#define DBG3(slot)                                                                       \
  if (kDbg && p.dbg && blockIdx.x == 0 && dbg_i < 128) p.dbg[dbg_i * 16 + (slot)] = clk();

// First conditioner layer on CUDA cores: 8 hidden units (n0 .. n0+7) for ONE row whose inputs are in xv.
template <int KINP>
__device__ __forceinline__ void first_chunk(const float* __restrict__ lc, int lc_b0, const float* __restrict__ xin,
                                            int kin, int trow, int n0, float* out) {
  float xv[KINP];
#pragma unroll
  for (int k = 0; k < KINP; ++k) xv[k] = (k < kin) ? xin[k * kTileM + trow] : 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int n = n0 + e;
    const float4* wr = reinterpret_cast<const float4*>(lc + (size_t)n * KINP);
    float a = lc[lc_b0 + n];
#pragma unroll
    for (int k4 = 0; k4 < KINP / 4; ++k4) {
      const float4 w = wr[k4];
      a = fmaf(w.x, xv[k4 * 4 + 0], a);
      a = fmaf(w.y, xv[k4 * 4 + 1], a);
      a = fmaf(w.z, xv[k4 * 4 + 2], a);
      a = fmaf(w.w, xv[k4 * 4 + 3], a);
    }
    out[e] = a;
  }
}

// kMode: 0 = affine, 1 = rational-quadratic spline with K = 8 (two lanes per row), 2 = any other spline.  One instantiation
// per mode keeps the transforms the launch cannot use out of the binary image of the kernel (the ncu report of the
// single-instantiation version showed a 96 % instruction-cache hit rate and 0.26 no-instruction stalls per issue).
template <bool kDbg, int kMode>
__global__ void __launch_bounds__(kV3Threads, 1) flow_tc_inv3_kernel(const __grid_constant__ KParamsInv p,
                                                                      const __grid_constant__ IoArgs io, int n_groups) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);            // [nslots] TMA -> issuers
  uint64_t* w_empty = w_full + 8;                                   // [nslots] count = kChains (one commit per issuer)
  uint64_t* xw_full = w_empty + 8;                                  // [kXSlots] small ring (first-layer images)
  uint64_t* xw_empty = xw_full + kXSlots;                           // [kXSlots] count = kChains
  uint64_t* bar_acc = xw_empty + kXSlots;                           // [kChains] issuer -> epilogue warps
  uint64_t* lc_full = bar_acc + kChains;                            // [2]
  uint64_t* lc_empty = lc_full + 2;                                 // [2], count = kV3EpiWarps
  uint64_t* a_ready = lc_empty + 2;                                 // [kChains][3: buffers 0, 1, A_X][kV3MaxSlices], count = 4 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + kChains * 3 * kV3MaxSlices);
  float* xin = reinterpret_cast<float*>(smem + p.off_xin);          // [kin][128]: ctx rows, then x rows
  float* lcs = reinterpret_cast<float*>(smem + p.off_lc);           // [2][lc_floats]
  float* ycur = reinterpret_cast<float*>(smem + p.off_y);           // [D][128]
  float* xorig = reinterpret_cast<float*>(smem + p.off_xo);         // [D][128]
  float* ljac = reinterpret_cast<float*>(smem + p.off_misc);        // [128]
  float* scratch = reinterpret_cast<float*>(smem + p.off_scratch);  // [32][128] (generic spline only)
  uint8_t* ring = smem + p.off_ring;
  uint8_t* xring = smem + p.off_xring;
  constexpr uint32_t ax_img_bytes = 16 * kChainRows * 2;            // A_X: one K = 16 slice [2 chunks][64 rows][8 halves]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int D = p.D, C = p.C, M = p.M;
  const uint32_t a_img_bytes = (uint32_t)p.kr_max * kChainRows * 2;   // one fp16 image (hi or lo) of an A block
  const uint32_t a_buf_bytes = 2 * a_img_bytes;

  if (tid == 0) {
    for (int i = 0; i < p.nslots; ++i) { tcx::mbar_init(w_full + i, 1); tcx::mbar_init(w_empty + i, kChains); }
    for (int i = 0; i < kChains; ++i) tcx::mbar_init(bar_acc + i, 1);
    for (int i = 0; i < 2; ++i) { tcx::mbar_init(lc_full + i, 1); tcx::mbar_init(lc_empty + i, kV3EpiWarps); }
    for (int i = 0; i < kXSlots; ++i) { tcx::mbar_init(xw_full + i, 1); tcx::mbar_init(xw_empty + i, kChains); }
    // slice 0 of every A block (and the A_X slice) also collects one arrival from each NON-producing warp of the chain:
    // a warp that waits on the accumulator barrier must be needed for the next MMA, otherwise the issuer could complete
    // two accumulator phases before a late warp (cold instruction / constant cache) has observed the first one, and the
    // parity wait would never return
    for (int i = 0; i < kChains * 3 * kV3MaxSlices; ++i)
      tcx::mbar_init(a_ready + i, (i % kV3MaxSlices == 0) ? 4 * kV3Parts : 4);
    tcx::mbar_fence_init();
  }
  if (warp == 0) tcx::tmem_alloc(tmem_slot, kTmemCols);
  for (uint32_t i = tid; i < (kChains * 2 * a_buf_bytes) / 16; i += kV3Threads)
    reinterpret_cast<uint4*>(smem + p.off_h)[i] = make_uint4(0, 0, 0, 0);
  for (uint32_t i = tid; i < (kChains * 2 * ax_img_bytes) / 16; i += kV3Threads)
    reinterpret_cast<uint4*>(smem + p.off_ax)[i] = make_uint4(0, 0, 0, 0);
  tcx::fence_async_smem();
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n_tiles = (io.N + kTileM - 1) / kTileM;
  const long long n_items = (long long)n_tiles * n_groups;

  if (warp == kV3Producer) {
    // ===================== TMA producer: weight images + layer constants =====================
    if (lane == 0) {
      uint32_t cnt = 0, xcnt = 0, lcnt = 0;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = (int)(item / n_tiles);
        for (int si = grp; si < io.s_count; si += n_groups) {
          const uint8_t* wdraw = p.wimg + (size_t)(io.s_begin + si) * p.draw_bytes;
          const float* lcdraw = p.lc + (size_t)(io.s_begin + si) * p.L * p.lc_floats;
          for (int li = 0; li < p.L; ++li) {
            const int l = p.L - 1 - li;
            {
              const uint32_t b = lcnt & 1, use = lcnt >> 1;
              tcx::mbar_wait(lc_empty + b, (use & 1) ^ 1);
              tcx::mbar_expect_tx(lc_full + b, (uint32_t)p.lc_floats * 4);
              tcx::bulk_g2s(lcs + (size_t)b * p.lc_floats, lcdraw + (size_t)l * p.lc_floats, (uint32_t)p.lc_floats * 4, lc_full + b);
              ++lcnt;
            }
            const uint8_t* wl = wdraw + (size_t)l * p.layer_bytes;
            for (int st = 0; st < p.nsteps; ++st) {
              const uint32_t wb = p.steps[st].w_bytes;
              if (wb == 0) continue;
              if (p.steps[st].a_buf == A_X) {
                const uint32_t slot = xcnt % kXSlots, use = xcnt / kXSlots;
                tcx::mbar_wait(xw_empty + slot, (use & 1) ^ 1);
                tcx::mbar_expect_tx(xw_full + slot, wb);
                tcx::bulk_g2s(xring + (size_t)slot * p.xslot_bytes, wl + p.steps[st].w_off, wb, xw_full + slot);
                ++xcnt;
                continue;
              }
              const uint32_t slot = cnt % p.nslots, use = cnt / p.nslots;
              tcx::mbar_wait(w_empty + slot, (use & 1) ^ 1);
              tcx::mbar_expect_tx(w_full + slot, wb);
              tcx::bulk_g2s(ring + (size_t)slot * kSlotBytes, wl + p.steps[st].w_off, wb, w_full + slot);
              ++cnt;
            }
          }
        }
      }
    }
  } else if (warp >= kV3Issuer0) {
    // ===================== MMA issuer of chain `ch` =====================
    // Whole warp convergent, tcgen05 instructions predicated on the elected lane (descriptors stay uniform).
    const int ch = warp - kV3Issuer0;
    const uint32_t elected = tcx::elect_one();
    const uint32_t ring_a = tcx::smem_u32(ring), xring_a = tcx::smem_u32(xring);
    const uint32_t a_base = tcx::smem_u32(smem + p.off_h) + (uint32_t)ch * 2 * a_buf_bytes;
    const uint32_t ax_base = tcx::smem_u32(smem + p.off_ax) + (uint32_t)ch * 2 * ax_img_bytes;
    constexpr uint32_t lbo_a = kChainRows * 16;
    constexpr uint64_t dhi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
    const uint32_t d_lane = tmem + ((uint32_t)(ch * 16) << 16);
    uint64_t* my_acc = bar_acc + ch;
    uint64_t* my_ready = a_ready + ch * 3 * kV3MaxSlices;
    uint32_t slot = 0, use = 0, xslot = 0, xuse = 0, buf = 0, apar = 0;   // apar: one parity bit per (buffer, slice) barrier
    int dbg_i = 0;
    const bool dbg_me = (ch == 0 && lane == 0);
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int grp = (int)(item / n_tiles);
      for (int si = grp; si < io.s_count; si += n_groups) {
        for (int li = 0; li < p.L; ++li) {
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_wbytes = p.steps[st].w_bytes;
            if (s_wbytes == 0) { ++dbg_i; continue; }
            const uint32_t s_n = p.steps[st].n, s_ncrit = p.steps[st].n_crit, s_dcol = p.steps[st].d_col;
            const int ksteps = p.steps[st].ksteps;
            const uint32_t s_acc = p.steps[st].accumulate, s_last = (p.steps[st].epi != EPI_NONE);
            const uint32_t slice0 = p.steps[st].a_chunk0 >> 1;
            const bool is_x = (p.steps[st].a_buf == A_X);
            const uint32_t bsel = is_x ? 2u : buf;
            const uint32_t n_rest = s_n - s_ncrit;
            const uint32_t idesc_c = tcx::make_idesc_f16_m64(s_ncrit);
            const uint32_t idesc_r = tcx::make_idesc_f16_m64(n_rest);
            const uint32_t lbo_b = s_n * 16;
            const uint32_t b_hi = is_x ? (xring_a + xslot * (uint32_t)p.xslot_bytes) : (ring_a + slot * kSlotBytes);
            const uint32_t b_lo = b_hi + (s_wbytes >> 1);
            const uint32_t lbo_b_hi16 = (lbo_b >> 4) << 16, lbo_a_hi16 = (lbo_a >> 4) << 16;
            const uint32_t a_hi = is_x ? ax_base : (a_base + buf * a_buf_bytes);
            const uint32_t a_lo = a_hi + (is_x ? ax_img_bytes : a_img_bytes);
            const uint32_t d_c = d_lane + s_dcol, d_r = d_c + s_ncrit;
            const uint32_t ro = s_ncrit;              // image row n_crit = byte offset n_crit * 16, >> 4
            uint64_t* wf = is_x ? (xw_full + xslot) : (w_full + slot);
            uint64_t* we = is_x ? (xw_empty + xslot) : (w_empty + slot);
            if (dbg_me) { DBG3(8) }
            tcx::mbar_wait(wf, (is_x ? xuse : use) & 1);
            if (dbg_me) { DBG3(9) }
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t sl = slice0 + k, bit = 1u << (bsel * kV3MaxSlices + sl);
              tcx::mbar_wait(my_ready + bsel * kV3MaxSlices + sl, (apar & bit) ? 1u : 0u);
              apar ^= bit;
              tcx::tc_fence_after();
              if (dbg_me && k == 0) { DBG3(10) }
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
            tcx::mma_commit_elect(we, elected);   // the slot is free once these MMAs retire
            if (dbg_me) { DBG3(11) }
            if (is_x) { if (++xslot == kXSlots) { xslot = 0; ++xuse; } }
            else {
              if (++slot == (uint32_t)p.nslots) { slot = 0; ++use; }
              if (s_last) buf ^= 1;
            }
            ++dbg_i;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps: 4 per chain, one per TMEM lane quadrant =====================
    const int ch = warp / (4 * kV3Parts), part = (warp >> 2) % kV3Parts, q = warp & 3;
    const int hw = lane >> 4, lr = lane & 15;
    const int crow = q * 16 + lr;                  // row inside the chain's 64-row sub-tile
    const int trow = ch * kChainRows + crow;       // row inside the 128-point tile
    const int wrow0 = ch * kChainRows + q * 16;    // first tile row owned by this warp
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32 + ch * 16) << 16);
    constexpr bool spline = kMode != 0;
    constexpr bool fast_rqs = kMode == 1;
    const bool rows_mine = (part == 0);            // this warp owns the per-row state of its 16 rows
    const bool owner = rows_mine && (hw == 0);
    const bool xf = p.xf != 0;
    float* scr = scratch + trow;
    auto raw = [&](int m) { return scr[m * kTileM]; };
    auto setw = [&](int m, float v) { scr[m * kTileM] = v; };
    uint8_t* a_chain = smem + p.off_h + (size_t)ch * 2 * a_buf_bytes;
    uint8_t* ax_chain = smem + p.off_ax + (size_t)ch * 2 * ax_img_bytes;
    uint64_t* my_acc = bar_acc + ch;
    uint64_t* my_ready = a_ready + ch * 3 * kV3MaxSlices;
    uint32_t par_acc = 0, lcnt = 0, buf = 0;
    int dbg_i = 0;
    const bool dbg_me = (tid == 0);
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
      if (lane == 0) tcx::mbar_arrive(my_ready + buf * kV3MaxSlices + sl);
    };
    // one element (this row, input column col) of the first-layer A operand [ctx | x | 1]
    auto ax_store = [&](int col, float v) {
      const float c = fminf(fmaxf(v, -65504.f), 65504.f);
      const __half h = __float2half_rn(c);
      const __half l = __float2half_rn(c - __half2float(h));
      uint8_t* dst = ax_chain + ((size_t)(col >> 3) * kChainRows + crow) * 16 + (col & 7) * 2;
      *reinterpret_cast<__half*>(dst) = h;
      *reinterpret_cast<__half*>(dst + ax_img_bytes) = l;
    };

    if (ch == 1 && p.phase_delay > 0) { const long long t0 = clk(); while (clk() - t0 < p.phase_delay) { } }
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = (int)(item % n_tiles), grp = (int)(item / n_tiles);
      const int n0 = tile * kTileM;
      const int npts = min(kTileM, io.N - n0);
      float run_m = -INFINITY, run_s = 0.f;
      // ---- tile load: the 16 rows of this warp ----
      __syncwarp();
      for (int i = lane; i < 16 * C && rows_mine; i += 32) {
        const int pt = wrow0 + i / C, c = i % C;
        float v = 0.f;
        if (pt < npts) v = io.ctx[((io.ctx_rows == 1) ? 0 : (size_t)(n0 + pt)) * C + c];
        xin[c * kTileM + pt] = v;
      }
      for (int i = lane; i < 16 * D && rows_mine; i += 32) {
        const int pt = wrow0 + i / D, d = i % D;
        xorig[d * kTileM + pt] = (pt < npts) ? io.x[(size_t)(n0 + pt) * D + d] : 0.f;
      }
      __syncwarp();
      if (owner) {
        float lj = 0.f;
        if (io.lo != nullptr && trow < npts)
          for (int d = 0; d < D; ++d) xorig[d * kTileM + trow] = nazb::bound_fwd(xorig[d * kTileM + trow], io.lo[d], io.hi[d], lj);
        ljac[trow] = lj;
        if (xf) {   // constant part of the first-layer A operand: context and the bias column
          for (int c = 0; c < C; ++c) ax_store(c, xin[c * kTileM + trow]);
          ax_store(p.kin, 1.f);
        }
      }

      for (int si = grp; si < io.s_count; si += n_groups) {
        // ---- draw start ----
        float ld_acc = 0.f;
        if (owner)
          for (int d = 0; d < D; ++d) {
            ycur[d * kTileM + trow] = xorig[d * kTileM + trow];
            xin[(C + d) * kTileM + trow] = 0.f;
            if (xf) ax_store(C + d, 0.f);
          }
        __syncwarp();

        for (int li = 0; li < p.L; ++li) {
          const int l = p.L - 1 - li;
          const int* perm = p.perm + l * D;
          const float* lc = lcs + (size_t)(lcnt & 1) * p.lc_floats;
          tcx::mbar_wait(lc_full + (lcnt & 1), (lcnt >> 1) & 1);
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_epi = p.steps[st].epi;
            if (s_epi == EPI_NONE) { ++dbg_i; continue; }   // K-split sub-step: nothing to do on this side
            const uint32_t s_ecol = p.steps[st].e_col, s_encols = p.steps[st].e_ncols, s_eaux = p.steps[st].e_aux;
            const uint32_t s_stage = p.steps[st].stage, s_flags = p.steps[st].flags;
            if (dbg_me) { DBG3(0) }
            if (p.steps[st].w_bytes) {
              tcx::mbar_wait(my_acc, par_acc);
              par_acc ^= 1;
              tcx::tc_fence_after();
            }
            if (dbg_me) { DBG3(1) }
            if (s_epi == EPI_TANH) {
              // block of nch 8-column chunks = nsl K slices; slice s = chunks {2s (half-warp 0), 2s+1 (half-warp 1)}
              const int nch = s_encols >> 3, nsl = (nch + 1) >> 1;
              const bool prescaled = (s_flags & 4) != 0;   // first-layer block: bias and tanh scale folded into the image
              const int nmine = (nsl - part + kV3Parts - 1) / kV3Parts;   // slices part, part + kV3Parts, ...
              if (part != 0 && lane == 0) tcx::mbar_arrive(my_ready + buf * kV3MaxSlices);   // observer arrival on slice 0
              for (int j0 = 0; j0 < nmine; j0 += 2) {
                uint32_t r[16];
                const int nj = min(2, nmine - j0);
                const int sl0 = part + j0 * kV3Parts;
                const uint32_t ta = lane_base + s_ecol + sl0 * 16;
                tcx::tmem_ld16x2_8<8>(ta, r);
                if (nj > 1) tcx::tmem_ld16x2_8<8>(ta + 16 * kV3Parts, r + 8);
                tcx::tmem_ld_wait();
                tcx::tc_fence_before();
                if (dbg_me && j0 == 0) { DBG3(2) }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  if (u < nj) {
                    const int sl = sl0 + u * kV3Parts;
                    const int c = sl * 2 + hw;
                    const int cl = min(c, nch - 1);
                    const uint32_t* ru = r + 8 * u;
                    uint64_t s2[4];
                    if (prescaled) {
#pragma unroll
                      for (int i = 0; i < 4; ++i) s2[i] = tcx::pk2(__uint_as_float(ru[2 * i]), __uint_as_float(ru[2 * i + 1]));
                    } else {
                      const ulonglong2* bv = reinterpret_cast<const ulonglong2*>(lc + s_eaux + cl * 8);
                      const ulonglong2 b0 = bv[0], b1 = bv[1];   // biases already multiplied by 2 log2 e
                      s2[0] = tcx::fma2(tcx::pk2(__uint_as_float(ru[0]), __uint_as_float(ru[1])), scale2, b0.x);
                      s2[1] = tcx::fma2(tcx::pk2(__uint_as_float(ru[2]), __uint_as_float(ru[3])), scale2, b0.y);
                      s2[2] = tcx::fma2(tcx::pk2(__uint_as_float(ru[4]), __uint_as_float(ru[5])), scale2, b1.x);
                      s2[3] = tcx::fma2(tcx::pk2(__uint_as_float(ru[6]), __uint_as_float(ru[7])), scale2, b1.y);
                    }
                    uint4 hi4, lo4;
                    tcx::tanh8_scaled(s2, hi4, lo4);
                    if (c >= nch) { hi4 = make_uint4(0, 0, 0, 0); lo4 = hi4; }   // K padding chunk
                    if (dbg_me && j0 == 0 && u == 0) { DBG3(4) }
                    store_chunk(c, hi4, lo4);
                    if (dbg_me && j0 == 0 && u == 0) { DBG3(5) }
                    publish(sl);
                    if (dbg_me && j0 == 0 && u == 0) { DBG3(6) }
                  }
                }
              }
              buf ^= 1;
            } else if (s_epi == EPI_FIRST) {
              if (xf) {
                // tensor-core first layer: the A operand [ctx | x | 1] was updated in place by the row owners
                if (rows_mine) {
                  tcx::fence_async_smem();
                  __syncwarp();
                }
                if (lane == 0) tcx::mbar_arrive(my_ready + 2 * kV3MaxSlices);   // other parts: observer arrival
              } else {
                const int nch = s_encols >> 3, nsl = (nch + 1) >> 1;
                const int u0 = s_eaux;
                if (part != 0 && lane == 0) tcx::mbar_arrive(my_ready + buf * kV3MaxSlices);   // observer arrival on slice 0
                for (int sl = 0; sl < nsl && rows_mine; ++sl) {   // (the row's inputs live with the part-0 warp)
                  const int c = sl * 2 + hw;
                  uint4 hi4 = make_uint4(0, 0, 0, 0), lo4 = hi4;
                  if (c < nch) {
                    float ra[8];
                    switch (p.kinp) {   // W0 and b0 are stored multiplied by 2 log2 e
                      case 4: first_chunk<4>(lc, p.lc_b0, xin, p.kin, trow, u0 + c * 8, ra); break;
                      case 8: first_chunk<8>(lc, p.lc_b0, xin, p.kin, trow, u0 + c * 8, ra); break;
                      case 12: first_chunk<12>(lc, p.lc_b0, xin, p.kin, trow, u0 + c * 8, ra); break;
                      default: first_chunk<16>(lc, p.lc_b0, xin, p.kin, trow, u0 + c * 8, ra); break;
                    }
                    uint64_t s2[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) s2[i] = tcx::pk2(ra[2 * i], ra[2 * i + 1]);
                    tcx::tanh8_scaled(s2, hi4, lo4);
                  }
                  store_chunk(c, hi4, lo4);
                  publish(sl);
                }
                buf ^= 1;
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
                if (dbg_me) { DBG3(2) }
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
                ld_acc += ld;
                xin[(C + d) * kTileM + trow] = xv;
                if (r == D - 1) {
                  // end of this flow layer: x becomes the y of the next (earlier) layer, x restarts at 0
                  for (int dd = 0; dd < D; ++dd) {
                    ycur[dd * kTileM + trow] = xin[(C + dd) * kTileM + trow];
                    xin[(C + dd) * kTileM + trow] = 0.f;
                    if (xf) ax_store(C + dd, 0.f);
                  }
                } else if (xf) {
                  ax_store(C + d, xv);
                }
              }
              __syncwarp();   // the row's second thread reads xin in the next FIRST phase
            }
            if (dbg_me) { DBG3(3) }
            ++dbg_i;
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
