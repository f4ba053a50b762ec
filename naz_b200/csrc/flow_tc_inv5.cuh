// v5 inverse (`log_prob` direction) kernel of the tcgen05 engine: ONE 128-row chain per CTA, M = 128 MMAs.
// Included by flow_tc.cu after flow_tc_inv4.cuh (shares KParamsInv4, mbar_wait4, the fold kernel and the helpers).
//
// Why (measured on B200 with the event log of this file, tools/inv5_timeline.py, on the v4 kernel): the two 64-row chains
// of v3 / v4 run in lockstep, so they do not hide each other's latencies, while every M = 64 MMA occupies the tensor pipe
// and the shared-memory operand path as long as an M = 128 one.  A push of 4 K slices took ~550-850 cycles per slice in
// the issuer (6 small dependent MMAs per slice and chain, ~4 KB of operand reads each against 128 B/clk of shared-memory
// bandwidth), and that backlog — not the tanh epilogue — was the longest segment of every phase.  Here:
//   * one chain of 128 rows: half the MMAs and half the operand bytes per row; all 16 epilogue warps work on the same
//     phase (4 TMEM lane quadrants x 2 row halves x 2 slice lanes);
//   * pushes are issued unsplit (one accumulator region of N % 16 == 0 columns, N >= 112 for the wide ones, so consecutive
//     accumulating MMAs pipeline instead of waiting on each other), guarded by a fresh elect.sync so that ptxas keeps the
//     descriptors in uniform registers (UIADD3 + UTCHMMA back to back; the v3 / v4 form re-broadcast 7-8 operands per MMA);
//   * everything else as v4: context fold, first conditioner layer on CUDA cores, watchdog'd waits, draw-group gate.
// Row mapping: epilogue warp w: quadrant q = w & 3, part = w >> 2; rows 32 q + 16 (part & 1) + (lane & 15), two lanes per
// row (tcgen05.ld.16x32bx2); slice lane = part >> 1 takes K slices (part >> 1), (part >> 1) + 2, ...; the warps with
// part < 2 own the per-row state (spline, x, outputs).
#pragma once

constexpr int kV5EpiWarps = 16;
constexpr int kV5Issuer = kV5EpiWarps;
constexpr int kV5Producer = kV5EpiWarps + 1;
constexpr int kV5Threads = (kV5EpiWarps + 2) * 32;
constexpr int kV5MaxSlices = 8;
constexpr int kV5MaxPairs = kV5MaxSlices / 2;   // a_ready barriers: one per PAIR of K slices (2 j, 2 j + 1), 16 arrivals each
constexpr int kDbgEvents = 4096;
// Debug event log (kDbg instantiation only): CTA 0, three logger warps (slot 0 = epilogue warp 0 (q 0, rows 0-15, slice lane 0),
// slot 1 = epilogue warp 8 (same rows, slice lane 1), slot 2 = the issuer), each writing (clock, step << 8 | event) pairs.
#define LOG5(slot, ev)                                                                                    \
  if (kDbg && dbg_on && lane == 0 && dbg_n < kDbgEvents) {                                                \
    p.dbg[((size_t)(slot) * kDbgEvents + dbg_n) * 2] = clk();                                             \
    p.dbg[((size_t)(slot) * kDbgEvents + dbg_n) * 2 + 1] = ((long long)st << 8) | (ev);                   \
    ++dbg_n;                                                                                              \
  }

// kMode: 0 = affine, 1 = rational-quadratic spline with K = 8 (two lanes per row), 2 = any other spline.
// kATmem: the A operand (fp16 hi / lo of the activations) lives in tensor memory (written with tcgen05.st, read by the MMA)
// instead of shared memory: measured (tools/tc_probe_rate3.cu) an M = 128 MMA costs max(N/2, 32 + N/4) cycles with A in shared
// memory (operand fetch at 128 B/clk) and the N/2 pipe floor with A in TMEM; in the kernel the shared-memory A path ran at
// ~130 cycles per MMA because the epilogue warps' own traffic shares that port.
// kSplit (A in TMEM, programs with split pushes): 1 = single A buffer, writers of an A block wait for the a_free barrier that
// the issuer commits behind the trailing MMAs; 2 = DOUBLE-buffered A (needs 2 kr_max free columns), the trailing MMAs are held
// back until every epilogue warp has finished reading its accumulators (ld_done barrier in the a_free slot): the
// accumulator reads of the critical path no longer share the TMEM ports with MMAs.  A template parameter, not a runtime
// flag: the mere presence of the hand-shake code cost 4 % on programs that never use it (81.4 -> 77.8).
// kCtx: the program adds a per-point context in the first-layer phase (false for context-folded programs: their instantiation
// carries no context code — its presence alone moved the folded cfg3 path by 1.6 %).
template <bool kDbg, int kMode, bool kATmem, int kSplit = 0, bool kCtx = true>
__global__ void __launch_bounds__(kV5Threads, 1) flow_tc_inv5_kernel(const __grid_constant__ KParamsInv4 p,
                                                                      const __grid_constant__ IoArgs io, int n_groups) {
  // kSplit = 3: double-buffered A, trailing MMAs right behind the critical ones, NO hand-shake: the writers of block k + 2
  // start after the accumulator barrier of push k + 1, whose critical MMAs sit behind the trailing ones of push k in the
  // in-order tensor pipe, so the buffer they overwrite is no longer read
  constexpr bool kAFree = (kSplit == 1), kDefer = (kSplit == 2), kDbl = (kSplit >= 2);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);            // [nslots] TMA -> issuers
  uint64_t* w_empty = w_full + 8;                                   // [nslots] count = 1 (the issuer's commit)
  uint64_t* bar_acc = w_empty + 8;                                  // [1] issuer -> epilogue warps
  uint64_t* lc_full = bar_acc + 2;                                  // [2]
  uint64_t* lc_empty = lc_full + 2;                                 // [2], count = kV5EpiWarps
  uint64_t* a_ready = lc_empty + 2;                                 // [2 barrier sets][kV5MaxPairs]
  uint64_t* a_free = a_ready + 2 * kV5MaxPairs;                     // [1] A in TMEM: commit behind the last MMA that reads an A block
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 1);
  volatile int* xflag = reinterpret_cast<volatile int*>(tmem_slot + 4);   // [8] x publications of each 16-row group (owner warp -> partner warp)
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
  const uint32_t a_img_bytes = (uint32_t)p.kr_max * kTileM * 2;   // one fp16 image (hi or lo) of an A block: [chunk][128 rows][8 halves]
  const uint32_t a_buf_bytes = 2 * a_img_bytes;

  if (tid == 0) {
    for (int i = 0; i < p.nslots; ++i) { tcx::mbar_init(w_full + i, 1); tcx::mbar_init(w_empty + i, 1); }
    tcx::mbar_init(bar_acc, 1);
    for (int i = 0; i < 2; ++i) { tcx::mbar_init(lc_full + i, 1); tcx::mbar_init(lc_empty + i, kV5EpiWarps); }
    // slice 0 of every A block also collects one arrival from each NON-producing warp of the chain: a warp that waits
    // on the accumulator barrier must be needed for the next MMA, otherwise the issuer could complete two accumulator
    // phases before a late warp has observed the first one and its parity wait would never return
    // pair j = K slices 2 j (slice lane 0) and 2 j + 1 (slice lane 1): EVERY epilogue warp arrives exactly once per pair and push
    // (after publishing its slice, or as an observer when its lane has no slice in the pair).  A warp that waits on the
    // accumulator barrier is therefore always needed for the next MMA group and can never be lapped by two accumulator phases.
    for (int i = 0; i < 2 * kV5MaxPairs; ++i) tcx::mbar_init(a_ready + i, kV5EpiWarps);
    tcx::mbar_init(a_free, kDefer ? kV5EpiWarps : 1);   // kDefer: the same slot is the ld_done barrier (one arrival per epilogue warp)
    for (int i = 0; i < 8; ++i) xflag[i] = 0;
    tcx::mbar_fence_init();
  }
  if (warp == 0) tcx::tmem_alloc(tmem_slot, kTmemCols);
  for (uint32_t i = tid; i < (2 * a_buf_bytes) / 16; i += kV5Threads)
    reinterpret_cast<uint4*>(smem + p.off_h)[i] = make_uint4(0, 0, 0, 0);
  tcx::fence_async_smem();
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tmem != 0) __trap();   // 512 of 512 columns: the allocation is the whole tensor memory (the issuer assumes base 0)

  const int n_tiles = (io.N + kTileM - 1) / kTileM;
  const long long n_items = (long long)n_tiles * n_groups;

  if (warp == kV5Producer) {
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
              if (p.park == 2) mbar_wait4_sleepy(lc_empty + b, (use & 1) ^ 1, p.wd, WD_TAG(1)); else if (p.park == 1) mbar_wait4_parked(lc_empty + b, (use & 1) ^ 1, p.wd, WD_TAG(1)); else mbar_wait4(lc_empty + b, (use & 1) ^ 1, p.wd, WD_TAG(1));
              tcx::mbar_expect_tx(lc_full + b, (uint32_t)p.lc_floats * 4);
              tcx::bulk_g2s(lcs + (size_t)b * p.lc_floats, lcdraw + (size_t)l * p.lc_floats, (uint32_t)p.lc_floats * 4, lc_full + b);
              ++lcnt;
            }
            const uint8_t* wl = wdraw + (size_t)l * p.layer_bytes;
            for (int st = 0; st < p.nsteps; ++st) {
              const uint32_t wb = p.steps[st].w_bytes;
              if (wb == 0) continue;
              const uint32_t slot = cnt % p.nslots, use = cnt / p.nslots;
              if (p.park == 2) mbar_wait4_sleepy(w_empty + slot, (use & 1) ^ 1, p.wd, WD_TAG(2)); else if (p.park == 1) mbar_wait4_parked(w_empty + slot, (use & 1) ^ 1, p.wd, WD_TAG(2)); else mbar_wait4(w_empty + slot, (use & 1) ^ 1, p.wd, WD_TAG(2));
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
  } else if (warp == kV5Issuer) {
    // ===================== MMA issuer =====================
    // Whole warp convergent through the waits; each group of MMAs is guarded by a fresh elect.sync (CUTLASS idiom).
    const uint32_t ring_a = tcx::smem_u32(ring);
    const uint32_t a_base = tcx::smem_u32(smem + p.off_h);
    constexpr uint32_t lbo_a = kTileM * 16;
    uint32_t slot = 0, use = 0, buf = 0, apar = 0;   // apar: one parity bit per (buffer, slice) barrier
    [[maybe_unused]] uint32_t par_ld = 0;
    int dbg_n = 0;
    const int iss_slot = p.dbg_all ? kV5EpiWarps : 2;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int grp = (int)(item / n_tiles);
      for (int si = grp; si < io.s_count; si += n_groups) {
        for (int li = 0; li < p.L; ++li) {
          const bool dbg_on = kDbg && p.dbg != nullptr && blockIdx.x == 0 && item == blockIdx.x && si != grp && li >= 2;
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_wbytes = p.steps[st].w_bytes;
            if (s_wbytes == 0) continue;
            const uint32_t s_n = p.steps[st].n, s_ncrit = p.steps[st].n_crit, s_dcol = p.steps[st].d_col;
            const int ksteps = p.steps[st].ksteps;
            const uint32_t s_acc = p.steps[st].accumulate, s_last = (p.steps[st].epi != EPI_NONE);
            const uint32_t slice0 = p.steps[st].a_chunk0 >> 1;
            const uint32_t n_rest = s_n - s_ncrit;
            const uint32_t idesc_c = tcx::make_idesc_f16(s_ncrit);
            const uint32_t idesc_r = tcx::make_idesc_f16(n_rest);
            // The issuer is ONE thread running a serial instruction stream (~5 cycles per dependent instruction): what it
            // executes per K slice bounds the MMA rate long before the tensor pipe does (tools/tc_probe_rate2.cu: 123 cycles
            // per MMA for a 25-instruction loop body whatever M, N or the operand source).  So the slice loop only waits,
            // adds constants to the low words of four running descriptors and fires the MMAs.
            constexpr uint32_t dhi32 = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1
            const uint32_t a_step = (2u * lbo_a) >> 4, b_step = (2u * s_n * 16u) >> 4;
            uint32_t da_h = (((a_base + buf * a_buf_bytes) + slice0 * 2u * lbo_a) >> 4) | ((lbo_a >> 4) << 16);
            uint32_t da_l = da_h + (a_img_bytes >> 4);
            uint32_t db_h = ((ring_a + slot * kSlotBytes) >> 4) | (((s_n * 16u) >> 4) << 16);
            uint32_t db_l = db_h + (s_wbytes >> 5);
            const uint32_t d_c = s_dcol, d_r = d_c + s_ncrit;   // TMEM base is 0 (checked above): lane 0, column d_col
            const uint32_t ro = s_ncrit;                        // image row n_crit = byte offset n_crit * 16, >> 4
            // A in TMEM: ONE buffer (every MMA of a push has retired before the epilogue that writes the next block passes
            // the accumulator barrier: pushes are unsplit), K slice s at columns t_a + 8 s (hi) / t_a + kr_max / 2 + 8 s (lo)
            uint32_t ta_h = p.t_a + (kDbl ? buf * (uint32_t)p.kr_max : 0u) + slice0 * 8u, ta_l = ta_h + (uint32_t)p.kr_max / 2u;
            const uint32_t pair0 = slice0 >> 1;                 // sub-steps of a K-split push start at an even slice
            uint64_t* rdy = a_ready + buf * kV5MaxPairs + pair0;
            uint32_t bit = 1u << (buf * kV5MaxPairs + pair0);
            auto desc = [](uint32_t lo) { return ((uint64_t)dhi32 << 32) | lo; };
            LOG5(iss_slot, 20)
            if (p.park == 2) mbar_wait4_sleepy(w_full + slot, use & 1, p.wd, WD_TAG(3)); else if (p.park == 1) mbar_wait4_parked(w_full + slot, use & 1, p.wd, WD_TAG(3)); else mbar_wait4(w_full + slot, use & 1, p.wd, WD_TAG(3));
            LOG5(iss_slot, 21)
            for (int k = 0; k < ksteps; k += 2) {
              if (p.park == 2) mbar_wait4_sleepy(rdy, (apar & bit) ? 1u : 0u, p.wd, WD_TAG(4)); else if (p.park == 1) mbar_wait4_parked(rdy, (apar & bit) ? 1u : 0u, p.wd, WD_TAG(4)); else mbar_wait4(rdy, (apar & bit) ? 1u : 0u, p.wd, WD_TAG(4));
              apar ^= bit;
              tcx::tc_fence_after();
              LOG5(iss_slot, 32 + k)
              const bool two = (k + 1 < ksteps);
              const bool lastk = (k + 2 >= ksteps);
              if (tcx::elect_one()) {
                // hi*hi + hi*lo + lo*hi for each K slice of the pair
                const uint32_t acc0 = (k == 0) ? s_acc : 1u;
                if (kATmem) {
                  tcx::mma_f16_ts(d_c, ta_h, desc(db_h), idesc_c, acc0);
                  tcx::mma_f16_ts(d_c, ta_h, desc(db_l), idesc_c, 1u);
                  tcx::mma_f16_ts(d_c, ta_l, desc(db_h), idesc_c, 1u);
                  if (two) {
                    tcx::mma_f16_ts(d_c, ta_h + 8, desc(db_h + b_step), idesc_c, 1u);
                    tcx::mma_f16_ts(d_c, ta_h + 8, desc(db_l + b_step), idesc_c, 1u);
                    tcx::mma_f16_ts(d_c, ta_l + 8, desc(db_h + b_step), idesc_c, 1u);
                  }
                } else {
                  tcx::mma_f16_ss(d_c, desc(da_h), desc(db_h), idesc_c, acc0);
                  tcx::mma_f16_ss(d_c, desc(da_h), desc(db_l), idesc_c, 1u);
                  tcx::mma_f16_ss(d_c, desc(da_l), desc(db_h), idesc_c, 1u);
                  if (two) {
                    tcx::mma_f16_ss(d_c, desc(da_h + a_step), desc(db_h + b_step), idesc_c, 1u);
                    tcx::mma_f16_ss(d_c, desc(da_h + a_step), desc(db_l + b_step), idesc_c, 1u);
                    tcx::mma_f16_ss(d_c, desc(da_l + a_step), desc(db_h + b_step), idesc_c, 1u);
                  }
                }
                if (s_last && lastk) tcx::mma_commit(bar_acc);
                if (n_rest && kATmem && kSplit != 0) {
                  // split pushes, A in TMEM: the remaining columns of ALL K slices of this step go behind the critical MMAs of
                  // its last pair (issued per pair they would sit in front of the next pair's critical MMAs in the in-order pipe);
                  // kDefer holds those of a push's last sub-step back further (below, after the ld_done barrier)
                  if (lastk && !(kDefer && s_last)) {
                    for (int kk = 0; kk < ksteps; ++kk) {
                      const uint32_t dk = (uint32_t)(kk - k);   // K slices relative to the running descriptors (modular arithmetic)
                      const uint32_t a0 = (kk == 0) ? s_acc : 1u;
                      tcx::mma_f16_ts(d_r, ta_h + 8u * dk, desc(db_h + b_step * dk + ro), idesc_r, a0);
                      tcx::mma_f16_ts(d_r, ta_h + 8u * dk, desc(db_l + b_step * dk + ro), idesc_r, 1u);
                      tcx::mma_f16_ts(d_r, ta_l + 8u * dk, desc(db_h + b_step * dk + ro), idesc_r, 1u);
                    }
                  }
                } else if (n_rest) {   // split pushes: the remaining columns, behind the critical ones
                  if (kATmem) {
                    tcx::mma_f16_ts(d_r, ta_h, desc(db_h + ro), idesc_r, acc0);
                    tcx::mma_f16_ts(d_r, ta_h, desc(db_l + ro), idesc_r, 1u);
                    tcx::mma_f16_ts(d_r, ta_l, desc(db_h + ro), idesc_r, 1u);
                    if (two) {
                      tcx::mma_f16_ts(d_r, ta_h + 8, desc(db_h + b_step + ro), idesc_r, 1u);
                      tcx::mma_f16_ts(d_r, ta_h + 8, desc(db_l + b_step + ro), idesc_r, 1u);
                      tcx::mma_f16_ts(d_r, ta_l + 8, desc(db_h + b_step + ro), idesc_r, 1u);
                    }
                  } else {
                    tcx::mma_f16_ss(d_r, desc(da_h), desc(db_h + ro), idesc_r, acc0);
                    tcx::mma_f16_ss(d_r, desc(da_h), desc(db_l + ro), idesc_r, 1u);
                    tcx::mma_f16_ss(d_r, desc(da_l), desc(db_h + ro), idesc_r, 1u);
                    if (two) {
                      tcx::mma_f16_ss(d_r, desc(da_h + a_step), desc(db_h + b_step + ro), idesc_r, 1u);
                      tcx::mma_f16_ss(d_r, desc(da_h + a_step), desc(db_l + b_step + ro), idesc_r, 1u);
                      tcx::mma_f16_ss(d_r, desc(da_l + a_step), desc(db_h + b_step + ro), idesc_r, 1u);
                    }
                  }
                }
                if (lastk && !(kDefer && s_last && n_rest)) tcx::mma_commit(w_empty + slot);   // the slot is free once these MMAs retire
                // A in TMEM is single-buffered: the writers of the next block wait for this before their first store
                // (with unsplit pushes it has completed before they pass the accumulator barrier; with split ones the
                // trailing MMAs are still reading the block at that point)
                if (kATmem && kAFree && s_last && lastk) tcx::mma_commit(a_free);
              }
              __syncwarp();
              if (kDefer && kATmem && lastk && s_last && n_rest) {
                // the epilogue warps have their accumulators in registers: now the trailing columns may use the TMEM ports
                if (p.park == 2) mbar_wait4_sleepy(a_free, par_ld, p.wd, WD_TAG(8)); else if (p.park == 1) mbar_wait4_parked(a_free, par_ld, p.wd, WD_TAG(8)); else mbar_wait4(a_free, par_ld, p.wd, WD_TAG(8));
                par_ld ^= 1;
                tcx::tc_fence_after();
                if (tcx::elect_one()) {
                  for (int kk = 0; kk < ksteps; ++kk) {
                    const uint32_t dk = (uint32_t)(kk - k);
                    const uint32_t a0 = (kk == 0) ? s_acc : 1u;
                    tcx::mma_f16_ts(d_r, ta_h + 8u * dk, desc(db_h + b_step * dk + ro), idesc_r, a0);
                    tcx::mma_f16_ts(d_r, ta_h + 8u * dk, desc(db_l + b_step * dk + ro), idesc_r, 1u);
                    tcx::mma_f16_ts(d_r, ta_l + 8u * dk, desc(db_h + b_step * dk + ro), idesc_r, 1u);
                  }
                  tcx::mma_commit(w_empty + slot);
                }
                __syncwarp();
              }
              LOG5(iss_slot, 48 + k)
              da_h += 2 * a_step; da_l += 2 * a_step; db_h += 2 * b_step; db_l += 2 * b_step;
              ta_h += 16; ta_l += 16;
              ++rdy; bit <<= 1;
            }
            LOG5(iss_slot, 23)
            if (++slot == (uint32_t)p.nslots) { slot = 0; ++use; }
            if (s_last) buf ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps: 4 TMEM lane quadrants x 2 row halves x 2 slice lanes =====================
    const int q = warp & 3, part = warp >> 2, rh = part & 1, sll = part >> 1;
    const int hw = lane >> 4, lr = lane & 15;
    const int trow = q * 32 + rh * 16 + lr;        // row inside the 128-point tile (= TMEM lane)
    const int crow = trow;
    const int wrow0 = q * 32 + rh * 16;            // first tile row of this warp
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32 + rh * 16) << 16);
    constexpr bool spline = kMode != 0;
    constexpr bool fast_rqs = kMode == 1;
    const bool rows_mine = (sll == 0);             // this warp owns the per-row state of its 16 rows
    const bool owner = rows_mine && (hw == 0);
    const bool add_ctx = kCtx && (C > 0) && !p.folded;
    const int pair_id = 1 + q * 2 + rh;            // named barrier of the two warps (slice lanes) sharing these 16 rows
    float* scr = scratch + trow;
    auto raw = [&](int m) { return scr[m * kTileM]; };
    auto setw = [&](int m, float v) { scr[m * kTileM] = v; };
    uint8_t* a_chain = smem + p.off_h;
    uint64_t* my_acc = bar_acc;
    uint64_t* my_ready = a_ready;
    uint32_t par_acc = 0, lcnt = 0, buf = 0;
    [[maybe_unused]] uint32_t par_free = 1;   // first A block: wait(parity 1) on the fresh barrier returns at once
    int dbg_n = 0;
    const int dbg_slot = p.dbg_all ? warp : sll;
    auto ld_done_arrive = [&]() {   // kDefer: this warp holds every accumulator it needs of the current step in registers
      __syncwarp();
      if (lane == 0) tcx::mbar_arrive(a_free);
    };
    auto await_a_free = [&](bool has_slices) {   // once per A block and warp, right before the first store
      if (kATmem && kAFree) {
        if (has_slices) { mbar_wait4(a_free, par_free, p.wd, WD_TAG(7)); tcx::tc_fence_after(); }
        par_free ^= 1;
      }
    };
    // x_r hand-over inside a 16-row group: the owner warp (slice lane 0) publishes every x it finalises by bumping a
    // shared-memory counter, its partner warp (slice lane 1) counts the same program steps and polls the counter before
    // a first-layer phase that needs x (a 64-thread bar.sync cost ~300 cycles here, on the critical path of every stage)
    volatile int* my_xflag = xflag + q * 2 + rh;
    int xcount = 0;
    float pend_n = 1.f, pend_d = 1.f;   // deferred log-dets: product of the pending spline derivative numerators / denominators
    auto flush_ld = [&](float& ld_acc) {
      // ld = log(num) - 2 log(den) of every pending inverse, taken off the critical path (after a publish)
      ld_acc += (tcx::lg2_approx(pend_n) - 2.f * tcx::lg2_approx(pend_d)) * 0.6931471805599453f;
      pend_n = 1.f; pend_d = 1.f;
    };
    auto publish_x = [&]() {   // owner warp, after the owner lanes stored x into xr
      ++xcount;
      __syncwarp();
      // release store / acquire load on the counter itself instead of two block fences (fences: -1.6 % on cfg3: MEMBAR.ALL.CTA
      // also waits for the thread's outstanding global stores)
      if (lane == 0) asm volatile("st.release.cta.shared::cta.u32 [%0], %1;\n" ::"r"(tcx::smem_u32((const void*)my_xflag)), "r"(xcount) : "memory");
    };
    auto await_x = [&]() {     // partner warp: every x published so far is visible
      if (lane == 0) {
        int seen;
        do {
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(seen) : "r"(tcx::smem_u32((const void*)my_xflag)) : "memory");
        } while (seen - xcount < 0);
      }
      __syncwarp();
    };
    const uint64_t scale2 = tcx::pk2(kTanhScale, kTanhScale);

    // write this thread's 8-column chunk `c` (hi / lo fp16) of the A block: layout [chunk][64 rows][8 halves]
    auto store_chunk = [&](int c, const uint4& hi4, const uint4& lo4) {
      if (kATmem) {
        // chunk c = K elements [8c, 8c + 8) = TMEM columns [4c, 4c + 4) of the hi / lo image; the two half-warps of a row hold
        // chunks 2s and 2s + 1, i.e. the 8 columns of K slice s (the store is warp-collective: both halves always take part)
        const uint32_t ta = lane_base + p.t_a + (kDbl ? buf * (uint32_t)p.kr_max : 0u) + (uint32_t)(c >> 1) * 8u;
        tcx::tmem_st16x2_4<4>(ta, hi4.x, hi4.y, hi4.z, hi4.w);
        tcx::tmem_st16x2_4<4>(ta + (uint32_t)p.kr_max / 2u, lo4.x, lo4.y, lo4.z, lo4.w);
      } else {
        uint8_t* dst = a_chain + (size_t)buf * a_buf_bytes + ((size_t)c * kTileM + crow) * 16;
        *reinterpret_cast<uint4*>(dst) = hi4;
        *reinterpret_cast<uint4*>(dst + a_img_bytes) = lo4;
      }
    };
    // publish K slice `sl` of the A block under construction
    auto publish = [&](int sl) {
      if (kATmem) { tcx::tmem_st_wait(); tcx::tc_fence_before(); }
      else tcx::fence_async_smem();
      __syncwarp();
      if (lane == 0) tcx::mbar_arrive(my_ready + buf * kV5MaxPairs + (sl >> 1));
    };
    // pairs in which this warp's slice lane has no slice (only the last pair of a block with an odd slice count)
    auto observe = [&](int nsl) {
      const int npairs = (nsl + 1) >> 1;
      if (2 * (npairs - 1) + sll >= nsl && lane == 0) tcx::mbar_arrive(my_ready + buf * kV5MaxPairs + (npairs - 1));
    };

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = (int)(item % n_tiles), grp = (int)(item / n_tiles);
      const int n0 = tile * kTileM;
      const int npts = min(kTileM, io.N - n0);
      float run_m = -INFINITY, run_s = 0.f;
      // ---- tile load: the 16 rows of this quadrant (part-0 warp), then visible to the other parts ----
      pair_bar_sync(pair_id, 64);       // every part is done with the previous tile's rows
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
      pair_bar_sync(pair_id, 64);       // context rows are in shared memory

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
          const bool dbg_on = kDbg && p.dbg != nullptr && blockIdx.x == 0 && (p.dbg_all || (q == 0 && rh == 0)) && item == blockIdx.x && si != grp && li >= 2;
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_epi = p.steps[st].epi;
            if (s_epi == EPI_NONE) continue;   // K-split sub-step: nothing to do on this side
            LOG5(dbg_slot, 1)
            const uint32_t s_ecol = p.steps[st].e_col, s_encols = p.steps[st].e_ncols, s_eaux = p.steps[st].e_aux;
            const uint32_t s_stage = p.steps[st].stage, s_flags = p.steps[st].flags;
            if (p.steps[st].w_bytes) {
              mbar_wait4(my_acc, par_acc, p.wd, WD_TAG(6));
              par_acc ^= 1;
              tcx::tc_fence_after();
              LOG5(dbg_slot, 2)
            }
            // kDefer: the issuer holds the trailing MMAs of this (split) push back until all 16 warps have arrived here
            const bool s_split = kDefer && kATmem && p.steps[st].w_bytes && (p.steps[st].n != p.steps[st].n_crit);
            if (s_epi == EPI_TANH) {
              // block of nch 8-column chunks = nsl K slices; slice s = chunks {2s (half-warp 0), 2s+1 (half-warp 1)}
              const int nch = s_encols >> 3, nsl = (nch + 1) >> 1;
              const int nmine = (nsl - sll + 1) / 2;   // slices sll, sll + 2, ...
              observe(nsl);
              if (kAFree && nmine == 0) await_a_free(false);
              if (s_split && nmine == 0) ld_done_arrive();
              for (int j0 = 0; j0 < nmine; j0 += 2) {
                uint32_t r[16];
                const int nj = min(2, nmine - j0);
                const int sl0 = sll + j0 * 2;
                const uint32_t ta = lane_base + s_ecol + sl0 * 16;
                // software-pipelined accumulator reads (TMEM -> registers runs at ~64 B/clk for the whole SM: 450 cycles for a
                // 56-column block): the second slice of this warp is in flight while the first one goes through tanh
                tcx::tmem_ld16x2_8<8>(ta, r);
                if (kDefer && nj > 1) tcx::tmem_ld16x2_8<8>(ta + 32, r + 8);   // both slices at once: the ld_done arrival needs them
                tcx::tmem_ld_wait();
                if (!kDefer && nj > 1) tcx::tmem_ld16x2_8<8>(ta + 32, r + 8);
                tcx::tc_fence_before();
                if (s_split && j0 + 2 >= nmine) ld_done_arrive();
                LOG5(dbg_slot, 3)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  if (u < nj) {
                    if (!kDefer && u == 1) { tcx::tmem_ld_wait(); tcx::tc_fence_before(); }
                    const int sl = sl0 + u * 2;
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
                    if (kAFree && j0 == 0 && u == 0) await_a_free(true);
                    store_chunk(c, hi4, lo4);
                    publish(sl);
                    LOG5(dbg_slot, 8 + sl)
                  }
                }
              }
              buf ^= 1;
            } else if (s_epi == EPI_FIRST) {
              // first conditioner layer of block `stage` on CUDA cores: s = b'[n] + sum_c W0c[n][c] ctx_c + sum_{q < stage} W0x[n][q] x_q
              // (everything pre-multiplied by 2 log2 e), tanh, fp16 hi/lo A block.  x_q (by rank) were written by the row owners.
              const int r = (int)s_stage;
              if (r > 0 && !rows_mine) await_x();   // x_{r-1} of these rows is visible (the owner warp wrote it itself)
              LOG5(dbg_slot, 4)
              const int nch = s_encols >> 3, nsl = (nch + 1) >> 1;
              const int u0 = s_eaux;
              const int nmine = (nsl - sll + 1) / 2;
              observe(nsl);
              if (kAFree) await_a_free(nmine > 0);
              // q-major weights: lc_w0x[q][n] (x of rank q -> unit n), lc_w0c[c][n] (context c -> unit n), everything scaled by
              // 2 log2 e; a thread's 8 units are two 16-byte loads per input and the update is 4 packed FFMA2
              const int hp0 = p.dp4;   // v5: KParamsInv4::dp4 carries the padded width of hidden layer 0 = row stride of both tables
              for (int j = 0; j < nmine; ++j) {
                const int sl = sll + j * 2;
                const int c = sl * 2 + hw;
                uint4 hi4 = make_uint4(0, 0, 0, 0), lo4 = hi4;
                if (c < nch) {
                  const int nb = u0 + c * 8;
                  uint64_t s2[4];
                  {
                    const ulonglong2* bb = reinterpret_cast<const ulonglong2*>(lc + p.lc_b0 + nb);
                    const ulonglong2 b0 = bb[0], b1 = bb[1];
                    s2[0] = b0.x; s2[1] = b0.y; s2[2] = b1.x; s2[3] = b1.y;
                  }
                  if (add_ctx) {
                    // per-point context: same treatment as the x terms below (straight-line for C = 2 and C = 4, the bench shapes)
                    const float* cb = lc + p.lc_w0c + nb;
                    auto cterm = [&](int k) {
                      const float xc = xin[k * kTileM + trow];
                      const uint64_t x2 = tcx::pk2(xc, xc);
                      const ulonglong2* ww = reinterpret_cast<const ulonglong2*>(cb + (size_t)k * hp0);
                      const ulonglong2 w0 = ww[0], w1 = ww[1];
                      s2[0] = tcx::fma2(w0.x, x2, s2[0]); s2[1] = tcx::fma2(w0.y, x2, s2[1]);
                      s2[2] = tcx::fma2(w1.x, x2, s2[2]); s2[3] = tcx::fma2(w1.y, x2, s2[3]);
                    };
                    if (C == 2) { cterm(0); cterm(1); }
                    else if (C == 4) { cterm(0); cterm(1); cterm(2); cterm(3); }
                    else for (int k = 0; k < C; ++k) cterm(k);
                  }
                  // r = 1, 2, 3 (every stage of the 2-D ... 4-D bench flows) as straight-line code: all loads in flight together, no
                  // loop / reconvergence bookkeeping (event log: with the run-time loop ONE x term cost ~300 of a round's ~1050
                  // cycles, mostly branch-resolving and dependent shared-memory latency); deeper stages keep the loop
                  const float* wb = lc + p.lc_w0x + nb;
                  auto xterm = [&](int k, const ulonglong2& w0, const ulonglong2& w1, float xq) {
                    const uint64_t x2 = tcx::pk2(xq, xq);
                    s2[0] = tcx::fma2(w0.x, x2, s2[0]); s2[1] = tcx::fma2(w0.y, x2, s2[1]);
                    s2[2] = tcx::fma2(w1.x, x2, s2[2]); s2[3] = tcx::fma2(w1.y, x2, s2[3]);
                  };
                  if (r == 1) {
                    const float x0 = xr[trow];
                    const ulonglong2* a0 = reinterpret_cast<const ulonglong2*>(wb);
                    const ulonglong2 w00 = a0[0], w01 = a0[1];
                    xterm(0, w00, w01, x0);
                  } else if (r == 2) {
                    const float x0 = xr[trow], x1 = xr[kTileM + trow];
                    const ulonglong2* a0 = reinterpret_cast<const ulonglong2*>(wb);
                    const ulonglong2* a1 = reinterpret_cast<const ulonglong2*>(wb + hp0);
                    const ulonglong2 w00 = a0[0], w01 = a0[1], w10 = a1[0], w11 = a1[1];
                    xterm(0, w00, w01, x0);
                    xterm(1, w10, w11, x1);
                  } else if (r == 3) {
                    const float x0 = xr[trow], x1 = xr[kTileM + trow], x2v = xr[2 * kTileM + trow];
                    const ulonglong2* a0 = reinterpret_cast<const ulonglong2*>(wb);
                    const ulonglong2* a1 = reinterpret_cast<const ulonglong2*>(wb + hp0);
                    const ulonglong2* a2 = reinterpret_cast<const ulonglong2*>(wb + 2 * hp0);
                    const ulonglong2 w00 = a0[0], w01 = a0[1], w10 = a1[0], w11 = a1[1], w20 = a2[0], w21 = a2[1];
                    xterm(0, w00, w01, x0);
                    xterm(1, w10, w11, x1);
                    xterm(2, w20, w21, x2v);
                  } else {
                    for (int k = 0; k < r; ++k) {
                      const ulonglong2* ww = reinterpret_cast<const ulonglong2*>(wb + (size_t)k * hp0);
                      const ulonglong2 w0 = ww[0], w1 = ww[1];
                      xterm(k, w0, w1, xr[k * kTileM + trow]);
                    }
                  }
                  tcx::tanh8_scaled(s2, hi4, lo4);
                }
                if (kATmem) __syncwarp();   // the TMEM store is warp-collective: reconverge after the per-chunk branch
                store_chunk(c, hi4, lo4);
                publish(sl);
                LOG5(dbg_slot, 8 + sl)
              }
              if (fast_rqs && owner) flush_ld(ld_acc);   // log-dets of the inverses since the last flush, off the critical path
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
              if (rows_mine) publish_x(); else ++xcount;
            } else if (s_epi == EPI_XINV && !rows_mine) {
              if (s_split) ld_done_arrive();
              ++xcount;
            } else if (s_epi == EPI_XINV) {
              const int r = s_stage, d = perm[r];
              const float yv = ycur[d * kTileM + trow];
              const float* bo = lc + s_eaux;
              const bool has_acc = !(s_flags & 1);
              float xv = 0.f, ld = 0.f;
              if (!spline) {
                uint32_t rr[2] = {0u, 0u};
                if (has_acc) { tcx::tmem_ld16x2_2<0>(lane_base + s_ecol, rr); tcx::tmem_ld_wait(); }
                if (s_split) ld_done_arrive();
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
                if (s_split) ld_done_arrive();
                const float4* bw = reinterpret_cast<const float4*>(bo + hw * 8);
                const float4* bd = reinterpret_cast<const float4*>(bo + 16);
                const float4 w0 = bw[0], w1 = bw[1], e0 = bd[0], e1 = bd[1];
                const float own[8] = {__uint_as_float(ro[0]) + w0.x, __uint_as_float(ro[1]) + w0.y, __uint_as_float(ro[2]) + w0.z,
                                      __uint_as_float(ro[3]) + w0.w, __uint_as_float(ro[4]) + w1.x, __uint_as_float(ro[5]) + w1.y,
                                      __uint_as_float(ro[6]) + w1.z, __uint_as_float(ro[7]) + w1.w};
                const float dr[8] = {__uint_as_float(rd[0]) + e0.x, __uint_as_float(rd[1]) + e0.y, __uint_as_float(rd[2]) + e0.z,
                                     __uint_as_float(rd[3]) + e0.w, __uint_as_float(rd[4]) + e1.x, __uint_as_float(rd[5]) + e1.y,
                                     __uint_as_float(rd[6]) + e1.z, 0.f};
                float dn, dd;
                nazb::rqs8_inv_pair_nolog(yv, p.bound, own, dr, hw, xv, dn, dd);
                pend_n *= dn; pend_d *= dd;
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
                if (s_split) ld_done_arrive();
                if (owner) {
                  if (p.kind == NAZB_KIND_RQS) nazb::rational_spline<false>(yv, p.K, p.bound, true, raw, setw, xv, ld);
                  else nazb::rational_spline<true>(yv, p.K, p.bound, true, raw, setw, xv, ld);
                }
              }
              tcx::tc_fence_before();
              LOG5(dbg_slot, 5)
              if (owner) xr[r * kTileM + trow] = xv;   // published first: the next stage's first layer waits for it
              publish_x();
              if (owner) {
                ld_acc += ld;
                if (r == D - 1) {
                  // end of this flow layer: x becomes the y of the next (earlier) layer
                  for (int rr2 = 0; rr2 < D; ++rr2) ycur[perm[rr2] * kTileM + trow] = xr[rr2 * kTileM + trow];
                }
              }
            }
          }
          { const int st = 255; LOG5(dbg_slot, 6) }
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
            if (fast_rqs) flush_ld(ld_acc);
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

