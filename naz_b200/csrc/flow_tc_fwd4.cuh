// Two-tile forward (`sample` direction) kernel of the tcgen05 engine.  Included by flow_tc.cu after flow_tc_fwd3.cuh
// (same program, weight images and layer constants as the one-tile kernel there).
//
// The one-tile kernel leaves the tensor pipe idle whenever its single dependency chain is in an epilogue-only stretch
// (first slice of every layer, transform -> first-layer hand-over).  Two tiles would fill those holes, but a second A
// operand (fp16 hi/lo of a 128 x 160 activation block = 80 KB) does not fit in shared memory.  Here two 128-point
// tiles SHARE ONE A buffer as a ring of 16-column K slices:
//
//   tile 0 epilogue writes slice s -> issuer: tile 0 MMAs on slice s -> tcgen05.commit frees slice s
//     -> tile 1 epilogue writes slice s -> issuer: tile 1 MMAs on slice s -> commit frees slice s -> tile 0 (next layer) ...
//
// so in steady state the tiles are half a phase apart: tile 1's tanh epilogue is paced by (and hides under) tile 0's MMA
// stream and vice versa, and the issuer alternates T0.step, T1.step through the program.  Each tile keeps ONE
// pre-activation accumulator (160 columns) + the transform parameters (<= 96 columns) = one half of TMEM, therefore a
// layer's MMAs start only when ALL K slices of that tile have landed (its epilogue has then finished reading `pre`).
//   * warps 0-11 / 12-23: epilogue of tile 0 / 1 (4 TMEM quadrants x 3 parts, thread = row, part p takes slices p, p+3, ...)
//   * warp 24: MMA issuer, warp 25: TMA producer.  A gemm's weight images (one slot per K sub-step) are loaded ONCE and
//     used by tile 0 then tile 1; tile 1's tcgen05.commit frees the slot for the next gemm's sub-step.
#pragma once

constexpr int kF4Tiles = 2;
constexpr int kF4Parts = 3;
constexpr int kF4TileWarps = 4 * kF4Parts;
constexpr int kF4EpiWarps = kF4Tiles * kF4TileWarps;
constexpr int kF4Issuer = kF4EpiWarps;
constexpr int kF4Producer = kF4EpiWarps + 1;
constexpr int kF4Threads = (kF4EpiWarps + 2) * 32;
constexpr int kF4TileCols = kTmemCols / kF4Tiles;

__device__ __forceinline__ void f4_tile_sync(int t) { asm volatile("bar.sync %0, %1;\n" ::"r"(1 + t), "n"(kF4TileWarps * 32) : "memory"); }

#define DBG4(slot) if (kDbg && dbg_buf && blockIdx.x == 0 && dbg_i < 128) dbg_buf[dbg_i * 16 + (slot)] = clk();

template <bool kDbg>
__global__ void __launch_bounds__(kF4Threads, 1) flow_tc_fwd4_kernel(long long* dbg_buf, const __grid_constant__ KParamsFwd3 p,
                                                                      const __grid_constant__ IoArgs io, int n_groups) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);            // [nslots]
  uint64_t* w_empty = w_full + 8;                                   // [nslots]
  uint64_t* bar_acc = w_empty + 8;                                  // [2] issuer -> epilogue of tile t
  uint64_t* lc_full = bar_acc + kF4Tiles;                           // [2]
  uint64_t* lc_empty = lc_full + 2;                                 // [2], count = kF4EpiWarps
  uint64_t* ax_ready = lc_empty + 2;                                // [2] tile t -> issuer, count = kF4TileWarps
  uint64_t* a_ready = ax_ready + kF4Tiles;                          // [2][kF3MaxSlices]: slice written by tile t (count 4; slice 0: all 8 warps)
  uint64_t* a_free = a_ready + kF4Tiles * kF3MaxSlices;             // [kF3MaxSlices]: the MMAs that read the slice retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + kF3MaxSlices);
  float* lcs = reinterpret_cast<float*>(smem + p.off_lc);           // [2][lc_floats]
  uint8_t* ring = smem + p.off_ring;
  constexpr uint32_t ax_img_bytes = 16 * kTileM * 2;                // A_X: [2 chunks][128 rows][8 halves], hi then lo
  const uint32_t a_img_bytes = (uint32_t)p.hp_max * kTileM * 2;     // shared A ring: [hp_max / 8 chunks][128 rows][8 halves], hi then lo

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int D = p.D, C = p.C, M = p.M;

  if (tid == 0) {
    for (int i = 0; i < p.nslots; ++i) { tcx::mbar_init(w_full + i, 1); tcx::mbar_init(w_empty + i, 1); }
    for (int i = 0; i < kF4Tiles; ++i) { tcx::mbar_init(bar_acc + i, 1); tcx::mbar_init(ax_ready + i, kF4TileWarps); }
    for (int i = 0; i < 2; ++i) { tcx::mbar_init(lc_full + i, 1); tcx::mbar_init(lc_empty + i, kF4EpiWarps); }
    for (int i = 0; i < kF4Tiles * kF3MaxSlices; ++i) tcx::mbar_init(a_ready + i, (i % kF3MaxSlices == 0) ? kF4TileWarps : 4);
    for (int i = 0; i < kF3MaxSlices; ++i) tcx::mbar_init(a_free + i, 1);
    tcx::mbar_fence_init();
  }
  if (warp == 0) tcx::tmem_alloc(tmem_slot, kTmemCols);
  for (uint32_t i = tid; i < (kF4Tiles * 2 * ax_img_bytes) / 16; i += kF4Threads)
    reinterpret_cast<uint4*>(smem + p.off_ax)[i] = make_uint4(0, 0, 0, 0);
  for (uint32_t i = tid; i < (2 * a_img_bytes) / 16; i += kF4Threads)
    reinterpret_cast<uint4*>(smem + p.off_a)[i] = make_uint4(0, 0, 0, 0);
  tcx::fence_async_smem();
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n_pairs = (io.N + kF4Tiles * kTileM - 1) / (kF4Tiles * kTileM);
  const long long n_items = (long long)n_pairs * n_groups;

  if (warp == kF4Producer) {
    // ===================== TMA producer: every image once per tile, in issue order =====================
    if (lane == 0) {
      uint32_t wuse = 0, lcnt = 0;   // wuse: 4-bit use counter per weight slot
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = (int)(item / n_pairs);
        for (int si = grp; si < io.s_count; si += n_groups) {
          const uint8_t* wdraw = p.wimg + (size_t)(io.s_begin + si) * p.draw_bytes;
          const float* lcdraw = p.lc + (size_t)(io.s_begin + si) * p.L * p.lc_floats;
          for (int l = 0; l < p.L; ++l) {
            {
              const uint32_t b = lcnt & 1, use = lcnt >> 1;
              tcx::mbar_wait(lc_empty + b, (use & 1) ^ 1);
              tcx::mbar_expect_tx(lc_full + b, (uint32_t)p.lc_floats * 4);
              tcx::bulk_g2s(lcs + (size_t)b * p.lc_floats, lcdraw + (size_t)l * p.lc_floats, (uint32_t)p.lc_floats * 4, lc_full + b);
              ++lcnt;
            }
            const uint8_t* wl = wdraw + (size_t)l * p.layer_bytes;
            int st = 0;
            while (st < p.nsteps) {
              int st_end = st;                                   // [st, st_end]: the sub-steps of one gemm
              while (p.steps[st_end].epi == EPI_NONE) ++st_end;
              for (int s2 = st; s2 <= st_end; ++s2) {
                const uint32_t wb = p.steps[s2].w_bytes;
                const uint32_t slot = (uint32_t)(s2 - st), use = (wuse >> (4 * slot)) & 15u;   // per-slot use parity
                tcx::mbar_wait(w_empty + slot, (use & 1) ^ 1);
                tcx::mbar_expect_tx(w_full + slot, wb);
                tcx::bulk_g2s(ring + p.slot_off[slot], wl + p.steps[s2].w_off, wb, w_full + slot);
                wuse = (wuse & ~(15u << (4 * slot))) | (((use + 1) & 15u) << (4 * slot));
              }
              st = st_end + 1;
            }
          }
        }
      }
    }
  } else if (warp == kF4Issuer) {
    // ===================== MMA issuer: T0.gemm, T1.gemm, next gemm ... =====================
    const uint32_t elected = tcx::elect_one();
    const uint32_t ring_a = tcx::smem_u32(ring);
    const uint32_t a_base = tcx::smem_u32(smem + p.off_a), ax_base0 = tcx::smem_u32(smem + p.off_ax);
    constexpr uint32_t lbo_a = kTileM * 16;
    constexpr uint64_t dhi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
    uint32_t wuse = 0, xpar = 0;   // wuse: 4-bit use counter per weight slot
    uint32_t apar[kF4Tiles] = {0u, 0u};   // one parity bit per (tile, slice) a_ready barrier
    int dbg_i = 0;
    const bool dbg_me = (lane == 0);
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int grp = (int)(item / n_pairs);
      for (int si = grp; si < io.s_count; si += n_groups) {
        for (int l = 0; l < p.L; ++l) {
          int st = 0;
          while (st < p.nsteps) {
            int st_end = st, nsl_total = 0;
            while (p.steps[st_end].epi == EPI_NONE) { nsl_total += p.steps[st_end].ksteps; ++st_end; }
            nsl_total += p.steps[st_end].ksteps;
            const bool is_x = (p.steps[st].a_buf == A_X);
            const bool to_pre = (p.steps[st].d_col == 0xFFFF);
            const bool no_wait = (p.steps[st].flags & 8) != 0 || (p.steps[st_end].flags & 8) != 0;
#pragma unroll 1
            for (int t = 0; t < kF4Tiles; ++t) {
              const uint32_t d_addr = tmem + (uint32_t)t * kF4TileCols + (to_pre ? 0u : (uint32_t)p.t_pre1);
              if (dbg_me) { DBG4(8 + 3 * t) }
              if (is_x) {
                tcx::mbar_wait(ax_ready + t, (xpar >> t) & 1);
                xpar ^= 1u << t;
              } else if (!no_wait) {
                // the tile's epilogue has finished reading `pre` only when ALL its K slices have landed
                for (int sl = 0; sl < nsl_total; ++sl) {
                  const uint32_t bit = 1u << sl;
                  tcx::mbar_wait(a_ready + t * kF3MaxSlices + sl, (apar[t] & bit) ? 1u : 0u);
                  apar[t] ^= bit;
                }
              }
              tcx::tc_fence_after();
              if (dbg_me) { DBG4(9 + 3 * t) }
              for (int s2 = st; s2 <= st_end; ++s2) {
                const uint32_t s_wbytes = p.steps[s2].w_bytes, s_n = p.steps[s2].n;
                const int ksteps = p.steps[s2].ksteps;
                const uint32_t s_acc = p.steps[s2].accumulate;
                const uint32_t slice0 = p.steps[s2].a_chunk0 >> 1;
                const uint32_t idesc = tcx::make_idesc_f16(s_n);
                const uint32_t lbo_b = s_n * 16;
                const uint32_t slot = (uint32_t)(s2 - st), use = (wuse >> (4 * slot)) & 15u;
                const uint32_t b_hi = ring_a + p.slot_off[slot], b_lo = b_hi + (s_wbytes >> 1);
                const uint32_t lbo_b_hi16 = (lbo_b >> 4) << 16, lbo_a_hi16 = (lbo_a >> 4) << 16;
                const uint32_t a_hi = is_x ? (ax_base0 + (uint32_t)t * 2 * ax_img_bytes) : a_base;
                const uint32_t a_lo = a_hi + (is_x ? ax_img_bytes : a_img_bytes);
                tcx::mbar_wait(w_full + slot, use & 1);
                for (int k = 0; k < ksteps; ++k) {
                  const uint32_t sl = slice0 + k;
                  const uint32_t ao = sl * 2 * lbo_a, bo = (uint32_t)k * 2 * lbo_b;
                  const uint64_t da_h = dhi | (((a_hi + ao) >> 4) | lbo_a_hi16), da_l = dhi | (((a_lo + ao) >> 4) | lbo_a_hi16);
                  const uint64_t db_h = dhi | (((b_hi + bo) >> 4) | lbo_b_hi16), db_l = dhi | (((b_lo + bo) >> 4) | lbo_b_hi16);
                  tcx::mma_f16_ss_elect(d_addr, da_h, db_h, idesc, (k == 0) ? s_acc : 1u, elected);
                  tcx::mma_f16_ss_elect(d_addr, da_h, db_l, idesc, 1u, elected);
                  tcx::mma_f16_ss_elect(d_addr, da_l, db_h, idesc, 1u, elected);
                  if (!is_x) tcx::mma_commit_elect(a_free + sl, elected);   // the other tile may overwrite this slice
                }
                if (s2 == st_end) tcx::mma_commit_elect(bar_acc + t, elected);
                if (t == kF4Tiles - 1) {   // the last tile's MMAs on this slot retire -> the producer may refill it
                  tcx::mma_commit_elect(w_empty + slot, elected);
                  wuse = (wuse & ~(15u << (4 * slot))) | (((use + 1) & 15u) << (4 * slot));
                }
              }
              if (dbg_me) { DBG4(10 + 3 * t) }
            }
            ++dbg_i;
            st = st_end + 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps of tile t =====================
    const int t = warp / kF4TileWarps, part = (warp >> 2) % kF4Parts, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)t * kF4TileCols;
    const bool spline = p.kind != NAZB_KIND_AFFINE;
    float* xcur = reinterpret_cast<float*>(smem + p.off_x) + (size_t)t * D * kTileM;           // [D][128]
    float* ctxs = reinterpret_cast<float*>(smem + p.off_ctx) + (size_t)t * (C > 0 ? C : 1) * kTileM;
    float* ldpart = reinterpret_cast<float*>(smem + p.off_misc) + (size_t)t * kF4Parts * kTileM;   // [parts][128]
    uint8_t* a_buf = smem + p.off_a;
    uint8_t* ax_buf = smem + p.off_ax + (size_t)t * 2 * ax_img_bytes;
    uint64_t* my_ready = a_ready + t * kF3MaxSlices;
    uint32_t par_acc = 0, lcnt = 0;
    uint32_t nwrites = 0;                       // A blocks produced so far by this tile (ring protocol parity)
    int dbg_i = 0;
    const bool dbg_me = (warp % kF4TileWarps == 0 && lane == 0);
    const int dbg_o = t * 4;
    const uint64_t scale2 = tcx::pk2(kTanhScale, kTanhScale);

    auto ax_store = [&](int col, float v) {
      const float c = fminf(fmaxf(v, -65504.f), 65504.f);
      const __half h = __float2half_rn(c);
      const __half l = __float2half_rn(c - __half2float(h));
      uint8_t* dst = ax_buf + ((size_t)(col >> 3) * kTileM + row) * 16 + (col & 7) * 2;
      *reinterpret_cast<__half*>(dst) = h;
      *reinterpret_cast<__half*>(dst + ax_img_bytes) = l;
    };

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int pair = (int)(item % n_pairs), grp = (int)(item / n_pairs);
      const int n0 = (pair * kF4Tiles + t) * kTileM;
      const int npts = max(0, min(kTileM, io.N - n0));
      const bool valid = row < npts;
      f4_tile_sync(t);
      if (part == 0) {
        for (int c = 0; c < C; ++c) {
          float v = valid ? io.ctx[((io.ctx_rows == 1) ? 0 : (size_t)(n0 + row)) * C + c] : 0.f;
          ctxs[c * kTileM + row] = v;
          ax_store(c, v);
        }
        ax_store(p.kin, 1.f);
      }
      for (int si = grp; si < io.s_count; si += n_groups) {
        if (part == 0) {
          const float* zs = io.x + (size_t)si * io.x_draw_stride;
          for (int d = 0; d < D; ++d) {
            const float v = valid ? zs[(size_t)(n0 + row) * D + d] : 0.f;
            xcur[d * kTileM + row] = v;
            ax_store(C + d, v);
          }
        }
        float ld_acc = 0.f;
        f4_tile_sync(t);

        for (int l = 0; l < p.L; ++l) {
          const int* perm = p.perm + l * D;
          const float* lc = lcs + (size_t)(lcnt & 1) * p.lc_floats;
          tcx::mbar_wait(lc_full + (lcnt & 1), (lcnt >> 1) & 1);
          tcx::fence_async_smem();
          __syncwarp();
          if (lane == 0) tcx::mbar_arrive(ax_ready + t);     // first-layer operand of this flow layer is in place
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_epi = p.steps[st].epi;
            if (s_epi == EPI_NONE) continue;
            const uint32_t s_encols = p.steps[st].e_ncols, s_eaux = p.steps[st].e_aux;
            const uint32_t s_stage = p.steps[st].stage, s_nranks = p.steps[st].nranks, s_flags = p.steps[st].flags;
            if (dbg_me) { DBG4(dbg_o + 0) }
            tcx::mbar_wait(bar_acc + t, par_acc);
            par_acc ^= 1;
            tcx::tc_fence_after();
            if (dbg_me) { DBG4(dbg_o + 1) }
            if (s_epi == EPI_TANH) {
              const int nsl = (int)(s_encols + 15) >> 4;
              const bool prescaled = (s_flags & 4) != 0;
              if (part != 0 && lane == 0) tcx::mbar_arrive(my_ready);   // observer arrival on slice 0
              // ring protocol: tile 0's k-th block follows tile 1's (k-1)-th read (odd a_free phases), tile 1's k-th
              // block follows tile 0's k-th read (even phases); tile 0's very first block finds the ring empty
              const bool need_free = (t == 1) || (nwrites > 0);
              const uint32_t free_par = (t == 1) ? 0u : 1u;
              for (int sl = part; sl < nsl; sl += kF4Parts) {
                uint32_t r[16];
                tcx::tmem_ld16(lane_base + sl * 16, r);
                tcx::tmem_ld_wait();
                tcx::tc_fence_before();
                uint4 hi4[2], lo4[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const uint32_t* ru = r + 8 * u;
                  uint64_t s2[4];
                  if (prescaled) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) s2[i] = tcx::pk2(__uint_as_float(ru[2 * i]), __uint_as_float(ru[2 * i + 1]));
                  } else {
                    const ulonglong2* bv = reinterpret_cast<const ulonglong2*>(lc + s_eaux + sl * 16 + u * 8);
                    const ulonglong2 b0 = bv[0], b1 = bv[1];
                    s2[0] = tcx::fma2(tcx::pk2(__uint_as_float(ru[0]), __uint_as_float(ru[1])), scale2, b0.x);
                    s2[1] = tcx::fma2(tcx::pk2(__uint_as_float(ru[2]), __uint_as_float(ru[3])), scale2, b0.y);
                    s2[2] = tcx::fma2(tcx::pk2(__uint_as_float(ru[4]), __uint_as_float(ru[5])), scale2, b1.x);
                    s2[3] = tcx::fma2(tcx::pk2(__uint_as_float(ru[6]), __uint_as_float(ru[7])), scale2, b1.y);
                  }
                  tcx::tanh8_scaled(s2, hi4[u], lo4[u]);
                }
                if (need_free) tcx::mbar_wait(a_free + sl, free_par);   // the other tile's MMAs on this slice retired
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  uint8_t* dst = a_buf + ((size_t)(sl * 2 + u) * kTileM + row) * 16;
                  *reinterpret_cast<uint4*>(dst) = hi4[u];
                  *reinterpret_cast<uint4*>(dst + a_img_bytes) = lo4[u];
                }
                tcx::fence_async_smem();
                __syncwarp();
                if (lane == 0) tcx::mbar_arrive(my_ready + sl);
                if (dbg_me && sl == part) { DBG4(dbg_o + 2) }
              }
              ++nwrites;
            } else {   // EPI_XFWD: transform the dims of ranks [stage, stage + nranks)
              const float* bo = lc + s_eaux;
              const uint32_t ocol = lane_base + (uint32_t)p.t_pre1;
              for (int i = part; i < (int)s_nranks; i += kF4Parts) {
                const int d = perm[s_stage + i];
                const float xv = xcur[d * kTileM + row];
                float yv, ld;
                if (!spline) {
                  uint32_t rr[2];
                  tcx::tmem_ld2(ocol + i * M, rr);
                  tcx::tmem_ld_wait();
                  const float mu = __uint_as_float(rr[0]) + bo[i * M];
                  const float sc = fminf(fmaxf(__uint_as_float(rr[1]) + bo[i * M + 1], p.clip_lo), p.clip_hi);
                  yv = mu + xv * expf(sc);
                  ld = sc;
                } else {
                  uint32_t rr[24];
                  const uint32_t ta = ocol + i * M;   // M = 23 (K = 8 quadratic spline)
                  tcx::tmem_ld8(ta, rr);
                  tcx::tmem_ld8(ta + 8, rr + 8);
                  tcx::tmem_ld8(ta + 16, rr + 16);
                  tcx::tmem_ld_wait();
                  float rf[24];
#pragma unroll
                  for (int e = 0; e < 23; ++e) rf[e] = __uint_as_float(rr[e]) + bo[i * M + e];
                  rf[23] = 0.f;
                  nazb::rqs_fast<8>(xv, p.bound, false, rf, yv, ld);
                }
                ld_acc += ld;
                xcur[d * kTileM + row] = yv;
                ax_store(C + d, yv);
              }
              tcx::tc_fence_before();
            }
            if (dbg_me) { DBG4(dbg_o + 3) }
            ++dbg_i;
          }
          __syncwarp();
          if (lane == 0) tcx::mbar_arrive(lc_empty + (lcnt & 1));
          ++lcnt;
        }

        // ---- draw end ----
        ldpart[part * kTileM + row] = ld_acc;
        f4_tile_sync(t);
        if (part == 0 && valid) {
          if (io.out_l) {
            float a = 0.f;
            for (int pp = 0; pp < kF4Parts; ++pp) a += ldpart[pp * kTileM + row];
            io.out_l[(size_t)si * io.N + n0 + row] = a;
          }
          float* dst = io.out_x + ((size_t)si * io.N + n0 + row) * D;
          for (int d = 0; d < D; ++d) {
            float v = xcur[d * kTileM + row];
            if (io.lo != nullptr) v = nazb::bound_inv(v, io.lo[d], io.hi[d]);
            dst[d] = v;
          }
        }
        f4_tile_sync(t);
      }
    }
  }
  tcx::tc_fence_before();
  __syncthreads();
  if (warp == 0) tcx::tmem_dealloc(tmem, kTmemCols);
}
