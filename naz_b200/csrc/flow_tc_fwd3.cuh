// Forward (`sample` direction) kernel of the tcgen05 engine, slice-pipelined.  Included by flow_tc.cu inside its
// anonymous namespace (shares Step / IoArgs / helpers with the inverse kernel).
//
// One CTA = one 128-point tile (M = 128 MMAs, point = TMEM lane) looping over the draws of its group.  Per flow layer
// the conditioner is a dense chain  [ctx | x | 1] -> h1 -> ... -> h_nh -> spline / affine parameters -> y:
//
//   XF    pre[b]   = A_X (K = 16 slice [ctx | x | 1]) . W0img      (bias in the `1` column, values * 2 log2 e)
//   Lj    pre[b^1] = h_{j-1} . W_j^T      (K = H in 16-column slices, 3 MMAs per slice: hi*hi + hi*lo + lo*hi)
//   OUT   out      = h_nh . Wout^T        (chunked by whole dimensions)
//
// and every `tanh` epilogue writes the next A operand (fp16 hi / lo) slice by slice: the issuer fires the MMAs of a
// K slice as soon as its mbarrier completes, so the tensor pipe trails the MUFU-bound epilogue by one slice instead of
// waiting for the whole layer.  The pre-activation accumulator is double-buffered in TMEM (the MMAs of layer j+1 write
// one buffer while the epilogue still reads the other); the A operand needs a single buffer because the epilogue of a
// layer starts only after all of that layer's MMAs (the last readers of A) have retired.
//   * 16 epilogue warps = 4 TMEM quadrants x 4 parts; thread = one row; part p takes K slices p, p+4, ... (16 columns each).
//   * 1 MMA-issuer warp (tcgen05 instructions predicated on the elected lane), 1 TMA producer warp.
#pragma once

constexpr int kF3Parts = 4;
constexpr int kF3EpiWarps = 4 * kF3Parts;
constexpr int kF3Issuer = kF3EpiWarps;
constexpr int kF3Producer = kF3EpiWarps + 1;
constexpr int kF3Threads = (kF3EpiWarps + 2) * 32;
constexpr int kF3MaxSlices = 16;                 // hidden width <= 256

struct KParamsFwd3 {
  Step steps[kMaxSteps];
  int nsteps;
  const uint8_t* wimg;
  unsigned long long draw_bytes, layer_bytes;
  const float* lc;                 // [S][L][lc_floats]: hidden biases * 2 log2 e (layers 1..nh-1), output bias (rank-major, stride M)
  int lc_floats;
  const int* perm;
  int D, C, L, M, K, kind, kin, hp_max, nslots, t_pre1;
  float bound, clip_lo, clip_hi;
  uint32_t off_ax, off_a, off_x, off_ctx, off_misc, off_lc, off_ring;
  uint32_t slot_off[8];            // two-tile kernel: byte offset of the weight slot of sub-step j of a gemm (shared by both tiles)
};

__device__ __forceinline__ void f3_epi_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kF3EpiWarps * 32) : "memory"); }

__global__ void __launch_bounds__(kF3Threads, 1) flow_tc_fwd3_kernel(const __grid_constant__ KParamsFwd3 p,
                                                                      const __grid_constant__ IoArgs io, int n_groups) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);            // [nslots]
  uint64_t* w_empty = w_full + 8;                                   // [nslots]
  uint64_t* bar_acc = w_empty + 8;                                  // issuer -> epilogue
  uint64_t* lc_full = bar_acc + 1;                                  // [2]
  uint64_t* lc_empty = lc_full + 2;                                 // [2], count = kF3EpiWarps
  uint64_t* ax_ready = lc_empty + 2;                                // epilogue -> issuer, count = kF3EpiWarps
  uint64_t* a_ready = ax_ready + 1;                                 // [kF3MaxSlices], count = 4 (the quadrant warps of one part)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + kF3MaxSlices);
  float* xcur = reinterpret_cast<float*>(smem + p.off_x);           // [D][128]
  float* ctxs = reinterpret_cast<float*>(smem + p.off_ctx);         // [C][128]
  float* ldpart = reinterpret_cast<float*>(smem + p.off_misc);      // [kF3Parts][128]
  float* lcs = reinterpret_cast<float*>(smem + p.off_lc);           // [2][lc_floats]
  uint8_t* ring = smem + p.off_ring;
  constexpr uint32_t ax_img_bytes = 16 * kTileM * 2;                // A_X: [2 chunks][128 rows][8 halves]
  const uint32_t a_img_bytes = (uint32_t)p.hp_max * kTileM * 2;     // A: [hp_max / 8 chunks][128 rows][8 halves]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int D = p.D, C = p.C, M = p.M;

  if (tid == 0) {
    for (int i = 0; i < p.nslots; ++i) { tcx::mbar_init(w_full + i, 1); tcx::mbar_init(w_empty + i, 1); }
    tcx::mbar_init(bar_acc, 1);
    for (int i = 0; i < 2; ++i) { tcx::mbar_init(lc_full + i, 1); tcx::mbar_init(lc_empty + i, kF3EpiWarps); }
    tcx::mbar_init(ax_ready, kF3EpiWarps);
    // slice 0 also collects one arrival from every warp that does not produce it, so that no epilogue warp can be
    // lapped on the accumulator barrier when a layer has fewer K slices than there are parts
    for (int i = 0; i < kF3MaxSlices; ++i) tcx::mbar_init(a_ready + i, i == 0 ? kF3EpiWarps : 4);
    tcx::mbar_fence_init();
  }
  if (warp == 0) tcx::tmem_alloc(tmem_slot, kTmemCols);
  for (uint32_t i = tid; i < (2 * ax_img_bytes) / 16; i += kF3Threads)
    reinterpret_cast<uint4*>(smem + p.off_ax)[i] = make_uint4(0, 0, 0, 0);
  for (uint32_t i = tid; i < (2 * a_img_bytes) / 16; i += kF3Threads)
    reinterpret_cast<uint4*>(smem + p.off_a)[i] = make_uint4(0, 0, 0, 0);
  tcx::fence_async_smem();
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n_tiles = (io.N + kTileM - 1) / kTileM;
  const long long n_items = (long long)n_tiles * n_groups;

  if (warp == kF3Producer) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t cnt = 0, lcnt = 0;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = (int)(item / n_tiles);
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
            for (int st = 0; st < p.nsteps; ++st) {
              const uint32_t wb = p.steps[st].w_bytes;
              const uint32_t slot = cnt % p.nslots, use = cnt / p.nslots;
              tcx::mbar_wait(w_empty + slot, (use & 1) ^ 1);
              tcx::mbar_expect_tx(w_full + slot, wb);
              tcx::bulk_g2s(ring + (size_t)slot * kFwdSlotBytes, wl + p.steps[st].w_off, wb, w_full + slot);
              ++cnt;
            }
          }
        }
      }
    }
  } else if (warp == kF3Issuer) {
    // ===================== MMA issuer =====================
    const uint32_t elected = tcx::elect_one();
    const uint32_t ring_a = tcx::smem_u32(ring);
    const uint32_t a_base = tcx::smem_u32(smem + p.off_a), ax_base = tcx::smem_u32(smem + p.off_ax);
    constexpr uint32_t lbo_a = kTileM * 16;
    constexpr uint64_t dhi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;   // SBO = 128 B, descriptor version 1
    uint32_t slot = 0, use = 0, pbuf = 0, apar = 0, xpar = 0;   // apar: one parity bit per A slice barrier
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int grp = (int)(item / n_tiles);
      for (int si = grp; si < io.s_count; si += n_groups) {
        for (int l = 0; l < p.L; ++l) {
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_wbytes = p.steps[st].w_bytes, s_n = p.steps[st].n;
            const int ksteps = p.steps[st].ksteps;
            const uint32_t s_acc = p.steps[st].accumulate, s_epi = p.steps[st].epi, s_flags = p.steps[st].flags;
            const uint32_t slice0 = p.steps[st].a_chunk0 >> 1;
            const bool is_x = (p.steps[st].a_buf == A_X);
            const bool to_pre = (p.steps[st].d_col == 0xFFFF);      // pre-activation buffer chosen by phase parity
            const uint32_t idesc = tcx::make_idesc_f16(s_n);
            const uint32_t lbo_b = s_n * 16;
            const uint32_t b_hi = ring_a + slot * kFwdSlotBytes, b_lo = b_hi + (s_wbytes >> 1);
            const uint32_t lbo_b_hi16 = (lbo_b >> 4) << 16, lbo_a_hi16 = (lbo_a >> 4) << 16;
            const uint32_t a_hi = is_x ? ax_base : a_base;
            const uint32_t a_lo = a_hi + (is_x ? ax_img_bytes : a_img_bytes);
            const uint32_t d_addr = tmem + (to_pre ? pbuf * (uint32_t)p.t_pre1 : (uint32_t)p.steps[st].d_col);
            tcx::mbar_wait(w_full + slot, use & 1);
            if (is_x) { tcx::mbar_wait(ax_ready, xpar); xpar ^= 1; tcx::tc_fence_after(); }
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t sl = slice0 + k;
              if (!is_x && !(s_flags & 8)) {              // bit 3: the A slices were already waited for (later output chunks)
                const uint32_t bit = 1u << sl;
                tcx::mbar_wait(a_ready + sl, (apar & bit) ? 1u : 0u);
                apar ^= bit;
                tcx::tc_fence_after();
              }
              const uint32_t ao = sl * 2 * lbo_a, bo = (uint32_t)k * 2 * lbo_b;
              const uint64_t da_h = dhi | (((a_hi + ao) >> 4) | lbo_a_hi16), da_l = dhi | (((a_lo + ao) >> 4) | lbo_a_hi16);
              const uint64_t db_h = dhi | (((b_hi + bo) >> 4) | lbo_b_hi16), db_l = dhi | (((b_lo + bo) >> 4) | lbo_b_hi16);
              tcx::mma_f16_ss_elect(d_addr, da_h, db_h, idesc, (k == 0) ? s_acc : 1u, elected);
              tcx::mma_f16_ss_elect(d_addr, da_h, db_l, idesc, 1u, elected);
              tcx::mma_f16_ss_elect(d_addr, da_l, db_h, idesc, 1u, elected);
            }
            if (s_epi != EPI_NONE) tcx::mma_commit_elect(bar_acc, elected);
            tcx::mma_commit_elect(w_empty + slot, elected);
            if (++slot == (uint32_t)p.nslots) { slot = 0; ++use; }
            if (s_epi == EPI_TANH) pbuf ^= 1;                        // the next layer accumulates into the other buffer
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int part = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const bool spline = p.kind != NAZB_KIND_AFFINE;
    uint8_t* a_buf = smem + p.off_a;
    uint8_t* ax_buf = smem + p.off_ax;
    uint32_t par_acc = 0, lcnt = 0, pbuf = 0;
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
      const int tile = (int)(item % n_tiles), grp = (int)(item / n_tiles);
      const int n0 = tile * kTileM;
      const int npts = min(kTileM, io.N - n0);
      const bool valid = row < npts;
      // ---- tile load: context (and the base noise when it is shared by all draws) ----
      f3_epi_sync();
      if (part == 0) {
        for (int c = 0; c < C; ++c) {
          float v = valid ? io.ctx[((io.ctx_rows == 1) ? 0 : (size_t)(n0 + row)) * C + c] : 0.f;
          ctxs[c * kTileM + row] = v;
          ax_store(c, v);
        }
        ax_store(p.kin, 1.f);
      }
      for (int si = grp; si < io.s_count; si += n_groups) {
        // ---- draw start: base noise -> x, first-layer operand ----
        if (part == 0) {
          const float* zs = io.x + (size_t)si * io.x_draw_stride;
          for (int d = 0; d < D; ++d) {
            const float v = valid ? zs[(size_t)(n0 + row) * D + d] : 0.f;
            xcur[d * kTileM + row] = v;
            ax_store(C + d, v);
          }
        }
        float ld_acc = 0.f;
        f3_epi_sync();

        for (int l = 0; l < p.L; ++l) {
          const int* perm = p.perm + l * D;
          const float* lc = lcs + (size_t)(lcnt & 1) * p.lc_floats;
          tcx::mbar_wait(lc_full + (lcnt & 1), (lcnt >> 1) & 1);
          // publish the first-layer operand of this flow layer
          tcx::fence_async_smem();
          __syncwarp();
          if (lane == 0) tcx::mbar_arrive(ax_ready);
          for (int st = 0; st < p.nsteps; ++st) {
            const uint32_t s_epi = p.steps[st].epi;
            if (s_epi == EPI_NONE) continue;
            const uint32_t s_ecol = p.steps[st].e_col, s_encols = p.steps[st].e_ncols, s_eaux = p.steps[st].e_aux;
            const uint32_t s_stage = p.steps[st].stage, s_nranks = p.steps[st].nranks, s_flags = p.steps[st].flags;
            tcx::mbar_wait(bar_acc, par_acc);
            par_acc ^= 1;
            tcx::tc_fence_after();
            if (s_epi == EPI_TANH) {
              const int nsl = (int)(s_encols + 15) >> 4;
              const bool prescaled = (s_flags & 4) != 0;
              const uint32_t tbase = lane_base + pbuf * (uint32_t)p.t_pre1;
              if (part != 0 && lane == 0) tcx::mbar_arrive(a_ready);   // observer arrival on slice 0
              for (int sl = part; sl < nsl; sl += kF3Parts) {
                uint32_t r[16];
                tcx::tmem_ld16(tbase + sl * 16, r);
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
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  uint8_t* dst = a_buf + ((size_t)(sl * 2 + u) * kTileM + row) * 16;
                  *reinterpret_cast<uint4*>(dst) = hi4[u];
                  *reinterpret_cast<uint4*>(dst + a_img_bytes) = lo4[u];
                }
                tcx::fence_async_smem();
                __syncwarp();
                if (lane == 0) tcx::mbar_arrive(a_ready + sl);
              }
              pbuf ^= 1;
            } else {   // EPI_XFWD: transform the dims of ranks [stage, stage + nranks)
              const float* bo = lc + s_eaux;
              for (int i = part; i < (int)s_nranks; i += kF3Parts) {
                const int d = perm[s_stage + i];
                const float xv = xcur[d * kTileM + row];
                float yv, ld;
                if (!spline) {
                  uint32_t rr[2];
                  tcx::tmem_ld2(lane_base + s_ecol + i * M, rr);
                  tcx::tmem_ld_wait();
                  const float mu = __uint_as_float(rr[0]) + bo[i * M];
                  const float sc = fminf(fmaxf(__uint_as_float(rr[1]) + bo[i * M + 1], p.clip_lo), p.clip_hi);
                  yv = mu + xv * expf(sc);
                  ld = sc;
                } else {
                  uint32_t rr[24];
                  const uint32_t ta = lane_base + s_ecol + i * M;   // M = 23 (K = 8 quadratic spline)
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
          }
          __syncwarp();
          if (lane == 0) tcx::mbar_arrive(lc_empty + (lcnt & 1));
          ++lcnt;
        }

        // ---- draw end ----
        ldpart[part * kTileM + row] = ld_acc;
        f3_epi_sync();
        if (part == 0 && valid) {
          if (io.out_l) {
            float a = 0.f;
            for (int pp = 0; pp < kF3Parts; ++pp) a += ldpart[pp * kTileM + row];
            io.out_l[(size_t)si * io.N + n0 + row] = a;
          }
          float* dst = io.out_x + ((size_t)si * io.N + n0 + row) * D;
          for (int d = 0; d < D; ++d) {
            float v = xcur[d * kTileM + row];
            if (io.lo != nullptr) v = nazb::bound_inv(v, io.lo[d], io.hi[d]);
            dst[d] = v;
          }
        }
        f3_epi_sync();   // xcur / ldpart are rewritten at the next draw start
      }
    }
  }
  tcx::tc_fence_before();
  __syncthreads();
  if (warp == 0) tcx::tmem_dealloc(tmem, kTmemCols);
}
