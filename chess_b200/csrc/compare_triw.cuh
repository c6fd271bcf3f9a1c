// The -a 3 / 5 / 9 / 11 work-item kernel over the upper triangle: compare_tri.cuh's three packed row
// bands with the loop of win_rows<G, H> (compare_kernels.cuh; separate running sums for the 2H+1 SSIM
// window and the 7x7 SN window, both fed from one ring of x and y).  Weights as in compare_tri7.cuh.
// Windows 9 and 11 widen the halo around a band to H rows / columns and need strips of G >= H columns,
// so that a window still reaches into one neighbour lane only.
#pragma once
#include "compare_tri.cuh"

namespace {

// NW = warps per CTA: 4, or 1 for windows 9 / 11, whose ring of 9 / 11 rows of x and y (32 / 39 KB per warp at 7
// columns per lane) lets only ONE 4-warp CTA onto an SM -- one-warp CTAs pack 7 / 5 of them
template <int G, int H, int NW, bool BATCH>
__global__ void __launch_bounds__(32 * NW, (NW == 1 ? 1 : (G <= 4 ? 3 : 2)))
compare_win_tri_kernel(chess_item_args a) {
  static_assert((H == 1 || H == 2 || H == 4 || H == 5) && G >= H, "window 3 / 5 / 9 / 11 (window 7 has its own kernel)");
  constexpr int HALO = H > 3 ? H : 3;           // halo rows / columns around a band: the wider of the two windows
  constexpr int D = 2 * H + 1 > 7 ? 2 * H + 1 : 7;  // ring depth in rows
  __shared__ int s_susp[NW][kSuspCap];
  __shared__ int s_susp_n[NW];
  __shared__ double s_part[NW][4];   // per warp: totals of the parts of the current item (pair items)
  constexpr bool HYB = G <= 4;      // pair items only for the narrow strips (see compare_tri.cuh)
  extern __shared__ __align__(16) double ringbuf[];  // per warp: D rows of x and y, D*2*G*32 doubles

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *susp = s_susp[warp];
  int *susp_n = &s_susp_n[warp];
  constexpr int RING = D * 2 * G * 32;
  double *myring = ringbuf + (size_t)warp * RING + lane;
  const int NCOL = a.ncol_pad;
  short *kmap = reinterpret_cast<short *>(ringbuf + (size_t)NW * RING) + (size_t)warp * NCOL;
  unsigned char *mflag = reinterpret_cast<unsigned char *>(reinterpret_cast<short *>(ringbuf + (size_t)NW * RING) +
                                                          (size_t)NW * NCOL) + (size_t)warp * NCOL;
  const chess_params &p = a.p;
  const double qnan = CUDART_NAN;
  if (blockIdx.x == 0 && threadIdx.x == 0) { *a.counter_next = 0ull; *a.retry_next = 0; }
  const long long n_pair_items = HYB ? a.pair_items : 0;
  const long long n_items = n_pair_items + 2 * (a.n_pairs - n_pair_items);

  // the first item of every warp is its own index: no storm of atomics on one counter when the grid
  // starts; later items come from the counter, offset by the number of warps
  const long long n_warps = (long long)gridDim.x * NW;
  bool first_item = true;
  while (true) {
    long long item = (long long)blockIdx.x * NW + warp;
    if (!first_item) {
      if (lane == 0) item = n_warps + (long long)atomicAdd(a.counter_cur, 1ull);
      item = __shfl_sync(0xffffffffu, item, 0);
    }
    first_item = false;
    if (item >= n_items) break;
    TriItem it;
    const bool PAIR = HYB && item < n_pair_items;
    const long long half = item - n_pair_items;   // index among the half items
    if (!tri_item_begin<BATCH>(a, PAIR ? item : n_pair_items + (half >> 1), PAIR ? 0 : (int)(half & 1), 31 * G, lane, kmap,
                        mflag, it))
      continue;
    const long long pi = it.pi;
    const int part0 = it.part, n = it.n, np_ = it.np_, flip = it.flip;
    const double *R = it.R, *Q = it.Q;
    const size_t rp = it.rp, qp = it.qp;
    if (lane == 0) *susp_n = 0;
    __syncwarp();

    const int win = 2 * H + 1;
    if (win > np_) {
      if (part0 == 0 && lane == 0) { a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_WINDOW_EXCEEDS; }
      continue;
    }
    const int m = np_ - win + 1;
    const bool with_sn = p.compute_sn != 0;

    // ---- data range from the running extrema, else min / max from the data
    double xmax, xmin;
    // C1 / C2 are needed before the first window: the rare data scan runs up front
    if (p.data_range > 0.0) { xmax = p.data_range; xmin = 0.0; }   // fixed range (chess_params.data_range): no look-up
    else if (tri_extrema<BATCH>(a, it, lane, mflag, xmax, xmin)) tri_scan_extrema(it, lane, kmap, xmax, xmin);
    const double drange = xmax - xmin;
    const double C1 = (0.01 * drange) * (0.01 * drange), C2 = (0.03 * drange) * (0.03 * drange);

    if (!(C2 > 0.0)) {
      // constant matrices (zero data range): the SSIM quotient degenerates to x / 0; the generic kernel evaluates
      // it with a true division, as numpy does
      if (part0 == 0 && lane == 0) {
        a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
        atomicOr(a.retry_cur, 1);
      }
      continue;
    }

    // ---- the parts of this item (tri_rows.cuh: bands cut at multiples of G, lane-uniform weights)
    int B1, B2, L1, L2;
    const int mode = tri_mode(np_, G, B1, B2, L1, L2);   // 1: all three bands in one pass (the first part does it all)
    double sn_sum = 0.0, sn_cnt = 0.0, ss_sum = 0.0;
    bool overflow = false;
    constexpr double NP2 = (double)((2 * H + 1) * (2 * H + 1)) * (double)((2 * H + 1) * (2 * H + 1));
    const double c1 = C1 * NP2, c2 = C2 * NP2;
    for (int part = part0; part <= (PAIR ? 1 : part0); ++part) {
    if (lane == 0) *susp_n = 0;
    __syncwarp();
    const int pcode = (mode == 1 && part == 0) ? 2 : part;
    const TriLane ln = tri_lane(pcode, lane, G, np_, B1, B2, L1, L2, HALO);
    const int rows_max = pcode == 0 ? B1 : pcode == 1 ? (mode == 2 ? max(B2 - B1, np_ - B2) : 0)
                                                      : max(B1, max(B2 - B1, np_ - B2) + HALO);
    if (rows_max > 0) {
      tri_win_part<G, H>(R, Q, (unsigned)rp, (unsigned)qp, n, np_, flip, kmap, myring, ln, pcode != 1, rows_max, with_sn, c1,
                         c2, susp, susp_n, ss_sum, sn_sum, sn_cnt);
      if (with_sn) {
        __syncwarp();
        const int ns = *susp_n;
        if (ns > 0 && ns <= kSuspCap) {  // cold: numpy's two passes for the cancelled windows
          if (lane == 0) {               // fixed order, so the additions are reproducible
            for (int i = 1; i < ns; ++i) {
              const int key = susp[i];
              int j = i - 1;
              while (j >= 0 && susp[j] > key) { susp[j + 1] = susp[j]; --j; }
              susp[j + 1] = key;
            }
          }
          __syncwarp();
          for (int e = lane; e < ns; e += 32) {
            const int rc = susp[e] >> 16, c = susp[e] & 0xffff;
            const double cnt = (double)((min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1) *
                                        (min(c + 3, np_ - 1) - max(c - 3, 0) + 1));
            const double gq = sn_window_exact(R, Q, rp, qp, kmap, n, flip, rc, c, np_, cnt);
            const int band_hi = rc < B1 ? B1 : (rc < B2 ? B2 : np_);  // end of the row's diagonal block
            if (gq == gq) { sn_sum += c >= band_hi ? gq + gq : gq; sn_cnt += c >= band_hi ? 2.0 : 1.0; }
          }
        } else if (ns > kSuspCap) {
          overflow = true;               // too many: the generic kernel redoes the pair exactly
        }
        __syncwarp();
      }
    }

    // this part's totals, reduced over the warp: exactly the numbers a half item writes to its record, added up
    // in the order the combine adds two records -- a pair gives the same bits whichever way it ran
    {
      const double ws = warp_sum(sn_sum), wc = warp_sum(sn_cnt), ww = warp_sum(ss_sum);
      if (lane == 0) {
        double *ptot = s_part[warp];
        const bool first_part = part == part0;
        ptot[0] = first_part ? ws : ptot[0] + ws;
        ptot[1] = first_part ? wc : ptot[1] + wc;
        ptot[2] = first_part ? ww : ptot[2] + ww;
      }
      sn_sum = 0.0; sn_cnt = 0.0; ss_sum = 0.0;
    }
    }  // parts
    __syncwarp();
    const double PS = s_part[warp][0], PC = s_part[warp][1], PW = s_part[warp][2];

    if (PAIR) {  // the whole pair was walked by this warp: verify the predicted masks and finish
      bool bad = overflow || !tri_masks_hold<BATCH>(a, it, lane, kmap, mflag);
      bad = __any_sync(0xffffffffu, bad);
      if (lane == 0) {
        if (bad) {
          a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
          atomicOr(a.retry_cur, 1);
        } else {
          a.ssim[pi] = PW / ((double)m * (double)m);
          a.sn[pi] = with_sn ? PS / PC : CUDART_NAN;
          a.status[pi] = CHESS_PAIR_OK;
        }
      }
      continue;
    }

    // ---- record of this item
    double *rec = a.scratch + ((size_t)pi * 2 + part0) * kTriRec;
    if (lane == 0) {
      rec[5] = xmax; rec[6] = xmin; rec[7] = PS; rec[8] = PC;
      rec[14] = overflow ? 1.0 : 0.0;
      rec[15] = PW;
    }
    __threadfence();
    __syncwarp();
    int arrived = 0;
    if (lane == 0) arrived = atomicAdd(a.done + pi, 1);
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived != 1) continue;

    // ---- second item to arrive: verify the predicted masks, combine, write the result
    __threadfence();
    const double *recs = a.scratch + (size_t)pi * 2 * kTriRec;
    bool bad = __ldcg(recs + 14) != 0.0 || __ldcg(recs + kTriRec + 14) != 0.0;
    bad |= !tri_masks_hold<BATCH>(a, it, lane, kmap, mflag);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      if (bad) {
        a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
        atomicOr(a.retry_cur, 1);
      } else {
        const double ss = __ldcg(recs + 7) + __ldcg(recs + kTriRec + 7);
        const double sc = __ldcg(recs + 8) + __ldcg(recs + kTriRec + 8);
        const double sm = __ldcg(recs + 15) + __ldcg(recs + kTriRec + 15);
        a.ssim[pi] = sm / ((double)m * (double)m);
        a.sn[pi] = with_sn ? ss / sc : CUDART_NAN;
        a.status[pi] = CHESS_PAIR_OK;
      }
      a.done[pi] = 0;
    }
  }
}

template <int G, int H, bool BATCH>
int launch_triw_impl(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  constexpr int NW = H >= 4 ? 1 : 4;
  auto kern = compare_win_tri_kernel<G, H, NW, BATCH>;
  constexpr int D = 2 * H + 1 > 7 ? 2 * H + 1 : 7;
  constexpr int RING = D * 2 * G * 32;
  const size_t smem = (size_t)NW * RING * sizeof(double) + (size_t)NW * a.ncol_pad * 3;
  const size_t smem_max = (size_t)NW * RING * sizeof(double) + (size_t)NW * 256 * 3;
  static bool configured[64] = {false};
  if (ctx->device >= 64 || !configured[ctx->device]) {
    CHESS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    if (ctx->device < 64) configured[ctx->device] = true;
  }
  static int per_sm = 0, per_sm_ncol = -1;
  if (per_sm_ncol != a.ncol_pad) {
    CHESS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * NW, smem));
    if (per_sm < 1) per_sm = 1;
    per_sm_ncol = a.ncol_pad;
  }
  long long grid = (long long)ctx->sm_count * per_sm;
  // narrow strips: whole waves of pairs run as pair items (compare_tri.cuh: launch_tri)
  const long long slots = grid * NW;
  long long pair_items = G > 4 ? 0 : (a.n_pairs >= 8 * slots ? a.n_pairs : (a.n_pairs / slots) * slots);
  static const int pair_env = getenv("CHESS_TRI_PAIR") ? atoi(getenv("CHESS_TRI_PAIR")) : -1;  // A/B runs
  if (pair_env == 0) pair_items = 0;
  if (pair_env == 1 && G <= 4) pair_items = a.n_pairs;
  int b1, b2, l1, l2;
  if (G <= 4 && tri_mode(a.max_n, G, b1, b2, l1, l2) != 2) pair_items = a.n_pairs;   // one pass per pair: nothing to halve
  a.pair_items = pair_items;
  const long long need = (pair_items + 2 * (a.n_pairs - pair_items) + NW - 1) / NW;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, 32 * NW, smem, stream>>>(a);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

// set-up look-ups batched or one group at a time (compare_tri7.cuh: launch_tri7); CHESS_TRIW_BATCH=0 / 1: A/B
template <int G, int H>
int launch_triw(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  static const int batch_env = getenv("CHESS_TRIW_BATCH") ? atoi(getenv("CHESS_TRIW_BATCH")) : -1;
  if constexpr (G >= 5) {
    if (batch_env == 0) return launch_triw_impl<G, H, false>(ctx, a, stream);
  }
  return launch_triw_impl<G, H, true>(ctx, a, stream);
}

}  // namespace
