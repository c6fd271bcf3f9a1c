// The -r 1 work-item kernel for matrices wider than one warp (222 .. 1024 bins): the compacted pair
// is cut into nb x nb equal square blocks, and -- both matrices being symmetric -- only the blocks on
// and above the diagonal are computed, the ones above it counting twice (compare_tri.cuh explains the
// symmetry argument).  A work item is one block: its rows plus three halo rows either side, its
// columns plus three halo columns either side (weight 0), lane l owning G adjacent columns.  Compared
// with the full-square walk (two row bands x column blocks) a 601-bin pair costs 6 blocks of 207
// rows at G = 7 instead of 6 strips of 306 rows at G = 8.
#pragma once
#include <type_traits>
#include "compare_tri.cuh"

namespace {

template <int G, bool WITH_SN>
__global__ void __launch_bounds__(128, 2)
compare_r1_blk_kernel(chess_item_args a) {
  __shared__ int s_susp[4][kSuspCap];
  __shared__ int s_susp_n[4];
  extern __shared__ __align__(16) double ringbuf[];  // per warp: 7*G*32 doubles of d = x - y

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *susp = s_susp[warp];
  int *susp_n = &s_susp_n[warp];
  constexpr int RING = 7 * G * 32;
  double *myring = ringbuf + (size_t)warp * RING + lane;
  const int NCOL = a.ncol_pad;
  short *kmap = reinterpret_cast<short *>(ringbuf + (size_t)4 * RING) + (size_t)warp * NCOL;
  unsigned char *mflag = reinterpret_cast<unsigned char *>(reinterpret_cast<short *>(ringbuf + (size_t)4 * RING) +
                                                          (size_t)4 * NCOL) + (size_t)warp * NCOL;
  const chess_params &p = a.p;
  const double qnan = CUDART_NAN;
  if (blockIdx.x == 0 && threadIdx.x == 0) { *a.counter_next = 0ull; *a.retry_next = 0; }
  const int nb = a.blocks;                       // blocks per side
  const int ipp = nb * (nb + 1) / 2;             // items per pair: blocks on and above the diagonal
  const long long n_items = a.n_pairs * ipp;

  // the first item of every warp is its own index: no storm of atomics on one counter when the grid
  // starts; later items come from the counter, offset by the number of warps
  const long long n_warps = (long long)gridDim.x * 4;
  bool first_item = true;
  while (true) {
    long long item = (long long)blockIdx.x * 4 + warp;
    if (!first_item) {
      if (lane == 0) item = n_warps + (long long)atomicAdd(a.counter_cur, 1ull);
      item = __shfl_sync(0xffffffffu, item, 0);
    }
    first_item = false;
    if (item >= n_items) break;
    TriItem it;
    if (!tri_item_begin(a, item / ipp, (int)(item % ipp), nb * 29 * G, lane, kmap, mflag, it)) continue;
    const long long pi = it.pi;
    const int part = it.part, n = it.n, np_ = it.np_, flip = it.flip;
    const double *R = it.R, *Q = it.Q;
    const size_t rp = it.rp, qp = it.qp;
    if (lane == 0) *susp_n = 0;
    __syncwarp();

    int win = (np_ & 1) ? np_ : np_ - 1;
    if (win < 3) win = 3;
    if (win > np_) {
      if (part == 0 && lane == 0) { a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_WINDOW_EXCEEDS; }
      continue;
    }
    const int m = np_ - win + 1;
    if ((np_ + nb - 1) / nb < 8 || (((np_ + nb - 1) / nb + G - 1) / G) > 29) {  // blocks narrower than their halos (almost everything masked): generic kernel
      if (part == 0 && lane == 0) {
        a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
        atomicOr(a.retry_cur, 1);
      }
      continue;
    }

    // ---- data range from the running extrema, else min / max from the data
    double xmax, xmin;
    const bool fixed_range = p.data_range > 0.0;     // chess_params.data_range: no look-up, no scan
    const bool inloop_minmax = !fixed_range && tri_extrema(a, it, lane, mflag, xmax, xmin);  // rare in aligned samples, common against
                                                                      // background regions (a bin masked on one side)
    if (inloop_minmax) { xmax = -CUDART_INF; xmin = CUDART_INF; }   // neutral for an item without rows
    if (fixed_range) { xmax = p.data_range; xmin = 0.0; }

    // ---- this item's block (bi, bj), bj >= bi, of nb x nb square blocks of the compacted pair (tri_rows.cuh:
    // blk_lane): `side` rows, a multiple of the strip width, so that every lane is block or halo as a whole
    int bi = 0, bj = 0;
    {
      int left = part;
      while (left >= nb - bi) { left -= nb - bi; ++bi; }
      bj = bi + left;
    }
    const int side = (((np_ + nb - 1) / nb + G - 1) / G) * G;
    const TriLane ln = blk_lane(bi, bj, side, lane, G, np_, WITH_SN ? 3 : 0);
    const int rows_max = (min(bj * side, np_) < np_) ? ln.hi - ln.lo : 0;
    const double wblock = bi == bj ? 1.0 : 2.0;   // a block above the diagonal stands for its mirror image too
    double *rec = a.scratch + ((size_t)pi * ipp + part) * kTriRec;

    // first row totals (first item) / last row totals (last item) for the 2x2 SSIM map
    double rt[5] = {0, 0, 0, 0, 0};
    if (m == 2 && (part == 0 || part == ipp - 1)) {
      const int kr = kmap[part == 0 ? 0 : np_ - 1];
      const double *rowR = R + (size_t)kr * rp;
      const double *rowQ = Q + (size_t)(flip ? n - 1 - kr : kr) * qp;
      double t5[5] = {0, 0, 0, 0, 0};
      for (int c = lane; c < np_; c += 32) {
        const int kc = kmap[c];
        const double x = __ldg(rowR + kc), y = __ldg(rowQ + (flip ? n - 1 - kc : kc));
        t5[0] += x; t5[1] += y; t5[2] = fma(x, x, t5[2]); t5[3] = fma(y, y, t5[3]); t5[4] = fma(x, y, t5[4]);
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) rt[k] = warp_sum(t5[k]);
    }

    double T0 = 0, T1 = 0, T2 = 0, T3 = 0, T4 = 0;
    double sn_sum = 0.0, sn_cnt = 0.0;
    bool overflow = false;

    if (rows_max > 0) {
      TriPartSums ps;
      tri_r1_part<G, WITH_SN>(R, Q, (unsigned)rp, (unsigned)qp, n, np_, flip, kmap, myring, ln, bi == 0, rows_max,
                              inloop_minmax, susp, susp_n, ps);
      T0 = ps.T[0]; T1 = ps.T[1]; T2 = ps.T[2]; T3 = ps.T[3]; T4 = ps.T[4];
      sn_sum = ps.sn_sum; sn_cnt = ps.sn_cnt;
      if (inloop_minmax) { xmax = warp_max(ps.vmax); xmin = warp_min(ps.vmin); }
      if (WITH_SN) {
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
            if (gq == gq) { sn_sum += wblock * gq; sn_cnt += wblock; }
          }
        } else if (ns > kSuspCap) {
          overflow = true;               // too many: the generic kernel redoes the pair exactly
        }
        __syncwarp();
      }
    }

    // ---- record of this item
    {
      const double t0 = warp_sum(T0), t1 = warp_sum(T1), t2 = warp_sum(T2), t3 = warp_sum(T3), t4 = warp_sum(T4);
      const double ss = warp_sum(sn_sum);
      const double sc = warp_sum(sn_cnt);
      if (lane == 0) {
        rec[0] = t0; rec[1] = t1; rec[2] = t2; rec[3] = t3; rec[4] = t4;
        rec[5] = xmax; rec[6] = xmin; rec[7] = ss; rec[8] = sc;
#pragma unroll
        for (int k = 0; k < 5; ++k) rec[9 + k] = rt[k];
        rec[14] = overflow ? 1.0 : 0.0;
      }
    }
    __threadfence();
    __syncwarp();
    int arrived = 0;
    if (lane == 0) arrived = atomicAdd(a.done + pi, 1);
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived != ipp - 1) continue;

    // ---- last item to arrive: verify the predicted masks, combine in block order, write the result
    __threadfence();
    const double *recs = a.scratch + (size_t)pi * ipp * kTriRec;
    bool bad = false;
    for (int b = lane; b < ipp; b += 32) bad |= __ldcg(recs + (size_t)b * kTriRec + 14) != 0.0;
    bad |= !tri_masks_hold(a, it, lane, kmap, mflag);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      if (bad) {
        a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
        atomicOr(a.retry_cur, 1);
      } else {
        double t[5] = {0, 0, 0, 0, 0}, mx = -CUDART_INF, mn = CUDART_INF, ss = 0.0, sc = 0.0;
        for (int b = 0; b < ipp; ++b) {
          const double *rb = recs + (size_t)b * kTriRec;
#pragma unroll
          for (int k = 0; k < 5; ++k) t[k] += __ldcg(rb + k);
          mx = fmax(mx, __ldcg(rb + 5)); mn = fmin(mn, __ldcg(rb + 6));
          ss += __ldcg(rb + 7); sc += __ldcg(rb + 8);
        }
        double row0[5], rowL[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          row0[k] = __ldcg(recs + 9 + k);
          rowL[k] = __ldcg(recs + (size_t)(ipp - 1) * kTriRec + 9 + k);
        }
        finish_pair(t, mx, mn, ss, sc, row0, rowL, np_, win, m, kmap, R, Q, rp, qp, n, flip, WITH_SN,
                    a.ssim + pi, a.sn + pi, a.status + pi);
      }
      a.done[pi] = 0;
    }
  }
}

template <int G, bool WITH_SN>
int launch_blk(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  auto kern = compare_r1_blk_kernel<G, WITH_SN>;
  constexpr int RING = 7 * G * 32;
  const size_t smem = (size_t)4 * RING * sizeof(double) + (size_t)4 * a.ncol_pad * 3;
  const size_t smem_max = (size_t)4 * RING * sizeof(double) + (size_t)4 * 1024 * 3;
  static bool configured[64] = {false};  // per device: the attribute belongs to the device's copy of the kernel
  if (ctx->device >= 64 || !configured[ctx->device]) {
    CHESS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    if (ctx->device < 64) configured[ctx->device] = true;
  }
  static int per_sm = 0, per_sm_ncol = -1;
  if (per_sm_ncol != a.ncol_pad) {
    CHESS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
    if (per_sm < 1) per_sm = 1;
    per_sm_ncol = a.ncol_pad;
  }
  long long grid = (long long)ctx->sm_count * per_sm;
  const long long need = (a.n_pairs * (a.blocks * (a.blocks + 1) / 2) + 3) / 4;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, 128, smem, stream>>>(a);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

}  // namespace
