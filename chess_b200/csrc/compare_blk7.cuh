// The -a 7 kernel of compare_tri7.cuh on the block decomposition of compare_blk.cuh (matrices wider than
// one warp): equal square blocks on and above the diagonal, the ones above it counting twice.
#pragma once
#include "compare_tri.cuh"

namespace {

template <int G>
__global__ void __launch_bounds__(128, 2)
compare_win7_blk_kernel(chess_item_args a) {
  __shared__ int s_susp[4][kSuspCap];
  __shared__ int s_susp_n[4];
  extern __shared__ __align__(16) double ringbuf[];  // per warp: 7 rows of x and y, 7*2*G*32 doubles

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *susp = s_susp[warp];
  int *susp_n = &s_susp_n[warp];
  constexpr int RING = 7 * 2 * G * 32;
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

    const int win = 7;
    if (win > np_) {
      if (part == 0 && lane == 0) { a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_WINDOW_EXCEEDS; }
      continue;
    }
    const int m = np_ - win + 1;
    const bool with_sn = p.compute_sn != 0;

    // ---- data range from the running extrema, else min / max from the data
    double xmax, xmin;
    // C1 / C2 are needed before the first window: the rare data scan runs up front
    if (p.data_range > 0.0) { xmax = p.data_range; xmin = 0.0; }   // fixed range (chess_params.data_range): no look-up
    else if (tri_extrema(a, it, lane, mflag, xmax, xmin)) tri_scan_extrema(it, lane, kmap, xmax, xmin);
    const double drange = xmax - xmin;
    const double C1 = (0.01 * drange) * (0.01 * drange), C2 = (0.03 * drange) * (0.03 * drange);

    // ---- this item's block (bi, bj), bj >= bi, of nb x nb equal blocks of the compacted pair:
    // rows [lo, hi), emitted columns [clo, chi), HALO rows / columns either side
    if ((np_ + nb - 1) / nb < 8 + 3 || !(C2 > 0.0) || (((np_ + nb - 1) / nb + G - 1) / G) > 29) {  // blocks narrower than their halos; constant matrices (almost everything masked)
      if (part == 0 && lane == 0) {
        a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
        atomicOr(a.retry_cur, 1);
      }
      continue;
    }
    int bi = 0, bj = 0;
    {
      int left = part;
      while (left >= nb - bi) { left -= nb - bi; ++bi; }
      bj = bi + left;
    }
    const int side = (((np_ + nb - 1) / nb + G - 1) / G) * G;
    const TriLane ln = blk_lane(bi, bj, side, lane, G, np_, 3);
    const int rows_max = (min(bj * side, np_) < np_) ? ln.hi - ln.lo : 0;
    const double wblock = bi == bj ? 1.0 : 2.0;   // a block above the diagonal stands for its mirror image too
    double *rec = a.scratch + ((size_t)pi * ipp + part) * kTriRec;
    double sn_sum = 0.0, sn_cnt = 0.0, ss_sum = 0.0;
    bool overflow = false;

    if (rows_max > 0) {
      tri_w7_part<G>(R, Q, (unsigned)rp, (unsigned)qp, n, np_, flip, kmap, myring, ln, bi == 0, rows_max, with_sn,
                     C1 * 2401.0, C2 * 2401.0, susp, susp_n, ss_sum, sn_sum, sn_cnt);
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
      const double ss = warp_sum(sn_sum);
      const double sc = warp_sum(sn_cnt);
      const double sw = warp_sum(ss_sum);
      if (lane == 0) {
        rec[5] = xmax; rec[6] = xmin; rec[7] = ss; rec[8] = sc;
        rec[14] = overflow ? 1.0 : 0.0;
        rec[15] = sw;
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
        double ss = 0.0, sc = 0.0, sm = 0.0;
        for (int b = 0; b < ipp; ++b) {
          const double *rb = recs + (size_t)b * kTriRec;
          ss += __ldcg(rb + 7); sc += __ldcg(rb + 8); sm += __ldcg(rb + 15);
        }
        a.ssim[pi] = sm / ((double)m * (double)m);
        a.sn[pi] = with_sn ? ss / sc : CUDART_NAN;
        a.status[pi] = CHESS_PAIR_OK;
      }
      a.done[pi] = 0;
    }
  }
}

template <int G>
int launch_blk7(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  auto kern = compare_win7_blk_kernel<G>;
  constexpr int RING = 7 * 2 * G * 32;
  const size_t smem = (size_t)4 * RING * sizeof(double) + (size_t)4 * a.ncol_pad * 3;
  const size_t smem_max = (size_t)4 * RING * sizeof(double) + (size_t)4 * 1024 * 3;
  static bool configured[64] = {false};
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
