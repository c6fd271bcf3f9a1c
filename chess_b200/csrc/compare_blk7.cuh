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
    if (!tri_item_begin(a, item / ipp, (int)(item % ipp), nb * (31 * G - 2 * 3), lane, kmap, mflag, it)) continue;
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
    if (tri_extrema(a, it, lane, mflag, xmax, xmin)) tri_scan_extrema(it, lane, kmap, xmax, xmin);
    const double drange = xmax - xmin;
    const double C1 = (0.01 * drange) * (0.01 * drange), C2 = (0.03 * drange) * (0.03 * drange);

    // ---- this item's block (bi, bj), bj >= bi, of nb x nb equal blocks of the compacted pair:
    // rows [lo, hi), emitted columns [clo, chi), HALO rows / columns either side
    if ((np_ + nb - 1) / nb < 8 + 3) {  // blocks narrower than their halos (almost everything masked)
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
    const int side = (np_ + nb - 1) / nb;
    const int lo = min(bi * side, np_), hi = min(lo + side, np_);
    const int clo = min(bj * side, np_), chi = min(clo + side, np_);
    const int cbase = clo - 3 + G * lane;
    const int rows_max = (chi > clo) ? hi - lo : 0;
    const double wblock = bi == bj ? 1.0 : 2.0;   // a block above the diagonal stands for its mirror image too
    unsigned xo[G], yo[G];
    bool valid[G];
    double wc[G];   // weight of the column: 0 in the halo, the block's weight inside
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int c = cbase + g;
      valid[g] = c >= 0 && c < np_ && c < chi + 3;
      wc[g] = (c >= clo && c < chi) ? wblock : 0.0;
      const int kc = valid[g] ? (int)kmap[c] : 0;
      xo[g] = (unsigned)kc;
      yo[g] = (unsigned)(flip ? n - 1 - kc : kc);
    }
    double *rec = a.scratch + ((size_t)pi * ipp + part) * kTriRec;
    double sn_sum = 0.0, sn_cnt = 0.0, ss_sum = 0.0;
    bool overflow = false;

    if (rows_max > 0) {
      const int first = max(lo - 3, 0);                           // the top blocks start at row 0
      const int t_emit0 = bi == 0 ? 3 : 6;                        // first iteration with a complete window
      const int T = rows_max + (bi == 0 ? 3 : 6);
      const double inv_np = 1.0 / 49.0;
      const double cov_norm = 49.0 / 48.0;
      double ccnt[G], wss[G];
      double Vx[G], Vy[G], Vq[G], Vxy[G];
      unsigned hist[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = cbase + g;
        ccnt[g] = (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1);
        wss[g] = (c >= 3 && c <= np_ - 4) ? wc[g] : 0.0;          // SSIM windows need 3 columns either side
        Vx[g] = Vy[g] = Vq[g] = Vxy[g] = 0.0;
        hist[g] = 0u;
#pragma unroll
        for (int k = 0; k < 7; ++k) { myring[((k * G + g) * 2) * 32] = 0.0; myring[((k * G + g) * 2 + 1) * 32] = 0.0; }
      }
      const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
      int slot = 0;
      double xn[G], yn[G];
      auto fetch = [&](int row) {
#pragma unroll
        for (int g = 0; g < G; ++g) { xn[g] = 0.0; yn[g] = 0.0; }
        if (row < np_) {
          const unsigned kr = (unsigned)kmap[row];
          const unsigned ro = kr * (unsigned)rp;
          const unsigned qo = (flip ? (unsigned)(n - 1) - kr : kr) * (unsigned)qp;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (valid[g]) { xn[g] = __ldg(R + (ro + xo[g])); yn[g] = __ldg(Q + (qo + yo[g])); }
          }
        }
      };
      int row = first;
      fetch(row);
      for (int t = 0; t < T; ++t, ++row) {
        double x[G], y[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { x[g] = xn[g]; y[g] = yn[g]; }
        if (t + 1 < T) fetch(row + 1);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          double *cell = myring + ((slot * G + g) * 2) * 32;   // the slot being overwritten holds row - 7
          const double xold = cell[0], yold = cell[32];
          cell[0] = x[g];
          cell[32] = y[g];
          Vx[g] += x[g]; Vx[g] -= xold;
          Vy[g] += y[g]; Vy[g] -= yold;
          Vq[g] = fma(x[g], x[g], fma(y[g], y[g], Vq[g]));
          Vq[g] = fma(-xold, xold, fma(-yold, yold, Vq[g]));
          Vxy[g] = fma(x[g], y[g], Vxy[g]); Vxy[g] = fma(-xold, yold, Vxy[g]);
          hist[g] = ((hist[g] << 1) & 0x7fu) | (x[g] != y[g] ? 1u : 0u);
          if (hist[g] == 0u) { Vy[g] = Vx[g]; Vq[g] = 2.0 * Vxy[g]; }
        }
        slot = slot == 6 ? 0 : slot + 1;
        if (slot == 0) {
          // the ring has just wrapped: rebuild the four column sums from its seven rows.  A running sum keeps the
          // rounding of every value that ever passed through it; rebuilt once per window height it is exact to a
          // few ulps of the current window (sum d*d = Wq - 2 Wxy magnifies whatever error Vq and Vxy carry)
#pragma unroll
          for (int g = 0; g < G; ++g) {
            double vq = 0.0, vxy = 0.0;   // the quadratic sums only: errors in sum x, sum y enter linearly
#pragma unroll
            for (int k = 0; k < 7; ++k) {
              const double xk = myring[((k * G + g) * 2) * 32], yk = myring[((k * G + g) * 2 + 1) * 32];
              vq = fma(xk, xk, fma(yk, yk, vq));
              vxy = fma(xk, yk, vxy);
            }
            Vq[g] = vq; Vxy[g] = vxy;
            if (hist[g] == 0u) { Vy[g] = Vx[g]; Vq[g] = 2.0 * Vxy[g]; }
          }
        }

        if (t >= t_emit0) {
          const int rc = row - 3;   // centre row of the SSIM window and SN pixel row
          const bool emit_row = rc >= lo && rc < hi;
          const bool ssim_row = emit_row && rc >= 3 && rc <= np_ - 4;
          double W4[4][G];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double *V = q == 0 ? Vx : q == 1 ? Vy : q == 2 ? Vq : Vxy;
            double P[G + 1];
            P[0] = 0.0; P[1] = V[0];
#pragma unroll
            for (int g = 1; g < G; ++g) P[g + 1] = P[g] + V[g];
            double up[3], dn[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double sfx = j == 0 ? V[G - 1] : (G - 1 - j == 0 ? P[G] : P[G] - P[G - 1 - j]);
              up[j] = __shfl_sync(0xffffffffu, sfx, up_lane);
              dn[j] = __shfl_sync(0xffffffffu, P[j + 1], dn_lane);
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const int a0 = g - 3 < 0 ? 0 : g - 3;
              const int b0 = g + 3 > G - 1 ? G - 1 : g + 3;
              double w = a0 == 0 ? P[b0 + 1] : P[b0 + 1] - P[a0];
              if (g - 3 < 0) w += up[2 - g];
              if (g + 3 > G - 1) w += dn[g + 3 - G];
              W4[q][g] = w;
            }
          }
          const double rcnt = (double)(min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1);
          unsigned suspect = 0u;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            {
              const double ux = W4[0][g] * inv_np, uy = W4[1][g] * inv_np;
              const double uq = W4[2][g] * inv_np, uxy = W4[3][g] * inv_np;
              const double m2s = fma(ux, ux, uy * uy);                  // ux^2 + uy^2
              const double vsum = cov_norm * (uq - m2s);                // vx + vy
              const double vxy = cov_norm * fma(-ux, uy, uxy);
              const double A1 = fma(2.0 * ux, uy, C1), A2 = fma(2.0, vxy, C2);
              const double B1s = m2s + C1, B2s = vsum + C2;
              const double den = B1s * B2s;
              const double num = A1 * A2;
              double S = num * fast_rcp(den);
              if (den == 0.0) S = num * CUDART_INF;  // zero data range: +-inf or NaN exactly like num / 0
              if (ssim_row && wss[g] != 0.0) ss_sum = fma(wss[g], S, ss_sum);
            }
            if (with_sn) {
              const double w1 = W4[0][g] - W4[1][g];
              const double w2 = fma(-2.0, W4[3][g], W4[2][g]);
              const bool own = emit_row & (wc[g] != 0.0);
              const double cnt = rcnt * ccnt[g];
              const double m2 = w2 * cnt;
              const double den = fma(-w1, w1, m2);
              // trusted: one-pass variance well conditioned AND the q - 2xy subtraction kept >= 10 digits
              const bool good = own & (den > fma(m2, kVarGuard, 1e-290)) & (w2 > 1e-6 * W4[2][g]);
              const double gq = fabs(w1) * cnt * fast_rcp(den);
              sn_sum = fma(wc[g], good ? gq : 0.0, sn_sum);
              sn_cnt += good ? wc[g] : 0.0;
              if (!good && own && w2 != 0.0) suspect |= 1u << g;  // rare: cancellation
            }
          }
          if (suspect) {  // one test per row instead of one per column: exact recompute after the loop
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (suspect & (1u << g)) {
                const int e = atomicAdd(susp_n, 1);
                if (e < kSuspCap) susp[e] = (rc << 16) | (cbase + g);
              }
            }
          }
        }
      }
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
