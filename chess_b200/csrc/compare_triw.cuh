// The -a 3 / 5 / 9 / 11 work-item kernel over the upper triangle: compare_tri.cuh's three packed row
// bands with the loop of win_rows<G, H> (compare_kernels.cuh; separate running sums for the 2H+1 SSIM
// window and the 7x7 SN window, both fed from one ring of x and y).  Weights as in compare_tri7.cuh.
// Windows 9 and 11 widen the halo around a band to H rows / columns and need strips of G >= H columns,
// so that a window still reaches into one neighbour lane only.
#pragma once
#include "compare_tri.cuh"

namespace {

template <int G, int H>
__global__ void __launch_bounds__(128, (G <= 4 ? 3 : 2))
compare_win_tri_kernel(chess_item_args a) {
  static_assert((H == 1 || H == 2 || H == 4 || H == 5) && G >= H, "window 3 / 5 / 9 / 11 (window 7 has its own kernel)");
  constexpr int HALO = H > 3 ? H : 3;           // halo rows / columns around a band: the wider of the two windows
  constexpr int D = 2 * H + 1 > 7 ? 2 * H + 1 : 7;  // ring depth in rows
  __shared__ int s_susp[4][kSuspCap];
  __shared__ int s_susp_n[4];
  __shared__ double s_part[4][4];   // per warp: totals of the parts of the current item (pair items)
  constexpr bool HYB = G <= 4;      // pair items only for the narrow strips (see compare_tri.cuh)
  extern __shared__ __align__(16) double ringbuf[];  // per warp: D rows of x and y, D*2*G*32 doubles

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *susp = s_susp[warp];
  int *susp_n = &s_susp_n[warp];
  constexpr int RING = D * 2 * G * 32;
  double *myring = ringbuf + (size_t)warp * RING + lane;
  const int NCOL = a.ncol_pad;
  short *kmap = reinterpret_cast<short *>(ringbuf + (size_t)4 * RING) + (size_t)warp * NCOL;
  unsigned char *mflag = reinterpret_cast<unsigned char *>(reinterpret_cast<short *>(ringbuf + (size_t)4 * RING) +
                                                          (size_t)4 * NCOL) + (size_t)warp * NCOL;
  const chess_params &p = a.p;
  const double qnan = CUDART_NAN;
  if (blockIdx.x == 0 && threadIdx.x == 0) { *a.counter_next = 0ull; *a.retry_next = 0; }
  const long long n_pair_items = HYB ? a.pair_items : 0;
  const long long n_items = n_pair_items + 2 * (a.n_pairs - n_pair_items);

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
    const bool PAIR = HYB && item < n_pair_items;
    const long long half = item - n_pair_items;   // index among the half items
    if (!tri_item_begin(a, PAIR ? item : n_pair_items + (half >> 1), PAIR ? 0 : (int)(half & 1), 31 * G, lane, kmap,
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
    if (tri_extrema(a, it, lane, mflag, xmax, xmin)) tri_scan_extrema(it, lane, kmap, xmax, xmin);
    const double drange = xmax - xmin;
    const double C1 = (0.01 * drange) * (0.01 * drange), C2 = (0.03 * drange) * (0.03 * drange);

    // ---- this item's lanes: rows [lo, hi), first column cbase
    int B1, B2, L1, L2;
    tri_split_halo(np_, G, HALO, B1, B2, L1, L2);
    double sn_sum = 0.0, sn_cnt = 0.0, ss_sum = 0.0;
    bool overflow = false;
    for (int part = part0; part <= (PAIR ? 1 : part0); ++part) {
    if (PAIR && part == 1) {
      if (lane == 0) *susp_n = 0;
      __syncwarp();
    }
    int lo = 0, hi = 0, cbase = 1 << 20;
    if (part == 0) {
      lo = 0; hi = B1; cbase = G * lane;
    } else if (lane < L1) {
      lo = B1; hi = B2; cbase = B1 - HALO + G * lane;
    } else if (lane > L1 && lane <= L1 + L2) {
      lo = B2; hi = np_; cbase = B2 - HALO + G * (lane - L1 - 1);
    }
    const int rows_max = part == 0 ? B1 : max(B2 - B1, np_ - B2);
    unsigned xo[G], yo[G];
    bool valid[G];
    double wc[G];   // column weight: 0 halo / missing, 1 inside the band's diagonal block, 2 right of it
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int c = cbase + g;
      valid[g] = c < np_;
      wc[g] = (!valid[g] || c < lo) ? 0.0 : (c < hi ? 1.0 : 2.0);
      const int kc = valid[g] ? (int)kmap[c] : 0;
      xo[g] = (unsigned)kc;
      yo[g] = (unsigned)(flip ? n - 1 - kc : kc);
    }

    if (rows_max > 0) {
      const int first = part == 0 ? 0 : max(lo - HALO, 0);       // idle lanes have lo = 0
      const int t_own0 = part == 0 ? 0 : HALO;                    // iteration of the band's first own row
      const int T = rows_max + (part == 0 ? HALO : 2 * HALO);
      constexpr int WIN = 2 * H + 1;
      const double inv_np = 1.0 / (double)(WIN * WIN);
      const double cov_norm = (double)(WIN * WIN) / (double)(WIN * WIN - 1);
      double ccnt[G], wss[G];
      double Sd[G], Se[G], Vx[G], Vy[G], Vxx[G], Vyy[G], Vxy[G];
      unsigned hist[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = cbase + g;
        ccnt[g] = (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1);
        wss[g] = (c >= H && c <= np_ - 1 - H) ? wc[g] : 0.0;      // SSIM windows need H columns either side
        Sd[g] = Se[g] = Vx[g] = Vy[g] = Vxx[g] = Vyy[g] = Vxy[g] = 0.0;
        hist[g] = 0u;
#pragma unroll
        for (int k = 0; k < D; ++k) { myring[((k * G + g) * 2) * 32] = 0.0; myring[((k * G + g) * 2 + 1) * 32] = 0.0; }
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
        // ring positions of the rows that leave the 7-row and the WIN-row windows
        const int s7 = slot + D - 7 >= D ? slot - 7 : slot + D - 7;        // row - 7
        const int sw = slot + D - WIN >= D ? slot - WIN : slot + D - WIN;  // row - WIN
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const double xo7 = myring[((s7 * G + g) * 2) * 32], yo7 = myring[((s7 * G + g) * 2 + 1) * 32];
          const double xow = myring[((sw * G + g) * 2) * 32], yow = myring[((sw * G + g) * 2 + 1) * 32];
          myring[((slot * G + g) * 2) * 32] = x[g];
          myring[((slot * G + g) * 2 + 1) * 32] = y[g];
          const double dv = x[g] - y[g], dold = xo7 - yo7;
          Sd[g] += dv; Sd[g] -= dold;
          Se[g] = fma(dv, dv, Se[g]); Se[g] = fma(-dold, dold, Se[g]);
          hist[g] = ((hist[g] << 1) & 0x7fu) | (x[g] != y[g] ? 1u : 0u);
          Se[g] = hist[g] == 0u ? 0.0 : Se[g];
          Vx[g] += x[g]; Vx[g] -= xow;
          Vy[g] += y[g]; Vy[g] -= yow;
          Vxx[g] = fma(x[g], x[g], Vxx[g]); Vxx[g] = fma(-xow, xow, Vxx[g]);
          Vyy[g] = fma(y[g], y[g], Vyy[g]); Vyy[g] = fma(-yow, yow, Vyy[g]);
          Vxy[g] = fma(x[g], y[g], Vxy[g]); Vxy[g] = fma(-xow, yow, Vxy[g]);
        }
        slot = slot == D - 1 ? 0 : slot + 1;
        if (slot == 0) {
          // the ring has just wrapped: rebuild the 7-row difference sums from its most recent seven rows
          // (slots D-7 .. D-1).  A running sum keeps the rounding of every value that ever passed through it;
          // rebuilt once per wrap it is exact to a few ulps of the current window
#pragma unroll
          for (int g = 0; g < G; ++g) {
            double s2 = 0.0;   // sum d*d only: an error in sum d enters the variance linearly and is harmless
#pragma unroll
            for (int k = D - 7; k < D; ++k) {
              const double dk = myring[((k * G + g) * 2) * 32] - myring[((k * G + g) * 2 + 1) * 32];
              s2 = fma(dk, dk, s2);
            }
            Se[g] = hist[g] == 0u ? 0.0 : s2;
          }
        }


        // ---- SSIM windows whose centre row is row - H
        if (t >= t_own0 + H) {
          const int rcs = row - H;
          const bool ssim_row = rcs >= lo && rcs < hi && rcs >= H && rcs <= np_ - 1 - H;
          double W5[5][G];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const double *V = q == 0 ? Vx : q == 1 ? Vy : q == 2 ? Vxx : q == 3 ? Vyy : Vxy;
            double P[G + 1];
            P[0] = 0.0; P[1] = V[0];
#pragma unroll
            for (int g = 1; g < G; ++g) P[g + 1] = P[g] + V[g];
            double up[H], dn[H];
#pragma unroll
            for (int j = 0; j < H; ++j) {
              const double sfx = j == 0 ? V[G - 1] : (G - 1 - j == 0 ? P[G] : P[G] - P[G - 1 - j]);
              up[j] = __shfl_sync(0xffffffffu, sfx, up_lane);
              dn[j] = __shfl_sync(0xffffffffu, P[j + 1], dn_lane);
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const int a0 = g - H < 0 ? 0 : g - H;
              const int b0 = g + H > G - 1 ? G - 1 : g + H;
              double w = a0 == 0 ? P[b0 + 1] : P[b0 + 1] - P[a0];
              if (g - H < 0) w += up[H - 1 - g];
              if (g + H > G - 1) w += dn[g + H - G];
              W5[q][g] = w;
            }
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const double ux = W5[0][g] * inv_np, uy = W5[1][g] * inv_np;
            const double uxx = W5[2][g] * inv_np, uyy = W5[3][g] * inv_np, uxy = W5[4][g] * inv_np;
            const double vx = cov_norm * fma(-ux, ux, uxx);
            const double vy = cov_norm * fma(-uy, uy, uyy);
            const double vxy = cov_norm * fma(-ux, uy, uxy);
            const double A1 = fma(2.0 * ux, uy, C1), A2 = fma(2.0, vxy, C2);
            const double B1s = fma(ux, ux, fma(uy, uy, C1)), B2s = vx + vy + C2;
            const double den = B1s * B2s;
            const double num = A1 * A2;
            double S = num * fast_rcp(den);
            if (den == 0.0) S = num * CUDART_INF;  // zero data range: +-inf or NaN exactly like num / 0
            if (ssim_row && wss[g] != 0.0) ss_sum = fma(wss[g], S, ss_sum);
          }
        }

        // ---- SN pixels of row - 3
        if (with_sn && t >= t_own0 + 3) {
          const int rc = row - 3;
          const bool emit_row = rc >= lo && rc < hi;
          const double rcnt = (double)(min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1);
          double P1[G + 1], P2[G + 1];
          P1[0] = 0.0; P2[0] = 0.0;
          P1[1] = Sd[0]; P2[1] = Se[0];
#pragma unroll
          for (int g = 1; g < G; ++g) { P1[g + 1] = P1[g] + Sd[g]; P2[g + 1] = P2[g] + Se[g]; }
          // suffix sums as well: a window that ends at the strip's right edge is a suffix, and no window of 7
          // starts after column 0 without reaching that edge (G <= 8), so nothing is ever subtracted -- a
          // difference of prefixes would carry the rounding of every column left of the window
          double X1[G], X2[G];
          X1[G - 1] = Sd[G - 1]; X2[G - 1] = Se[G - 1];
#pragma unroll
          for (int g = G - 2; g >= 0; --g) { X1[g] = X1[g + 1] + Sd[g]; X2[g] = X2[g + 1] + Se[g]; }
          unsigned suspect = 0u;
          double up1[3], up2[3], dn1[3], dn2[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double s1 = X1[G - 1 - j < 0 ? 0 : G - 1 - j];
            const double s2 = X2[G - 1 - j < 0 ? 0 : G - 1 - j];
            up1[j] = __shfl_sync(0xffffffffu, s1, up_lane);
            up2[j] = __shfl_sync(0xffffffffu, s2, up_lane);
            dn1[j] = __shfl_sync(0xffffffffu, P1[j + 1], dn_lane);
            dn2[j] = __shfl_sync(0xffffffffu, P2[j + 1], dn_lane);
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int a0 = g - 3 < 0 ? 0 : g - 3;
            const int b0 = g + 3 > G - 1 ? G - 1 : g + 3;
            double w1 = a0 == 0 ? P1[b0 + 1] : X1[a0];
            double w2 = a0 == 0 ? P2[b0 + 1] : X2[a0];
            if (g - 3 < 0) { w1 += up1[2 - g]; w2 += up2[2 - g]; }
            if (g + 3 > G - 1) { w1 += dn1[g + 3 - G]; w2 += dn2[g + 3 - G]; }
            const bool own = emit_row & (wc[g] != 0.0);
            const double cnt = rcnt * ccnt[g];
            const double m2 = w2 * cnt;
            const double den = fma(-w1, w1, m2);
            const bool good = own & (den > fma(m2, kVarGuard, 1e-290));
            const double gq = fabs(w1) * cnt * fast_rcp(den);
            sn_sum = fma(wc[g], good ? gq : 0.0, sn_sum);
            sn_cnt += good ? wc[g] : 0.0;
            if (!good && own && w2 != 0.0) suspect |= 1u << g;  // rare: cancellation
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
            const int band_hi = part == 0 ? B1 : (rc < B2 ? B2 : np_);  // end of the row's diagonal block
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
      bool bad = overflow || !tri_masks_hold(a, it, lane, kmap, mflag);
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
    bad |= !tri_masks_hold(a, it, lane, kmap, mflag);
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

template <int G, int H>
int launch_triw(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  auto kern = compare_win_tri_kernel<G, H>;
  constexpr int D = 2 * H + 1 > 7 ? 2 * H + 1 : 7;
  constexpr int RING = D * 2 * G * 32;
  const size_t smem = (size_t)4 * RING * sizeof(double) + (size_t)4 * a.ncol_pad * 3;
  const size_t smem_max = (size_t)4 * RING * sizeof(double) + (size_t)4 * 256 * 3;
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
  // narrow strips: whole waves of pairs run as pair items (compare_tri.cuh: launch_tri)
  const long long slots = grid * 4;
  long long pair_items = G > 4 ? 0 : (a.n_pairs >= 8 * slots ? a.n_pairs : (a.n_pairs / slots) * slots);
  static const int pair_env = getenv("CHESS_TRI_PAIR") ? atoi(getenv("CHESS_TRI_PAIR")) : -1;  // A/B runs
  if (pair_env == 0) pair_items = 0;
  if (pair_env == 1 && G <= 4) pair_items = a.n_pairs;
  a.pair_items = pair_items;
  const long long need = (pair_items + 2 * (a.n_pairs - pair_items) + 3) / 4;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, 128, smem, stream>>>(a);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

}  // namespace
