// Device code shared by the tuned compare kernels (compare_fast.cu and the items_h*.cu
// translation units, which are compiled in parallel): band loops, the work-item kernel and its
// launcher template.  Everything lives in an anonymous namespace, one copy per translation unit.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace {

// Reciprocal: rcp.approx.ftz.f64 (MUFU.RCP64H, measured max relative error 9.9e-7 on B200,
// profiles/r01_microbench.txt) refined by one Newton step -> 1e-12, five orders below the 1e-6
// parity budget, at 1 MUFU + 2 DFMA instead of the ~13 issue slots of an IEEE division.
__device__ __forceinline__ double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  return fma(y, e, y);
}

// numpy's pairwise summation for n = 49 contiguous doubles (n < 128: eight running accumulators
// over blocks of eight, combined as a tree, then the tail).  Deliberately not unrolled: this is the
// cold path and must stay small in registers, because its frame adds to the hot kernel's.
__device__ __noinline__ double np_sum49(const double *a) {
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
#pragma unroll 1
  for (int i = 8; i < 48; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  return __dadd_rn(res, a[48]);
}

// Exact (numpy-order, two-pass) |nanmean| / nanvar of the 7x7 window at (rc, c),
// read from memory through the compaction map.
__device__ __noinline__ double sn_window_exact(const double *R, const double *Q, size_t rp, size_t qp,
                                               const short *kmap, int n, int flip, int rc, int c,
                                               int np_, double cnt) {
  double val[49];
  unsigned long long okmask = 0ull;
#pragma unroll 1
  for (int wi = 0; wi < 7; ++wi) {
    const int rr = rc - 3 + wi;
    const bool rok = rr >= 0 && rr < np_;
    const int kr = rok ? kmap[rr] : 0;
#pragma unroll 1
    for (int wj = 0; wj < 7; ++wj) {
      const int cc = c - 3 + wj;
      const bool v = rok && cc >= 0 && cc < np_;
      double dv = 0.0;
      if (v) {
        const int kc = kmap[cc];
        const double x = __ldg(R + (size_t)kr * rp + kc);
        const double y = flip ? __ldg(Q + (size_t)(n - 1 - kr) * qp + (n - 1 - kc))
                              : __ldg(Q + (size_t)kr * qp + kc);
        dv = __dsub_rn(x, y);
        okmask |= 1ull << (wi * 7 + wj);
      }
      val[wi * 7 + wj] = dv;
    }
  }
  const double avg = __ddiv_rn(np_sum49(val), cnt);
#pragma unroll 1
  for (int s = 0; s < 49; ++s) {
    const double dv = ((okmask >> s) & 1ull) ? __dsub_rn(val[s], avg) : 0.0;
    val[s] = __dmul_rn(dv, dv);
  }
  const double var = __ddiv_rn(np_sum49(val), cnt);
  return __ddiv_rn(fabs(avg), var);
}

__device__ __forceinline__ bool is_nonzero(double v) {
  return ((unsigned long long)__double_as_longlong(v) << 1) != 0ull;  // +-0 -> false, everything else true
}

constexpr int kSuspCap = 24;

// The one-pass variance of a 7x7 window, cnt*sum(d^2) - sum(d)^2, is trusted when it keeps at least this
// fraction of cnt*sum(d^2): the running column sums and prefix differences carry some hundred ulps, so the
// result is good to ~1e-14 / kVarGuard.  Windows below it are recomputed in numpy's own two-pass order.
constexpr double kVarGuard = 1e-4;

struct Band2 {
  double T0, T1, T2, T3, T4, vmax, vmin, sn_sum;
  int sn_cnt;
};

// Per-lane view of the compacted pair: column offsets of the lane's G columns in both matrices.
template <int G, bool MULTI>
struct LaneCols {
  unsigned xo[G], yo[G];   // element offsets inside a row of the reference / query view
  bool valid[G];           // column exists (0 <= c < np): it is loaded and feeds the window sums
  bool emit[G];            // column belongs to this column block: it produces outputs / totals
  int cbase;               // compacted index of the lane's first column (negative in a left halo)
  // single-block kernels (MULTI = false) emit every valid column: the compiler sees one predicate
  __device__ __forceinline__ bool em(int g) const { return MULTI ? emit[g] : valid[g]; }
};

// Matrices wider than one warp's 31*G columns are swept in column blocks: a block starts three
// columns before and ends three columns after the columns it emits, so the 7-column (2H+1-column)
// sums of its own columns see real neighbours; the halo columns are loaded and summed but emit
// nothing.  A matrix that fits one block has cstart = 0 and emits every column.
template <int G, bool MULTI>
__device__ __forceinline__ void lane_cols(LaneCols<G, MULTI> &lc, const short *kmap, int n, int np_, int flip,
                                          int lane, int cstart, int estart, int eend) {
  lc.cbase = cstart + G * lane;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = lc.cbase + g;
    lc.valid[g] = c >= 0 && c < np_;
    lc.emit[g] = MULTI ? (lc.valid[g] && c >= estart && c < eend) : lc.valid[g];
    const int kc = lc.valid[g] ? (int)kmap[c] : 0;
    lc.xo[g] = (unsigned)kc;
    lc.yo[g] = (unsigned)(flip ? n - 1 - kc : kc);
  }
}

// Loop A of pass 2: moments of x and y and min / max over the rows [lo, hi) of the band.
// COLSUM additionally accumulates the lane's column sums over those rows (the item kernel uses
// them to verify its predicted masks).
template <int G, bool COLSUM, bool MINMAX, bool MULTI>
__device__ __forceinline__ void moments_rows(const double *__restrict__ R, const double *__restrict__ Q,
                                             size_t rp, size_t qp, const short *kmap, int n, int flip,
                                             int lo, int hi, const LaneCols<G, MULTI> &lc, Band2 &acc,
                                             double *cs_r, double *cs_q) {
  double T0 = 0, T1 = 0, T2 = 0, T3 = 0, T4 = 0;
  double vmax = -CUDART_INF, vmin = CUDART_INF;
#pragma unroll 4
  for (int rr = lo; rr < hi; ++rr) {
    // all offsets inside one pair fit 32 bits: one IMAD per row, one add + one wide MAD per load
    const unsigned kr = (unsigned)kmap[rr];
    const unsigned ro = kr * (unsigned)rp;
    const unsigned qo = (flip ? (unsigned)(n - 1) - kr : kr) * (unsigned)qp;
    double x[G], y[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      x[g] = 0.0; y[g] = 0.0;
      if (lc.em(g)) { x[g] = __ldg(R + (ro + lc.xo[g])); y[g] = __ldg(Q + (qo + lc.yo[g])); }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      T0 += x[g]; T1 += y[g];
      T2 = fma(x[g], x[g], T2); T3 = fma(y[g], y[g], T3); T4 = fma(x[g], y[g], T4);
      if (COLSUM) { cs_r[g] += x[g]; cs_q[g] += y[g]; }
      if (MINMAX) {
        // branch-free compare-selects (the data are finite: no NaN handling needed)
        const bool xgt = x[g] > y[g];
        const double hi2 = xgt ? x[g] : y[g];
        const double lo2 = xgt ? y[g] : x[g];
        vmax = (lc.em(g) & (hi2 > vmax)) ? hi2 : vmax;
        vmin = (lc.em(g) & (lo2 < vmin)) ? lo2 : vmin;
      }
    }
  }
  acc.T0 = T0; acc.T1 = T1; acc.T2 = T2; acc.T3 = T3; acc.T4 = T4;
  acc.vmax = vmax; acc.vmin = vmin;
}

// Loop B of pass 2: the SN map of the band rows [lo, hi) (three halo rows either side feed the
// 7-row column sums).
// MOMENTS fuses loop A (without min / max) and the column sums into this loop, so the pair is read
// once: used by the item kernel when the data range comes from the running-extrema arrays.
template <int G, bool MOMENTS, bool MULTI>
__device__ __forceinline__ void sn_rows(const double *__restrict__ R, const double *__restrict__ Q,
                                        size_t rp, size_t qp, const short *kmap, int n, int np_, int flip,
                                        int lo, int hi, int lane, const LaneCols<G, MULTI> &lc, double *myring,
                                        int *susp, int *susp_n, Band2 &acc, double *cs_r, double *cs_q) {
  const int c0 = lc.cbase;
  double ccnt[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = c0 + g;
    ccnt[g] = (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1);
  }
  double sn_sum = 0.0;
  int sn_cnt = 0;
  double T0 = 0, T1 = 0, T2 = 0, T3 = 0, T4 = 0;
  double Sd[G], Se[G];
  unsigned hist[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    Sd[g] = 0.0; Se[g] = 0.0; hist[g] = 0u;
#pragma unroll
    for (int k = 0; k < 7; ++k) myring[(k * G + g) * 32] = 0.0;
  }
  const int first = lo - 3, last = hi + 3;  // exclusive
  const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
  double *ringrow = myring;
  int slot = 0;
  // software pipeline: the loads of row rr+1 are issued before row rr is processed
  double xn[G], yn[G];
  auto fetch = [&](int row) {
#pragma unroll
    for (int g = 0; g < G; ++g) { xn[g] = 0.0; yn[g] = 0.0; }
    if (row >= 0 && row < np_) {
      const unsigned kr = (unsigned)kmap[row];
      const unsigned ro = kr * (unsigned)rp;
      const unsigned qo = (flip ? (unsigned)(n - 1) - kr : kr) * (unsigned)qp;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (lc.valid[g]) { xn[g] = __ldg(R + (ro + lc.xo[g])); yn[g] = __ldg(Q + (qo + lc.yo[g])); }
      }
    }
  };
  fetch(first);
  for (int rr = first; rr < last; ++rr) {
    double dv[G];
    bool nzv[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { dv[g] = xn[g] - yn[g]; nzv[g] = xn[g] != yn[g]; }
    if (MOMENTS && rr >= lo && rr < hi) {  // the band's own rows (halo rows belong to the neighbours)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (lc.em(g)) {
          T0 += xn[g]; T1 += yn[g];
          T2 = fma(xn[g], xn[g], T2); T3 = fma(yn[g], yn[g], T3); T4 = fma(xn[g], yn[g], T4);
          cs_r[g] += xn[g]; cs_q[g] += yn[g];
        }
      }
    }
    if (rr + 1 < last) fetch(rr + 1);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      double *cell = ringrow + g * 32;
      const double dold = *cell;
      *cell = dv[g];
      Sd[g] += dv[g]; Sd[g] -= dold;
      Se[g] = fma(dv[g], dv[g], Se[g]); Se[g] = fma(-dold, dold, Se[g]);
      // 7-row history of "d != 0": an all-zero column window is snapped to exactly 0, so an
      // all-zero 7x7 window gives w2 == 0 exactly (0/0 -> NaN -> skipped by np.nanmean)
      hist[g] = ((hist[g] << 1) & 0x7fu) | (nzv[g] ? 1u : 0u);
      Se[g] = hist[g] == 0u ? 0.0 : Se[g];
    }
    if (slot == 6) { slot = 0; ringrow = myring; } else { slot += 1; ringrow += G * 32; }
    if (slot == 0) {
      // the ring has just wrapped: rebuild the column sums from its seven rows.  A running sum keeps the
      // rounding of every value that ever passed through it; rebuilt once per window height it is exact
      // to a few ulps of the current window, which the quiet windows after a loud stretch need
#pragma unroll
      for (int g = 0; g < G; ++g) {
        double s2 = 0.0;   // sum d*d only: an error in sum d enters the variance linearly and is harmless
#pragma unroll
        for (int k = 0; k < 7; ++k) { const double v = myring[(k * G + g) * 32]; s2 = fma(v, v, s2); }
        Se[g] = hist[g] == 0u ? 0.0 : s2;
      }
    }
    const int rc = rr - 3;
    if (rc >= lo && rc < hi) {
      const double rcnt = (double)(min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1);
      // strip prefix sums, then three partial sums from each neighbour lane (lane 31 is never
      // valid -- the launcher guarantees n <= 31*G -- so the rotation wraps onto zeros)
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
      double up1[3], up2[3], dn1[3], dn2[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        // suffix of j+1 columns of the previous lane, prefix of j+1 columns of the next lane
        const double s1 = X1[G - 1 - j < 0 ? 0 : G - 1 - j];
        const double s2 = X2[G - 1 - j < 0 ? 0 : G - 1 - j];
        up1[j] = __shfl_sync(0xffffffffu, s1, up_lane);
        up2[j] = __shfl_sync(0xffffffffu, s2, up_lane);
        dn1[j] = __shfl_sync(0xffffffffu, P1[j + 1], dn_lane);
        dn2[j] = __shfl_sync(0xffffffffu, P2[j + 1], dn_lane);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int a = g - 3 < 0 ? 0 : g - 3;
        const int b = g + 3 > G - 1 ? G - 1 : g + 3;
        double w1 = a == 0 ? P1[b + 1] : X1[a];
        double w2 = a == 0 ? P2[b + 1] : X2[a];
        if (g - 3 < 0) { w1 += up1[2 - g]; w2 += up2[2 - g]; }
        if (g + 3 > G - 1) { w1 += dn1[g + 3 - G]; w2 += dn2[g + 3 - G]; }
        const double cnt = rcnt * ccnt[g];
        const double m2 = w2 * cnt;
        const double den = fma(-w1, w1, m2);
        // one-pass variance is trusted when var > 1e-6 * E[d^2] (and E[d^2] is not denormal-small)
        const bool good = lc.em(g) & (den > fma(m2, kVarGuard, 1e-290));
        const double gq = fabs(w1) * cnt * fast_rcp(den);
        sn_sum += good ? gq : 0.0;
        sn_cnt += good ? 1 : 0;
        if (!good && lc.em(g) && w2 != 0.0) {  // rare: cancellation -> exact recompute after the loop
          const int e = atomicAdd(susp_n, 1);
          if (e < kSuspCap) susp[e] = (rc << 16) | (c0 + g);
        }
      }
    }
  }
  acc.sn_sum = sn_sum;
  acc.sn_cnt = sn_cnt;
  if (MOMENTS) { acc.T0 = T0; acc.T1 = T1; acc.T2 = T2; acc.T3 = T3; acc.T4 = T4; }
}

// Band loop of the windowed regime (-a 3 / 5 / 7 ...: SSIM window 2H+1 much smaller than the
// matrix, chess/sim.py:356-365), one pass over the pair:
//   * per column, running sums over the last 2H+1 rows of x, y, xx, yy, xy (the row that leaves
//     the window comes from a ring of x and y in shared memory), and over the last 7 rows of d and
//     d*d for SN exactly as in sn_rows (same ring: d = x - y is recomputed bit-identically);
//   * per emitted row, sums over 2H+1 (7 for SN) columns by exchanging H (3) strip prefix / suffix
//     sums with the two neighbour lanes, then skimage's SSIM formula on every valid window and
//     |mean| / var on every pixel;
//   * the band's own rows also feed the column sums that verify the predicted masks.
// C1 / C2 need the data range before the first window: it comes from the running-extrema arrays
// (or a min / max pre-loop when those cannot be used).
template <int G, int H, bool MULTI>
__device__ __forceinline__ void win_rows(const double *__restrict__ R, const double *__restrict__ Q,
                                         size_t rp, size_t qp, const short *kmap, int n, int np_, int flip,
                                         int lo, int hi, int lane, const LaneCols<G, MULTI> &lc, double *myring,
                                         int *susp, int *susp_n, bool with_sn, double C1, double C2,
                                         Band2 &acc, double &ssim_sum, double *cs_r, double *cs_q) {
  static_assert(H >= 1 && H <= 3 && G >= 3, "neighbour exchange covers at most 3 columns per side");
  constexpr int WIN = 2 * H + 1;
  constexpr int D = WIN > 7 ? WIN : 7;  // ring depth in rows
  const int c0 = lc.cbase;
  const double inv_np = 1.0 / (double)(WIN * WIN);
  const double cov_norm = (double)(WIN * WIN) / (double)(WIN * WIN - 1);
  double ccnt[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = c0 + g;
    ccnt[g] = (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1);
  }
  double sn_sum = 0.0, ss_sum = 0.0;
  int sn_cnt = 0;
  double Sd[G], Se[G], Vx[G], Vy[G], Vxx[G], Vyy[G], Vxy[G];
  unsigned hist[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    Sd[g] = Se[g] = Vx[g] = Vy[g] = Vxx[g] = Vyy[g] = Vxy[g] = 0.0;
    hist[g] = 0u;
#pragma unroll
    for (int k = 0; k < D; ++k) { myring[((k * G + g) * 2) * 32] = 0.0; myring[((k * G + g) * 2 + 1) * 32] = 0.0; }
  }
  const int first = lo - 3, last = hi + 3;  // exclusive; 3 >= H
  const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
  int slot = 0;
  double xn[G], yn[G];
  auto fetch = [&](int row) {
#pragma unroll
    for (int g = 0; g < G; ++g) { xn[g] = 0.0; yn[g] = 0.0; }
    if (row >= 0 && row < np_) {
      const unsigned kr = (unsigned)kmap[row];
      const unsigned ro = kr * (unsigned)rp;
      const unsigned qo = (flip ? (unsigned)(n - 1) - kr : kr) * (unsigned)qp;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (lc.valid[g]) { xn[g] = __ldg(R + (ro + lc.xo[g])); yn[g] = __ldg(Q + (qo + lc.yo[g])); }
      }
    }
  };
  fetch(first);
  for (int rr = first; rr < last; ++rr) {
    double x[G], y[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { x[g] = xn[g]; y[g] = yn[g]; }
    if (rr >= lo && rr < hi) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (lc.em(g)) { cs_r[g] += x[g]; cs_q[g] += y[g]; }
      }
    }
    if (rr + 1 < last) fetch(rr + 1);
    // ring positions of the rows that leave the 7-row and the WIN-row windows
    const int s7 = slot + D - 7 >= D ? slot - 7 : slot + D - 7;       // row rr - 7
    const int sw = slot + D - WIN >= D ? slot - WIN : slot + D - WIN;  // row rr - WIN
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const double xo7 = myring[((s7 * G + g) * 2) * 32], yo7 = myring[((s7 * G + g) * 2 + 1) * 32];
      double xow = xo7, yow = yo7;
      if (WIN != 7) { xow = myring[((sw * G + g) * 2) * 32]; yow = myring[((sw * G + g) * 2 + 1) * 32]; }
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


    // ---- SSIM windows whose centre row is rr - H
    const int rcs = rr - H;
    if (rcs >= lo && rcs < hi && rcs >= H && rcs <= np_ - 1 - H) {
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
          const int a = g - H < 0 ? 0 : g - H;
          const int b = g + H > G - 1 ? G - 1 : g + H;
          double w = a == 0 ? P[b + 1] : P[b + 1] - P[a];
          if (g - H < 0) w += up[H - 1 - g];
          if (g + H > G - 1) w += dn[g + H - G];
          W5[q][g] = w;
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = c0 + g;
        const double ux = W5[0][g] * inv_np, uy = W5[1][g] * inv_np;
        const double uxx = W5[2][g] * inv_np, uyy = W5[3][g] * inv_np, uxy = W5[4][g] * inv_np;
        const double vx = cov_norm * fma(-ux, ux, uxx);
        const double vy = cov_norm * fma(-uy, uy, uyy);
        const double vxy = cov_norm * fma(-ux, uy, uxy);
        const double A1 = fma(2.0 * ux, uy, C1), A2 = fma(2.0, vxy, C2);
        const double B1 = fma(ux, ux, fma(uy, uy, C1)), B2 = vx + vy + C2;
        const double den = B1 * B2;
        const double num = A1 * A2;
        double S = num * fast_rcp(den);
        if (den == 0.0) S = num * CUDART_INF;  // zero data range: +-inf or NaN exactly like num / 0
        if (lc.em(g) && c >= H && c <= np_ - 1 - H) ss_sum += S;
      }
    }

    // ---- SN pixels of row rr - 3 (identical to sn_rows)
    const int rc = rr - 3;
    if (with_sn && rc >= lo && rc < hi) {
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
        const int a = g - 3 < 0 ? 0 : g - 3;
        const int b = g + 3 > G - 1 ? G - 1 : g + 3;
        double w1 = a == 0 ? P1[b + 1] : X1[a];
        double w2 = a == 0 ? P2[b + 1] : X2[a];
        if (g - 3 < 0) { w1 += up1[2 - g]; w2 += up2[2 - g]; }
        if (g + 3 > G - 1) { w1 += dn1[g + 3 - G]; w2 += dn2[g + 3 - G]; }
        const double cnt = rcnt * ccnt[g];
        const double m2 = w2 * cnt;
        const double den = fma(-w1, w1, m2);
        const bool good = lc.em(g) & (den > fma(m2, kVarGuard, 1e-290));
        const double gq = fabs(w1) * cnt * fast_rcp(den);
        sn_sum += good ? gq : 0.0;
        sn_cnt += good ? 1 : 0;
        if (!good && lc.em(g) && w2 != 0.0) {
          const int e = atomicAdd(susp_n, 1);
          if (e < kSuspCap) susp[e] = (rc << 16) | (c0 + g);
        }
      }
    }
  }
  acc.sn_sum = sn_sum;
  acc.sn_cnt = sn_cnt;
  ssim_sum = ss_sum;
}

// Window 7 (-a 7) is also the SN window, so one set of sums serves both: running 7-row column
// sums of x, y, q = xx + yy and xy (the SSIM formula needs the two variances only as a sum), four
// quantities through the neighbour exchange instead of seven, and
//     sum d = Wx - Wy,   sum d*d = Wq - 2 Wxy.
// An all-zero difference window must give exactly 0 for sum d*d (0/0 -> NaN -> skipped): a column
// whose last seven rows all have x == y is snapped to Vy = Vx and Vq = 2 Vxy, and doubling commutes
// with every rounding in the prefix / suffix arithmetic, so Wq - 2 Wxy cancels exactly.  Windows
// where the subtraction loses more than ~6 digits (sum d*d < 1e-6 sum q) go to the exact recompute.
template <int G, bool MULTI>
__device__ __forceinline__ void win7_rows(const double *__restrict__ R, const double *__restrict__ Q,
                                          size_t rp, size_t qp, const short *kmap, int n, int np_, int flip,
                                          int lo, int hi, int lane, const LaneCols<G, MULTI> &lc,
                                          double *myring, int *susp, int *susp_n, bool with_sn, double C1,
                                          double C2, Band2 &acc, double &ssim_sum, double *cs_r,
                                          double *cs_q) {
  constexpr int H = 3, D = 7;
  const int c0 = lc.cbase;
  const double inv_np = 1.0 / 49.0;
  const double cov_norm = 49.0 / 48.0;
  double ccnt[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = c0 + g;
    ccnt[g] = (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1);
  }
  double sn_sum = 0.0, ss_sum = 0.0;
  int sn_cnt = 0;
  double Vx[G], Vy[G], Vq[G], Vxy[G];
  unsigned hist[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    Vx[g] = Vy[g] = Vq[g] = Vxy[g] = 0.0;
    hist[g] = 0u;
#pragma unroll
    for (int k = 0; k < D; ++k) { myring[((k * G + g) * 2) * 32] = 0.0; myring[((k * G + g) * 2 + 1) * 32] = 0.0; }
  }
  const int first = lo - 3, last = hi + 3;  // exclusive
  const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
  int slot = 0;
  double xn[G], yn[G];
  auto fetch = [&](int row) {
#pragma unroll
    for (int g = 0; g < G; ++g) { xn[g] = 0.0; yn[g] = 0.0; }
    if (row >= 0 && row < np_) {
      const unsigned kr = (unsigned)kmap[row];
      const unsigned ro = kr * (unsigned)rp;
      const unsigned qo = (flip ? (unsigned)(n - 1) - kr : kr) * (unsigned)qp;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (lc.valid[g]) { xn[g] = __ldg(R + (ro + lc.xo[g])); yn[g] = __ldg(Q + (qo + lc.yo[g])); }
      }
    }
  };
  fetch(first);
  for (int rr = first; rr < last; ++rr) {
    double x[G], y[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { x[g] = xn[g]; y[g] = yn[g]; }
    if (rr >= lo && rr < hi) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (lc.em(g)) { cs_r[g] += x[g]; cs_q[g] += y[g]; }
      }
    }
    if (rr + 1 < last) fetch(rr + 1);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      double *cell = myring + ((slot * G + g) * 2) * 32;   // the slot being overwritten holds row rr - 7
      const double xo = cell[0], yo = cell[32];
      cell[0] = x[g];
      cell[32] = y[g];
      Vx[g] += x[g]; Vx[g] -= xo;
      Vy[g] += y[g]; Vy[g] -= yo;
      Vq[g] = fma(x[g], x[g], fma(y[g], y[g], Vq[g]));
      Vq[g] = fma(-xo, xo, fma(-yo, yo, Vq[g]));
      Vxy[g] = fma(x[g], y[g], Vxy[g]); Vxy[g] = fma(-xo, yo, Vxy[g]);
      hist[g] = ((hist[g] << 1) & 0x7fu) | (x[g] != y[g] ? 1u : 0u);
      if (hist[g] == 0u) { Vy[g] = Vx[g]; Vq[g] = 2.0 * Vxy[g]; }
    }
    slot = slot == D - 1 ? 0 : slot + 1;
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


    const int rc = rr - 3;   // centre row of the SSIM window and SN pixel row
    if (rc >= lo && rc < hi) {
      const bool ssim_row = rc >= H && rc <= np_ - 1 - H;
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
          const int a = g - 3 < 0 ? 0 : g - 3;
          const int b = g + 3 > G - 1 ? G - 1 : g + 3;
          double w = a == 0 ? P[b + 1] : P[b + 1] - P[a];
          if (g - 3 < 0) w += up[2 - g];
          if (g + 3 > G - 1) w += dn[g + 3 - G];
          W4[q][g] = w;
        }
      }
      const double rcnt = (double)(min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = c0 + g;
        if (ssim_row) {
          const double ux = W4[0][g] * inv_np, uy = W4[1][g] * inv_np;
          const double uq = W4[2][g] * inv_np, uxy = W4[3][g] * inv_np;
          const double m2s = fma(ux, ux, uy * uy);                  // ux^2 + uy^2
          const double vsum = cov_norm * (uq - m2s);                // vx + vy
          const double vxy = cov_norm * fma(-ux, uy, uxy);
          const double A1 = fma(2.0 * ux, uy, C1), A2 = fma(2.0, vxy, C2);
          const double B1 = m2s + C1, B2 = vsum + C2;
          const double den = B1 * B2;
          const double num = A1 * A2;
          double S = num * fast_rcp(den);
          if (den == 0.0) S = num * CUDART_INF;  // zero data range: +-inf or NaN exactly like num / 0
          if (lc.em(g) && c >= H && c <= np_ - 1 - H) ss_sum += S;
        }
        if (with_sn) {
          const double w1 = W4[0][g] - W4[1][g];
          const double w2 = fma(-2.0, W4[3][g], W4[2][g]);
          const double cnt = rcnt * ccnt[g];
          const double m2 = w2 * cnt;
          const double den = fma(-w1, w1, m2);
          // trusted: one-pass variance well conditioned AND the q - 2xy subtraction kept >= 10 digits
          const bool good = lc.em(g) & (den > fma(m2, kVarGuard, 1e-290)) & (w2 > 1e-6 * W4[2][g]);
          const double gq = fabs(w1) * cnt * fast_rcp(den);
          sn_sum += good ? gq : 0.0;
          sn_cnt += good ? 1 : 0;
          if (!good && lc.em(g) && w2 != 0.0) {
            const int e = atomicAdd(susp_n, 1);
            if (e < kSuspCap) susp[e] = (rc << 16) | (c0 + g);
          }
        }
      }
    }
  }
  acc.sn_sum = sn_sum;
  acc.sn_cnt = sn_cnt;
  ssim_sum = ss_sum;
}

// =====================================================================================
// compare_r1_items_kernel -- the same computation as compare_r1_kernel, reorganised so that
// nothing waits on anything:
//   * a work item is (pair, row band); warps take items from a global counter, so a launch of
//     P pairs is 4P independent pieces and the tail of the grid costs a quarter pair, not a pair;
//   * the masks are *predicted* in O(1) per column from a per-bin index built once per sample
//     (nearest bin with a non-default entry on either side, chess_band_index_build): a column is
//     all-default -- hence masked, chess/sim.py:82-85 -- iff that neighbour lies outside the
//     window.  This removes the separate column-sum pass and every CTA barrier;
//   * each band writes its partial results (moments, min / max, SN sum / count, first / last row
//     totals, its share of every column sum) to a scratch record; the band that arrives last
//     (atomic counter, no spinning) checks the prediction against the column sums -- a column
//     that is not all-default is masked only if its sum hits n*default exactly; such a
//     coincidence is confirmed with the reference's own sequential sum and sends the pair to the
//     generic kernel -- and otherwise combines the bands in fixed order and writes the result.
constexpr int kBands = 2;
constexpr int kItemHdr = 24;



__device__ __forceinline__ void finish_pair(const double *t, double mx, double mn, double ss, double sc,
                                            const double *row0, const double *rowL, int np_, int win,
                                            int m, const short *kmap, const double *R, const double *Q,
                                            size_t rp, size_t qp, int n, int flip, bool with_sn,
                                            double *ssim_out, double *sn_out, int *status_out) {
  const double drange = mx - mn;
  const double NP = (double)win * (double)win;
  const double cov_norm = NP / (NP - 1.0);
  const double C1 = (0.01 * drange) * (0.01 * drange);
  const double C2 = (0.03 * drange) * (0.03 * drange);
  double acc = 0.0;
  for (int a = 0; a < m; ++a) {
    for (int b = 0; b < m; ++b) {
      double h[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) h[k] = t[k];
      if (m == 2) {
        // window (a, b) leaves out one border row and one border column; column totals = row
        // totals because every moment map of a symmetric pair is symmetric
        const bool ra = a == 0, rb = b == 0;  // true: the last row / column is left out
        const int er = ra ? np_ - 1 : 0, ec = rb ? np_ - 1 : 0;
        const int kr = kmap[er], kcc = kmap[ec];
        const double x = __ldg(R + (size_t)kr * rp + kcc);
        const double y = flip ? __ldg(Q + (size_t)(n - 1 - kr) * qp + (n - 1 - kcc))
                              : __ldg(Q + (size_t)kr * qp + kcc);
        const double corner[5] = {x, y, x * x, y * y, x * y};
#pragma unroll
        for (int k = 0; k < 5; ++k)
          h[k] = h[k] - (ra ? rowL[k] : row0[k]) - (rb ? rowL[k] : row0[k]) + corner[k];
      }
      const double ux = h[0] / NP, uy = h[1] / NP;
      const double uxx = h[2] / NP, uyy = h[3] / NP, uxy = h[4] / NP;
      const double vx = cov_norm * (uxx - ux * ux);
      const double vy = cov_norm * (uyy - uy * uy);
      const double vxy = cov_norm * (uxy - ux * uy);
      const double A1 = 2.0 * ux * uy + C1, A2 = 2.0 * vxy + C2;
      const double B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
      acc += (A1 * A2) / (B1 * B2);
    }
  }
  *ssim_out = acc / (double)(m * m);
  *sn_out = with_sn ? ss / sc : CUDART_NAN;
  *status_out = CHESS_PAIR_OK;
}

// H = 0: the -r 1 regime (window = whole matrix); H = 1..3: windowed regime with window 2H+1,
// where WITH_SN is ignored and a.p.compute_sn decides at run time.
// MULTI = true sweeps matrices wider than 31*G columns in column blocks (n <= 1024).
template <int G, bool WITH_SN, int H, bool MULTI>
__global__ void __launch_bounds__(128, (G <= 4 ? (H == 0 ? 4 : 3) : 2))
compare_r1_items_kernel(chess_item_args a) {
  __shared__ int s_susp[4][kSuspCap];
  __shared__ int s_susp_n[4];
  extern __shared__ __align__(16) double ringbuf[];  // per warp: 7*G*32 doubles (d), or D*2G*32 (x, y)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *susp = s_susp[warp];
  int *susp_n = &s_susp_n[warp];
  constexpr int RING = H == 0 ? 7 * G * 32 : (2 * H + 1 > 7 ? 2 * H + 1 : 7) * 2 * G * 32;
  double *myring = ringbuf + (size_t)warp * RING + lane;
  // behind the rings: per warp, the compaction map and the mask flags of the current pair
  const int NCOL = a.ncol_pad;
  const size_t STRIDE = (size_t)kItemHdr + 2 * (size_t)NCOL;
  short *kmap = reinterpret_cast<short *>(ringbuf + (size_t)4 * RING) + (size_t)warp * NCOL;
  unsigned char *mflag = reinterpret_cast<unsigned char *>(reinterpret_cast<short *>(ringbuf + (size_t)4 * RING) +
                                                          (size_t)4 * NCOL) + (size_t)warp * NCOL;
  const bool with_sn = H == 0 ? WITH_SN : a.p.compute_sn != 0;
  const chess_params &p = a.p;
  const double qnan = CUDART_NAN;
  const bool use_masks = p.mappability_cutoff > 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { *a.counter_next = 0ull; *a.retry_next = 0; }
  const long long n_items = a.n_pairs * kBands;

  while (true) {
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(a.counter_cur, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const long long pi = item / kBands;
    const int band = (int)(item - pi * kBands);
    const chess_pair d = a.pairs[pi];
    int status = CHESS_PAIR_OK;
    if (d.ref_n <= 0 || d.qry_n <= 0) status = CHESS_PAIR_NOT_FOUND;
    else if (d.qry_n < p.min_bins || d.ref_n < p.min_bins) status = CHESS_PAIR_TOO_SMALL;
    else if (d.ref_n != d.qry_n || d.ref_n > (MULTI ? NCOL : 31 * G)) status = CHESS_PAIR_RETRY;
    if (status != CHESS_PAIR_OK) {
      if (band == 0 && lane == 0) {
        a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = status;
        if (status == CHESS_PAIR_RETRY) atomicOr(a.retry_cur, 1);
      }
      continue;
    }
    const int n = d.ref_n, flip = d.flip;
    const double *R = a.ref_band + d.ref_off;
    const double *Q = a.qry_band + d.qry_off;
    const size_t rp = (size_t)d.ref_pitch, qp = (size_t)d.qry_pitch;
    // first bins of the two windows, recovered from the band view (off = i0*(2W-1) + W-1, pitch = 2W-2)
    const long long Wr = d.ref_pitch / 2 + 1, Wq = d.qry_pitch / 2 + 1;
    const int i0r = (int)((d.ref_off - (Wr - 1)) / (2 * Wr - 1));
    const int i0q = (int)((d.qry_off - (Wq - 1)) / (2 * Wq - 1));

    // ---- predicted masks -> cutoff test -> compaction map (every band of a pair repeats this; O(n))
    int np_ = n, tot_r = 0, tot_q = 0, tot_one_sided = 0;
    if (use_masks) {
      for (int base = 0; base < n; base += 32) {
        const int c = base + lane;
        bool fr = false, fq = false;
        if (c < n) {
          const int jr = i0r + c, jq = i0q + (flip ? n - 1 - c : c);
          fr = __ldg(a.ref_lo + jr) < i0r && __ldg(a.ref_hi + jr) >= i0r + n;
          fq = __ldg(a.qry_lo + jq) < i0q && __ldg(a.qry_hi + jq) >= i0q + n;
          mflag[c] = (unsigned char)((fr ? 1 : 0) | (fq ? 2 : 0));
        }
        tot_r += __popc(__ballot_sync(0xffffffffu, fr));
        tot_q += __popc(__ballot_sync(0xffffffffu, fq));
        tot_one_sided += __popc(__ballot_sync(0xffffffffu, fr != fq));
      }
      if ((double)tot_r / (double)n > p.mappability_cutoff || (double)tot_q / (double)n > p.mappability_cutoff) {
        if (band == 0 && lane == 0) { a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_UNMAPPABLE; }
        continue;  // NB: a coincidentally masked column could only add to the count; see verify below
      }
    }
    __syncwarp();
    if (use_masks && !p.keep_unmappable && (tot_r + tot_q) > 0) {
      int carry = 0;
      for (int base = 0; base < n; base += 32) {
        const int c = base + lane;
        const bool keep = c < n && mflag[c] == 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) kmap[carry + __popc(bal & ((1u << lane) - 1u))] = (short)c;
        carry += __popc(bal);
      }
      np_ = carry;
    } else {
      for (int c = lane; c < n; c += 32) kmap[c] = (short)c;
    }
    if (lane == 0) *susp_n = 0;
    __syncwarp();

    int win = (np_ & 1) ? np_ : np_ - 1;
    if (win < 3) win = 3;
    if (H > 0) win = 2 * H + 1;
    if (win > np_) {
      if (band == 0 && lane == 0) { a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_WINDOW_EXCEEDS; }
      continue;
    }
    const int m = np_ - win + 1;

    // ---- this band
    const int B2 = (np_ + kBands - 1) / kBands;
    const int lo = band * B2, hi = min(np_, lo + B2);
    double *rec = a.scratch + ((size_t)pi * kBands + band) * STRIDE;
    // data range: n lookups in the running-extrema arrays give max / min of the uncompacted pair.
    // They are also the extrema of the compacted pair if every removed cell holds the default value
    // in both matrices (bins masked in one matrix only are removed from the other, where they carry
    // data) and the default is neither the max nor the min; otherwise min / max are taken in a loop
    bool inloop_minmax = true;
    double xmax = -CUDART_INF, xmin = CUDART_INF;
    if (a.ref_pmax != nullptr) {
      for (int r = lane; r < n; r += 32) {
        const size_t ir = (size_t)(i0r + r) * (size_t)Wr + (size_t)(n - 1 - r);
        const size_t iq = (size_t)(i0q + r) * (size_t)Wq + (size_t)(n - 1 - r);
        xmax = fmax(xmax, fmax(__ldg(a.ref_pmax + ir), __ldg(a.qry_pmax + iq)));
        xmin = fmin(xmin, fmin(__ldg(a.ref_pmin + ir), __ldg(a.qry_pmin + iq)));
      }
      xmax = warp_max(xmax);
      xmin = warp_min(xmin);
      inloop_minmax = np_ < n && (tot_one_sided > 0 || !(xmax > p.default_value && xmin < p.default_value));
    }
    // column blocks: one if the compacted matrix fits the warp (np <= 31*G), else blocks that emit
    // 31*G - 6 columns each and carry three halo columns either side
    constexpr int CAP = 31 * G;
    const int SPAN = (!MULTI || np_ <= CAP) ? np_ : CAP - 6;
    const int nblk = (!MULTI || np_ <= CAP) ? 1 : (np_ + SPAN - 1) / SPAN;
    Band2 acc;
    acc.T0 = acc.T1 = acc.T2 = acc.T3 = acc.T4 = 0.0;
    acc.vmax = -CUDART_INF; acc.vmin = CUDART_INF; acc.sn_sum = 0.0; acc.sn_cnt = 0;
    double rt[5] = {0, 0, 0, 0, 0};
    double ssim_sum = 0.0;
    if (H > 0 && inloop_minmax && lo < hi) {
      // the range must be complete before the first window: every band scans the whole pair
      double bmax = -CUDART_INF, bmin = CUDART_INF;
      for (int blk = 0; blk < nblk; ++blk) {
        LaneCols<G, MULTI> lc;
        lane_cols<G, MULTI>(lc, kmap, n, np_, flip, lane, blk * SPAN, blk * SPAN, min(np_, (blk + 1) * SPAN));
        Band2 mm;
        moments_rows<G, false, true>(R, Q, rp, qp, kmap, n, flip, 0, np_, lc, mm, nullptr, nullptr);
        bmax = fmax(bmax, mm.vmax); bmin = fmin(bmin, mm.vmin);
      }
      xmax = warp_max(bmax); xmin = warp_min(bmin);
    } else if (inloop_minmax) {
      xmax = -CUDART_INF; xmin = CUDART_INF;
    }
    for (int blk = 0; blk < nblk; ++blk) {
      const int estart = blk * SPAN, eend = min(np_, estart + SPAN);
      const int cstart = nblk == 1 ? 0 : estart - 3;
      LaneCols<G, MULTI> lc;
      lane_cols<G, MULTI>(lc, kmap, n, np_, flip, lane, cstart, estart, eend);
      Band2 part;
      part.T0 = part.T1 = part.T2 = part.T3 = part.T4 = 0.0;
      part.vmax = -CUDART_INF; part.vmin = CUDART_INF; part.sn_sum = 0.0; part.sn_cnt = 0;
      double cs_r[G], cs_q[G];
#pragma unroll
      for (int g = 0; g < G; ++g) { cs_r[g] = 0.0; cs_q[g] = 0.0; }
      if (H == 0 && m == 2 && (band == 0 || band == kBands - 1)) {  // first / last row totals (2x2 SSIM map)
        const int kr = kmap[band == 0 ? 0 : np_ - 1];
        const double *rowR = R + (size_t)kr * rp;
        const double *rowQ = Q + (size_t)(flip ? n - 1 - kr : kr) * qp;
        double t5[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (lc.em(g)) {
            const double x = __ldg(rowR + lc.xo[g]), y = __ldg(rowQ + lc.yo[g]);
            t5[0] += x; t5[1] += y; t5[2] = fma(x, x, t5[2]); t5[3] = fma(y, y, t5[3]); t5[4] = fma(x, y, t5[4]);
          }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) rt[k] += warp_sum(t5[k]);
      }
      if (use_masks && band == 0 && np_ < n) {
        // the column sums must cover all n rows of the uncompacted matrices: band 0 adds the few
        // removed rows (a row removed because of the *other* matrix is not all-default in this one)
        for (int r = 0; r < n; ++r) {
          if (mflag[r] != 0) {
            const double *rowR = R + (size_t)r * rp;
            const double *rowQ = Q + (size_t)(flip ? n - 1 - r : r) * qp;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (lc.em(g)) { cs_r[g] += __ldg(rowR + lc.xo[g]); cs_q[g] += __ldg(rowQ + lc.yo[g]); }
            }
          }
        }
      }
      if (lo < hi) {
        if (lane == 0) *susp_n = 0;
        __syncwarp();
        double ssum_blk = 0.0;
        if (H > 0) {
          const double drange = xmax - xmin;
          const double C1 = (0.01 * drange) * (0.01 * drange), C2 = (0.03 * drange) * (0.03 * drange);
          if (H == 3)
            win7_rows<G>(R, Q, rp, qp, kmap, n, np_, flip, lo, hi, lane, lc, myring, susp, susp_n, with_sn, C1,
                         C2, part, ssum_blk, cs_r, cs_q);
          else
            win_rows<G, (H == 1 || H == 2 ? H : 1)>(R, Q, rp, qp, kmap, n, np_, flip, lo, hi, lane, lc, myring,
                                                    susp, susp_n, with_sn, C1, C2, part, ssum_blk, cs_r, cs_q);
          ssim_sum += ssum_blk;
        } else if (WITH_SN) {
          if (inloop_minmax) {
            Band2 mm;
            moments_rows<G, false, true>(R, Q, rp, qp, kmap, n, flip, lo, hi, lc, mm, nullptr, nullptr);
            xmax = fmax(xmax, mm.vmax); xmin = fmin(xmin, mm.vmin);
          }
          sn_rows<G, true>(R, Q, rp, qp, kmap, n, np_, flip, lo, hi, lane, lc, myring, susp, susp_n, part, cs_r,
                           cs_q);
        } else {
          Band2 mm;
          if (inloop_minmax) {
            moments_rows<G, true, true>(R, Q, rp, qp, kmap, n, flip, lo, hi, lc, mm, cs_r, cs_q);
            xmax = fmax(xmax, mm.vmax); xmin = fmin(xmin, mm.vmin);
          } else {
            moments_rows<G, true, false>(R, Q, rp, qp, kmap, n, flip, lo, hi, lc, mm, cs_r, cs_q);
          }
          part.T0 = mm.T0; part.T1 = mm.T1; part.T2 = mm.T2; part.T3 = mm.T3; part.T4 = mm.T4;
        }
        if (with_sn) {
          __syncwarp();
          const int ns = *susp_n;
          if (ns > 0 && ns <= kSuspCap) {  // cold: cancelled one-pass variances -> numpy's two passes
            // the queue was filled in arrival order: sort the (at most 24) codes first so that the
            // lane an entry lands on, and with it the order of the additions, is reproducible
            if (lane == 0) {
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
              if (gq == gq) { part.sn_sum += gq; part.sn_cnt += 1; }
            }
          } else if (ns > kSuspCap) {  // list overflow: redo this block of the band exactly
            part.sn_sum = 0.0; part.sn_cnt = 0;
            for (int rc = lo; rc < hi; ++rc) {
              for (int c = estart + lane; c < eend; c += 32) {
                const double cnt = (double)((min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1) *
                                            (min(c + 3, np_ - 1) - max(c - 3, 0) + 1));
                const double gq = sn_window_exact(R, Q, rp, qp, kmap, n, flip, rc, c, np_, cnt);
                if (gq == gq) { part.sn_sum += gq; part.sn_cnt += 1; }
              }
            }
          }
          __syncwarp();
        }
      }
      acc.T0 += part.T0; acc.T1 += part.T1; acc.T2 += part.T2; acc.T3 += part.T3; acc.T4 += part.T4;
      acc.sn_sum += part.sn_sum; acc.sn_cnt += part.sn_cnt;
      if (use_masks) {  // this band's share of the column sums of the block's own columns
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (lc.em(g)) {
            rec[kItemHdr + lc.cbase + g] = cs_r[g];
            rec[kItemHdr + NCOL + lc.cbase + g] = cs_q[g];
          }
        }
      }
    }
    acc.vmax = xmax; acc.vmin = xmin;
    // ---- partial record of this band
    {
      const double T0 = warp_sum(acc.T0), T1 = warp_sum(acc.T1), T2 = warp_sum(acc.T2);
      const double T3 = warp_sum(acc.T3), T4 = warp_sum(acc.T4);
      const double vmax = warp_max(acc.vmax), vmin = warp_min(acc.vmin);
      const double ss = warp_sum(acc.sn_sum);
      const int sc = warp_sum_int(acc.sn_cnt);
      if (lane == 0) {
        rec[0] = T0; rec[1] = T1; rec[2] = T2; rec[3] = T3; rec[4] = T4;
        rec[5] = vmax; rec[6] = vmin; rec[7] = ss; rec[8] = (double)sc;
#pragma unroll
        for (int k = 0; k < 5; ++k) rec[9 + k] = rt[k];
      }
      if (H > 0) {
        const double sw = warp_sum(ssim_sum);
        if (lane == 0) rec[14] = sw;
      }
    }
    __threadfence();
    __syncwarp();
    int arrived = 0;
    if (lane == 0) arrived = atomicAdd(a.done + pi, 1);
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived != kBands - 1) continue;

    // ---- last band to arrive: verify the prediction, combine in band order, write the result
    __threadfence();
    const double *recs = a.scratch + (size_t)pi * kBands * STRIDE;
    bool bad = false;
    if (use_masks) {
      const double target = (double)n * p.default_value;
      for (int c = lane; c < np_; c += 32) {
        double sr = 0.0, sq = 0.0;
#pragma unroll
        for (int b = 0; b < kBands; ++b) {
          sr += __ldcg(recs + (size_t)b * STRIDE + kItemHdr + c);
          sq += __ldcg(recs + (size_t)b * STRIDE + kItemHdr + NCOL + c);
        }
        const int kc = kmap[c];
        if (!(mflag[kc] & 1) && fabs(sr - target) <= 1e-7 * (fabs(sr) + fabs(target))) {
          double s2 = 0.0;  // the reference's own order: rows 0 .. n-1 of the uncompacted matrix
          for (int r = 0; r < n; ++r) s2 += __ldg(R + (size_t)r * rp + kc);
          bad |= s2 == target;
        }
        if (!(mflag[kc] & 2) && fabs(sq - target) <= 1e-7 * (fabs(sq) + fabs(target))) {
          double s2 = 0.0;
          const int cq = flip ? n - 1 - kc : kc;
          for (int r = 0; r < n; ++r) s2 += __ldg(Q + (size_t)(flip ? n - 1 - r : r) * qp + cq);
          bad |= s2 == target;
        }
      }
      bad = __any_sync(0xffffffffu, bad);
    }
    if (lane == 0) {
      if (bad) {  // a coincidentally masked column: the generic kernel redoes this pair exactly
        a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
        atomicOr(a.retry_cur, 1);
      } else {
        double t[5] = {0, 0, 0, 0, 0}, mx = -CUDART_INF, mn = CUDART_INF, ss = 0.0, sc = 0.0;
        for (int b = 0; b < kBands; ++b) {
          const double *rb = recs + (size_t)b * STRIDE;
#pragma unroll
          for (int k = 0; k < 5; ++k) t[k] += __ldcg(rb + k);
          mx = fmax(mx, __ldcg(rb + 5)); mn = fmin(mn, __ldcg(rb + 6));
          ss += __ldcg(rb + 7); sc += __ldcg(rb + 8);
        }
        double row0[5], rowL[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          row0[k] = __ldcg(recs + 9 + k);
          rowL[k] = __ldcg(recs + (size_t)(kBands - 1) * STRIDE + 9 + k);
        }
        if (H == 0) {
          finish_pair(t, mx, mn, ss, sc, row0, rowL, np_, win, m, kmap, R, Q, rp, qp, n, flip, with_sn,
                      a.ssim + pi, a.sn + pi, a.status + pi);
        } else {  // mean of the SSIM map: band sums in band order
          double sm = 0.0;
          for (int b = 0; b < kBands; ++b) sm += __ldcg(recs + (size_t)b * STRIDE + 14);
          a.ssim[pi] = sm / ((double)m * (double)m);
          a.sn[pi] = with_sn ? ss / sc : CUDART_NAN;
          a.status[pi] = CHESS_PAIR_OK;
        }
      }
      a.done[pi] = 0;  // leave the arrival counters clean for the next launch
    }
  }
}

template <int G, bool WITH_SN, int H, bool MULTI>
int launch_items_sn(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  auto kern = compare_r1_items_kernel<G, WITH_SN, H, MULTI>;
  constexpr int RING = H == 0 ? 7 * G * 32 : (2 * H + 1 > 7 ? 2 * H + 1 : 7) * 2 * G * 32;
  // rings (always reserved: the kernel addresses the maps behind them) + per-warp map and flags
  const size_t smem = (size_t)4 * RING * sizeof(double) + (size_t)4 * a.ncol_pad * 3;
  const size_t smem_max = (size_t)4 * RING * sizeof(double) + (size_t)4 * 1024 * 3;
  static bool configured = false;
  if (!configured) {
    CHESS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    configured = true;
  }
  static int per_sm = 0, per_sm_ncol = -1;  // the occupancy query costs microseconds of host time per call
  if (per_sm_ncol != a.ncol_pad) {
    CHESS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
    if (per_sm < 1) per_sm = 1;
    per_sm_ncol = a.ncol_pad;
  }
  long long grid = (long long)ctx->sm_count * per_sm;
  const long long need = (a.n_pairs * kBands + 3) / 4;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, 128, smem, stream>>>(a);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}


}  // namespace
