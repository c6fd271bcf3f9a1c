// Row loop of the -r 1 triangle kernels (compare_tri.cuh), second form.
//
// What changed against the first form, each point from the SASS / stall profile of the first
// (profiles/r02_hg38_r1_*: 613 warp instructions per row at 7 columns per lane, of which 210 FP64;
// issue slots 50 % busy, the rest mostly `wait`):
//
//  * Lane-uniform weights.  The three row bands are cut at multiples of the strip width G
//    (tri_split_aligned), and the bands of item B start one whole lane left of their diagonal block.
//    A cell of the band counts 0 (left halo lane), 1 (inside the band's diagonal block: the block is
//    symmetric, so its full square is simply summed once) or 2 (right of the block: its mirror image
//    is never visited) -- a constant of the LANE.  The moments and the SN sums are accumulated
//    unweighted and the lane's totals multiplied by the weight once, after the loop: no per-pixel
//    selects, no per-pixel weight multiplications (5 instead of 7 FP64 per pixel in the moments).
//  * Columns beyond the matrix are never loaded: their registers are zeroed once before the loop
//    and the loads are predicated on a loop-invariant count, instead of zero + predicated load per
//    row (28 CS2R per row).  Rows beyond the matrix (the three SN rows below the last band) are
//    handled by a separate short tail loop, so the main loop has no row test in its loads.
//  * Row base pointers are formed once per row; a load is one IMAD.WIDE + LDG.
//  * The SN quotient is accumulated under a predicate (no selects); membership tests of the
//    cancellation queue are one AND per row with a lane-constant mask.
#pragma once
#include <type_traits>
#include "compare_kernels.cuh"

namespace {

// Three row bands cut at multiples of G.  Band 0 = rows [0, B1) x all columns (item A); bands 1 and 2 =
// rows [B1, B2) x columns >= B1 - G and rows [B2, np) x columns >= B2 - G, side by side in one warp
// (item B): band 1 on lanes [0, L1), an idle lane, band 2 on lanes (L1, L1 + L2].  False (single band,
// item B empty) when the matrix is too small to be worth cutting.
__host__ __device__ __forceinline__ bool tri_split_aligned(int np_, int G, int &B1, int &B2, int &L1, int &L2) {
  B1 = np_; B2 = np_; L1 = 0; L2 = 0;
  const int L = (np_ + G - 1) / G;          // lanes that hold columns
  if (np_ < 24 || L < 4) return false;
  int k1 = (L + 1) / 3, k2 = (2 * L + 1) / 3;
  if (k1 < 1) k1 = 1;
  if (k2 <= k1) k2 = k1 + 1;
  if (k2 > L - 1) k2 = L - 1;
  if (k1 >= k2) k1 = k2 - 1;
  // lanes of item B: (L - k1 + 1) + 1 + (L - k2 + 1) <= 31 -- lane 31 stays idle, so that the neighbour exchange
  // of the last lane of band 2 wraps onto zeros, not onto band 1; move the cuts down the diagonal until it fits
  while (2 * L - k1 - k2 + 3 > 31) {
    if (k2 < L - 1 && (k2 - k1) <= k1) ++k2;
    else if (k1 < k2 - 1) ++k1;
    else if (k2 < L - 1) ++k2;
    else return false;
  }
  B1 = G * k1; B2 = G * k2; L1 = L - k1 + 1; L2 = L - k2 + 1;
  return true;
}

// A pointer as one 64-bit register value (an opaque move: keeps the two halves in an aligned register pair, which
// IMAD.WIDE needs for its 64-bit addend -- computed piecewise, ptxas may leave them in unrelated registers and then
// splits every address into a multiply and a 64-bit add).
__device__ __forceinline__ const double *as_reg_pair(const double *p) {
  unsigned long long v;
  asm("mov.b64 %0, %1;" : "=l"(v) : "l"(p));
  return reinterpret_cast<const double *>(v);
}

// One element of a matrix row: row base + 32-bit element offset as ONE IMAD.WIDE, then the load.  The loads
// are volatile asm so that they stay where they are written (at the top of the iteration that precedes their
// use); left to the compiler they sink towards their first use and the row they overwrite gets copied.
__device__ __forceinline__ double ld_row(const double *row, unsigned off) {
  unsigned long long addr;
  asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(addr) : "r"(off), "l"(row));
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(addr));
  return v;
}
// The same with a byte step of +8 or -8 per column: the query row of a flipped pair is read backwards from its
// last element, so one offset array serves both matrices.
__device__ __forceinline__ double ld_row_step(const double *row, unsigned off, int step) {
  unsigned long long addr;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(addr) : "r"((int)off), "r"(step), "l"(row));
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(addr));
  return v;
}

// Scheduling fence for values: an empty volatile asm that names them as inputs.  Volatile asms keep their order,
// so every volatile load written after the fence is issued after the named values exist -- i.e. after everything
// that still reads the registers those loads overwrite.  Without it the compiler sinks such computations towards
// their uses, below the loads, and has to copy the whole row first (28 moves per row at G = 7).  No instruction.
template <int G>
__device__ __forceinline__ void pin_before_loads(const double (&v)[G]) {
#pragma unroll
  for (int g = 0; g < G; ++g) asm volatile("" ::"d"(v[g]));
}
__device__ __forceinline__ void pin_before_loads(double a, double b, double c, double d, double e) {
  asm volatile("" ::"d"(a), "d"(b), "d"(c), "d"(d), "d"(e));
}
// ... and stores written before it (the ring slots that take the current row) stay before it, too
__device__ __forceinline__ void pin_stores_before_loads() { asm volatile("" ::: "memory"); }
template <int G>
__device__ __forceinline__ void pin_before_loads(const unsigned (&v)[G]) {
#pragma unroll
  for (int g = 0; g < G; ++g) asm volatile("" ::"r"(v[g]));
}

// How a pair is walked: 0 = one band (too small to cut), 1 = all three bands side by side in ONE pass of one warp
// (small matrices: band 0 on lanes [0, L), an idle lane, band 1, an idle lane, band 2 -- half as many warp-rows as
// two parts, no records to combine), 2 = two parts (item A = band 0, item B = bands 1 and 2).
__host__ __device__ __forceinline__ int tri_mode(int np_, int G, int &B1, int &B2, int &L1, int &L2) {
  if (!tri_split_aligned(np_, G, B1, B2, L1, L2)) return 0;
  const int L = (np_ + G - 1) / G;
  // strips of up to 4 columns only (16 warps per SM): measured at 81 bins, one pass of 7-column strips at 8 warps
  // per SM loses to two parts of 3-column strips (3.6 against 2.9 ms per 100 k pairs); at 41 bins one pass of
  // 4-column strips wins (1.59 against 1.76 ms)
  return (G <= 4 && L + 1 + L1 + 1 + L2 <= 31) ? 1 : 2;
}

// Geometry of one lane in one part of a pair.  (tri_lane / blk_lane compile for the host too:
// tests/tri_geometry_harness.cu checks on the CPU that the lanes of every layout tile the triangle exactly.)
struct TriLane {
  int lo, hi;      // rows of the lane's band
  int cbase;       // first column of the lane's strip (1 << 20: idle lane)
  int first;       // first row the lane reads (lo - 3, clamped)
  int cend;        // columns are read up to here (the matrix edge, or the end of a block's right halo lane)
  double wl;       // weight of the lane's columns: 0 halo, 1 diagonal block, 2 right of it
};

__host__ __device__ __forceinline__ TriLane tri_lane(int part, int lane, int G, int np_, int B1, int B2, int L1, int L2,
                                            int halo) {
  TriLane g;
  g.lo = 0; g.hi = 0; g.cbase = 1 << 20;
  if (part == 0) {
    g.lo = 0; g.hi = B1; g.cbase = G * lane;
  } else {
    // part 1: bands 1 and 2 from lane 0; part 2 (all bands in one pass): behind band 0 and an idle lane
    const int base = part == 2 ? (np_ + G - 1) / G + 1 : 0;
    const int l = lane - base;
    if (part == 2 && lane < base - 1) {
      g.lo = 0; g.hi = B1; g.cbase = G * lane;
    } else if (l >= 0 && l < L1) {
      g.lo = B1; g.hi = B2; g.cbase = B1 - G + G * l;
    } else if (l > L1 && l <= L1 + L2) {
      g.lo = B2; g.hi = np_; g.cbase = B2 - G + G * (l - L1 - 1);
    }
  }
  g.first = g.lo - halo < 0 ? 0 : g.lo - halo;
  g.cend = np_;
  g.wl = g.cbase < g.lo ? 0.0 : (g.cbase < g.hi ? 1.0 : 2.0);
  return g;
}

// Matrices wider than one warp: nb x nb square blocks of `side` rows (a multiple of G, at most 29 G), the blocks
// on and above the diagonal are the work items.  Block (bi, bj): rows [bi side, ...), columns [bj side, ...) plus
// one whole lane of halo columns on either side that exists; the block's lanes weigh 1 on the diagonal (its full
// square is summed once) and 2 above it (it stands for its mirror image), halo lanes 0.
__host__ __device__ __forceinline__ TriLane blk_lane(int bi, int bj, int side, int lane, int G, int np_, int halo) {
  TriLane g;
  g.lo = min(bi * side, np_); g.hi = min(g.lo + side, np_);
  const int clo = min(bj * side, np_), chi = min(clo + side, np_);
  g.cbase = (clo == 0 ? 0 : clo - G) + G * lane;
  g.cend = min(np_, chi + G);
  if (g.cbase >= g.cend) g.cbase = 1 << 20;          // idle lane
  g.first = g.lo - halo < 0 ? 0 : g.lo - halo;
  g.wl = (g.cbase >= clo && g.cbase < chi) ? (bi == bj ? 1.0 : 2.0) : 0.0;
  return g;
}

// Per-lane results of one part, already multiplied by the lane weight.
struct TriPartSums {
  double T[5];       // sum x, sum y, sum x^2, sum y^2, sum xy over the lane's share of the triangle
  double sn_sum;     // sum of |mean| / var over the lane's SN pixels
  double sn_cnt;     // number of those pixels (weighted)
  double vmax, vmin; // MINMAX instance only
};

// One part (item A, or the two bands of item B) of a pair: moments of the weighted triangle and,
// WITH_SN, the 7x7 difference-window sums.  `rows_max` = longest band of the part.  Windows whose
// one-pass variance shows cancellation are queued in susp / susp_n (codes (row << 16) | column) for the
// caller's exact recompute; the caller adds them with weight (column >= end of the row's block ? 2 : 1).
template <int G, bool WITH_SN>
__device__ __forceinline__ void tri_r1_part(const double *__restrict__ R, const double *__restrict__ Q, unsigned rp,
                                            unsigned qp, int n, int np_, int flip, const short *kmap,
                                            double *myring, const TriLane &ln, bool top, int rows_max,
                                            bool minmax, int *susp, int *susp_n, TriPartSums &out) {
  // `top`: the rows start at row 0 of the matrix (no halo rows above them)
  constexpr int HALO = WITH_SN ? 3 : 0;
  const int lo = ln.lo, hi = ln.hi, cbase = ln.cbase;
  const int t_emit0 = top ? HALO : 2 * HALO;           // first iteration with a complete window
  const int T = rows_max + (top ? HALO : 2 * HALO);
  // number of columns of the strip that exist; the others keep x = y = 0 for the whole loop
  int nv = ln.cend - cbase;
  nv = nv < 0 ? 0 : (nv > G ? G : nv);
  unsigned xo[G];
  double ccnt[G];
  const int ystep = flip ? -8 : 8;       // query column kc of a flipped pair is n - 1 - kc: read the row backwards
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = cbase + g;
    const int kc = g < nv ? (int)kmap[c] : 0;
    xo[g] = (unsigned)kc;
    // a column that does not exist gets count 0: its one-pass variance is -w1^2 <= 0, never "good"
    ccnt[g] = g < nv ? (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1) : 0.0;
  }
  // columns that may emit SN pixels / be queued: the lane's real columns unless it is a halo lane
  const unsigned ownmask = ln.wl > 0.0 ? ((1u << nv) - 1u) : 0u;

  double T0 = 0, T1 = 0, T2 = 0, T3 = 0, T4 = 0;
  double sn_sum = 0.0;
  int sn_cnt = 0;
  double vmax = -CUDART_INF, vmin = CUDART_INF;
  double Sd[G], Se[G];
  unsigned hist[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    Sd[g] = 0.0; Se[g] = 0.0; hist[g] = 0u;
    if (WITH_SN) {
#pragma unroll
      for (int k = 0; k < 7; ++k) myring[(k * G + g) * 32] = 0.0;
    }
  }
  const int up_lane = (threadIdx.x + 31) & 31, dn_lane = (threadIdx.x + 1) & 31;
  double *ringrow = myring;
  int slot = 0;
  double xn[G], yn[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { xn[g] = 0.0; yn[g] = 0.0; }

  // main loop: every lane's NEXT row exists, the prefetch is unconditional.  tail (a handful of iterations):
  // rows past the last one read as zeros (the SN windows of the last three rows of the matrix).
  int t_tail = nv > 0 ? np_ - ln.first : (1 << 20);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t_tail = min(t_tail, __shfl_xor_sync(0xffffffffu, t_tail, o));
  const int t_main = min(T, t_tail) - 1;   // iterations whose prefetch is unconditional

  auto fetch = [&](int row, auto tail_tag) {
    constexpr bool TAIL = decltype(tail_tag)::value;
    const bool vr = !TAIL || row < np_;
    const unsigned kr = (unsigned)kmap[TAIL ? (vr ? row : 0) : row];
    const double *rowR = as_reg_pair(R + (size_t)(kr * rp));
    const double *rowQ = as_reg_pair(Q + (size_t)((flip ? (unsigned)(n - 1) - kr : kr) * qp) + (flip ? n - 1 : 0));
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (g < nv) {
        if (!TAIL || vr) {
          xn[g] = ld_row(rowR, xo[g]);
          yn[g] = ld_row_step(rowQ, xo[g], ystep);
        } else {
          xn[g] = 0.0; yn[g] = 0.0;
        }
      }
    }
  };

  int row = ln.first;
  if (T > 0) fetch(row, std::false_type());
  for (int t = 0; t < T; ++t, ++row) {
    double dv[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      dv[g] = xn[g] - yn[g];
      if (WITH_SN) hist[g] = ((hist[g] << 1) & 0x7fu) | (xn[g] != yn[g] ? 1u : 0u);
    }
    if (row >= lo && row < hi) {  // the band's own rows: moments, unweighted (columns that do not exist hold 0)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        T0 += xn[g]; T1 += yn[g];
        T2 = fma(xn[g], xn[g], T2); T3 = fma(yn[g], yn[g], T3); T4 = fma(xn[g], yn[g], T4);
      }
      if (minmax) {  // the data range could not be looked up: every loaded element of an own row belongs to the pair
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const bool xgt = xn[g] > yn[g];
          const double hi2 = xgt ? xn[g] : yn[g];
          const double lo2 = xgt ? yn[g] : xn[g];
          vmax = ((g < nv) & (hi2 > vmax)) ? hi2 : vmax;
          vmin = ((g < nv) & (lo2 < vmin)) ? lo2 : vmin;
        }
      }
    }
    // the row is used up: the next one travels while the SN part below runs
    pin_before_loads<G>(dv);
    if (WITH_SN) pin_before_loads<G>(hist);
    pin_before_loads(T0, T1, T2, T3, T4);
    pin_before_loads(vmax, vmin, T0, T1, T2);
    if (t < t_main) fetch(row + 1, std::false_type());
    else if (t + 1 < T) fetch(row + 1, std::true_type());
    if (WITH_SN) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        double *cell = ringrow + g * 32;
        const double dold = *cell;
        *cell = dv[g];
        Sd[g] += dv[g]; Sd[g] -= dold;
        Se[g] = fma(dv[g], dv[g], Se[g]); Se[g] = fma(-dold, dold, Se[g]);
        Se[g] = hist[g] == 0u ? 0.0 : Se[g];     // seven equal rows: exactly 0, so that 0 / 0 is skipped like NaN
      }
      if (slot == 6) { slot = 0; ringrow = myring; } else { slot += 1; ringrow += G * 32; }
      if (slot == 0) {
        // the ring has just wrapped: rebuild the column sums of d*d from its seven rows.  A running sum keeps the
        // rounding of every value that ever passed through it; rebuilt once per window height it is exact
        // to a few ulps of the current window, which the quiet windows after a loud stretch need
#pragma unroll
        for (int g = 0; g < G; ++g) {
          double s2 = 0.0;
#pragma unroll
          for (int k = 0; k < 7; ++k) { const double v = myring[(k * G + g) * 32]; s2 = fma(v, v, s2); }
          Se[g] = hist[g] == 0u ? 0.0 : s2;
        }
      }
      if (t >= t_emit0) {
        const int rc = row - 3;
        // prefix and suffix sums of the strip: a window that ends at the strip's right edge is a suffix, and no
        // window of 7 starts after column 0 without reaching that edge (G <= 8), so nothing is ever subtracted
        double P1[G + 1], P2[G + 1];
        P1[0] = 0.0; P2[0] = 0.0;
        P1[1] = Sd[0]; P2[1] = Se[0];
#pragma unroll
        for (int g = 1; g < G; ++g) { P1[g + 1] = P1[g] + Sd[g]; P2[g + 1] = P2[g] + Se[g]; }
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
        if (rc >= lo && rc < hi) {   // SN pixels of the band's own rows
          const double rcnt = (double)(min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1);
          unsigned bad = 0u;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int a0 = g - 3 < 0 ? 0 : g - 3;
            const int b0 = g + 3 > G - 1 ? G - 1 : g + 3;
            double w1 = a0 == 0 ? P1[b0 + 1] : X1[a0];
            double w2 = a0 == 0 ? P2[b0 + 1] : X2[a0];
            if (g - 3 < 0) { w1 += up1[2 - g]; w2 += up2[2 - g]; }
            if (g + 3 > G - 1) { w1 += dn1[g + 3 - G]; w2 += dn2[g + 3 - G]; }
            const double cnt = rcnt * ccnt[g];
            const double m2 = w2 * cnt;
            const double den = fma(-w1, w1, m2);
            const bool good = den > fma(m2, kVarGuard, 1e-290);
            const double gq = fabs(w1) * cnt * fast_rcp(den);
            if (good) { sn_sum += gq; sn_cnt += 1; }
            if (!good && w2 != 0.0) bad |= 1u << g;   // cancellation (or a column that is not the lane's)
          }
          bad &= ownmask;
          if (bad) {  // rare: exact recompute after the loop
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (bad & (1u << g)) {
                const int e = atomicAdd(susp_n, 1);
                if (e < kSuspCap) susp[e] = (rc << 16) | (cbase + g);
              }
            }
          }
        }
      }
    }
  }
  out.T[0] = ln.wl * T0; out.T[1] = ln.wl * T1; out.T[2] = ln.wl * T2; out.T[3] = ln.wl * T3; out.T[4] = ln.wl * T4;
  out.sn_sum = ln.wl * sn_sum;
  out.sn_cnt = ln.wl * (double)sn_cnt;
  out.vmax = vmax; out.vmin = vmin;
}


// ---------------------------------------------------------------------------------------------------
// -a 7: window 7 is also the SN window, so one set of running 7-row column sums serves both: of x, y,
// q = x^2 + y^2 and xy (the SSIM formula needs the two variances only as a sum); sum d = Wx - Wy,
// sum d^2 = Wq - 2 Wxy.  Same lane geometry, buffers and loads as tri_r1_part.
//
// The SSIM value is formed on window SUMS: with c1 = 49^2 C1, c2 = 49^2 C2, cn = 49/48,
//     p = Wx Wy,  s = Wx^2 + Wy^2,  t = 49 Wxy - p,  u = 49 Wq - s,
//     S = (2 p + c1) (2 cn t + c2) / ((s + c1) (cn u + c2))
// which is (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)) with every factor scaled by 49^2:
// 13 FP64 operations and one reciprocal instead of 20.
template <int G>
__device__ __forceinline__ void tri_w7_part(const double *__restrict__ R, const double *__restrict__ Q, unsigned rp,
                                            unsigned qp, int n, int np_, int flip, const short *kmap,
                                            double *myring, const TriLane &ln, bool top, int rows_max,
                                            bool with_sn, double c1, double c2, int *susp, int *susp_n,
                                            double &ss_out, double &sn_out, double &cnt_out) {
  const int lo = ln.lo, hi = ln.hi, cbase = ln.cbase;
  const int t_emit0 = top ? 3 : 6;           // first iteration with a complete window
  const int T = rows_max + (top ? 3 : 6);
  int nv = ln.cend - cbase;
  nv = nv < 0 ? 0 : (nv > G ? G : nv);
  unsigned xo[G];
  double ccnt[G];
  const int ystep = flip ? -8 : 8;
  unsigned ssmask = 0u;                            // columns whose 7-column window lies inside the matrix
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = cbase + g;
    const int kc = g < nv ? (int)kmap[c] : 0;
    xo[g] = (unsigned)kc;
    ccnt[g] = g < nv ? (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1) : 0.0;
    if (g < nv && c >= 3 && c <= np_ - 4) ssmask |= 1u << g;
  }
  const unsigned ownmask = ln.wl > 0.0 ? ((1u << nv) - 1u) : 0u;
  if (!(ln.wl > 0.0)) ssmask = 0u;

  double ss_sum = 0.0, sn_sum = 0.0;
  int sn_cnt = 0;
  double Vx[G], Vy[G], Vq[G], Vxy[G];
  unsigned hist[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    Vx[g] = Vy[g] = Vq[g] = Vxy[g] = 0.0;
    hist[g] = 0u;
#pragma unroll
    for (int k = 0; k < 7; ++k) { myring[((k * G + g) * 2) * 32] = 0.0; myring[((k * G + g) * 2 + 1) * 32] = 0.0; }
  }
  const int up_lane = (threadIdx.x + 31) & 31, dn_lane = (threadIdx.x + 1) & 31;
  double *ringrow = myring;
  int slot = 0;
  double x[G], y[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { x[g] = 0.0; y[g] = 0.0; }
  int t_tail = nv > 0 ? np_ - ln.first : (1 << 20);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t_tail = min(t_tail, __shfl_xor_sync(0xffffffffu, t_tail, o));
  const int t_main = min(T, t_tail) - 1;
  const double cn = 49.0 / 48.0;

  auto fetch = [&](int row, auto tail_tag) {
    constexpr bool TAIL = decltype(tail_tag)::value;
    const bool vr = !TAIL || row < np_;
    const unsigned kr = (unsigned)kmap[TAIL ? (vr ? row : 0) : row];
    const double *rowR = as_reg_pair(R + (size_t)(kr * rp));
    const double *rowQ = as_reg_pair(Q + (size_t)((flip ? (unsigned)(n - 1) - kr : kr) * qp) + (flip ? n - 1 : 0));
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (g < nv) {
        if (!TAIL || vr) { x[g] = ld_row(rowR, xo[g]); y[g] = ld_row_step(rowQ, xo[g], ystep); }
        else { x[g] = 0.0; y[g] = 0.0; }
      }
    }
  };

  int row = ln.first;
  if (T > 0) fetch(row, std::false_type());
  for (int t = 0; t < T; ++t, ++row) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      double *cell = ringrow + (g * 2) * 32;   // the slot being overwritten holds row - 7
      const double xold = cell[0], yold = cell[32];
      cell[0] = x[g];
      cell[32] = y[g];
      Vx[g] += x[g]; Vx[g] -= xold;
      Vy[g] += y[g]; Vy[g] -= yold;
      Vq[g] = fma(x[g], x[g], fma(y[g], y[g], Vq[g]));
      Vq[g] = fma(-xold, xold, fma(-yold, yold, Vq[g]));
      Vxy[g] = fma(x[g], y[g], Vxy[g]); Vxy[g] = fma(-xold, yold, Vxy[g]);
      hist[g] = ((hist[g] << 1) & 0x7fu) | (x[g] != y[g] ? 1u : 0u);
    }
    // the row is used up (ring, column sums, history): the next one travels while the rest of the iteration runs
    pin_before_loads<G>(Vx); pin_before_loads<G>(Vy); pin_before_loads<G>(Vq); pin_before_loads<G>(Vxy);
    pin_before_loads<G>(hist);
    pin_stores_before_loads();
    if (t < t_main) fetch(row + 1, std::false_type());
    else if (t + 1 < T) fetch(row + 1, std::true_type());
#pragma unroll
    for (int g = 0; g < G; ++g) {
      // seven equal rows: sum d and sum d^2 of the column are exactly 0 (doubling commutes with every rounding of
      // the prefix / suffix arithmetic below, so Wq - 2 Wxy cancels exactly)
      if (hist[g] == 0u) { Vy[g] = Vx[g]; Vq[g] = 2.0 * Vxy[g]; }
    }
    if (slot == 6) { slot = 0; ringrow = myring; } else { slot += 1; ringrow += 2 * G * 32; }
    if (slot == 0) {
      // the ring has just wrapped: rebuild the quadratic column sums from its seven rows (a running sum keeps the
      // rounding of every value that ever passed through it; sum d^2 = Wq - 2 Wxy magnifies whatever they carry)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        double vq = 0.0, vxy = 0.0;
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
      double W4[4][G];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double *V = q == 0 ? Vx : q == 1 ? Vy : q == 2 ? Vq : Vxy;
        double P[G + 1], X[G];
        P[0] = 0.0; P[1] = V[0];
#pragma unroll
        for (int g = 1; g < G; ++g) P[g + 1] = P[g] + V[g];
        X[G - 1] = V[G - 1];
#pragma unroll
        for (int g = G - 2; g >= 0; --g) X[g] = X[g + 1] + V[g];
        double up[3], dn[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          up[j] = __shfl_sync(0xffffffffu, X[G - 1 - j < 0 ? 0 : G - 1 - j], up_lane);
          dn[j] = __shfl_sync(0xffffffffu, P[j + 1], dn_lane);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int a0 = g - 3 < 0 ? 0 : g - 3;
          const int b0 = g + 3 > G - 1 ? G - 1 : g + 3;
          double w = a0 == 0 ? P[b0 + 1] : X[a0];
          if (g - 3 < 0) w += up[2 - g];
          if (g + 3 > G - 1) w += dn[g + 3 - G];
          W4[q][g] = w;
        }
      }
      if (rc >= lo && rc < hi) {
        if (rc >= 3 && rc <= np_ - 4) {   // SSIM windows need three rows either side
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const double wx = W4[0][g], wy = W4[1][g];
            const double p = wx * wy;
            const double s2 = fma(wx, wx, wy * wy);
            const double a1 = fma(2.0, p, c1), b1 = s2 + c1;
            const double tt = fma(49.0, W4[3][g], -p);
            const double a2 = fma(2.0 * cn, tt, c2);
            const double u = fma(49.0, W4[2][g], -s2);
            const double b2 = fma(cn, u, c2);
            const double num = a1 * a2, den = b1 * b2;
            const double S = num * fast_rcp(den);   // den > 0: a zero data range never gets here (caller)
            if (ssmask & (1u << g)) ss_sum += S;
          }
        }
        if (with_sn) {
          const double rcnt = (double)(min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1);
          unsigned bad = 0u;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const double w1 = W4[0][g] - W4[1][g];
            const double w2 = fma(-2.0, W4[3][g], W4[2][g]);
            const double cnt = rcnt * ccnt[g];
            const double m2 = w2 * cnt;
            const double den = fma(-w1, w1, m2);
            // trusted: one-pass variance well conditioned AND the q - 2xy subtraction kept >= 10 digits
            const bool good = (den > fma(m2, kVarGuard, 1e-290)) & (w2 > 1e-6 * W4[2][g]);
            const double gq = fabs(w1) * cnt * fast_rcp(den);
            if (good) { sn_sum += gq; sn_cnt += 1; }
            if (!good && w2 != 0.0) bad |= 1u << g;
          }
          bad &= ownmask;
          if (bad) {  // rare: exact recompute after the loop
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (bad & (1u << g)) {
                const int e = atomicAdd(susp_n, 1);
                if (e < kSuspCap) susp[e] = (rc << 16) | (cbase + g);
              }
            }
          }
        }
      }
    }
  }
  ss_out = ln.wl * ss_sum;
  sn_out = ln.wl * sn_sum;
  cnt_out = ln.wl * (double)sn_cnt;
}


// ---------------------------------------------------------------------------------------------------
// -a 3 / 5 / 9 / 11 (H = 1, 2, 4, 5): the SSIM window (2H+1 rows) and the SN window (7 rows) differ, so the
// column sums are kept separately -- Vx, Vy, Vq = sum x^2 + y^2, Vxy over the last 2H+1 rows, Sd, Se over the
// last 7 -- all fed from one ring of x and y that is max(2H+1, 7) rows deep.  Geometry, loads and fences as in
// tri_r1_part; the bands carry a halo of max(H, 3) rows, and one whole lane of columns left of their block.
template <int G, int H>
__device__ __forceinline__ void tri_win_part(const double *__restrict__ R, const double *__restrict__ Q, unsigned rp,
                                             unsigned qp, int n, int np_, int flip, const short *kmap,
                                             double *myring, const TriLane &ln, bool top, int rows_max,
                                             bool with_sn, double c1, double c2, int *susp, int *susp_n,
                                             double &ss_out, double &sn_out, double &cnt_out) {
  static_assert(H >= 1 && H <= 5 && H != 3 && G >= H, "window 3 / 5 / 9 / 11");
  constexpr int HALO = H > 3 ? H : 3;
  constexpr int WIN = 2 * H + 1;
  constexpr int D = WIN > 7 ? WIN : 7;          // ring depth in rows
  const int lo = ln.lo, hi = ln.hi, cbase = ln.cbase;
  const int t_own0 = top ? 0 : HALO;      // iteration of the band's first own row
  const int T = rows_max + (top ? HALO : 2 * HALO);
  int nv = ln.cend - cbase;
  nv = nv < 0 ? 0 : (nv > G ? G : nv);
  unsigned xo[G];
  double ccnt[G];
  const int ystep = flip ? -8 : 8;
  unsigned ssmask = 0u;                         // columns whose SSIM window lies inside the matrix
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = cbase + g;
    const int kc = g < nv ? (int)kmap[c] : 0;
    xo[g] = (unsigned)kc;
    ccnt[g] = g < nv ? (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1) : 0.0;
    if (g < nv && c >= H && c <= np_ - 1 - H) ssmask |= 1u << g;
  }
  const unsigned ownmask = ln.wl > 0.0 ? ((1u << nv) - 1u) : 0u;
  if (!(ln.wl > 0.0)) ssmask = 0u;

  double ss_sum = 0.0, sn_sum = 0.0;
  int sn_cnt = 0;
  double Sd[G], Se[G], Vx[G], Vy[G], Vq[G], Vxy[G];
  unsigned hist[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    Sd[g] = Se[g] = Vx[g] = Vy[g] = Vq[g] = Vxy[g] = 0.0;
    hist[g] = 0u;
#pragma unroll
    for (int k = 0; k < D; ++k) { myring[((k * G + g) * 2) * 32] = 0.0; myring[((k * G + g) * 2 + 1) * 32] = 0.0; }
  }
  const int up_lane = (threadIdx.x + 31) & 31, dn_lane = (threadIdx.x + 1) & 31;
  int slot = 0;
  double x[G], y[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { x[g] = 0.0; y[g] = 0.0; }
  int t_tail = nv > 0 ? np_ - ln.first : (1 << 20);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t_tail = min(t_tail, __shfl_xor_sync(0xffffffffu, t_tail, o));
  const int t_main = min(T, t_tail) - 1;
  constexpr double NP = (double)(WIN * WIN);
  constexpr double cn = NP / (NP - 1.0);

  auto fetch = [&](int row, auto tail_tag) {
    constexpr bool TAIL = decltype(tail_tag)::value;
    const bool vr = !TAIL || row < np_;
    const unsigned kr = (unsigned)kmap[TAIL ? (vr ? row : 0) : row];
    const double *rowR = as_reg_pair(R + (size_t)(kr * rp));
    const double *rowQ = as_reg_pair(Q + (size_t)((flip ? (unsigned)(n - 1) - kr : kr) * qp) + (flip ? n - 1 : 0));
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (g < nv) {
        if (!TAIL || vr) { x[g] = ld_row(rowR, xo[g]); y[g] = ld_row_step(rowQ, xo[g], ystep); }
        else { x[g] = 0.0; y[g] = 0.0; }
      }
    }
  };

  int row = ln.first;
  if (T > 0) fetch(row, std::false_type());
  for (int t = 0; t < T; ++t, ++row) {
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
      Vx[g] += x[g]; Vx[g] -= xow;
      Vy[g] += y[g]; Vy[g] -= yow;
      Vq[g] = fma(x[g], x[g], fma(y[g], y[g], Vq[g]));
      Vq[g] = fma(-xow, xow, fma(-yow, yow, Vq[g]));
      Vxy[g] = fma(x[g], y[g], Vxy[g]); Vxy[g] = fma(-xow, yow, Vxy[g]);
    }
    // the row is used up (ring, column sums, history): the next one travels while the rest of the iteration runs
    pin_before_loads<G>(Sd); pin_before_loads<G>(Se); pin_before_loads<G>(Vx); pin_before_loads<G>(Vy);
    pin_before_loads<G>(Vq); pin_before_loads<G>(Vxy); pin_before_loads<G>(hist);
    pin_stores_before_loads();
    if (t < t_main) fetch(row + 1, std::false_type());
    else if (t + 1 < T) fetch(row + 1, std::true_type());
#pragma unroll
    for (int g = 0; g < G; ++g) Se[g] = hist[g] == 0u ? 0.0 : Se[g];   // seven equal rows: exactly 0
    slot = slot == D - 1 ? 0 : slot + 1;
    if (slot == 0) {
      // the ring has just wrapped: rebuild the 7-row sums of d*d from its most recent seven rows (slots D-7 .. D-1)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        double s2 = 0.0;
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
      double W4[4][G];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double *V = q == 0 ? Vx : q == 1 ? Vy : q == 2 ? Vq : Vxy;
        double P[G + 1], X[G];
        P[0] = 0.0; P[1] = V[0];
#pragma unroll
        for (int g = 1; g < G; ++g) P[g + 1] = P[g] + V[g];
        X[G - 1] = V[G - 1];
#pragma unroll
        for (int g = G - 2; g >= 0; --g) X[g] = X[g + 1] + V[g];
        double up[H], dn[H];
#pragma unroll
        for (int j = 0; j < H; ++j) {
          up[j] = __shfl_sync(0xffffffffu, X[G - 1 - j < 0 ? 0 : G - 1 - j], up_lane);
          dn[j] = __shfl_sync(0xffffffffu, P[j + 1], dn_lane);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int a0 = g - H < 0 ? 0 : g - H;
          const int b0 = g + H > G - 1 ? G - 1 : g + H;
          double w;
          if (a0 == 0) w = P[b0 + 1];
          else if (b0 == G - 1) w = X[a0];
          else {            // the window lies inside the strip (H <= 2 only): its 2H+1 columns, added up
            w = V[a0];
#pragma unroll
            for (int j = a0 + 1; j <= b0; ++j) w += V[j];
          }
          if (g - H < 0) w += up[H - 1 - g];
          if (g + H > G - 1) w += dn[g + H - G];
          W4[q][g] = w;
        }
      }
      if (rcs >= lo && rcs < hi && rcs >= H && rcs <= np_ - 1 - H) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const double wx = W4[0][g], wy = W4[1][g];
          const double p = wx * wy;
          const double s2 = fma(wx, wx, wy * wy);
          const double a1 = fma(2.0, p, c1), b1 = s2 + c1;
          const double tt = fma(NP, W4[3][g], -p);
          const double a2 = fma(2.0 * cn, tt, c2);
          const double u = fma(NP, W4[2][g], -s2);
          const double b2 = fma(cn, u, c2);
          const double S = (a1 * a2) * fast_rcp(b1 * b2);   // b1 b2 > 0: a zero data range never gets here (caller)
          if (ssmask & (1u << g)) ss_sum += S;
        }
      }
    }

    // ---- SN pixels of row - 3
    if (with_sn && t >= t_own0 + 3) {
      const int rc = row - 3;
      double P1[G + 1], P2[G + 1];
      P1[0] = 0.0; P2[0] = 0.0;
      P1[1] = Sd[0]; P2[1] = Se[0];
#pragma unroll
      for (int g = 1; g < G; ++g) { P1[g + 1] = P1[g] + Sd[g]; P2[g + 1] = P2[g] + Se[g]; }
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
      if (rc >= lo && rc < hi) {
        const double rcnt = (double)(min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1);
        unsigned bad = 0u;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int a0 = g - 3 < 0 ? 0 : g - 3;
          const int b0 = g + 3 > G - 1 ? G - 1 : g + 3;
          double w1 = a0 == 0 ? P1[b0 + 1] : X1[a0];
          double w2 = a0 == 0 ? P2[b0 + 1] : X2[a0];
          if (g - 3 < 0) { w1 += up1[2 - g]; w2 += up2[2 - g]; }
          if (g + 3 > G - 1) { w1 += dn1[g + 3 - G]; w2 += dn2[g + 3 - G]; }
          const double cnt = rcnt * ccnt[g];
          const double m2 = w2 * cnt;
          const double den = fma(-w1, w1, m2);
          const bool good = den > fma(m2, kVarGuard, 1e-290);
          const double gq = fabs(w1) * cnt * fast_rcp(den);
          if (good) { sn_sum += gq; sn_cnt += 1; }
          if (!good && w2 != 0.0) bad |= 1u << g;
        }
        bad &= ownmask;
        if (bad) {  // rare: exact recompute after the loop
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (bad & (1u << g)) {
              const int e = atomicAdd(susp_n, 1);
              if (e < kSuspCap) susp[e] = (rc << 16) | (cbase + g);
            }
          }
        }
      }
    }
  }
  ss_out = ln.wl * ss_sum;
  sn_out = ln.wl * sn_sum;
  cnt_out = ln.wl * (double)sn_cnt;
}

}  // namespace
