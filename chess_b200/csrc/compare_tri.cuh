// The -r 1 work-item kernel over the UPPER TRIANGLE of the pair.
//
// Both matrices of a band-resident pair are symmetric (and stay so under the 180-degree flip and
// the compaction, which removes the same rows and columns), hence so are their moment maps and the
// 7x7 difference-window map of SN.  Every off-diagonal pixel therefore counts twice and only
// pixels with column >= row are computed.  A lane still owns G adjacent columns and walks down
// rows, so the triangle is cut into three row bands whose column ranges shrink:
//
//        band 0: rows [0, B1)    columns [0, np)        all lanes            -> item A
//        band 1: rows [B1, B2)   columns [B1-3, np)     lanes [0, L1)        \  item B: the two narrow
//        band 2: rows [B2, np)   columns [B2-3, np)     lanes (L1, L1+L2]    /  bands side by side
//
// (three halo columns left of the diagonal feed the 7-column sums; an idle lane separates the two
// bands of item B, so the neighbour exchange reads zeros across the seam).  With B1 = np - 2x,
// B2 = np - x, x = (np-3)/3 both items run x + 6 rows: 2(x+6) = 64 warp-rows for np = 81 instead
// of the 87 of the two-band rectangle split, 146 instead of 207 for np = 201.
//
// The column sums that verify the predicted masks (chess/sim.py:74-87) no longer fall out of the
// loop (a lane sees only part of a column); they are screened with one subtraction per column in a
// per-sample prefix-sum array over the band rows (chess_band_prefix_build), and only a column
// whose approximate sum hits n*default is summed exactly, in the reference's order.
#pragma once
#include <cstdlib>
#include <type_traits>
#include "compare_kernels.cuh"

namespace {

constexpr int kTriRec = 16;  // doubles per item record

// lanes of item B: band 1 on [0, L1), band 2 on (L1, L1 + L2]; false if the split does not fit
__host__ __device__ __forceinline__ bool tri_split(int np_, int G, int &B1, int &B2, int &L1, int &L2) {
  B1 = np_; B2 = np_; L1 = 0; L2 = 0;
  if (np_ < 24) return false;
  const int x = (np_ - 3) / 3;
  const int b1 = np_ - 2 * x, b2 = np_ - x;
  const int l1 = (np_ - (b1 - 3) + G - 1) / G, l2 = (np_ - (b2 - 3) + G - 1) / G;
  if (l1 + l2 > 30) return false;
  B1 = b1; B2 = b2; L1 = l1; L2 = l2;
  return true;
}

// the same split with a halo of `halo` >= 3 rows / columns around every band (SSIM windows 9 and 11)
__host__ __device__ __forceinline__ bool tri_split_halo(int np_, int G, int halo, int &B1, int &B2, int &L1,
                                                        int &L2) {
  B1 = np_; B2 = np_; L1 = 0; L2 = 0;
  if (np_ < 24) return false;
  const int x = (np_ - halo) / 3;
  const int b1 = np_ - 2 * x, b2 = np_ - x;
  const int l1 = (np_ - (b1 - halo) + G - 1) / G, l2 = (np_ - (b2 - halo) + G - 1) / G;
  if (l1 + l2 > 30) return false;
  B1 = b1; B2 = b2; L1 = l1; L2 = l2;
  return true;
}

// ---- pieces shared by the triangle kernels (-r 1 here, windows in compare_tri7.cuh / compare_triw.cuh)

// One work item's view of its pair.
struct TriItem {
  long long pi;            // pair index in the launch
  int part;                // 0 = item A (band 0), 1 = item B (bands 1 and 2)
  int n, np_, flip;        // bins before / after the compaction, 180-degree flip of the query
  const double *R, *Q;     // first element of the two band views
  size_t rp, qp;           // their row pitches
  long long Wr, Wq;        // half-widths of the two bands
  int i0r, i0q;            // first bin of the two windows
  int tot_one_sided;       // bins masked in one matrix only
};

// Descriptor, soft failures, predicted masks, cutoff test and compaction map (chess/sim.py:319-353).
// Returns false when the item is already done (a status was written by item A, or nothing is left
// for this item to do).
//   pi        pair index in the launch;   part      which of the pair's items this is (0 writes the soft failures)
//   n_limit   widest matrix the calling kernel can hold
__device__ __forceinline__ bool tri_item_begin(const chess_item_args &a, long long pi, int part, int n_limit, int lane,
                                               short *kmap, unsigned char *mflag, TriItem &it) {
  const chess_params &p = a.p;
  const double qnan = CUDART_NAN;
  const bool use_masks = p.mappability_cutoff > 0;
  const int NCOL = a.ncol_pad;
  const chess_pair d = a.pairs[pi];
  int status = CHESS_PAIR_OK;
  if (d.ref_n <= 0 || d.qry_n <= 0) status = CHESS_PAIR_NOT_FOUND;
  else if (d.qry_n < p.min_bins || d.ref_n < p.min_bins) status = CHESS_PAIR_TOO_SMALL;
  else if (d.ref_n != d.qry_n || d.ref_n > n_limit || d.ref_n > NCOL) status = CHESS_PAIR_RETRY;
  if (status != CHESS_PAIR_OK) {
    if (part == 0 && lane == 0) {
      a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = status;
      if (status == CHESS_PAIR_RETRY) atomicOr(a.retry_cur, 1);
    }
    return false;
  }
  const int n = d.ref_n, flip = d.flip;
  // first bins of the two windows, recovered from the band view (off = i0*(2W-1) + W-1, pitch = 2W-2)
  const long long Wr = d.ref_pitch / 2 + 1, Wq = d.qry_pitch / 2 + 1;
  const int i0r = (int)((d.ref_off - (Wr - 1)) / (2 * Wr - 1));
  const int i0q = (int)((d.qry_off - (Wq - 1)) / (2 * Wq - 1));
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
      if (part == 0 && lane == 0) { a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_UNMAPPABLE; }
      return false;  // NB: a coincidentally masked column could only add to the count; see tri_masks_hold
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
  it.pi = pi; it.part = part; it.n = n; it.np_ = np_; it.flip = flip;
  it.R = a.ref_band + d.ref_off; it.Q = a.qry_band + d.qry_off;
  it.rp = (size_t)d.ref_pitch; it.qp = (size_t)d.qry_pitch;
  it.Wr = Wr; it.Wq = Wq; it.i0r = i0r; it.i0q = i0q; it.tot_one_sided = tot_one_sided;
  return true;
}

// Data range of the pair from the running-extrema arrays: n lookups give max / min of the
// uncompacted pair.  They are the extrema of the compacted pair too if every removed cell holds the
// default value in both matrices and the default is neither the max nor the min.  Otherwise (a bin
// masked in one matrix only is removed from the other, where it carries data: the norm when a
// region is compared with background regions) the removed rows alone are scanned -- k*n cells for
// k removed bins: if none of them reaches the uncompacted extremes, those are attained by cells
// that stay, and the lookup stands.  Returns true only when that fails too and the caller has to
// take min / max from all the data.
__device__ __forceinline__ bool tri_extrema(const chess_item_args &a, const TriItem &it, int lane,
                                            const unsigned char *mflag, double &xmax, double &xmin) {
  xmax = -CUDART_INF; xmin = CUDART_INF;
  if (a.ref_pmax == nullptr) return true;
  const int n = it.n;
  for (int r = lane; r < n; r += 32) {
    const size_t ir = (size_t)(it.i0r + r) * (size_t)it.Wr + (size_t)(n - 1 - r);
    const size_t iq = (size_t)(it.i0q + r) * (size_t)it.Wq + (size_t)(n - 1 - r);
    xmax = fmax(xmax, fmax(__ldg(a.ref_pmax + ir), __ldg(a.qry_pmax + iq)));
    xmin = fmin(xmin, fmin(__ldg(a.ref_pmin + ir), __ldg(a.qry_pmin + iq)));
  }
  xmax = warp_max(xmax);
  xmin = warp_min(xmin);
  if (it.np_ == n) return false;                         // nothing was removed
  if (it.tot_one_sided == 0 && xmax > a.p.default_value && xmin < a.p.default_value) return false;
  // removed rows (= removed columns: both matrices are symmetric), all n cells of each, both matrices
  double rmax = -CUDART_INF, rmin = CUDART_INF;
  for (int r = 0; r < n; ++r) {
    if (mflag[r] == 0) continue;
    const double *rowR = it.R + (size_t)r * it.rp;
    const double *rowQ = it.Q + (size_t)(it.flip ? n - 1 - r : r) * it.qp;
    for (int c = lane; c < n; c += 32) {
      const double x = __ldg(rowR + c), y = __ldg(rowQ + c);   // the order inside a row does not matter here
      rmax = fmax(rmax, fmax(x, y)); rmin = fmin(rmin, fmin(x, y));
    }
  }
  rmax = warp_max(rmax);
  rmin = warp_min(rmin);
  return !(rmax < xmax && rmin > xmin);
}

// Min / max of the compacted pair from the data (rare: a bin masked in one matrix only, or a
// default value that could be the extreme).  The upper triangle holds every value.
__device__ __forceinline__ void tri_scan_extrema(const TriItem &it, int lane, const short *kmap, double &xmax,
                                                 double &xmin) {
  double bmax = -CUDART_INF, bmin = CUDART_INF;
  for (int r = 0; r < it.np_; ++r) {
    const int kr = kmap[r];
    const double *rowR = it.R + (size_t)kr * it.rp;
    const double *rowQ = it.Q + (size_t)(it.flip ? it.n - 1 - kr : kr) * it.qp;
    for (int c = r + lane; c < it.np_; c += 32) {
      const int kc = kmap[c];
      const double x = __ldg(rowR + kc), y = __ldg(rowQ + (it.flip ? it.n - 1 - kc : kc));
      bmax = fmax(bmax, fmax(x, y)); bmin = fmin(bmin, fmin(x, y));
    }
  }
  xmax = warp_max(bmax); xmin = warp_min(bmin);
}

// Verification of the predicted masks by the item that arrives last: a column that was NOT
// predicted as masked must not sum to n*default.  The sum is screened with one subtraction in the
// row-prefix arrays; a hit within 1e-7 is summed exactly in the reference's order (rows 0 .. n-1 of
// the uncompacted matrix).  Returns false if some column is masked after all (per lane; combine
// with __any_sync).
__device__ __forceinline__ bool tri_masks_hold(const chess_item_args &a, const TriItem &it, int lane,
                                               const short *kmap, const unsigned char *mflag) {
  if (!(a.p.mappability_cutoff > 0)) return true;
  const int n = it.n, flip = it.flip;
  const double target = (double)n * a.p.default_value;
  bool bad = false;
  for (int c = lane; c < it.np_; c += 32) {
    const int kc = kmap[c];
    if (!(mflag[kc] & 1)) {
      const double *pr = a.ref_pre + (size_t)(it.i0r + kc) * (size_t)(2 * it.Wr) + (size_t)(it.Wr - 1 - kc);
      const double sr = __ldg(pr + n) - __ldg(pr);
      if (fabs(sr - target) <= 1e-7 * (fabs(sr) + fabs(target))) {
        double s2 = 0.0;
        for (int r = 0; r < n; ++r) s2 += __ldg(it.R + (size_t)r * it.rp + kc);
        bad |= s2 == target;
      }
    }
    if (!(mflag[kc] & 2)) {
      const int cq = flip ? n - 1 - kc : kc;
      const double *pq = a.qry_pre + (size_t)(it.i0q + cq) * (size_t)(2 * it.Wq) + (size_t)(it.Wq - 1 - cq);
      const double sq = __ldg(pq + n) - __ldg(pq);
      if (fabs(sq - target) <= 1e-7 * (fabs(sq) + fabs(target))) {
        double s2 = 0.0;
        for (int r = 0; r < n; ++r) s2 += __ldg(it.Q + (size_t)(flip ? n - 1 - r : r) * it.qp + cq);
        bad |= s2 == target;
      }
    }
  }
  return !bad;
}

// The first a.pair_items pairs of a launch are "pair items": one warp walks item A, then item B, with one
// set-up, no record and no arrival counter.  The remaining pairs are two independent items each, which keeps
// the tail of the grid short.  The launcher makes as many pairs as fill whole waves of warps pair items.
template <int G, bool WITH_SN>
__global__ void __launch_bounds__(128, (G <= 4 ? 4 : 2))
compare_r1_tri_kernel(chess_item_args a) {
  __shared__ int s_susp[4][kSuspCap];
  __shared__ int s_susp_n[4];
  __shared__ double s_part[4][8];   // per warp: totals of the parts of the current item
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
  const long long n_pair_items = a.pair_items;
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
    const bool PAIR = item < n_pair_items;
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

    int win = (np_ & 1) ? np_ : np_ - 1;
    if (win < 3) win = 3;
    if (win > np_) {
      if (part0 == 0 && lane == 0) { a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_WINDOW_EXCEEDS; }
      continue;
    }
    const int m = np_ - win + 1;

    // ---- data range from the running extrema, else min / max from the data
    double xmax, xmin;
    const bool inloop_minmax = tri_extrema(a, it, lane, mflag, xmax, xmin);  // rare in aligned samples, common against
                                                                      // background regions (a bin masked on one side)
    if (inloop_minmax) { xmax = -CUDART_INF; xmin = CUDART_INF; }   // neutral for an item without rows

    double T0 = 0, T1 = 0, T2 = 0, T3 = 0, T4 = 0;
    double sn_sum = 0.0;
    int sn_cnt = 0;
    bool overflow = false;
    double vmax = -CUDART_INF, vmin = CUDART_INF;   // used by the min / max instance of the row loop only
    int B1, B2, L1, L2;
    tri_split(np_, G, B1, B2, L1, L2);
    for (int part = part0; part <= (PAIR ? 1 : part0); ++part) {
    // ---- this item's lanes: rows [lo, hi), first column cbase
    if (PAIR && part == 1) {
      if (lane == 0) *susp_n = 0;
      __syncwarp();
    }
    int lo = 0, hi = 0, cbase = 1 << 20;
    if (part == 0) {
      lo = 0; hi = B1; cbase = G * lane;
    } else if (lane < L1) {
      lo = B1; hi = B2; cbase = B1 - 3 + G * lane;
    } else if (lane > L1 && lane <= L1 + L2) {
      lo = B2; hi = np_; cbase = B2 - 3 + G * (lane - L1 - 1);
    }
    const int rows_max = part == 0 ? B1 : max(B2 - B1, np_ - B2);
    unsigned xo[G], yo[G];
    bool valid[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int c = cbase + g;
      valid[g] = c < np_;
      const int kc = valid[g] ? (int)kmap[c] : 0;
      xo[g] = (unsigned)kc;
      yo[g] = (unsigned)(flip ? n - 1 - kc : kc);
    }
    if (rows_max > 0) {
      // per-lane row cursor: halo rows above the band only where the SN window needs them
      const int halo = WITH_SN ? 3 : 0;
      const int first = lo - halo < 0 ? 0 : lo - halo;            // part 0 starts at row 0
      const int t_emit0 = part == 0 ? halo : 2 * halo;           // first iteration with a complete window
      const int T = rows_max + (part == 0 ? halo : 2 * halo);
      double ccnt[G];
      double Sd[G], Se[G];
      unsigned hist[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = cbase + g;
        ccnt[g] = (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1);
        Sd[g] = 0.0; Se[g] = 0.0; hist[g] = 0u;
        if (WITH_SN) {
#pragma unroll
          for (int k = 0; k < 7; ++k) myring[(k * G + g) * 32] = 0.0;
        }
      }
      const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
      double *ringrow = myring;
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
      // two instances of the row loop: the usual one, and one that also tracks min / max of the data
      auto run = [&](auto minmax_tag) {
      constexpr bool MINMAX = decltype(minmax_tag)::value;
      // strips of 3 / 4 columns run 16 warps per SM at 128 registers: the row is loaded where it is used, the
      // other warps cover the latency; wider strips (8 warps per SM) keep the one-row software prefetch
      constexpr bool PREFETCH = G >= 5;
      int row = first;
      if (PREFETCH) fetch(row);
      for (int t = 0; t < T; ++t, ++row) {
        if (!PREFETCH) fetch(row);
        double dv[G];
        bool nzv[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { dv[g] = xn[g] - yn[g]; nzv[g] = xn[g] != yn[g]; }
        if (row >= lo && row < hi) {  // the band's own rows: weighted moments of the upper triangle
          const int d0 = cbase - row;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int dd = d0 + g;
            const double wt = dd > 0 ? 2.0 : (dd == 0 ? 1.0 : 0.0);
            const double wx = wt * xn[g], wy = wt * yn[g];  // columns that do not exist were fetched as 0
            T0 += wx; T1 += wy;
            T2 = fma(wx, xn[g], T2); T3 = fma(wy, yn[g], T3); T4 = fma(wx, yn[g], T4);
            if (MINMAX) {  // every loaded element of an own row belongs to the compacted pair
              const bool xgt = xn[g] > yn[g];
              const double hi2 = xgt ? xn[g] : yn[g];
              const double lo2 = xgt ? yn[g] : xn[g];
              vmax = (valid[g] & (hi2 > vmax)) ? hi2 : vmax;
              vmin = (valid[g] & (lo2 < vmin)) ? lo2 : vmin;
            }
          }
        }
        if (PREFETCH && t + 1 < T) fetch(row + 1);
        if (WITH_SN) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            double *cell = ringrow + g * 32;
            const double dold = *cell;
            *cell = dv[g];
            Sd[g] += dv[g]; Sd[g] -= dold;
            Se[g] = fma(dv[g], dv[g], Se[g]); Se[g] = fma(-dold, dold, Se[g]);
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
          if (t >= t_emit0) {
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
            const int d0 = cbase - rc;
            unsigned suspect = 0u;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const int a0 = g - 3 < 0 ? 0 : g - 3;
              const int b0 = g + 3 > G - 1 ? G - 1 : g + 3;
              double w1 = a0 == 0 ? P1[b0 + 1] : X1[a0];
              double w2 = a0 == 0 ? P2[b0 + 1] : X2[a0];
              if (g - 3 < 0) { w1 += up1[2 - g]; w2 += up2[2 - g]; }
              if (g + 3 > G - 1) { w1 += dn1[g + 3 - G]; w2 += dn2[g + 3 - G]; }
              const int dd = d0 + g;                      // column - row of this pixel
              const bool own = valid[g] & emit_row & (dd >= 0);
              const double cnt = rcnt * ccnt[g];
              const double m2 = w2 * cnt;
              const double den = fma(-w1, w1, m2);
              const bool good = own & (den > fma(m2, kVarGuard, 1e-290));
              const double gq = fabs(w1) * cnt * fast_rcp(den);
              sn_sum += good ? (dd > 0 ? gq + gq : gq) : 0.0;
              sn_cnt += good ? (dd > 0 ? 2 : 1) : 0;
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
      }
      };
      if (inloop_minmax) run(std::true_type());
      else run(std::false_type());
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
            if (gq == gq) { sn_sum += c > rc ? gq + gq : gq; sn_cnt += c > rc ? 2 : 1; }
          }
        } else if (ns > kSuspCap) {
          overflow = true;               // too many: the generic kernel redoes the pair exactly
        }
        __syncwarp();
      }
    }

    // this part's totals, reduced over the warp: exactly the numbers a half item writes to its record, and
    // added up in the order the combine adds two records -- a pair gives the same bits whichever way it ran
    // (kept in shared memory between the two parts of a pair item: seven doubles that would stay live
    // through the second row loop)
    {
      const double w0 = warp_sum(T0), w1 = warp_sum(T1), w2 = warp_sum(T2), w3 = warp_sum(T3), w4 = warp_sum(T4);
      const double ws = warp_sum(sn_sum), wc = (double)warp_sum_int(sn_cnt);
      if (lane == 0) {
        double *ptot = s_part[warp];
        const bool first_part = part == part0;
        ptot[0] = first_part ? w0 : ptot[0] + w0; ptot[1] = first_part ? w1 : ptot[1] + w1;
        ptot[2] = first_part ? w2 : ptot[2] + w2; ptot[3] = first_part ? w3 : ptot[3] + w3;
        ptot[4] = first_part ? w4 : ptot[4] + w4; ptot[5] = first_part ? ws : ptot[5] + ws;
        ptot[6] = first_part ? wc : ptot[6] + wc;
      }
      T0 = T1 = T2 = T3 = T4 = 0.0; sn_sum = 0.0; sn_cnt = 0;
    }
    }  // parts
    __syncwarp();
    const double P0 = s_part[warp][0], P1 = s_part[warp][1], P2 = s_part[warp][2], P3 = s_part[warp][3];
    const double P4 = s_part[warp][4], PS = s_part[warp][5], PC = s_part[warp][6];
    if (inloop_minmax) { xmax = warp_max(vmax); xmin = warp_min(vmin); }

    // first / last row totals for the 2x2 SSIM map (np even), after the row loops so that they are not live there
    double rt0[5] = {0, 0, 0, 0, 0}, rtL[5] = {0, 0, 0, 0, 0};
    if (m == 2) {
      for (int which = (PAIR || part0 == 0) ? 0 : 1; which <= ((PAIR || part0 == 1) ? 1 : 0); ++which) {
        const int kr = kmap[which == 0 ? 0 : np_ - 1];
        const double *rowR = R + (size_t)kr * rp;
        const double *rowQ = Q + (size_t)(flip ? n - 1 - kr : kr) * qp;
        double t5[5] = {0, 0, 0, 0, 0};
        for (int c = lane; c < np_; c += 32) {
          const int kc = kmap[c];
          const double x = __ldg(rowR + kc), y = __ldg(rowQ + (flip ? n - 1 - kc : kc));
          t5[0] += x; t5[1] += y; t5[2] = fma(x, x, t5[2]); t5[3] = fma(y, y, t5[3]); t5[4] = fma(x, y, t5[4]);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double v = warp_sum(t5[k]);
          if (which == 0) rt0[k] = v; else rtL[k] = v;
        }
      }
    }

    if (PAIR) {  // the whole pair was walked by this warp: verify the predicted masks and finish
      bool bad = overflow || !tri_masks_hold(a, it, lane, kmap, mflag);
      bad = __any_sync(0xffffffffu, bad);
      if (lane == 0) {
        if (bad) {
          a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
          atomicOr(a.retry_cur, 1);
        } else {
          const double t[5] = {P0, P1, P2, P3, P4};
          finish_pair(t, xmax, xmin, PS, PC, rt0, rtL, np_, win, m, kmap, R, Q, rp, qp, n, flip, WITH_SN,
                      a.ssim + pi, a.sn + pi, a.status + pi);
        }
      }
      continue;
    }

    // ---- record of this item
    double *rec = a.scratch + ((size_t)pi * 2 + part0) * kTriRec;
    const double *rt = part0 == 0 ? rt0 : rtL;
    {
      if (lane == 0) {
        rec[0] = P0; rec[1] = P1; rec[2] = P2; rec[3] = P3; rec[4] = P4;
        rec[5] = xmax; rec[6] = xmin; rec[7] = PS; rec[8] = PC;
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
        double t[5], mx, mn, ss, sc;
#pragma unroll
        for (int k = 0; k < 5; ++k) t[k] = __ldcg(recs + k) + __ldcg(recs + kTriRec + k);
        mx = fmax(__ldcg(recs + 5), __ldcg(recs + kTriRec + 5));
        mn = fmin(__ldcg(recs + 6), __ldcg(recs + kTriRec + 6));
        ss = __ldcg(recs + 7) + __ldcg(recs + kTriRec + 7);
        sc = __ldcg(recs + 8) + __ldcg(recs + kTriRec + 8);
        double row0[5], rowL[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { row0[k] = __ldcg(recs + 9 + k); rowL[k] = __ldcg(recs + kTriRec + 9 + k); }
        finish_pair(t, mx, mn, ss, sc, row0, rowL, np_, win, m, kmap, R, Q, rp, qp, n, flip, WITH_SN,
                    a.ssim + pi, a.sn + pi, a.status + pi);
      }
      a.done[pi] = 0;
    }
  }
}

template <int G, bool WITH_SN>
int launch_tri(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  auto kern = compare_r1_tri_kernel<G, WITH_SN>;
  constexpr int RING = 7 * G * 32;
  const size_t smem = (size_t)4 * RING * sizeof(double) + (size_t)4 * a.ncol_pad * 3;
  const size_t smem_max = (size_t)4 * RING * sizeof(double) + (size_t)4 * 256 * 3;
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
  // pair items for as many pairs as fill whole waves of warps (all of them once the launch is many waves
  // long: the tail no longer matters); two half items per pair for the rest
  const long long slots = grid * 4;
  static const int pair_env = getenv("CHESS_TRI_PAIR") ? atoi(getenv("CHESS_TRI_PAIR")) : -1;  // A/B runs
  long long pair_items = a.n_pairs >= 8 * slots ? a.n_pairs : (a.n_pairs / slots) * slots;
  if (pair_env == 0) pair_items = 0;
  if (pair_env == 1) pair_items = a.n_pairs;
  a.pair_items = pair_items;
  const long long need = (pair_items + 2 * (a.n_pairs - pair_items) + 3) / 4;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, 128, smem, stream>>>(a);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

// The same kernel with two independent half items per pair only (no pair items): the form used for strips of
// 5 .. 8 columns, where the extra bookkeeping of the hybrid scheduling above costs more registers than the
// saved set-up returns (hg38 @ 10 kb, 7 columns per lane: 14.7 ms against 15.3 ms).
template <int G, bool WITH_SN>
__global__ void __launch_bounds__(128, (G <= 4 ? 4 : 2))
compare_r1_tri_half_kernel(chess_item_args a) {
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
  const long long n_items = a.n_pairs * 2;

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
    if (!tri_item_begin(a, item >> 1, (int)(item & 1), 31 * G, lane, kmap, mflag, it)) continue;
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

    // ---- data range from the running extrema, else min / max from the data
    double xmax, xmin;
    const bool inloop_minmax = tri_extrema(a, it, lane, mflag, xmax, xmin);  // rare in aligned samples, common against
                                                                      // background regions (a bin masked on one side)
    if (inloop_minmax) { xmax = -CUDART_INF; xmin = CUDART_INF; }   // neutral for an item without rows

    // ---- this item's lanes: rows [lo, hi), first column cbase
    int B1, B2, L1, L2;
    tri_split(np_, G, B1, B2, L1, L2);
    int lo = 0, hi = 0, cbase = 1 << 20;
    if (part == 0) {
      lo = 0; hi = B1; cbase = G * lane;
    } else if (lane < L1) {
      lo = B1; hi = B2; cbase = B1 - 3 + G * lane;
    } else if (lane > L1 && lane <= L1 + L2) {
      lo = B2; hi = np_; cbase = B2 - 3 + G * (lane - L1 - 1);
    }
    const int rows_max = part == 0 ? B1 : max(B2 - B1, np_ - B2);
    unsigned xo[G], yo[G];
    bool valid[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int c = cbase + g;
      valid[g] = c < np_;
      const int kc = valid[g] ? (int)kmap[c] : 0;
      xo[g] = (unsigned)kc;
      yo[g] = (unsigned)(flip ? n - 1 - kc : kc);
    }
    double *rec = a.scratch + ((size_t)pi * 2 + part) * kTriRec;

    // first (part 0) / last (part 1) row totals for the 2x2 SSIM map
    double rt[5] = {0, 0, 0, 0, 0};
    if (m == 2) {
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
    double sn_sum = 0.0;
    int sn_cnt = 0;
    bool overflow = false;

    if (rows_max > 0) {
      // per-lane row cursor: halo rows above the band only where the SN window needs them
      const int halo = WITH_SN ? 3 : 0;
      const int first = lo - halo < 0 ? 0 : lo - halo;            // part 0 starts at row 0
      const int t_emit0 = part == 0 ? halo : 2 * halo;           // first iteration with a complete window
      const int T = rows_max + (part == 0 ? halo : 2 * halo);
      double ccnt[G];
      double Sd[G], Se[G];
      unsigned hist[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = cbase + g;
        ccnt[g] = (double)(min(c + 3, np_ - 1) - max(c - 3, 0) + 1);
        Sd[g] = 0.0; Se[g] = 0.0; hist[g] = 0u;
        if (WITH_SN) {
#pragma unroll
          for (int k = 0; k < 7; ++k) myring[(k * G + g) * 32] = 0.0;
        }
      }
      const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
      double *ringrow = myring;
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
      // two instances of the row loop: the usual one, and one that also tracks min / max of the data
      double vmax = -CUDART_INF, vmin = CUDART_INF;
      auto run = [&](auto minmax_tag) {
      constexpr bool MINMAX = decltype(minmax_tag)::value;
      // strips of 3 / 4 columns run 16 warps per SM at 128 registers: the row is loaded where it is used, the
      // other warps cover the latency; wider strips (8 warps per SM) keep the one-row software prefetch
      constexpr bool PREFETCH = G >= 5;
      int row = first;
      if (PREFETCH) fetch(row);
      for (int t = 0; t < T; ++t, ++row) {
        if (!PREFETCH) fetch(row);
        double dv[G];
        bool nzv[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { dv[g] = xn[g] - yn[g]; nzv[g] = xn[g] != yn[g]; }
        if (row >= lo && row < hi) {  // the band's own rows: weighted moments of the upper triangle
          const int d0 = cbase - row;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int dd = d0 + g;
            const double wt = dd > 0 ? 2.0 : (dd == 0 ? 1.0 : 0.0);
            const double wx = wt * xn[g], wy = wt * yn[g];  // columns that do not exist were fetched as 0
            T0 += wx; T1 += wy;
            T2 = fma(wx, xn[g], T2); T3 = fma(wy, yn[g], T3); T4 = fma(wx, yn[g], T4);
            if (MINMAX) {  // every loaded element of an own row belongs to the compacted pair
              const bool xgt = xn[g] > yn[g];
              const double hi2 = xgt ? xn[g] : yn[g];
              const double lo2 = xgt ? yn[g] : xn[g];
              vmax = (valid[g] & (hi2 > vmax)) ? hi2 : vmax;
              vmin = (valid[g] & (lo2 < vmin)) ? lo2 : vmin;
            }
          }
        }
        if (PREFETCH && t + 1 < T) fetch(row + 1);
        if (WITH_SN) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            double *cell = ringrow + g * 32;
            const double dold = *cell;
            *cell = dv[g];
            Sd[g] += dv[g]; Sd[g] -= dold;
            Se[g] = fma(dv[g], dv[g], Se[g]); Se[g] = fma(-dold, dold, Se[g]);
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
          if (t >= t_emit0) {
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
            const int d0 = cbase - rc;
            unsigned suspect = 0u;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const int a0 = g - 3 < 0 ? 0 : g - 3;
              const int b0 = g + 3 > G - 1 ? G - 1 : g + 3;
              double w1 = a0 == 0 ? P1[b0 + 1] : X1[a0];
              double w2 = a0 == 0 ? P2[b0 + 1] : X2[a0];
              if (g - 3 < 0) { w1 += up1[2 - g]; w2 += up2[2 - g]; }
              if (g + 3 > G - 1) { w1 += dn1[g + 3 - G]; w2 += dn2[g + 3 - G]; }
              const int dd = d0 + g;                      // column - row of this pixel
              const bool own = valid[g] & emit_row & (dd >= 0);
              const double cnt = rcnt * ccnt[g];
              const double m2 = w2 * cnt;
              const double den = fma(-w1, w1, m2);
              const bool good = own & (den > fma(m2, kVarGuard, 1e-290));
              const double gq = fabs(w1) * cnt * fast_rcp(den);
              sn_sum += good ? (dd > 0 ? gq + gq : gq) : 0.0;
              sn_cnt += good ? (dd > 0 ? 2 : 1) : 0;
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
      }
      };
      if (inloop_minmax) {
        run(std::true_type());
        xmax = warp_max(vmax); xmin = warp_min(vmin);
      } else {
        run(std::false_type());
      }
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
            if (gq == gq) { sn_sum += c > rc ? gq + gq : gq; sn_cnt += c > rc ? 2 : 1; }
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
      const int sc = warp_sum_int(sn_cnt);
      if (lane == 0) {
        rec[0] = t0; rec[1] = t1; rec[2] = t2; rec[3] = t3; rec[4] = t4;
        rec[5] = xmax; rec[6] = xmin; rec[7] = ss; rec[8] = (double)sc;
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
        double t[5], mx, mn, ss, sc;
#pragma unroll
        for (int k = 0; k < 5; ++k) t[k] = __ldcg(recs + k) + __ldcg(recs + kTriRec + k);
        mx = fmax(__ldcg(recs + 5), __ldcg(recs + kTriRec + 5));
        mn = fmin(__ldcg(recs + 6), __ldcg(recs + kTriRec + 6));
        ss = __ldcg(recs + 7) + __ldcg(recs + kTriRec + 7);
        sc = __ldcg(recs + 8) + __ldcg(recs + kTriRec + 8);
        double row0[5], rowL[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { row0[k] = __ldcg(recs + 9 + k); rowL[k] = __ldcg(recs + kTriRec + 9 + k); }
        finish_pair(t, mx, mn, ss, sc, row0, rowL, np_, win, m, kmap, R, Q, rp, qp, n, flip, WITH_SN,
                    a.ssim + pi, a.sn + pi, a.status + pi);
      }
      a.done[pi] = 0;
    }
  }
}

template <int G, bool WITH_SN>
int launch_tri_half(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  auto kern = compare_r1_tri_half_kernel<G, WITH_SN>;
  constexpr int RING = 7 * G * 32;
  const size_t smem = (size_t)4 * RING * sizeof(double) + (size_t)4 * a.ncol_pad * 3;
  const size_t smem_max = (size_t)4 * RING * sizeof(double) + (size_t)4 * 256 * 3;
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
  const long long need = (a.n_pairs * 2 + 3) / 4;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, 128, smem, stream>>>(a);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

}  // namespace
