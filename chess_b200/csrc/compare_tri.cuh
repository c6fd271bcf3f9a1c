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
#include "tri_rows.cuh"

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
template <bool BATCH = true>
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
    if constexpr (BATCH) {
      // four groups of 32 columns at a time: their 16 look-ups per lane are in flight together (one per
      // iteration, each waiting for the one before, was 2 % of the kernel's stall samples at 201 bins)
      for (int base = 0; base < n; base += 128) {
        int lo_r[4], hi_r[4], lo_q[4], hi_q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = base + 32 * u + lane;
          lo_r[u] = lo_q[u] = 0x7fffffff; hi_r[u] = hi_q[u] = -1;      // "not masked" for columns beyond n
          if (c < n) {
            const int jr = i0r + c, jq = i0q + (flip ? n - 1 - c : c);
            lo_r[u] = __ldg(a.ref_lo + jr); hi_r[u] = __ldg(a.ref_hi + jr);
            lo_q[u] = __ldg(a.qry_lo + jq); hi_q[u] = __ldg(a.qry_hi + jq);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = base + 32 * u + lane;
          const bool fr = lo_r[u] < i0r && hi_r[u] >= i0r + n;
          const bool fq = lo_q[u] < i0q && hi_q[u] >= i0q + n;
          if (c < n) mflag[c] = (unsigned char)((fr ? 1 : 0) | (fq ? 2 : 0));
          tot_r += __popc(__ballot_sync(0xffffffffu, fr));
          tot_q += __popc(__ballot_sync(0xffffffffu, fq));
          tot_one_sided += __popc(__ballot_sync(0xffffffffu, fr != fq));
        }
      }
    } else {
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
template <bool BATCH = true>
__device__ __forceinline__ bool tri_extrema(const chess_item_args &a, const TriItem &it, int lane,
                                            const unsigned char *mflag, double &xmax, double &xmin) {
  xmax = -CUDART_INF; xmin = CUDART_INF;
  if (a.ref_pmax == nullptr) return true;
  const int n = it.n;
  if constexpr (BATCH) {
    for (int r0 = lane; r0 < n; r0 += 128) {     // 16 look-ups per lane in flight together
      double v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + 32 * u;
        v[u][0] = v[u][1] = -CUDART_INF; v[u][2] = v[u][3] = CUDART_INF;
        if (r < n) {
          const size_t ir = (size_t)(it.i0r + r) * (size_t)it.Wr + (size_t)(n - 1 - r);
          const size_t iq = (size_t)(it.i0q + r) * (size_t)it.Wq + (size_t)(n - 1 - r);
          v[u][0] = __ldg(a.ref_pmax + ir); v[u][1] = __ldg(a.qry_pmax + iq);
          v[u][2] = __ldg(a.ref_pmin + ir); v[u][3] = __ldg(a.qry_pmin + iq);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xmax = fmax(xmax, fmax(v[u][0], v[u][1]));
        xmin = fmin(xmin, fmin(v[u][2], v[u][3]));
      }
    }
  } else {
    for (int r = lane; r < n; r += 32) {
      const size_t ir = (size_t)(it.i0r + r) * (size_t)it.Wr + (size_t)(n - 1 - r);
      const size_t iq = (size_t)(it.i0q + r) * (size_t)it.Wq + (size_t)(n - 1 - r);
      xmax = fmax(xmax, fmax(__ldg(a.ref_pmax + ir), __ldg(a.qry_pmax + iq)));
      xmin = fmin(xmin, fmin(__ldg(a.ref_pmin + ir), __ldg(a.qry_pmin + iq)));
    }
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
template <bool BATCH = true>
__device__ __forceinline__ bool tri_masks_hold(const chess_item_args &a, const TriItem &it, int lane,
                                               const short *kmap, const unsigned char *mflag) {
  if (!(a.p.mappability_cutoff > 0)) return true;
  const int n = it.n, flip = it.flip;
  const double target = (double)n * a.p.default_value;
  bool bad = false;
  if constexpr (BATCH) {
    for (int c0 = lane; c0 < it.np_; c0 += 128) {     // the prefix look-ups of four columns per lane together
      double sr[4], sq[4];
      int kcs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        double r1 = 0.0, r0 = 0.0, q1 = 0.0, q0 = 0.0;
        kcs[u] = -1;
        if (c < it.np_) {
          const int kc = kmap[c];
          kcs[u] = kc;
          const int cq = flip ? n - 1 - kc : kc;
          const double *pr = a.ref_pre + (size_t)(it.i0r + kc) * (size_t)(2 * it.Wr) + (size_t)(it.Wr - 1 - kc);
          const double *pq = a.qry_pre + (size_t)(it.i0q + cq) * (size_t)(2 * it.Wq) + (size_t)(it.Wq - 1 - cq);
          r1 = __ldg(pr + n); r0 = __ldg(pr); q1 = __ldg(pq + n); q0 = __ldg(pq);
        }
        sr[u] = r1 - r0; sq[u] = q1 - q0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kc = kcs[u];
        if (kc < 0) continue;
        if (!(mflag[kc] & 1) && fabs(sr[u] - target) <= 1e-7 * (fabs(sr[u]) + fabs(target))) {
          double s2 = 0.0;
          for (int r = 0; r < n; ++r) s2 += __ldg(it.R + (size_t)r * it.rp + kc);
          bad |= s2 == target;
        }
        if (!(mflag[kc] & 2) && fabs(sq[u] - target) <= 1e-7 * (fabs(sq[u]) + fabs(target))) {
          const int cq = flip ? n - 1 - kc : kc;
          double s2 = 0.0;
          for (int r = 0; r < n; ++r) s2 += __ldg(it.Q + (size_t)(flip ? n - 1 - r : r) * it.qp + cq);
          bad |= s2 == target;
        }
      }
    }
  } else {
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
  }
  return !bad;
}

// One part of a pair on one warp: lane geometry, the row loop (tri_rows.cuh), the exact recompute of the
// queued windows, then the warp totals.  tot[0..4] = moments of the weighted triangle, tot[5] = SN sum,
// tot[6] = SN count; vmax / vmin are reduced only when `minmax` (the data range could not be looked up).
// Returns false if the cancellation queue overflowed (the generic kernel then redoes the pair exactly).
template <int G, bool WITH_SN>
__device__ __forceinline__ bool tri_r1_run_part(const TriItem &it, int part, int lane, const short *kmap,
                                                double *myring, int *susp, int *susp_n, bool minmax,
                                                double (&tot)[7], double &vmax, double &vmin) {
  const int np_ = it.np_;
  constexpr int HALO = WITH_SN ? 3 : 0;
  int B1, B2, L1, L2;
  const int mode = tri_mode(np_, G, B1, B2, L1, L2);
  // part: 0 = band 0 (or the whole matrix), 1 = bands 1 and 2, 2 = all three bands in one pass (mode 1)
  const TriLane ln = tri_lane(part, lane, G, np_, B1, B2, L1, L2, HALO);
  const int rows_max = part == 0 ? B1 : part == 1 ? (mode == 2 ? max(B2 - B1, np_ - B2) : 0)
                                                  : max(B1, max(B2 - B1, np_ - B2) + HALO);
  TriPartSums ps;
#pragma unroll
  for (int k = 0; k < 5; ++k) ps.T[k] = 0.0;
  ps.sn_sum = 0.0; ps.sn_cnt = 0.0; ps.vmax = -CUDART_INF; ps.vmin = CUDART_INF;
  bool ok = true;
  if (rows_max > 0) {
    if (lane == 0) *susp_n = 0;
    __syncwarp();
    tri_r1_part<G, WITH_SN>(it.R, it.Q, (unsigned)it.rp, (unsigned)it.qp, it.n, np_, it.flip, kmap, myring, ln,
                            part != 1, rows_max, minmax, susp, susp_n, ps);
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
          const double gq = sn_window_exact(it.R, it.Q, it.rp, it.qp, kmap, it.n, it.flip, rc, c, np_, cnt);
          const int band_hi = rc < B1 ? B1 : (rc < B2 ? B2 : np_);  // end of the row's diagonal block
          if (gq == gq) { ps.sn_sum += c >= band_hi ? gq + gq : gq; ps.sn_cnt += c >= band_hi ? 2.0 : 1.0; }
        }
      } else if (ns > kSuspCap) {
        ok = false;
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) tot[k] = warp_sum(ps.T[k]);
  tot[5] = warp_sum(ps.sn_sum);
  tot[6] = warp_sum(ps.sn_cnt);
  if (minmax) { vmax = warp_max(ps.vmax); vmin = warp_min(ps.vmin); }
  return ok;
}

// First (which = 0) / last (which = 1) row totals of the five moment maps, for the 2x2 SSIM map of an even np.
__device__ __forceinline__ void tri_edge_row(const TriItem &it, int which, int lane, const short *kmap, double *rt) {
  const int kr = kmap[which == 0 ? 0 : it.np_ - 1];
  const double *rowR = it.R + (size_t)kr * it.rp;
  const double *rowQ = it.Q + (size_t)(it.flip ? it.n - 1 - kr : kr) * it.qp;
  double t5[5] = {0, 0, 0, 0, 0};
  for (int c0 = lane; c0 < it.np_; c0 += 128) {     // eight loads per lane in flight; same order of additions
    double xs[4], ys[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 32 * u;
      xs[u] = 0.0; ys[u] = 0.0;
      if (c < it.np_) {
        const int kc = kmap[c];
        xs[u] = __ldg(rowR + kc); ys[u] = __ldg(rowQ + (it.flip ? it.n - 1 - kc : kc));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (c0 + 32 * u < it.np_) {
        const double x = xs[u], y = ys[u];
        t5[0] += x; t5[1] += y; t5[2] = fma(x, x, t5[2]); t5[3] = fma(y, y, t5[3]); t5[4] = fma(x, y, t5[4]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) rt[k] = warp_sum(t5[k]);
}

// Two records -> ssim, SN, status of the pair.  Deliberately not inlined: both ways a pair can run end here.
__device__ __noinline__ void tri_r1_finish(const double *recs, int np_, int win, int m, const short *kmap,
                                           const double *R, const double *Q, size_t rp, size_t qp, int n, int flip,
                                           bool with_sn, double *ssim_out, double *sn_out, int *status_out) {
  double t[5], row0[5], rowL[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    t[k] = recs[k] + recs[kTriRec + k];
    row0[k] = recs[9 + k];
    rowL[k] = recs[kTriRec + 9 + k];
  }
  const double mx = fmax(recs[5], recs[kTriRec + 5]), mn = fmin(recs[6], recs[kTriRec + 6]);
  const double ss = recs[7] + recs[kTriRec + 7], sc = recs[8] + recs[kTriRec + 8];
  finish_pair(t, mx, mn, ss, sc, row0, rowL, np_, win, m, kmap, R, Q, rp, qp, n, flip, with_sn, ssim_out, sn_out,
              status_out);
}

// The first a.pair_items pairs of a launch are "pair items": one warp walks item A, then item B, with one
// set-up, no record and no arrival counter.  The remaining pairs are two independent items each, which keeps
// the tail of the grid short.  The launcher makes as many pairs as fill whole waves of warps pair items
// (HYBRID; strips of 5 .. 8 columns run half items only: the bookkeeping costs them more registers than the
// saved set-up returns).  A pair gives the same bits whichever way it ran: the parts are reduced over the
// warp separately and added in the order the combine adds two records.
template <int G, bool WITH_SN, bool HYBRID, int MINB>
__global__ void __launch_bounds__(128, MINB)
compare_r1_tri_kernel(chess_item_args a) {
  __shared__ int s_susp[4][kSuspCap];
  __shared__ int s_susp_n[4];
  __shared__ double s_rec[4][2 * kTriRec];   // per warp: the two records of the pair being finished
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
  const double qnan = CUDART_NAN;
  if (blockIdx.x == 0 && threadIdx.x == 0) { *a.counter_next = 0ull; *a.retry_next = 0; }
  const long long n_pair_items = HYBRID ? a.pair_items : 0;
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
    const bool PAIR = HYBRID && item < n_pair_items;
    const long long half = item - n_pair_items;   // index among the half items
    if (!tri_item_begin(a, PAIR ? item : n_pair_items + (half >> 1), PAIR ? 0 : (int)(half & 1), 31 * G, lane, kmap,
                        mflag, it))
      continue;
    const long long pi = it.pi;
    const int part0 = it.part, np_ = it.np_;

    int win = (np_ & 1) ? np_ : np_ - 1;
    if (win < 3) win = 3;
    if (win > np_) {
      if (part0 == 0 && lane == 0) { a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_WINDOW_EXCEEDS; }
      continue;
    }
    const int m = np_ - win + 1;

    // ---- data range from the running extrema, else min / max from the data (rare in aligned samples, common
    // against background regions: a bin masked on one side)
    double xmax, xmin;
    const bool fixed_range = a.p.data_range > 0.0;   // chess_params.data_range: no look-up, no scan
    const bool inloop_minmax = !fixed_range && tri_extrema(a, it, lane, mflag, xmax, xmin);
    if (inloop_minmax) { xmax = -CUDART_INF; xmin = CUDART_INF; }   // neutral for an item without rows
    if (fixed_range) { xmax = a.p.data_range; xmin = 0.0; }

    // every part ends in a record -- in shared memory for a pair item, in the scratch array for a half item --
    // and ONE routine (tri_r1_finish, not inlined) turns two records into the result: a pair gives the same
    // bits whichever way it ran
    bool overflow = false;
    int tB1, tB2, tL1, tL2;
    const bool mono = tri_mode(np_, G, tB1, tB2, tL1, tL2) == 1;   // all bands in one pass: the first part does it all
    for (int part = part0; part <= (PAIR ? 1 : part0); ++part) {
      double tot[7], vmax = -CUDART_INF, vmin = CUDART_INF;
      overflow |= !tri_r1_run_part<G, WITH_SN>(it, (mono && part == 0) ? 2 : part, lane, kmap, myring, susp, susp_n,
                                               inloop_minmax, tot, vmax, vmin);
      double rt[5] = {0, 0, 0, 0, 0};     // first (part 0) / last (part 1) row totals for the 2x2 SSIM map (np even)
      if (m == 2) tri_edge_row(it, part, lane, kmap, rt);
      if (lane == 0) {
        double *rec = PAIR ? s_rec[warp] + part * kTriRec : a.scratch + ((size_t)pi * 2 + part) * kTriRec;
#pragma unroll
        for (int k = 0; k < 5; ++k) rec[k] = tot[k];
        rec[5] = inloop_minmax ? vmax : xmax; rec[6] = inloop_minmax ? vmin : xmin;
        rec[7] = tot[5]; rec[8] = tot[6];
#pragma unroll
        for (int k = 0; k < 5; ++k) rec[9 + k] = rt[k];
        rec[14] = overflow ? 1.0 : 0.0;
      }
    }
    if (!PAIR) {
      __threadfence();
      __syncwarp();
      int arrived = 0;
      if (lane == 0) arrived = atomicAdd(a.done + pi, 1);
      arrived = __shfl_sync(0xffffffffu, arrived, 0);
      if (arrived != 1) continue;
      // second item to arrive: fetch both records
      __threadfence();
      const double *recs = a.scratch + (size_t)pi * 2 * kTriRec;
      s_rec[warp][lane] = __ldcg(recs + lane);
    }
    __syncwarp();
    // verify the predicted masks, combine, write the result
    bool bad = s_rec[warp][14] != 0.0 || s_rec[warp][kTriRec + 14] != 0.0;
    bad |= !tri_masks_hold(a, it, lane, kmap, mflag);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      if (bad) {
        a.ssim[pi] = qnan; a.sn[pi] = qnan; a.status[pi] = CHESS_PAIR_RETRY;
        atomicOr(a.retry_cur, 1);
      } else {
        tri_r1_finish(s_rec[warp], np_, win, m, kmap, it.R, it.Q, it.rp, it.qp, it.n, it.flip, WITH_SN, a.ssim + pi,
                      a.sn + pi, a.status + pi);
      }
      if (!PAIR) a.done[pi] = 0;
    }
    __syncwarp();
  }
}

template <int G, bool WITH_SN, bool HYBRID, int MINB>
int launch_tri_impl(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  auto kern = compare_r1_tri_kernel<G, WITH_SN, HYBRID, MINB>;
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
  long long pair_items = !HYBRID ? 0 : (a.n_pairs >= 8 * slots ? a.n_pairs : (a.n_pairs / slots) * slots);
  if (pair_env == 0) pair_items = 0;
  if (pair_env == 1 && HYBRID) pair_items = a.n_pairs;
  int b1, b2, l1, l2;
  if (HYBRID && tri_mode(a.max_n, G, b1, b2, l1, l2) != 2) pair_items = a.n_pairs;  // one pass per pair: nothing to halve
  a.pair_items = pair_items;
  const long long need = (pair_items + 2 * (a.n_pairs - pair_items) + 3) / 4;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, 128, smem, stream>>>(a);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

template <int G, bool WITH_SN>
int launch_tri(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  // strips of 3 / 4 columns: 128 registers and 16 warps per SM -- or 168 registers and 12 warps without spills for
  // the one-pass layout of small matrices in batches of many waves (41 bins, 100 k pairs: 1.60 -> 1.48 ms; two
  // parts at 81 bins lose 1 %, and a batch of one wave -- 2 402 pairs -- needs the 16 warps: 90 -> 104 us)
  static const int wide_env = getenv("CHESS_TRI_WIDE") ? atoi(getenv("CHESS_TRI_WIDE")) : -1;  // A/B runs
  int b1, b2, l1, l2;
  const bool mono = tri_mode(a.max_n, G, b1, b2, l1, l2) == 1;
  const bool wide = wide_env >= 0 ? wide_env != 0 : (mono && a.n_pairs >= 8LL * ctx->sm_count * 16);
  if (wide) return launch_tri_impl<G, WITH_SN, true, 3>(ctx, a, stream);
  return launch_tri_impl<G, WITH_SN, true, 4>(ctx, a, stream);
}

// Strips of 5 .. 8 columns: two independent half items per pair only (hg38 @ 10 kb, 7 columns per lane: the
// hybrid scheduling measured 15.3 ms against 14.7 ms).
template <int G, bool WITH_SN>
int launch_tri_half(chess_ctx *ctx, chess_item_args &a, cudaStream_t stream) {
  return launch_tri_impl<G, WITH_SN, false, 2>(ctx, a, stream);
}

}  // namespace
