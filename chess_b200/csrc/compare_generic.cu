// Generic (any n <= 1024, any odd window) per-pair comparison kernel.
//
// One CTA per region pair.  It follows compare_matrix_pair (chess/sim.py:306-380)
// step by step, reading the two submatrices straight from their storage (the
// symmetric O/E band or packed dense matrices, see chess_pair in
// include/chess_b200.h) -- the gather of sub_matrix_from_edges_dict
// (chess/helpers.py:286-306) is index arithmetic here, never a copy:
//
//   0. not found / < min_bins                         sim.py:198-200, 319-320
//   1. index maps: order-0 resize of the smaller side, 180-degree flip of the
//      query when the strands differ                  sim.py:323-333
//   2. column sums in row order (bit-identical to np.sum(m, 0)), masked
//      columns, cutoff test                           sim.py:74-87, 336-347
//   3. union of masked columns removed from both      sim.py:350-353
//   4. window rule, data range                        sim.py:356-370
//   5. SSIM: running sums of x, y, xx, yy, xy down the columns, block scan
//      across a row, skimage's formula on every valid window, mean
//   6. SN: direct (subtraction-free) 7x7 sums of d and d*d with the
//      NaN-padded border of filter_uniform_filter (sim.py:33-71); windows whose
//      one-pass variance suffers cancellation are recomputed two-pass in
//      numpy's own summation order from a 7-row ring in shared memory.
//
// This kernel is the exact, shape-agnostic path; compare_fast.cu holds the
// tuned kernels and falls back to this one.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kRed = 384;  // doubles of reduction / scan scratch (block reductions: 0..32, warp totals: 64..223,
                           // warp offsets: 224..383)

template <int BLOCK>
__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < BLOCK / 32 ? red[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) red[32] = v;
  }
  __syncthreads();
  v = red[32];
  __syncthreads();
  return v;
}
template <int BLOCK>
__device__ __forceinline__ double block_max(double v, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < BLOCK / 32 ? red[lane] : -CUDART_INF;
    v = warp_max(v);
    if (lane == 0) red[32] = v;
  }
  __syncthreads();
  v = red[32];
  __syncthreads();
  return v;
}

// numpy's pairwise summation for n = 49 contiguous doubles (n < 128: eight
// running accumulators over blocks of eight, combined as a tree, then the tail).
__device__ __forceinline__ double numpy_sum49(const double *a) {
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
#pragma unroll
  for (int i = 8; i < 48; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  return __dadd_rn(res, a[48]);
}

// |nanmean| / nanvar of one NaN-padded 7x7 window, evaluated the way numpy
// does it (two passes, np.nanmean / np.nanvar on the 49-slot axis).
__device__ __noinline__ double sn_window_twopass(const double *ring, int stride, int rc, int c,
                                                 int np_, double cnt) {
  double val[49];
  bool ok[49];
#pragma unroll
  for (int wi = 0; wi < 7; ++wi) {
    const int rr = rc - 3 + wi;
    const bool rok = rr >= 0 && rr < np_;
    const int slot = ((rr % 7) + 7) % 7;
#pragma unroll
    for (int wj = 0; wj < 7; ++wj) {
      const int cc = c - 3 + wj;
      const bool v = rok && cc >= 0 && cc < np_;
      ok[wi * 7 + wj] = v;
      val[wi * 7 + wj] = v ? ring[slot * stride + cc] : 0.0;
    }
  }
  const double avg = __ddiv_rn(numpy_sum49(val), cnt);
#pragma unroll
  for (int s = 0; s < 49; ++s) {
    const double dv = ok[s] ? __dsub_rn(val[s], avg) : 0.0;
    val[s] = __dmul_rn(dv, dv);
  }
  const double var = __ddiv_rn(numpy_sum49(val), cnt);
  return __ddiv_rn(fabs(avg), var);
}

// WITH_SN = false drops the SN phase from the instantiation (SSIM-only batches: background comparisons, the
// dense one-vs-many batches of configs[0]); its registers then allow MINB CTAs of BLOCK threads per SM.
template <int BLOCK, int CPT, bool WITH_SN, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
compare_generic_kernel(const double *__restrict__ ref_base, const double *__restrict__ qry_base,
                       const chess_pair *__restrict__ pairs, long long n_pairs, int ncap,
                       chess_params p, double *__restrict__ ssim_out, double *__restrict__ sn_out,
                       int *__restrict__ status_out, int retry_only, const int *retry_flag) {
  if (retry_flag != nullptr) {
    // retry pass: launched with programmatic stream serialisation so that its launch latency
    // overlaps the tail of the tuned kernel; wait here until that kernel's writes are visible
    cudaGridDependencySynchronize();
    if (*retry_flag == 0) return;  // nothing was left for the generic kernel
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *red = reinterpret_cast<double *>(smem_raw);
  double *buf = red + kRed;                       // max(10*(ncap+3), 9*ncap+12) doubles
  int *mr = reinterpret_cast<int *>(buf + 10 * (size_t)ncap + 44);
  int *mq = mr + ncap;                            // pre-compaction source indices
  int *kr = mq + ncap;
  int *kq = kr + ncap;                            // compacted source indices
  int *iscratch = kq + ncap;                      // 64 ints
  unsigned char *flag = reinterpret_cast<unsigned char *>(iscratch + 64);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double qnan = CUDART_NAN;

  // Pairs of this CTA.  Normal pass: blockIdx.x, + gridDim.x, ...  Retry pass: the CTAs scan the status
  // array 32 pairs at a time (one coalesced load and a ballot per group, identical in every warp of the
  // CTA) and walk only into the groups that hold a RETRY mark -- a 65 536-pair chunk with a handful of
  // marks costs microseconds, not the millisecond that one status load per pair and CTA did.
  long long group = blockIdx.x, cur_base = 0;
  unsigned pending = 0;
  long long pi = (long long)blockIdx.x - (long long)gridDim.x;
  while (true) {
    if (retry_only) {
      while (pending == 0 && group * 32 < n_pairs) {
        cur_base = group * 32;
        const long long q = cur_base + lane;
        pending = __ballot_sync(0xffffffffu, q < n_pairs && status_out[q] == CHESS_PAIR_RETRY);
        group += gridDim.x;
      }
      if (pending == 0) break;
      pi = cur_base + (__ffs(pending) - 1);
      pending &= pending - 1;
    } else {
      pi += gridDim.x;
      if (pi >= n_pairs) break;
    }
    const chess_pair d = pairs[pi];
    int status = CHESS_PAIR_OK;
    if (d.ref_n <= 0 || d.qry_n <= 0) status = CHESS_PAIR_NOT_FOUND;
    else if (d.qry_n < p.min_bins || d.ref_n < p.min_bins) status = CHESS_PAIR_TOO_SMALL;
    if (status != CHESS_PAIR_OK) {
      if (tid == 0) { ssim_out[pi] = qnan; sn_out[pi] = qnan; status_out[pi] = status; }
      continue;
    }
    const int n = max(d.ref_n, d.qry_n);
    const double *R = ref_base + d.ref_off;
    const double *Q = qry_base + d.qry_off;
    const size_t rp = (size_t)d.ref_pitch, qp = (size_t)d.qry_pitch;

    // ---- 1. index maps (resize + flip)
    for (int k = tid; k < n; k += BLOCK) {
      mr[k] = d.ref_n < n ? resize_src_index(k, d.ref_n, n) : k;
      const int t = d.flip ? n - 1 - k : k;
      mq[k] = d.qry_n < n ? resize_src_index(t, d.qry_n, n) : t;
    }
    __syncthreads();

    // ---- 2. column sums in row order, min / max of the uncompacted pair
    const bool use_masks = p.mappability_cutoff > 0;
    const double masked_sum = (double)n * p.default_value;
    double vmax = -CUDART_INF, vmin = CUDART_INF;
    int cnt_r = 0, cnt_q = 0;
    const bool need_pass = use_masks || !(p.data_range > 0.0);   // neither sums nor extremes are wanted otherwise
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tid + j * BLOCK;
      if (need_pass && c < n) {
        const int cr = mr[c], cq = mq[c];
        double sr = 0.0, sq = 0.0;
        for (int r = 0; r < n; ++r) {
          const double x = __ldg(R + (size_t)mr[r] * rp + cr);
          const double y = __ldg(Q + (size_t)mq[r] * qp + cq);
          sr += x;
          sq += y;
          vmax = fmax(vmax, fmax(x, y));
          vmin = fmin(vmin, fmin(x, y));
        }
        const bool fr = use_masks && sr == masked_sum;
        const bool fq = use_masks && sq == masked_sum;
        flag[c] = (unsigned char)((fr ? 1 : 0) | (fq ? 2 : 0));
        cnt_r += fr;
        cnt_q += fq;
      }
    }
    const int tot_r = (int)block_sum<BLOCK>((double)cnt_r, red);
    const int tot_q = (int)block_sum<BLOCK>((double)cnt_q, red);
    if (use_masks && ((double)tot_r / (double)n > p.mappability_cutoff ||
                      (double)tot_q / (double)n > p.mappability_cutoff)) {
      if (tid == 0) { ssim_out[pi] = qnan; sn_out[pi] = qnan; status_out[pi] = CHESS_PAIR_UNMAPPABLE; }
      __syncthreads();
      continue;
    }

    // ---- 3. remove the union of masked columns from rows and columns of both
    int np_ = n;
    const bool remove = use_masks && !p.keep_unmappable && (tot_r + tot_q) > 0;
    if (remove) {
      int carry = 0;
      for (int base = 0; base < n; base += BLOCK) {
        const int c = base + tid;
        const bool keep = c < n && flag[c] == 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) iscratch[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
        for (int w = 0; w < BLOCK / 32; ++w) {
          const int v = iscratch[w];
          if (w < warp) woff += v;
          total += v;
        }
        if (keep) {
          const int pos = carry + woff + __popc(bal & ((1u << lane) - 1u));
          kr[pos] = mr[c];
          kq[pos] = mq[c];
        }
        carry += total;
        __syncthreads();
      }
      np_ = carry;
    } else {
      for (int k = tid; k < n; k += BLOCK) { kr[k] = mr[k]; kq[k] = mq[k]; }
    }
    __syncthreads();

    // ---- 4. window rule and data range
    const int win = window_rule(np_, p.abs_window, p.rel_window);
    if (win > np_) {
      if (tid == 0) { ssim_out[pi] = qnan; sn_out[pi] = qnan; status_out[pi] = CHESS_PAIR_WINDOW_EXCEEDS; }
      __syncthreads();
      continue;
    }
    if (remove) {  // min / max over the compacted pair only
      vmax = -CUDART_INF;
      vmin = CUDART_INF;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = tid + j * BLOCK;
        if (c < np_) {
          const int cr = kr[c], cq = kq[c];
          for (int r = 0; r < np_; ++r) {
            const double x = __ldg(R + (size_t)kr[r] * rp + cr);
            const double y = __ldg(Q + (size_t)kq[r] * qp + cq);
            vmax = fmax(vmax, fmax(x, y));
            vmin = fmin(vmin, fmin(x, y));
          }
        }
      }
    }
    const double dmax = block_max<BLOCK>(vmax, red);
    const double dmin = -block_max<BLOCK>(-vmin, red);
    const double drange = p.data_range > 0.0 ? p.data_range : dmax - dmin;

    // ---- 5. SSIM over every valid window
    const int m = np_ - win + 1;
    const double NP = (double)win * (double)win;
    const double cov_norm = NP / (NP - 1.0);
    const double C1 = (0.01 * drange) * (0.01 * drange);
    const double C2 = (0.03 * drange) * (0.03 * drange);
    double acc = 0.0;
    {
      // Thread t owns the CPT adjacent compacted columns CPT*t .. CPT*t+CPT-1: five running sums down
      // each column (add the incoming row, subtract the row that left the window).  Across the row the
      // window sum is a difference of prefix sums; the prefix is built in registers -- over the thread's
      // columns, then a shuffle scan over the lanes -- and only the warp-level prefixes and the warp
      // totals go through shared memory, double-buffered by row parity: ONE barrier per row.
      constexpr int NW = BLOCK / 32;
      // column c of a prefix array sits at (c % CPT) * PS + c / CPT: the threads of a warp store and
      // load consecutive words (no bank conflicts at any CPT)
      const int PS = (ncap + CPT - 1) / CPT, PA = PS * CPT;
      double *pre = buf;            // [2][5][PA] inclusive prefix within the owning warp
      double *wtot = red + 64;      // [2][5][16]   totals of the warps
      double *woff = red + 224;     // [2][5][16]   sum of the warps left of a warp
      const int c0 = CPT * tid;
      double V[CPT][5];
#pragma unroll
      for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int k = 0; k < 5; ++k) V[j][k] = 0.0;
      int kcr[CPT], kcq[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        kcr[j] = c0 + j < np_ ? kr[c0 + j] : 0;
        kcq[j] = c0 + j < np_ ? kq[c0 + j] : 0;
      }
      // S = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)) with every mean written as
      // window sum / NP: numerator and denominator are scaled by NP^4, one division per window
      const double c1s = C1 * NP * NP, c2s = C2 * NP * NP;
      // rows enter and leave through registers loaded one iteration ahead (the loads of row r+1 are in
      // flight while row r goes through the scan and the barrier)
      double nx[CPT], ny[CPT], ox[CPT], oy[CPT];
      // dense, uncompacted, unresized views with an even pitch and a 16-byte aligned origin: the thread's
      // adjacent columns are adjacent in memory and come in 16-byte loads (half the L1 wavefronts)
      const bool vecx = np_ == n && d.ref_n == n && (rp & 1) == 0 && (reinterpret_cast<size_t>(R) & 15) == 0;
      const bool vecy = np_ == n && d.qry_n == n && !d.flip && (qp & 1) == 0 && (reinterpret_cast<size_t>(Q) & 15) == 0;
      auto load_cols = [&](const double *row, const int *kc, bool vec, bool on, double *out) {
        if constexpr (CPT % 2 == 0) {
#pragma unroll
          for (int i = 0; i < CPT; i += 2) {
            if (on && vec && c0 + i + 1 < np_) {
              const double2 v = __ldg(reinterpret_cast<const double2 *>(row + c0 + i));
              out[i] = v.x; out[i + 1] = v.y;
            } else {
              out[i] = on && c0 + i < np_ ? __ldg(row + kc[i]) : 0.0;
              out[i + 1] = on && c0 + i + 1 < np_ ? __ldg(row + kc[i + 1]) : 0.0;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < CPT; ++i) out[i] = on && c0 + i < np_ ? __ldg(row + kc[i]) : 0.0;
        }
      };
      load_cols(R + (size_t)kr[0] * rp, kcr, vecx, true, nx);
      load_cols(Q + (size_t)kq[0] * qp, kcq, vecy, true, ny);
#pragma unroll
      for (int j = 0; j < CPT; ++j) { ox[j] = 0.0; oy[j] = 0.0; }
      for (int r = 0; r < np_; ++r) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const double x = nx[j], y = ny[j], xo = ox[j], yo = oy[j];
          V[j][0] += x; V[j][1] += y; V[j][2] += x * x; V[j][3] += y * y; V[j][4] += x * y;
          V[j][0] -= xo; V[j][1] -= yo; V[j][2] -= xo * xo; V[j][3] -= yo * yo; V[j][4] -= xo * yo;
        }
        if (r + 1 < np_) {
          const bool sub = r + 1 >= win;
          load_cols(R + (size_t)kr[r + 1] * rp, kcr, vecx, true, nx);
          load_cols(Q + (size_t)kq[r + 1] * qp, kcq, vecy, true, ny);
          load_cols(R + (size_t)kr[sub ? r + 1 - win : 0] * rp, kcr, vecx, sub, ox);
          load_cols(Q + (size_t)kq[sub ? r + 1 - win : 0] * qp, kcq, vecy, sub, oy);
        }
        if (r >= win - 1) {   // uniform across the CTA
          const int par = r & 1;
          double W[5], E[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            W[k] = V[0][k];
#pragma unroll
            for (int j = 1; j < CPT; ++j) W[k] += V[j][k];
          }
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              const double t = __shfl_up_sync(0xffffffffu, W[k], o);
              if (lane >= o) W[k] += t;
            }
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const double t = __shfl_up_sync(0xffffffffu, W[k], 1);
            E[k] = lane ? t : 0.0;                       // exclusive prefix of this thread within its warp
            if (NW > 1 && lane == 31) wtot[(par * 5 + k) * 16 + warp] = W[k];
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            double run = E[k];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
              run += V[j][k];
              if (c0 + j < np_) pre[(par * 5 + k) * PA + j * PS + tid] = run;
            }
          }
          __syncthreads();
          // warp 0 turns the warp totals into the offsets of the warps (exclusive scan); everybody reads them
          // from shared memory after a second barrier (cheaper than every warp repeating the shuffle scan)
          if (NW > 1) {
            if (warp == 0) {
#pragma unroll
              for (int k = 0; k < 5; ++k) {
                double t = lane < NW ? wtot[(par * 5 + k) * 16 + lane] : 0.0;
#pragma unroll
                for (int o = 1; o < NW; o <<= 1) {
                  const double u = __shfl_up_sync(0xffffffffu, t, o);
                  if (lane >= o) t += u;
                }
                const double ex = __shfl_up_sync(0xffffffffu, t, 1);
                if (lane < NW) woff[(par * 5 + k) * 16 + lane] = lane ? ex : 0.0;
              }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 5; ++k) E[k] += woff[(par * 5 + k) * 16 + warp];   // prefix left of column c0
          }
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const int c = c0 + j;
            const int hi = min(c + win - 1, np_ - 1);
            const int whi = (hi / CPT) >> 5;             // warp that owns column hi
            if (c < m) {
              double h[5];
#pragma unroll
              for (int k = 0; k < 5; ++k) {
                const double up = pre[(par * 5 + k) * PA + (hi % CPT) * PS + hi / CPT] +
                                  (NW > 1 ? woff[(par * 5 + k) * 16 + whi] : 0.0);
                h[k] = up - E[k];
              }
              const double pxy = h[0] * h[1], pxx = h[0] * h[0], pyy = h[1] * h[1];
              const double a1 = 2.0 * pxy + c1s, b1 = pxx + pyy + c1s;
              const double a2 = 2.0 * cov_norm * (NP * h[4] - pxy) + c2s;
              const double b2 = cov_norm * ((NP * h[2] - pxx) + (NP * h[3] - pyy)) + c2s;
              acc += (a1 * a2) / (b1 * b2);
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) E[k] += V[j][k];   // prefix left of the next column
          }
        }
      }
      __syncthreads();   // the last row's reads of wtot / pre are done before red / buf are reused
    }
    const double ssim = block_sum<BLOCK>(acc, red) / ((double)m * (double)m);

    // ---- 6. SN
    double sn = qnan;
    if (WITH_SN && p.compute_sn) {
      double *ring = buf;                 // 7 * ncap
      double *a1 = buf + 7 * (size_t)ncap; // np_+6, shifted by 3
      double *a2 = a1 + ncap + 6;
      for (int k = tid; k < 3; k += BLOCK) {
        a1[k] = 0.0; a2[k] = 0.0; a1[np_ + 3 + k] = 0.0; a2[np_ + 3 + k] = 0.0;
      }
      double dreg[CPT][7];
#pragma unroll
      for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int k = 0; k < 7; ++k) dreg[j][k] = 0.0;
      double sacc = 0.0, scnt = 0.0;
      for (int r = 0; r < np_ + 3; ++r) {
        const bool live = r < np_;
        const size_t ro = live ? (size_t)kr[r] * rp : 0, qo = live ? (size_t)kq[r] * qp : 0;
        const int slot = r % 7;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int c = tid + j * BLOCK;
          if (c < np_) {
            const double dv = live ? __ldg(R + ro + kr[c]) - __ldg(Q + qo + kq[c]) : 0.0;
#pragma unroll
            for (int k = 6; k > 0; --k) dreg[j][k] = dreg[j][k - 1];
            dreg[j][0] = dv;
            ring[slot * ncap + c] = dv;
            double s1 = dreg[j][0], s2 = dreg[j][0] * dreg[j][0];
#pragma unroll
            for (int k = 1; k < 7; ++k) { s1 += dreg[j][k]; s2 += dreg[j][k] * dreg[j][k]; }
            a1[3 + c] = s1;
            a2[3 + c] = s2;
          }
        }
        __syncthreads();
        if (r >= 3) {
          const int rc = r - 3;
          const int rcnt = min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1;
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const int c = tid + j * BLOCK;
            if (c < np_) {
              double w1 = 0.0, w2 = 0.0;
#pragma unroll
              for (int t = 0; t < 7; ++t) { w1 += a1[c + t]; w2 += a2[c + t]; }
              if (w2 != 0.0) {  // w2 == 0 <=> every d in the window is 0 <=> 0/0 -> NaN, skipped
                const int ccnt = min(c + 3, np_ - 1) - max(c - 3, 0) + 1;
                const double cnt = (double)(rcnt * ccnt);
                const double mean = w1 / cnt, ms = w2 / cnt;
                const double var = ms - mean * mean;
                double g;
                if (var > 1e-6 * ms) g = fabs(mean) / var;
                else g = sn_window_twopass(ring, ncap, rc, c, np_, cnt);
                if (g == g) { sacc += g; scnt += 1.0; }
              }
            }
          }
        }
        __syncthreads();
      }
      const double tot = block_sum<BLOCK>(sacc, red);
      const double cnt = block_sum<BLOCK>(scnt, red);
      sn = tot / cnt;  // 0/0 -> NaN, like np.nanmean of an all-NaN map
    }
    if (tid == 0) { ssim_out[pi] = ssim; sn_out[pi] = sn; status_out[pi] = CHESS_PAIR_OK; }
    __syncthreads();
  }
}

size_t generic_smem_bytes(int ncap) {
  size_t doubles = kRed + 10 * (size_t)ncap + 44;
  size_t ints = 4 * (size_t)ncap + 64;
  size_t bytes = doubles * 8 + ints * 4 + (size_t)ncap;
  return (bytes + 15) & ~size_t(15);
}

template <int BLOCK, int CPT, bool WITH_SN = true, int MINB = 1>
int launch(chess_ctx *ctx, const double *ref_base, const double *qry_base, const chess_pair *pairs,
           int64_t n_pairs, int ncap, const chess_params &p, double *ssim, double *sn,
           int32_t *status, cudaStream_t stream, bool retry_only, const int *retry_flag) {
  auto kern = compare_generic_kernel<BLOCK, CPT, WITH_SN, MINB>;
  const size_t smem = generic_smem_bytes(ncap);
  if ((int)smem > ctx->max_smem_optin) {
    ctx->err = "compare_generic: shared memory demand exceeds the device limit";
    return CHESS_ERR_UNSUPPORTED;
  }
  CHESS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  CHESS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem));
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > n_pairs) grid = n_pairs;
  // the retry pass after a tuned kernel normally has nothing to do (it returns on the flag) and at
  // most a handful of pairs otherwise: a few CTAs keep its launch cost down
  if (retry_flag != nullptr) {
    long long want = (n_pairs + 255) / 256;   // 8 status groups per CTA
    if (want < 16) want = 16;
    if (grid > want) grid = want;
  }
  if (retry_flag != nullptr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(BLOCK);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CHESS_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, ref_base, qry_base, pairs, (long long)n_pairs, ncap, p, ssim,
                                       sn, status, retry_only ? 1 : 0, retry_flag));
  } else {
    kern<<<(unsigned)grid, BLOCK, smem, stream>>>(ref_base, qry_base, pairs, (long long)n_pairs, ncap, p,
                                                  ssim, sn, status, retry_only ? 1 : 0, retry_flag);
  }
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

}  // namespace

int chess_launch_compare_generic(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                                 const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                                 const chess_params &p, double *ssim, double *sn, int32_t *status,
                                 cudaStream_t stream, bool retry_only, const int *retry_flag) {
  if (n_pairs <= 0) return CHESS_OK;
  if (max_n > 1024) {
    ctx->err = "compare: submatrices larger than 1024 bins are not supported";
    return CHESS_ERR_UNSUPPORTED;
  }
  const int ncap = max_n < 32 ? 32 : max_n;
  // four adjacent columns per thread in the SSIM loop: a pair of <= 128 bins is one warp (no cross-warp step)
  // a retry pass holds a handful of pairs: what counts is the latency of one pair, so small matrices get a
  // thread per column (four warps) instead of one warp per pair
  const bool few_pairs = retry_only && ncap <= 128;
  if (!p.compute_sn && !few_pairs) {
    // SSIM only (background comparisons, the dense one-vs-many batches of configs[0]): adjacent columns per
    // thread amortise the prefix shuffles; without the SN phase the kernel fits 16 warps per SM
#define CHESS_GEN_ARGS ctx, ref_base, qry_base, pairs, n_pairs, ncap, p, ssim, sn, status, stream, retry_only, retry_flag
    // measured on the dense one-vs-many batches (profiles/r01_config0_*.txt): one warp per pair up to 128 bins,
    // two columns per thread and 16 warps per SM above
    // (more than a wave of them: the 168-register build, 12 warps per SM; a single wave is bound by the
    // latency of one warp, where the unconstrained build is faster: 0.198 against 0.231 ms at 1 001 pairs)
    if (ncap <= 128 && n_pairs > 9LL * ctx->sm_count) return launch<32, 4, false, 12>(CHESS_GEN_ARGS);
    if (ncap <= 128) return launch<32, 4, false, 1>(CHESS_GEN_ARGS);
    if (ncap <= 256) return launch<128, 2, false, 4>(CHESS_GEN_ARGS);
    if (ncap <= 512) return launch<256, 2, false, 2>(CHESS_GEN_ARGS);
    return launch<512, 2, false, 1>(CHESS_GEN_ARGS);
#undef CHESS_GEN_ARGS
  }
  // with SN the 7-row window phase wants a thread per column (measured: 573 us against 699 us on configs[1])
  if (ncap <= 128) return launch<128, 1, true, 3>(ctx, ref_base, qry_base, pairs, n_pairs, ncap, p, ssim, sn, status, stream, retry_only, retry_flag);
  if (ncap <= 256) return launch<256, 1>(ctx, ref_base, qry_base, pairs, n_pairs, ncap, p, ssim, sn, status, stream, retry_only, retry_flag);
  if (ncap <= 512) return launch<256, 2>(ctx, ref_base, qry_base, pairs, n_pairs, ncap, p, ssim, sn, status, stream, retry_only, retry_flag);
  return launch<256, 4>(ctx, ref_base, qry_base, pairs, n_pairs, ncap, p, ssim, sn, status, stream, retry_only, retry_flag);
}
