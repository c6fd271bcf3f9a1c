// Tuned compare kernels for sm_100a.
//
// compare_r1_kernel -- the default `chess sim` regime (-r 1: the SSIM window is
// the whole compacted matrix or one less, so the SSIM map is 1x1 or 2x2,
// chess/sim.py:356-365) on symmetric band data with equal-sized sides.
//
// Work decomposition (B200: 148 SMs, FP64 64 FMA/clk/SM, L2 ~70 B/clk/SM for
// 8-byte loads; the samples stay L2-resident, so the kernel is bound by issue
// slots and the FP64 pipe, not by HBM):
//   * one CTA per pair, NW warps; warp w owns a band of rows, lane l owns the
//     G adjacent columns G*l .. G*l+G-1 in registers (row loads are contiguous
//     across the warp: coalesced 8-byte loads straight from the band);
//   * pass 1 (masks): per-band partial column sums + an "any non-default"
//     bit per column; columns are classified exactly like
//     np.sum(m, 0) == n * default (chess/sim.py:82-85): all-default columns are
//     masked, columns whose blocked sum is not within 1e-7 of n * default are
//     not, and the (practically never occurring) near-coincidences are
//     re-summed sequentially in row order by one thread;
//   * pass 2: global moments of x, y (SSIM, with min / max for the data range)
//     and the SN map: running 7-row sums of d = x - y and d*d down each column
//     (a 7-deep register ring supplies the outgoing row; a per-column count of
//     non-zero d snaps all-zero windows to exactly 0, which is what makes
//     0/0 -> NaN -> skipped come out as in np.nanmean), 7-column sums across
//     lanes by three up / three down shuffles of strip prefix / suffix sums,
//     then |mean| / var with a Newton-refined reciprocal.  Windows whose
//     one-pass variance shows cancellation are recomputed two-pass in numpy's
//     summation order straight from memory.
//   * the 2x2 SSIM map needs only totals, the first / last row totals and four
//     corner elements, because every moment map of a symmetric pair is
//     symmetric (column totals = row totals).
// Pairs this kernel cannot take (unequal sizes, n > 32*G) are marked
// CHESS_PAIR_RETRY and handled by the generic kernel.
#include <cstdlib>

#include "compare_kernels.cuh"
#include "tri_rows.cuh"

namespace {

template <int G, int NW, bool WITH_SN>
__global__ void __launch_bounds__(32 * NW, (G <= 4 ? 4 : 2))
compare_r1_kernel(const double *__restrict__ ref_base, const double *__restrict__ qry_base,
                  const chess_pair *__restrict__ pairs, long long n_pairs, chess_params p,
                  double *__restrict__ ssim_out, double *__restrict__ sn_out,
                  int *__restrict__ status_out) {
  constexpr int NCOL = 32 * G;
  constexpr int NT = 32 * NW;
  __shared__ short kmap[NCOL];
  __shared__ double cpart[NW][2][NCOL];
  __shared__ unsigned char ndpart[NW][NCOL];
  __shared__ unsigned char mflag[NCOL];
  __shared__ double red[NW][12];
  __shared__ double rowtot[2][5];
  __shared__ int icount[4];
  __shared__ int wcount[NW];
  __shared__ int susp[NW][kSuspCap];
  __shared__ int susp_n[NW];
  extern __shared__ __align__(16) double ringbuf[];  // NW * 7 * G * 32 doubles (SN only)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double qnan = CUDART_NAN;
  const int c0 = G * lane;

  for (long long pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
    const chess_pair d = pairs[pi];
    int status = CHESS_PAIR_OK;
    if (d.ref_n <= 0 || d.qry_n <= 0) status = CHESS_PAIR_NOT_FOUND;
    else if (d.qry_n < p.min_bins || d.ref_n < p.min_bins) status = CHESS_PAIR_TOO_SMALL;
    else if (d.ref_n != d.qry_n || d.ref_n > 31 * G) status = CHESS_PAIR_RETRY;
    if (status != CHESS_PAIR_OK) {
      if (tid == 0) { ssim_out[pi] = qnan; sn_out[pi] = qnan; status_out[pi] = status; }
      continue;
    }
    const int n = d.ref_n;
    const int flip = d.flip;
    const double *R = ref_base + d.ref_off;
    const double *Q = qry_base + d.qry_off;
    const size_t rp = (size_t)d.ref_pitch, qp = (size_t)d.qry_pitch;
    const bool use_masks = p.mappability_cutoff > 0;
    int np_ = n;

    if (use_masks) {
      // ---- pass 1: blocked column sums + non-default bits, one row band per warp
      const int B1 = (n + NW - 1) / NW;
      const int r_lo = warp * B1, r_hi = min(n, r_lo + B1);
      const unsigned long long defbits = (unsigned long long)__double_as_longlong(p.default_value);
      double cs_r[G], cs_q[G];
      unsigned long long nd_r[G], nd_q[G];
#pragma unroll
      for (int g = 0; g < G; ++g) { cs_r[g] = 0.0; cs_q[g] = 0.0; nd_r[g] = 0ull; nd_q[g] = 0ull; }
      {
        const double *xp = R + (size_t)r_lo * rp + c0;
        const double *yp = flip ? Q + (size_t)(n - 1 - r_lo) * qp + (n - c0 - G) : Q + (size_t)r_lo * qp + c0;
        const ptrdiff_t ystep = flip ? -(ptrdiff_t)qp : (ptrdiff_t)qp;
        bool valid[G];
#pragma unroll
        for (int g = 0; g < G; ++g) valid[g] = c0 + g < n;
#pragma unroll 4
        for (int r = r_lo; r < r_hi; ++r) {
          double x[G], y[G];
#pragma unroll
          for (int g = 0; g < G; ++g) { x[g] = p.default_value; y[g] = p.default_value; }
#pragma unroll
          for (int g = 0; g < G; ++g)
            if (valid[g]) x[g] = __ldg(xp + g);
          if (!flip) {
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (valid[g]) y[g] = __ldg(yp + g);
          } else {
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (valid[g]) y[g] = __ldg(yp + (G - 1 - g));
          }
          xp += rp;
          yp += ystep;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            cs_r[g] += x[g];
            cs_q[g] += y[g];
            nd_r[g] |= (unsigned long long)__double_as_longlong(x[g]) ^ defbits;
            nd_q[g] |= (unsigned long long)__double_as_longlong(y[g]) ^ defbits;
          }
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = c0 + g;
        cpart[warp][0][c] = cs_r[g];
        cpart[warp][1][c] = cs_q[g];
        ndpart[warp][c] = (unsigned char)((nd_r[g] ? 1 : 0) | (nd_q[g] ? 2 : 0));
      }
      if (tid < 4) icount[tid] = 0;
      __syncthreads();
      // ---- classify columns
      const double target = (double)n * p.default_value;
      int cnt_r = 0, cnt_q = 0;
      for (int c = tid; c < n; c += NT) {
        double sr = cpart[0][0][c], sq = cpart[0][1][c];
        unsigned nd = ndpart[0][c];
#pragma unroll
        for (int w = 1; w < NW; ++w) { sr += cpart[w][0][c]; sq += cpart[w][1][c]; nd |= ndpart[w][c]; }
        bool fr, fq;
        if (!(nd & 1)) fr = true;  // every entry equals the default: n additions of 0 or 1 are exact
        else if (fabs(sr - target) > 1e-7 * (fabs(sr) + fabs(target))) fr = false;
        else {  // near-coincidence: decide with the reference's own summation order
          double s = 0.0;
          for (int r = 0; r < n; ++r) s += __ldg(R + (size_t)r * rp + c);
          fr = s == target;
        }
        if (!(nd & 2)) fq = true;
        else if (fabs(sq - target) > 1e-7 * (fabs(sq) + fabs(target))) fq = false;
        else {
          double s = 0.0;
          const int cq = flip ? n - 1 - c : c;
          for (int r = 0; r < n; ++r) s += __ldg(Q + (size_t)(flip ? n - 1 - r : r) * qp + cq);
          fq = s == target;
        }
        mflag[c] = (unsigned char)((fr ? 1 : 0) | (fq ? 2 : 0));
        cnt_r += fr;
        cnt_q += fq;
      }
      cnt_r = warp_sum_int(cnt_r);
      cnt_q = warp_sum_int(cnt_q);
      if (lane == 0 && (cnt_r | cnt_q)) { atomicAdd(&icount[0], cnt_r); atomicAdd(&icount[1], cnt_q); }
      __syncthreads();
      const int tot_r = icount[0], tot_q = icount[1];
      if ((double)tot_r / (double)n > p.mappability_cutoff || (double)tot_q / (double)n > p.mappability_cutoff) {
        if (tid == 0) { ssim_out[pi] = qnan; sn_out[pi] = qnan; status_out[pi] = CHESS_PAIR_UNMAPPABLE; }
        __syncthreads();
        continue;
      }
      // ---- compaction map
      if (!p.keep_unmappable && (tot_r + tot_q) > 0) {
        int carry = 0;
        for (int base = 0; base < n; base += NT) {
          const int c = base + tid;
          const bool keep = c < n && mflag[c] == 0;
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if (lane == 0) wcount[warp] = __popc(bal);
          __syncthreads();
          int woff = 0, total = 0;
#pragma unroll
          for (int w = 0; w < NW; ++w) {
            const int v = wcount[w];
            if (w < warp) woff += v;
            total += v;
          }
          if (keep) kmap[carry + woff + __popc(bal & ((1u << lane) - 1u))] = (short)c;
          carry += total;
          __syncthreads();
        }
        np_ = carry;
      } else {
        for (int c = tid; c < n; c += NT) kmap[c] = (short)c;
        __syncthreads();
      }
    } else {
      for (int c = tid; c < n; c += NT) kmap[c] = (short)c;
      __syncthreads();
    }

    // ---- window rule (-r 1): the whole matrix, made odd, at least 3
    int win = (np_ & 1) ? np_ : np_ - 1;
    if (win < 3) win = 3;
    if (win > np_) {
      if (tid == 0) { ssim_out[pi] = qnan; sn_out[pi] = qnan; status_out[pi] = CHESS_PAIR_WINDOW_EXCEEDS; }
      __syncthreads();
      continue;
    }
    const int m = np_ - win + 1;  // 1 or 2

    // ---- pass 2
    // first / last row totals of the five moments (needed for the 2x2 SSIM map only)
    if (m == 2 && warp < 2) {
      const int rr = warp == 0 ? 0 : np_ - 1;
      const int kr = kmap[rr];
      const double *rowR = R + (size_t)kr * rp;
      const double *rowQ = Q + (size_t)(flip ? n - 1 - kr : kr) * qp;
      double t[5] = {0, 0, 0, 0, 0};
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = c0 + g;
        if (c < np_) {
          const int kcc = kmap[c];
          const double x = __ldg(rowR + kcc), y = __ldg(rowQ + (flip ? n - 1 - kcc : kcc));
          t[0] += x; t[1] += y; t[2] = fma(x, x, t[2]); t[3] = fma(y, y, t[3]); t[4] = fma(x, y, t[4]);
        }
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        t[k] = warp_sum(t[k]);
        if (lane == 0) rowtot[warp][k] = t[k];
      }
    }

    const int B2 = (np_ + NW - 1) / NW;
    const int lo = warp * B2, hi = min(np_, lo + B2);
    Band2 acc;
    acc.T0 = acc.T1 = acc.T2 = acc.T3 = acc.T4 = 0.0;
    acc.vmax = -CUDART_INF; acc.vmin = CUDART_INF; acc.sn_sum = 0.0; acc.sn_cnt = 0;
    double *myring = ringbuf + (size_t)warp * 7 * G * 32 + lane;
    if (lane == 0) susp_n[warp] = 0;
    __syncwarp();
    if (lo < hi) {
      LaneCols<G, false> lc;
      lane_cols<G, false>(lc, kmap, n, np_, flip, lane, 0, 0, np_);
      moments_rows<G, false, true>(R, Q, rp, qp, kmap, n, flip, lo, hi, lc, acc, nullptr, nullptr);
      if (WITH_SN) sn_rows<G, false>(R, Q, rp, qp, kmap, n, np_, flip, lo, hi, lane, lc, myring, susp[warp],
                                     &susp_n[warp], acc, nullptr, nullptr);
      if (WITH_SN) {
        // cold: windows whose one-pass variance cancelled -> numpy's own two passes, after the hot loop
        __syncwarp();
        const int ns = susp_n[warp];
        if (ns > 0 && ns <= kSuspCap) {
          // sort the queued codes (arrival order is not reproducible; the addition order must be)
          if (lane == 0) {
            for (int i2 = 1; i2 < ns; ++i2) {
              const int key = susp[warp][i2];
              int j2 = i2 - 1;
              while (j2 >= 0 && susp[warp][j2] > key) { susp[warp][j2 + 1] = susp[warp][j2]; --j2; }
              susp[warp][j2 + 1] = key;
            }
          }
          __syncwarp();
          for (int e = lane; e < ns; e += 32) {
            const int rc = susp[warp][e] >> 16, c = susp[warp][e] & 0xffff;
            const double cnt = (double)((min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1) *
                                        (min(c + 3, np_ - 1) - max(c - 3, 0) + 1));
            const double gq = sn_window_exact(R, Q, rp, qp, kmap, n, flip, rc, c, np_, cnt);
            if (gq == gq) { acc.sn_sum += gq; acc.sn_cnt += 1; }
          }
        } else if (ns > kSuspCap) {  // list overflow: redo the whole band exactly
          acc.sn_sum = 0.0; acc.sn_cnt = 0;
          for (int rc = lo; rc < hi; ++rc) {
            for (int c = lane; c < np_; c += 32) {
              const double cnt = (double)((min(rc + 3, np_ - 1) - max(rc - 3, 0) + 1) *
                                          (min(c + 3, np_ - 1) - max(c - 3, 0) + 1));
              const double gq = sn_window_exact(R, Q, rp, qp, kmap, n, flip, rc, c, np_, cnt);
              if (gq == gq) { acc.sn_sum += gq; acc.sn_cnt += 1; }
            }
          }
        }
      }
    }
    double T0 = acc.T0, T1 = acc.T1, T2 = acc.T2, T3 = acc.T3, T4 = acc.T4;
    double vmax = acc.vmax, vmin = acc.vmin, sn_sum = acc.sn_sum;
    int sn_cnt = acc.sn_cnt;

    // ---- combine the row bands (fixed order -> deterministic)
    T0 = warp_sum(T0); T1 = warp_sum(T1); T2 = warp_sum(T2); T3 = warp_sum(T3); T4 = warp_sum(T4);
    vmax = warp_max(vmax); vmin = warp_min(vmin);
    sn_sum = warp_sum(sn_sum);
    sn_cnt = warp_sum_int(sn_cnt);
    if (lane == 0) {
      red[warp][0] = T0; red[warp][1] = T1; red[warp][2] = T2; red[warp][3] = T3; red[warp][4] = T4;
      red[warp][5] = vmax; red[warp][6] = vmin; red[warp][7] = sn_sum; red[warp][8] = (double)sn_cnt;
    }
    __syncthreads();
    if (tid == 0) {
      double t[5] = {0, 0, 0, 0, 0}, mx = -CUDART_INF, mn = CUDART_INF, ss = 0.0, sc = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
#pragma unroll
        for (int k = 0; k < 5; ++k) t[k] += red[w][k];
        mx = fmax(mx, red[w][5]); mn = fmin(mn, red[w][6]);
        ss += red[w][7]; sc += red[w][8];
      }
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
            // window (a, b) leaves out row er and column ec; column totals = row totals (symmetry)
            const int ia = a == 0 ? 1 : 0, ib = b == 0 ? 1 : 0;
            const int er = ia ? np_ - 1 : 0, ec = ib ? np_ - 1 : 0;
            const int kr = kmap[er], kcc = kmap[ec];
            const double x = __ldg(R + (size_t)kr * rp + kcc);
            const double y = flip ? __ldg(Q + (size_t)(n - 1 - kr) * qp + (n - 1 - kcc))
                                  : __ldg(Q + (size_t)kr * qp + kcc);
            const double corner[5] = {x, y, x * x, y * y, x * y};
#pragma unroll
            for (int k = 0; k < 5; ++k) h[k] = h[k] - rowtot[ia][k] - rowtot[ib][k] + corner[k];
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
      ssim_out[pi] = acc / (double)(m * m);
      sn_out[pi] = WITH_SN ? ss / sc : qnan;
      status_out[pi] = CHESS_PAIR_OK;
    }
    __syncthreads();
  }
}

template <int G, int NW, bool WITH_SN>
int launch_r1_sn(chess_ctx *ctx, const double *ref_base, const double *qry_base, const chess_pair *pairs,
                 int64_t n_pairs, const chess_params &p, double *ssim, double *sn, int32_t *status,
                 cudaStream_t stream) {
  auto kern = compare_r1_kernel<G, NW, WITH_SN>;
  const size_t smem = WITH_SN ? (size_t)NW * 7 * G * 32 * sizeof(double) : 0;
  static bool configured = false;  // static + dynamic shared memory can exceed the 48 KB default
  if (!configured && smem > 0) {
    CHESS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  long long grid = n_pairs;
  const long long cap = (long long)ctx->sm_count * 32;
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, 32 * NW, smem, stream>>>(ref_base, qry_base, pairs, (long long)n_pairs, p, ssim, sn,
                                                 status);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

template <int G, int NW>
int launch_r1(chess_ctx *ctx, const double *ref_base, const double *qry_base, const chess_pair *pairs,
              int64_t n_pairs, const chess_params &p, double *ssim, double *sn, int32_t *status,
              cudaStream_t stream) {
  if (p.compute_sn)
    return launch_r1_sn<G, NW, true>(ctx, ref_base, qry_base, pairs, n_pairs, p, ssim, sn, status, stream);
  return launch_r1_sn<G, NW, false>(ctx, ref_base, qry_base, pairs, n_pairs, p, ssim, sn, status, stream);
}

__global__ void band_index_kernel(const double *__restrict__ band, long long n_bins, int W, double dv,
                                  int *__restrict__ nd_lo, int *__restrict__ nd_hi) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_bins) return;
  const double *row = band + i * (2ll * W - 1) + (W - 1);  // row[k - i] = band(i, k)
  int lo = -1, hi = 0x7fffffff;
  for (int dist = 0; dist < W; ++dist) {
    if (i - dist < 0) break;
    if (row[-dist] != dv) { lo = (int)(i - dist); break; }
  }
  for (int dist = 0; dist < W; ++dist) {
    if (i + dist >= n_bins) break;
    if (row[dist] != dv) { hi = (int)(i + dist); break; }
  }
  nd_lo[i] = lo;
  nd_hi[i] = hi;
}

__global__ void band_extrema_kernel(const double *__restrict__ band, long long n_bins, int W,
                                    double *__restrict__ pmax, double *__restrict__ pmin) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_bins) return;
  const double *row = band + i * (2ll * W - 1) + (W - 1);  // row[d] = band(i, i+d)
  double mx = -CUDART_INF, mn = CUDART_INF;
  for (int d = 0; d < W; ++d) {
    const double v = row[d];  // cells beyond the last bin hold the fill value; they are never looked up
    mx = v > mx ? v : mx;
    mn = v < mn ? v : mn;
    pmax[i * W + d] = mx;
    pmin[i * W + d] = mn;
  }
}

size_t items_stride(int ncol_pad) { return (size_t)kItemHdr + 2 * (size_t)ncol_pad; }

}  // namespace

int chess_launch_compare_fast(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                              const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                              const chess_params &p, double *ssim, double *sn, int32_t *status,
                              cudaStream_t stream, bool *handled) {
  *handled = false;
  const bool symmetric = (p.flags & CHESS_FLAG_SYMMETRIC) != 0;
  const bool r1 = p.abs_window == 0 && p.rel_window == 1.0;
  const bool plain_default = p.default_value == 1.0 || p.default_value == 0.0;
  if (!symmetric || !r1 || !plain_default || max_n > 248) return CHESS_OK;
  int rc;
  if (max_n <= 93) rc = launch_r1<3, 4>(ctx, ref_base, qry_base, pairs, n_pairs, p, ssim, sn, status, stream);
  else if (max_n <= 124) rc = launch_r1<4, 4>(ctx, ref_base, qry_base, pairs, n_pairs, p, ssim, sn, status, stream);
  else if (max_n <= 155) rc = launch_r1<5, 4>(ctx, ref_base, qry_base, pairs, n_pairs, p, ssim, sn, status, stream);
  else if (max_n <= 186) rc = launch_r1<6, 4>(ctx, ref_base, qry_base, pairs, n_pairs, p, ssim, sn, status, stream);
  else if (max_n <= 217) rc = launch_r1<7, 4>(ctx, ref_base, qry_base, pairs, n_pairs, p, ssim, sn, status, stream);
  else rc = launch_r1<8, 4>(ctx, ref_base, qry_base, pairs, n_pairs, p, ssim, sn, status, stream);
  if (rc) return rc;
  *handled = true;
  if (!(p.flags & CHESS_FLAG_EQUAL_SIZES)) {
    // some pairs may need the resize path: let the generic kernel pick up the RETRY marks
    rc = chess_launch_compare_generic(ctx, ref_base, qry_base, pairs, n_pairs, max_n, p, ssim, sn, status,
                                      stream, /*retry_only=*/true);
  }
  return rc;
}

// ---- band-sample entry points -------------------------------------------------------------------
extern "C" int chess_band_index_build(chess_ctx *ctx, const double *band, int64_t n_bins, int32_t W,
                                      double default_value, int32_t *nd_lo, int32_t *nd_hi,
                                      void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (!band || n_bins <= 0 || W <= 0 || !nd_lo || !nd_hi) {
    ctx->err = "chess_band_index_build: bad arguments";
    return CHESS_ERR_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  band_index_kernel<<<(unsigned)((n_bins + 127) / 128), 128, 0, stream>>>(band, (long long)n_bins, W,
                                                                          default_value, nd_lo, nd_hi);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_band_extrema_build(chess_ctx *ctx, const double *band, int64_t n_bins, int32_t W,
                                        double *pmax, double *pmin, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (!band || n_bins <= 0 || W <= 0 || !pmax || !pmin) {
    ctx->err = "chess_band_extrema_build: bad arguments";
    return CHESS_ERR_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  band_extrema_kernel<<<(unsigned)((n_bins + 127) / 128), 128, 0, stream>>>(band, (long long)n_bins, W, pmax, pmin);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

namespace {
__global__ void __launch_bounds__(128)
band_prefix_kernel(const double *__restrict__ band, long long n_bins, int W, double *__restrict__ pre) {
  // one warp per band row: lanes take 32 consecutive entries at a time, an inclusive warp scan plus
  // the running carry gives the prefix (fixed order, so the array is reproducible)
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= n_bins) return;
  const double *src = band + row * (2LL * W - 1);
  double *dst = pre + row * (2LL * W);
  const int len = 2 * W - 1;
  double carry = 0.0;
  if (lane == 0) dst[0] = 0.0;
  for (int base = 0; base < len; base += 32) {
    const int k = base + lane;
    double v = k < len ? src[k] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    v += carry;
    if (k < len) dst[k + 1] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
}
}  // namespace

extern "C" int chess_band_prefix_build(chess_ctx *ctx, const double *band, int64_t n_bins, int32_t W,
                                       double *rowpre, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (!band || n_bins <= 0 || W <= 0 || !rowpre) {
    ctx->err = "chess_band_prefix_build: bad arguments";
    return CHESS_ERR_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  band_prefix_kernel<<<(unsigned)((n_bins + 3) / 4), 128, 0, stream>>>(band, (long long)n_bins, W, rowpre);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

// smallest strip width whose triangle split (compare_tri.cuh: tri_split) fits the warp, 0 if none
static int tri_strip_width(int n, int halo = 3) {
  static const int g_env = getenv("CHESS_TRI_G") ? atoi(getenv("CHESS_TRI_G")) : 3;  // experiments only
  static const int mono_env = getenv("CHESS_TRI_MONO") ? atoi(getenv("CHESS_TRI_MONO")) : 1;  // A/B runs
  const int g_min = g_env > halo ? g_env : halo;  // a window may reach into one neighbour lane only
  // small matrices: the narrowest strip with which all three bands fit one warp side by side (tri_rows.cuh:
  // tri_mode) -- one pass per pair instead of two items
  if (mono_env) {
    for (int G = g_min < 3 ? 3 : g_min; G <= 4; ++G) {
      int b1, b2, l1, l2;
      if (n <= 31 * G && tri_mode(n, G, b1, b2, l1, l2) == 1) return G;
    }
  }
  for (int G = g_min < 3 ? 3 : g_min; G <= 8; ++G) {
    if (n > 31 * G || n < 24) continue;
    const int x = (n - halo) / 3;
    const int b1 = n - 2 * x, b2 = n - x;
    const int l1 = (n - (b1 - halo) + G - 1) / G, l2 = (n - (b2 - halo) + G - 1) / G;
    if (l1 + l2 <= 30) return G;
  }
  return 0;
}

int chess_launch_compare_items(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                               const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                               const chess_params &p, double *ssim, double *sn, int32_t *status,
                               cudaStream_t stream, bool *handled) {
  *handled = false;
  const bool r1 = p.abs_window == 0 && p.rel_window == 1.0;
  // windowed regime: -a 3 / 5 / 7 (an even -a is lowered by one, values below 3 become 3; sim.py:360-365)
  int H = 0;
  if (p.abs_window > 0) {
    const int w = window_rule(0, p.abs_window, 1.0);
    H = (w - 1) / 2;
  }
  // windows 3 / 5 / 7 have a full-square and a triangle kernel, 9 / 11 the triangle kernel only
  // a sample without index arrays (dense matrices handed to chess_compare_batch) can still take the tuned
  // kernels when it needs none of them: no masks to predict or verify (mappability_cutoff <= 0, chess/sim.py:336)
  // and a fixed data range (chess_params.data_range) -- the dense one-vs-many batches of BASELINE configs[0]
  const bool no_masks = !(p.mappability_cutoff > 0);
  const bool have_index = ref->nd_lo && ref->nd_hi && qry->nd_lo && qry->nd_hi;
  const bool have_pm = ref->pmax && ref->pmin && qry->pmax && qry->pmin;
  const bool have_pre = ref->rowpre && qry->rowpre;
  const bool bare = no_masks && p.data_range > 0.0 && (p.flags & CHESS_FLAG_SYMMETRIC);
  const bool windowed = p.abs_window > 0 && H >= 1 && H <= 5 && (have_pm || bare);
  const bool plain_default = p.default_value == 1.0 || p.default_value == 0.0;
  // unequal-sized pairs would all bounce to the generic kernel: leave such batches to the other paths
  if (!(r1 || windowed) || !plain_default || max_n > 1024 ||
      !(p.flags & (CHESS_FLAG_EQUAL_SIZES | CHESS_FLAG_MOSTLY_EQUAL_SIZES)) || !(have_index || bare))
    return CHESS_OK;
  if (!bare && p.data_range > 0.0) return CHESS_OK;   // a fixed range on indexed samples: generic kernel, as before
  // column-strip width: the narrowest strip that holds the matrix in one block; wider matrices are
  // swept in blocks of 31*G - 6 columns, with the width that wastes the fewest lane-columns
  int G = 8;
  if (max_n <= 248) {
    G = max_n <= 93 ? 3 : max_n <= 124 ? 4 : max_n <= 155 ? 5 : max_n <= 186 ? 6 : max_n <= 217 ? 7 : 8;
  } else {
    long best = -1;
    for (int g = 5; g <= 8; ++g) {
      const int span = 31 * g - 6;
      const long cost = (long)((max_n + span - 1) / span) * g;
      if (best < 0 || cost <= best) { best = cost; G = g; }
    }
  }
  // -r 1 over the upper triangle when both samples carry their row prefix sums (kernel = 3 keeps the
  // full-square walk for A/B runs)
  const int tri_G = ((r1 || windowed) && p.kernel != 3 && (have_pre || bare))
                        ? tri_strip_width(max_n, H > 3 ? H : 3) : 0;
  const bool tri = tri_G != 0;
  if (tri) G = tri_G;
  // -r 1 on matrices wider than one warp: square blocks on and above the diagonal; strip width and block
  // count that cost the fewest lane-columns (blocks x rows per block x strip width)
  int blk_nb = 0;
  if (!tri && (r1 || windowed) && p.kernel != 3 && (have_pre || bare) && max_n > 93) {
    // blocks of `side` rows, a multiple of the strip width and at most 29 strips (one halo lane either side and
    // an idle lane for the neighbour exchange to wrap onto): cost = items x rows per item x strip width
    long best = -1;
    const int halo = H > 3 ? H : 3;
    for (int g = (H > 5 ? H : 5); g <= 8; ++g) {
      const int nb = (max_n + 29 * g - 1) / (29 * g);
      if (nb < 2 || nb > 7) continue;
      const int side = (((max_n + nb - 1) / nb + g - 1) / g) * g;
      const long cost = (long)(nb * (nb + 1) / 2) * (side + 2 * halo) * g;
      if (best < 0 || cost < best) { best = cost; G = g; blk_nb = nb; }
    }
  }
  const bool blk = blk_nb != 0;
  if (H > 3 && !tri && !blk) return CHESS_OK;  // windows 9 / 11 have no full-square kernel: generic
  if (bare && !tri && !blk) return CHESS_OK;   // the full-square work-item kernels need the index arrays
  const int ipp = tri ? 2 : blk ? blk_nb * (blk_nb + 1) / 2 : kBands;  // work items (records) per pair
  const int ncol_pad = ((max_n + 31) / 32) * 32;
  const size_t stride = (tri || blk) ? (size_t)16 : items_stride(ncol_pad);
  // pairs per launch: at most 16384 (size of the arrival-counter array; 64k-pair chunks measured slower -- their
  // band records no longer stay in L2) and at most ~256 MB of band records
  const int64_t chunk_max = 65536;  // arrival counters reserved in the workspace
  int64_t chunk = (int64_t)((size_t)(256u << 20) / ((size_t)ipp * stride * sizeof(double)));
  if (chunk > ((tri || blk) ? chunk_max : 16384)) chunk = (tri || blk) ? chunk_max : 16384;  // triangle records are 128 B
  if (chunk < 512) chunk = 512;
  const int64_t cap = n_pairs < chunk ? n_pairs : chunk;
  // workspace: [2 counters][2 retry flags + pad][done: chunk ints][scratch: cap * kBands * stride doubles];
  // the arrival counters sit at a fixed offset and size so that scratch records can never overlap them
  const size_t head = 64;
  const size_t done_bytes = (size_t)chunk_max * sizeof(int);
  const size_t need = head + done_bytes + (size_t)cap * ipp * stride * sizeof(double);
  if (ctx->d_items_bytes < need) {
    if (ctx->d_items) cudaFree(ctx->d_items);
    ctx->d_items = nullptr;
    ctx->d_items_bytes = 0;
    const size_t want = need + need / 4;
    if (cudaMalloc(&ctx->d_items, want) != cudaSuccess) {
      ctx->err = "allocation of the item workspace failed";
      return CHESS_ERR_NOMEM;
    }
    ctx->d_items_bytes = want;
    ctx->items_done_cap = 0;
  }
  char *base = reinterpret_cast<char *>(ctx->d_items);
  if (ctx->items_done_cap == 0) {  // fresh allocation: counters, flags and arrival counters start at zero
    CHESS_CUDA(ctx, cudaMemsetAsync(base, 0, head + done_bytes, stream));
    ctx->items_parity = 0;
    ctx->items_done_cap = (size_t)chunk_max;
  }
  unsigned long long *counters = reinterpret_cast<unsigned long long *>(base);
  int *flags = reinterpret_cast<int *>(base + 16);
  chess_item_args a;
  a.ref_band = ref->band; a.qry_band = qry->band;
  a.ref_lo = ref->nd_lo; a.ref_hi = ref->nd_hi; a.qry_lo = qry->nd_lo; a.qry_hi = qry->nd_hi;
  const bool have_extrema = ref->pmax && ref->pmin && qry->pmax && qry->pmin;
  a.ref_pmax = have_extrema ? ref->pmax : nullptr; a.ref_pmin = have_extrema ? ref->pmin : nullptr;
  a.qry_pmax = have_extrema ? qry->pmax : nullptr; a.qry_pmin = have_extrema ? qry->pmin : nullptr;
  a.ref_pre = ref->rowpre; a.qry_pre = qry->rowpre;
  a.p = p;
  a.ncol_pad = ncol_pad;
  a.max_n = max_n;
  a.blocks = blk_nb;
  a.done = reinterpret_cast<int *>(base + head);
  a.scratch = reinterpret_cast<double *>(base + head + done_bytes);
  for (int64_t off = 0; off < n_pairs; off += chunk) {
    const int64_t cnt = n_pairs - off < chunk ? n_pairs - off : chunk;
    const int par = ctx->items_parity;
    a.pairs = pairs + off; a.n_pairs = cnt;
    a.ssim = ssim + off; a.sn = sn + off; a.status = status + off;
    a.counter_cur = counters + par; a.counter_next = counters + (par ^ 1);
    a.retry_cur = flags + par; a.retry_next = flags + (par ^ 1);
    int rc;
    const int hh = r1 ? 0 : H;
    const bool multi = max_n > 31 * G;
    if (blk && r1) rc = chess_items_launch_blk(ctx, a, G, stream);
    else if (blk) rc = chess_items_launch_blkw(ctx, a, G, H, stream);
    else if (tri && r1) rc = chess_items_launch_tri(ctx, a, G, stream);
    else if (tri && H == 3) rc = chess_items_launch_tri7(ctx, a, G, stream);
    else if (tri) rc = chess_items_launch_triw(ctx, a, G, H, stream);
    else if (hh == 0) rc = chess_items_launch_h0(ctx, a, G, multi, stream);
    else if (hh == 1) rc = chess_items_launch_h1(ctx, a, G, multi, stream);
    else if (hh == 2) rc = chess_items_launch_h2(ctx, a, G, multi, stream);
    else rc = chess_items_launch_h3(ctx, a, G, multi, stream);
    if (rc) return rc;
    ctx->items_parity ^= 1;
    // pairs the item kernel could not finish (unequal sizes, coincidentally masked column): the
    // generic kernel picks them up; it returns at once when the flag is clear
    rc = chess_launch_compare_generic(ctx, ref->band, qry->band, pairs + off, cnt, max_n, p, ssim + off,
                                      sn + off, status + off, stream, /*retry_only=*/true, flags + par);
    if (rc) return rc;
  }
  *handled = true;
  return CHESS_OK;
}
