// Background statistics on the device (bin/chess:259-332): pair descriptors for the
// (pair x background region) grid are written by a kernel, compared by the band kernels and
// reduced to z_bg / p_bg per pair without leaving HBM.
#include <math_constants.h>

#include "common.cuh"

namespace {

// one descriptor per (pair, background region); pairs whose own ssim is NaN are not compared at all
// (ref_n = 0 -> status NOT_FOUND, NaN result; their statistics are NaN by bin/chess:319-321)
__global__ void __launch_bounds__(256)
bg_table_kernel(const int64_t *__restrict__ ref_first, const int32_t *__restrict__ ref_n,
                const int32_t *__restrict__ ref_strand, const double *__restrict__ pair_ssim,
                long long pair0, long long n_chunk_pairs, const int64_t *__restrict__ bg_first,
                const int32_t *__restrict__ bg_n, const int32_t *__restrict__ bg_strand, long long n_bg,
                long long Wr, long long Wq, chess_pair *__restrict__ out) {
  const long long total = n_chunk_pairs * n_bg;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / n_bg, j = t - k * n_bg;
    const long long pk = pair0 + k;
    const double s = pair_ssim[pk];
    int rn = ref_n[pk], qn = bg_n[j];
    if (!(s == s)) rn = 0;
    chess_pair d;
    d.ref_off = rn > 0 ? ref_first[pk] * (2 * Wr - 1) + Wr - 1 : 0;
    d.qry_off = qn > 0 ? bg_first[j] * (2 * Wq - 1) + Wq - 1 : 0;
    d.ref_pitch = (int32_t)(2 * Wr - 2);
    d.qry_pitch = (int32_t)(2 * Wq - 2);
    d.ref_n = rn > 0 ? rn : 0;
    d.qry_n = qn > 0 ? qn : 0;
    d.flip = ref_strand[pk] != bg_strand[j] ? 1 : 0;
    d.reserved = 0;
    out[t] = d;
  }
}

// Background regions generated from the query bins (bin/chess:272-298): every bin start x {+, -} with the size of
// the pair's query region, kept where it ends inside its chromosome (and, with --limit-background, lies on the
// pair's chromosome).  One thread per bin decides and looks the region up -- the bins that overlap
// [start - 1, start + size) in the half-open sense of chess/helpers.py:212-242, by two bisections inside the
// bin's chromosome (bins sorted and contiguous per chromosome) -- then a single-CTA scan packs the kept regions
// in bin order, '+' before '-', exactly the order of the reference's list.
__global__ void __launch_bounds__(256)
bg_candidates_kernel(const int64_t *__restrict__ bin_begin, const int64_t *__restrict__ bin_end,
                     const int32_t *__restrict__ bin_chrom, const int64_t *__restrict__ chrom_first,
                     const int64_t *__restrict__ chrom_end_bp, long long n_bins, long long size, int limit_chrom,
                     int64_t *__restrict__ first, int32_t *__restrict__ count, unsigned char *__restrict__ keep) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_bins) return;
  const int c = bin_chrom[i];
  const long long start = bin_begin[i] + 1;           // region = [start, start + size], looked up as [start-1, end)
  const bool ok = (limit_chrom < 0 || c == limit_chrom) && start + size <= chrom_end_bp[c];
  keep[i] = ok ? 1 : 0;
  if (!ok) return;
  const long long qb = start - 1, qe = start + size;
  long long lo = chrom_first[c], hi = chrom_first[c + 1];
  // first bin whose end > qb
  long long a = lo, b = hi;
  while (a < b) { const long long m = (a + b) >> 1; if (bin_end[m] > qb) b = m; else a = m + 1; }
  const long long f = a;
  // last bin whose begin < qe
  a = lo; b = hi;
  while (a < b) { const long long m = (a + b) >> 1; if (bin_begin[m] < qe) a = m + 1; else b = m; }
  const long long l = a - 1;
  first[i] = f <= l ? f : -1;
  count[i] = f <= l ? (int32_t)(l - f + 1) : 0;
}

__global__ void __launch_bounds__(1024)
bg_pack_kernel(const int64_t *__restrict__ first, const int32_t *__restrict__ count,
               const unsigned char *__restrict__ keep, long long n_bins, int64_t *__restrict__ out_first,
               int32_t *__restrict__ out_n, int32_t *__restrict__ out_strand, long long *__restrict__ n_out) {
  __shared__ long long sh[1024];
  const long long per = (n_bins + 1023) / 1024;
  const long long lo = per * threadIdx.x, hi = lo + per < n_bins ? lo + per : n_bins;
  long long s = 0;
  for (long long i = lo; i < hi; ++i) s += keep[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int t = 0; t < 1024; ++t) { const long long v = sh[t]; sh[t] = run; run += v; }
    *n_out = 2 * run;
  }
  __syncthreads();
  long long run = sh[threadIdx.x];
  for (long long i = lo; i < hi; ++i) {
    if (keep[i]) {
      out_first[2 * run] = first[i]; out_first[2 * run + 1] = first[i];
      out_n[2 * run] = count[i]; out_n[2 * run + 1] = count[i];
      out_strand[2 * run] = 0; out_strand[2 * run + 1] = 1;      // '+', then '-'
      run += 1;
    }
  }
}

__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// one CTA per pair; two passes like numpy's nanmean / nanstd (mean first, then squared deviations),
// fixed reduction order
__global__ void __launch_bounds__(256)
bg_stats_kernel(const double *__restrict__ vals, const int32_t *__restrict__ status, long long n_bg,
                const double *__restrict__ pair_ssim, long long pair0, const int32_t *__restrict__ ref_strand,
                const int64_t *__restrict__ own_first, const int32_t *__restrict__ own_n,
                const int32_t *__restrict__ own_strand, const int64_t *__restrict__ bg_first,
                const int32_t *__restrict__ bg_n, const int32_t *__restrict__ bg_strand,
                double *__restrict__ z_bg, double *__restrict__ p_bg, int32_t *__restrict__ n_exceeds) {
  __shared__ double red[33];
  const long long k = blockIdx.x;
  const double s = pair_ssim[pair0 + k];
  if (!(s == s)) {
    if (threadIdx.x == 0) { z_bg[pair0 + k] = CUDART_NAN; p_bg[pair0 + k] = CUDART_NAN; }
    return;
  }
  const double *v = vals + k * n_bg;
  const int32_t *st = status + k * n_bg;
  // A background region that IS the pair's own query region (same bins, same orientation) repeats
  // the pair's own comparison: the reference gets the bit-identical value there, and `ssim <= bg`
  // counts it.  Substitute the pair's value so that the tie does not depend on which kernel
  // (strip width, SN variant) produced either number.
  long long of = -1;
  int on = 0, oflip = 0;
  if (own_first != nullptr) {
    of = own_first[pair0 + k]; on = own_n[pair0 + k];
    oflip = ref_strand[pair0 + k] != own_strand[pair0 + k];
  }
  const int rs = ref_strand[pair0 + k];
  auto value = [&](long long j) {
    double x = v[j];
    if (on > 0 && bg_n[j] == on && bg_first[j] == of && (rs != bg_strand[j]) == (oflip != 0)) x = s;
    return x;
  };
  double cnt = 0.0, sum = 0.0, ge = 0.0;
  int exceeds = 0;
  for (long long j = threadIdx.x; j < n_bg; j += blockDim.x) {
    const double x = value(j);
    if (x == x) { cnt += 1.0; sum += x; }
    if (s <= x) ge += 1.0;
    if (st[j] == CHESS_PAIR_WINDOW_EXCEEDS) ++exceeds;
  }
  cnt = block_sum(cnt, red);
  sum = block_sum(sum, red);
  ge = block_sum(ge, red);
  if (n_exceeds != nullptr && exceeds > 0) atomicAdd(n_exceeds, exceeds);
  const double mean = sum / cnt;  // 0/0 -> NaN like nanmean of an all-NaN slice
  double ssq = 0.0;
  for (long long j = threadIdx.x; j < n_bg; j += blockDim.x) {
    const double x = value(j);
    if (x == x) { const double d = x - mean; ssq = fma(d, d, ssq); }
  }
  ssq = block_sum(ssq, red);
  if (threadIdx.x == 0) {
    const double sd = sqrt(ssq / cnt);
    z_bg[pair0 + k] = (s - mean) / sd;
    p_bg[pair0 + k] = (1.0 + ge) / (double)n_bg;
  }
}

}  // namespace

extern "C" int chess_background_regions(chess_ctx *ctx, const int64_t *bin_begin, const int64_t *bin_end,
                                        const int32_t *bin_chrom, const int64_t *chrom_first,
                                        const int64_t *chrom_end_bp, int64_t n_bins, int64_t size,
                                        int32_t limit_chrom, int64_t *bg_first, int32_t *bg_n, int32_t *bg_strand,
                                        int64_t *n_regions, void *workspace, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_bins < 0 || (n_bins > 0 && (!bin_begin || !bin_end || !bin_chrom || !chrom_first || !chrom_end_bp ||
                                    !bg_first || !bg_n || !bg_strand || !workspace)) || !n_regions) {
    ctx->err = "chess_background_regions: bad arguments";
    return CHESS_ERR_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  // workspace: first int64[n_bins] | count int32[n_bins] | keep u8[n_bins]  (13 bytes per bin)
  int64_t *first = reinterpret_cast<int64_t *>(workspace);
  int32_t *count = reinterpret_cast<int32_t *>(first + n_bins);
  unsigned char *keep = reinterpret_cast<unsigned char *>(count + n_bins);
  if (n_bins > 0) {
    bg_candidates_kernel<<<(unsigned)((n_bins + 255) / 256), 256, 0, stream>>>(
        bin_begin, bin_end, bin_chrom, chrom_first, chrom_end_bp, (long long)n_bins, (long long)size, limit_chrom,
        first, count, keep);
    ctx->launches += 1;
  }
  bg_pack_kernel<<<1, 1024, 0, stream>>>(first, count, keep, (long long)n_bins, bg_first, bg_n, bg_strand,
                                         reinterpret_cast<long long *>(n_regions));
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_background_batch(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                                      const int64_t *ref_first, const int32_t *ref_n,
                                      const int32_t *ref_strand, const int64_t *own_first,
                                      const int32_t *own_n, const int32_t *own_strand,
                                      const double *pair_ssim, int64_t n_pairs, const int64_t *bg_first, const int32_t *bg_n, const int32_t *bg_strand,
                                      int64_t n_bg, int32_t max_n, const chess_params *params, double *z_bg,
                                      double *p_bg, int32_t *n_exceeds, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (!params) { ctx->err = "params is NULL"; return CHESS_ERR_ARG; }
  if (!ref || !qry || !ref->band || !qry->band || n_pairs < 0 || n_bg < 0 ||
      (n_pairs > 0 && (!ref_first || !ref_n || !ref_strand || !pair_ssim || !z_bg || !p_bg)) ||
      (n_bg > 0 && (!bg_first || !bg_n || !bg_strand))) {
    ctx->err = "chess_background_batch: NULL buffer";
    return CHESS_ERR_ARG;
  }
  if (own_first != nullptr && (!own_n || !own_strand)) {
    ctx->err = "chess_background_batch: own_first needs own_n and own_strand";
    return CHESS_ERR_ARG;
  }
  if (n_pairs == 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_exceeds) CHESS_CUDA(ctx, cudaMemsetAsync(n_exceeds, 0, sizeof(int32_t), stream));
  chess_params p = *params;
  p.compute_sn = 0;  // bin/chess:306
  // chunk: whole pairs, about two million comparisons (descriptors 40 B + results 20 B each)
  const int64_t target = 2 << 20;
  int64_t ppc = n_bg > 0 ? target / n_bg : n_pairs;
  if (ppc < 1) ppc = 1;
  if (ppc > n_pairs) ppc = n_pairs;
  const size_t per = sizeof(chess_pair) + 2 * sizeof(double) + sizeof(int32_t);
  const size_t need = (size_t)ppc * (size_t)(n_bg > 0 ? n_bg : 1) * per + 256;
  if (ctx->d_bg_bytes < need) {
    if (ctx->d_bg) cudaFree(ctx->d_bg);
    ctx->d_bg = nullptr;
    ctx->d_bg_bytes = 0;
    if (cudaMalloc(&ctx->d_bg, need) != cudaSuccess) {
      ctx->err = "allocation of the background workspace failed";
      return CHESS_ERR_NOMEM;
    }
    ctx->d_bg_bytes = need;
  }
  const size_t cap = (size_t)ppc * (size_t)(n_bg > 0 ? n_bg : 1);
  chess_pair *table = reinterpret_cast<chess_pair *>(ctx->d_bg);
  double *vals = reinterpret_cast<double *>(table + cap);
  double *sn = vals + cap;
  int32_t *status = reinterpret_cast<int32_t *>(sn + cap);
  for (int64_t pair0 = 0; pair0 < n_pairs; pair0 += ppc) {
    const int64_t cnt = n_pairs - pair0 < ppc ? n_pairs - pair0 : ppc;
    const int64_t total = cnt * n_bg;
    if (total > 0) {
      long long grid = (total + 255) / 256;
      if (grid > 148 * 16) grid = 148 * 16;
      bg_table_kernel<<<(unsigned)grid, 256, 0, stream>>>(ref_first, ref_n, ref_strand, pair_ssim, pair0, cnt,
                                                          bg_first, bg_n, bg_strand, n_bg, ref->W, qry->W, table);
      ctx->launches += 1;
      CHESS_CUDA(ctx, cudaGetLastError());
      int rc = chess_compare_band_batch(ctx, ref, qry, table, total, max_n, &p, vals, sn, status, stream_);
      if (rc) return rc;
    }
    bg_stats_kernel<<<(unsigned)cnt, 256, 0, stream>>>(vals, status, n_bg, pair_ssim, pair0, ref_strand, own_first,
                                                       own_n, own_strand, bg_first, bg_n, bg_strand, z_bg, p_bg,
                                                       n_exceeds);
    ctx->launches += 1;
    CHESS_CUDA(ctx, cudaGetLastError());
  }
  return CHESS_OK;
}
