// K9: first stage of `chess extract` (chess/get_structures.py:99-121) on band-resident samples.
// Per pair: both submatrices are gathered with default 0, OR = log(query / reference) with NaN and
// +-inf replaced by 0, std = np.std(OR) (population, two passes), and the two thresholded maps
//     positive = where(OR >  0.5 std, OR, 0)      negative = |where(OR < -0.5 std, OR, 0)|
// are written together with their means (the reference feeds them to denoise_bilateral as
// sigma_color).  The image-processing stages after this one (bilateral / median / Otsu / labelling)
// are third-party skimage / scipy code and stay on the host.
#include <math_constants.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ double block_sum256(double v, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < 8 ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[8] = t;
  }
  __syncthreads();
  return red[8];
}

__global__ void __launch_bounds__(256)
log_ratio_kernel(const double *__restrict__ ref_base, const double *__restrict__ qry_base,
                 const chess_pair *__restrict__ pairs, long long n_pairs, long long tile_stride,
                 double *__restrict__ positive, double *__restrict__ negative, double *__restrict__ stats,
                 int *__restrict__ status) {
  __shared__ double red[9];
  for (long long pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
    const chess_pair d = pairs[pi];
    double *pos = positive + pi * tile_stride, *neg = negative + pi * tile_stride;
    double *st = stats + pi * 4;
    if (d.ref_n <= 0 || d.qry_n <= 0 || d.ref_n != d.qry_n) {
      // region not found (ValueError in sub_matrix_from_edges_dict) or shapes that numpy cannot divide
      if (threadIdx.x == 0) {
        st[0] = st[1] = st[2] = CUDART_NAN; st[3] = 0.0;
        status[pi] = d.ref_n <= 0 || d.qry_n <= 0 ? CHESS_PAIR_NOT_FOUND : CHESS_PAIR_RETRY;
      }
      continue;
    }
    const int n = d.ref_n;
    const long long nn = (long long)n * n;
    const double *R = ref_base + d.ref_off, *Q = qry_base + d.qry_off;
    double sum = 0.0;
    for (long long e = threadIdx.x; e < nn; e += 256) {
      const int r = (int)(e / n), c = (int)(e - (long long)r * n);
      const double x = R[(size_t)r * d.ref_pitch + c], y = Q[(size_t)r * d.qry_pitch + c];
      double v = log(y / x);                      // np.log(np.divide(query, reference))
      if (!(fabs(v) <= 1.7976931348623157e308)) v = 0.0;   // NaN, +inf, -inf -> 0
      pos[e] = v;
      sum += v;
    }
    const double mean = block_sum256(sum, red) / (double)nn;
    double ssq = 0.0;
    for (long long e = threadIdx.x; e < nn; e += 256) { const double dv = pos[e] - mean; ssq = fma(dv, dv, ssq); }
    const double sd = sqrt(block_sum256(ssq, red) / (double)nn);
    const double thr = 0.5 * sd;
    double sp = 0.0, sg = 0.0;
    for (long long e = threadIdx.x; e < nn; e += 256) {
      const double v = pos[e];
      const double p = v > thr ? v : 0.0;
      const double g = v < -thr ? -v : 0.0;       // np.abs(np.where(or < -0.5 std, or, 0))
      pos[e] = p;
      neg[e] = g;
      sp += p; sg += g;
    }
    sp = block_sum256(sp, red);
    sg = block_sum256(sg, red);
    if (threadIdx.x == 0) {
      st[0] = sd; st[1] = sp / (double)nn; st[2] = sg / (double)nn; st[3] = (double)n;
      status[pi] = CHESS_PAIR_OK;
    }
  }
}

}  // namespace

extern "C" int chess_log_ratio_batch(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                                     const chess_pair *pairs, int64_t n_pairs, int32_t max_n, double *positive,
                                     double *negative, double *stats, int32_t *status, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_pairs < 0 || max_n <= 0 ||
      (n_pairs > 0 && (!ref_base || !qry_base || !pairs || !positive || !negative || !stats || !status))) {
    ctx->err = "chess_log_ratio_batch: bad arguments";
    return CHESS_ERR_ARG;
  }
  if (n_pairs == 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  long long grid = n_pairs < (long long)ctx->sm_count * 8 ? n_pairs : (long long)ctx->sm_count * 8;
  log_ratio_kernel<<<(unsigned)grid, 256, 0, stream>>>(ref_base, qry_base, pairs, (long long)n_pairs,
                                                      (long long)max_n * max_n, positive, negative, stats, status);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}
