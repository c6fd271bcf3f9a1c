// K1 / K2: observed/expected on the device (replaces chess/oe.py:17-181) and
// the symmetric band store that replaces the dict-of-dict edges
// (chess/helpers.py:342-356) plus the per-cell fill of
// sub_matrix_from_edges_dict (chess/helpers.py:296-304).
//
// Expected values.  The reference accumulates, in file order and in Python
// floats, the marginal of every bin, the sum of every (chromosome, distance)
// diagonal and the inter-chromosomal sum (oe.py:98-117), then divides each
// diagonal by the number of bin pairs at that distance whose two bins both
// have a positive marginal (oe.py:17-75, 136-141).  Here every triple adds
// into 128-bit fixed-point accumulators with two 64-bit integer atomics.
// Integer addition commutes, so the sums are exact for the retained bits and
// independent of the order in which threads arrive: the expected values are
// bit-identical from run to run and on every GPU of a sharded job.  The unit
// is 2^(emax+1-94) where emax is the largest weight exponent: weights down to
// 2^-41 of the largest keep all 53 mantissa bits, and 2^32 triples cannot
// overflow.  The final int128 -> double conversion rounds once, so the result
// sits within 1 ulp of the exact sum (the reference's sequential sum carries
// ~nnz*eps of rounding noise itself).
#include <math_constants.h>

#include "common.cuh"

namespace {

struct Acc128 {
  unsigned long long lo;
  long long hi;
};

__device__ __forceinline__ void acc_add(Acc128 *a, __int128 v) {
  const unsigned long long lo = (unsigned long long)v;
  const long long hi = (long long)(v >> 64);
  const unsigned long long old = atomicAdd(&a->lo, lo);
  const unsigned long long carry = (old + lo) < old ? 1ull : 0ull;
  const unsigned long long add_hi = (unsigned long long)hi + carry;
  if (add_hi) atomicAdd(reinterpret_cast<unsigned long long *>(&a->hi), add_hi);
}

// weight -> fixed point with unit 2^(unit_exp)
__device__ __forceinline__ __int128 to_fixed(double w, int unit_exp) {
  if (w == 0.0) return 0;
  const long long bits = __double_as_longlong(w);
  const int e = (int)((bits >> 52) & 0x7ff);
  unsigned long long mant = (unsigned long long)bits & 0xfffffffffffffull;
  int ex;  // value = mant * 2^ex
  if (e == 0) ex = -1074;
  else { mant |= 1ull << 52; ex = e - 1075; }
  const int sh = ex - unit_exp;
  __int128 v;
  if (sh >= 0) v = (__int128)mant << sh;   // sh <= 41 by construction of unit_exp
  else if (sh > -64) v = (__int128)(mant >> (-sh));
  else v = 0;
  return bits < 0 ? -v : v;
}

__device__ __forceinline__ double from_fixed(const Acc128 &a, int unit_exp) {
  const __int128 v = ((__int128)a.hi << 64) | (__int128)a.lo;
  if (v == 0) return 0.0;
  const bool neg = v < 0;
  unsigned __int128 u = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
  // normalise to 64 significant bits, round to nearest even into 53
  int msb = 127;
  while (!((u >> msb) & 1)) --msb;
  int shift = msb - 63;
  unsigned long long top;
  bool sticky = false;
  if (shift > 0) {
    sticky = (u & ((((unsigned __int128)1) << shift) - 1)) != 0;
    top = (unsigned long long)(u >> shift);
  } else {
    top = (unsigned long long)(u << (-shift));
  }
  // top has its MSB set (bit 63); keep 53 bits
  unsigned long long keep = top >> 11;
  const unsigned long long rem = top & 0x7ff;
  if (rem > 0x400 || (rem == 0x400 && (sticky || (keep & 1)))) keep += 1;
  double r = ldexp((double)keep, shift + 11 + unit_exp);  // (double)keep exact (<= 2^53)
  return neg ? -r : r;
}

__global__ void max_exponent_kernel(const double *__restrict__ w, long long nnz, int *out_max_e,
                                    int *any_negative) {
  int e = -5000, neg = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = w[i];
    if (v != 0.0) {
      int ex;
      frexp(v, &ex);  // |v| in [2^(ex-1), 2^ex)
      e = max(e, ex - 1);
    }
    neg |= v < 0.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
    neg |= __shfl_xor_sync(0xffffffffu, neg, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out_max_e, e);
    if (neg) atomicOr(any_negative, 1);
  }
}

__global__ void accumulate_kernel(const int *__restrict__ src, const int *__restrict__ sink,
                                  const double *__restrict__ w, long long nnz,
                                  const int *__restrict__ bin_chrom,
                                  const long long *__restrict__ chrom_start, const int *max_e,
                                  Acc128 *marg, Acc128 *diag, Acc128 *inter,
                                  unsigned char *positive) {
  const int unit_exp = *max_e + 1 - 94;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x) {
    const int a = src[i], b = sink[i];
    const double v = w[i];
    const __int128 f = to_fixed(v, unit_exp);
    acc_add(&marg[a], f);
    acc_add(&marg[b], f);  // the diagonal adds twice, like oe.py:109-110
    if (v > 0.0) { positive[a] = 1; positive[b] = 1; }
    const int ca = bin_chrom[a];
    if (ca != bin_chrom[b]) acc_add(inter, f);
    else acc_add(&diag[chrom_start[ca] + (b - a)], f);
  }
}

__global__ void mappable_kernel(const Acc128 *marg, const unsigned char *positive,
                                const int *any_negative, long long n_bins, unsigned char *mappable) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_bins) return;
  unsigned char m;
  if (*any_negative == 0) m = positive[i];  // all weights >= 0: marginal > 0 <=> some weight > 0
  else {
    const Acc128 a = marg[i];
    m = (a.hi > 0 || (a.hi == 0 && a.lo != 0)) ? 1 : 0;
  }
  mappable[i] = m;
}

// possible[chrom_start[c] + d] = #{i : bins i and i+d of chromosome c both mappable}
__global__ void possible_kernel(const unsigned char *__restrict__ mappable,
                                const long long *__restrict__ chrom_start, int n_chrom,
                                long long *__restrict__ possible) {
  const int c = blockIdx.y;
  const long long lo = chrom_start[c], n = chrom_start[c + 1] - lo;
  for (long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x; d < n;
       d += (long long)gridDim.x * blockDim.x) {
    long long cnt = 0;
    const unsigned char *m = mappable + lo;
    for (long long i = 0; i + d < n; ++i) cnt += m[i] & m[i + d];
    possible[lo + d] = cnt;
  }
}

__global__ void finalize_kernel(const Acc128 *diag, const Acc128 *inter, const int *max_e,
                                const unsigned char *__restrict__ mappable,
                                const long long *__restrict__ possible,
                                const long long *__restrict__ chrom_start, int n_chrom,
                                long long n_bins, double *expected, double *inter_expected,
                                double *diag_sum) {
  const int unit_exp = *max_e + 1 - 94;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n_bins) {
    const long long cnt = possible[i];
    const double s = from_fixed(diag[i], unit_exp);
    expected[i] = cnt > 0 ? s / (double)cnt : 0.0;  // oe.py:136-141
    if (diag_sum) diag_sum[i] = s;
  }
  if (i == 0) {
    // inter_total = sum_{a<b} M_a * M_b over mappable-bin counts (oe.py:70-73)
    long long total = 0, seen = 0;
    for (int c = 0; c < n_chrom; ++c) {
      long long mc = 0;
      for (long long k = chrom_start[c]; k < chrom_start[c + 1]; ++k) mc += mappable[k];
      total += seen * mc;
      seen += mc;
    }
    *inter_expected = total == 0 ? 0.0 : from_fixed(*inter, unit_exp) / (double)total;  // oe.py:123
  }
}

__global__ void oe_values_kernel(const int *__restrict__ src, const int *__restrict__ sink,
                                 const double *__restrict__ w, long long nnz,
                                 const int *__restrict__ bin_chrom,
                                 const long long *__restrict__ chrom_start,
                                 const double *__restrict__ expected,
                                 const double *__restrict__ inter_expected, int apply_log2,
                                 double *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x) {
    const int a = src[i], b = sink[i];
    const int ca = bin_chrom[a];
    const double e = ca == bin_chrom[b] ? expected[chrom_start[ca] + (b - a)] : *inter_expected;
    double v = e == 0.0 ? 1.0 : __ddiv_rn(w[i], e);  // ZeroDivisionError -> 1, oe.py:158-161
    if (apply_log2) v = log2(v);
    out[i] = v;
  }
}

__global__ void band_fill_kernel(double *__restrict__ band, long long count, double fill) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // 16-byte stores where aligned
  double2 *b2 = reinterpret_cast<double2 *>(band);
  const long long pairs = count >> 1;
  const double2 f2 = make_double2(fill, fill);
  for (long long k = i; k < pairs; k += stride) b2[k] = f2;
  if (i == 0 && (count & 1)) band[count - 1] = fill;
}

__global__ void band_scatter_kernel(const int *__restrict__ src, const int *__restrict__ sink,
                                    const double *__restrict__ value, long long nnz, int W,
                                    double *__restrict__ band) {
  const long long pitch = 2ll * W - 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x) {
    const long long a = src[i], b = sink[i];
    const long long dist = b - a;
    if (dist < W && dist > -(long long)W) {
      const double v = value[i];
      band[a * pitch + dist + (W - 1)] = v;
      band[b * pitch - dist + (W - 1)] = v;
    }
  }
}

// The same for a band that holds bins [bin_lo, bin_lo + n_range) only (per-GPU placement of a pair-list shard):
// a contact is written where both of its bins lie in the range, at its position relative to bin_lo.
__global__ void band_scatter_range_kernel(const int *__restrict__ src, const int *__restrict__ sink,
                                          const double *__restrict__ value, long long nnz, int W, long long bin_lo,
                                          long long n_range, double *__restrict__ band) {
  const long long pitch = 2ll * W - 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x) {
    const long long a = (long long)src[i] - bin_lo, b = (long long)sink[i] - bin_lo;
    const long long dist = b - a;
    if (a >= 0 && b >= 0 && a < n_range && b < n_range && dist < W && dist > -(long long)W) {
      const double v = value[i];
      band[a * pitch + dist + (W - 1)] = v;
      band[b * pitch - dist + (W - 1)] = v;
    }
  }
}

// Facts about a COO list that decide how it may be used: smallest / largest index, entries with src > sink, and
// entries that do not come strictly after their predecessor in (src, sink) order (0 = sorted and duplicate-free).
__global__ void coo_check_kernel(const int *__restrict__ src, const int *__restrict__ sink, long long nnz,
                                 long long *__restrict__ out /* [4]: min, max, n_noncanonical, n_unordered */) {
  long long lo = 0x7fffffffffffffffll, hi = -1, nc = 0, un = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x) {
    const int a = src[i], b = sink[i];
    lo = min(lo, (long long)min(a, b));
    hi = max(hi, (long long)max(a, b));
    nc += a > b;
    if (i > 0) {
      const int pa = src[i - 1], pb = sink[i - 1];
      un += (pa > a) || (pa == a && pb >= b);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    nc += __shfl_xor_sync(0xffffffffu, nc, o);
    un += __shfl_xor_sync(0xffffffffu, un, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(out, lo);
    atomicMax(out + 1, hi);
    if (nc) atomicAdd(reinterpret_cast<unsigned long long *>(out + 2), (unsigned long long)nc);
    if (un) atomicAdd(reinterpret_cast<unsigned long long *>(out + 3), (unsigned long long)un);
  }
}

// out[0] |= 1 if any of `count` dense n x n matrices packed back to back differs from its transpose (bit
// pattern of the doubles; NaNs therefore count as equal to themselves)
__global__ void symmetric_check_kernel(const double *__restrict__ base, int n, long long count, int *out) {
  const long long per = (long long)n * n;
  const long long total = per * count;
  int bad = 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / per, r = t - k * per;
    const int i = (int)(r / n), j = (int)(r - (long long)i * n);
    if (j > i) {
      const double a = base[t], b = base[k * per + (long long)j * n + i];
      bad |= __double_as_longlong(a) != __double_as_longlong(b);
    }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(out, 1);
}

__global__ void coo_check_init_kernel(long long *out) {
  out[0] = 0x7fffffffffffffffll; out[1] = -1; out[2] = 0; out[3] = 0;
}

__global__ void gather_kernel(const double *__restrict__ base, long long off, int pitch, int n,
                              double *__restrict__ out) {
  const long long total = (long long)n * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n), c = (int)(i % n);
    out[i] = base[off + (long long)r * pitch + c];
  }
}

inline unsigned grid_for(long long n, int block, int sms) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)sms * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

extern "C" size_t chess_expected_workspace_bytes(int64_t n_bins) {
  // marginals + diagonals (Acc128 each) + inter + positive flags + 2 ints
  return (size_t)(2 * n_bins + 1) * sizeof(Acc128) + (size_t)n_bins + 64;
}

extern "C" int chess_expected_from_coo(chess_ctx *ctx, const int32_t *src, const int32_t *sink,
                                       const double *weight, int64_t nnz, const int32_t *bin_chrom,
                                       const int64_t *chrom_start, int32_t n_chrom, int64_t n_bins,
                                       double *expected, double *inter_expected, uint8_t *mappable,
                                       int64_t *possible, double *diag_sum, void *workspace,
                                       size_t workspace_bytes, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_bins <= 0 || n_chrom <= 0 || nnz < 0 || !workspace ||
      workspace_bytes < chess_expected_workspace_bytes(n_bins)) {
    ctx->err = "chess_expected_from_coo: bad arguments or workspace too small";
    return CHESS_ERR_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  Acc128 *marg = reinterpret_cast<Acc128 *>(workspace);
  Acc128 *diag = marg + n_bins;
  Acc128 *inter = diag + n_bins;
  unsigned char *positive = reinterpret_cast<unsigned char *>(inter + 1);
  int *ints = reinterpret_cast<int *>(positive + ((n_bins + 15) & ~15ll));
  CHESS_CUDA(ctx, cudaMemsetAsync(workspace, 0, chess_expected_workspace_bytes(n_bins), stream));
  const int init[2] = {-5000, 0};
  CHESS_CUDA(ctx, cudaMemcpyAsync(ints, init, sizeof(init), cudaMemcpyHostToDevice, stream));
  const int B = 256;
  if (nnz > 0) {
    max_exponent_kernel<<<grid_for(nnz, B, ctx->sm_count), B, 0, stream>>>(weight, nnz, ints, ints + 1);
    accumulate_kernel<<<grid_for(nnz, B, ctx->sm_count), B, 0, stream>>>(
        src, sink, weight, nnz, bin_chrom, (const long long *)chrom_start, ints, marg, diag, inter, positive);
    ctx->launches += 2;
  }
  mappable_kernel<<<(unsigned)((n_bins + B - 1) / B), B, 0, stream>>>(marg, positive, ints + 1, n_bins, mappable);
  dim3 pg(64, (unsigned)n_chrom);
  possible_kernel<<<pg, 128, 0, stream>>>(mappable, (const long long *)chrom_start, n_chrom,
                                          (long long *)possible);
  finalize_kernel<<<(unsigned)((n_bins + B - 1) / B), B, 0, stream>>>(
      diag, inter, ints, mappable, (const long long *)possible, (const long long *)chrom_start, n_chrom,
      n_bins, expected, inter_expected, diag_sum);
  ctx->launches += 3;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_possible_contacts(chess_ctx *ctx, const uint8_t *mappable,
                                       const int64_t *chrom_start, int32_t n_chrom, int64_t n_bins,
                                       int64_t *possible, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_bins <= 0 || n_chrom <= 0) { ctx->err = "chess_possible_contacts: bad arguments"; return CHESS_ERR_ARG; }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  dim3 pg(64, (unsigned)n_chrom);
  possible_kernel<<<pg, 128, 0, stream>>>(mappable, (const long long *)chrom_start, n_chrom,
                                          (long long *)possible);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_oe_values(chess_ctx *ctx, const int32_t *src, const int32_t *sink,
                               const double *weight, int64_t nnz, const int32_t *bin_chrom,
                               const int64_t *chrom_start, const double *expected,
                               const double *inter_expected, int32_t apply_log2, double *out,
                               void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (nnz <= 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  oe_values_kernel<<<grid_for(nnz, 256, ctx->sm_count), 256, 0, stream>>>(
      src, sink, weight, nnz, bin_chrom, (const long long *)chrom_start, expected, inter_expected,
      apply_log2, out);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_band_build(chess_ctx *ctx, const int32_t *src, const int32_t *sink,
                                const double *value, int64_t nnz, int64_t n_bins, int32_t W,
                                double fill, double *band, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_bins <= 0 || W <= 0 || !band) {
    ctx->err = "chess_band_build: bad arguments";
    return CHESS_ERR_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long count = (long long)n_bins * (2ll * W - 1);
  band_fill_kernel<<<grid_for(count / 2 + 1, 256, ctx->sm_count), 256, 0, stream>>>(band, count, fill);
  ctx->launches += 1;
  if (nnz > 0) {
    band_scatter_kernel<<<grid_for(nnz, 256, ctx->sm_count), 256, 0, stream>>>(src, sink, value, nnz, W, band);
    ctx->launches += 1;
  }
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_band_build_range(chess_ctx *ctx, const int32_t *src, const int32_t *sink,
                                      const double *value, int64_t nnz, int64_t bin_lo, int64_t n_range, int32_t W,
                                      double fill, double *band, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_range <= 0 || W <= 0 || !band || bin_lo < 0) {
    ctx->err = "chess_band_build_range: bad arguments";
    return CHESS_ERR_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long count = (long long)n_range * (2ll * W - 1);
  band_fill_kernel<<<grid_for(count / 2 + 1, 256, ctx->sm_count), 256, 0, stream>>>(band, count, fill);
  ctx->launches += 1;
  if (nnz > 0) {
    band_scatter_range_kernel<<<grid_for(nnz, 256, ctx->sm_count), 256, 0, stream>>>(src, sink, value, nnz, W, bin_lo,
                                                                                     n_range, band);
    ctx->launches += 1;
  }
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_symmetric_check(chess_ctx *ctx, const double *base, int32_t n, int64_t count, int32_t *out,
                                     void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (!base || !out || n <= 0 || count < 0) { ctx->err = "chess_symmetric_check: bad arguments"; return CHESS_ERR_ARG; }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  CHESS_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(int32_t), stream));
  if (count > 0) {
    symmetric_check_kernel<<<grid_for((long long)n * n * count, 256, ctx->sm_count), 256, 0, stream>>>(base, n, count,
                                                                                                    out);
    ctx->launches += 1;
  }
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_coo_check(chess_ctx *ctx, const int32_t *src, const int32_t *sink, int64_t nnz, int64_t *out4,
                               void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (!out4 || nnz < 0 || (nnz > 0 && (!src || !sink))) { ctx->err = "chess_coo_check: bad arguments"; return CHESS_ERR_ARG; }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  coo_check_init_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<long long *>(out4));
  ctx->launches += 1;
  if (nnz > 0) {
    coo_check_kernel<<<grid_for(nnz, 256, ctx->sm_count), 256, 0, stream>>>(src, sink, nnz,
                                                                            reinterpret_cast<long long *>(out4));
    ctx->launches += 1;
  }
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_gather_submatrix(chess_ctx *ctx, const double *base, int64_t off, int32_t pitch,
                                      int32_t n, double *out, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n <= 0 || !base || !out) { ctx->err = "chess_gather_submatrix: bad arguments"; return CHESS_ERR_ARG; }
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  gather_kernel<<<grid_for((long long)n * n, 256, ctx->sm_count), 256, 0, stream>>>(base, off, pitch, n, out);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}
