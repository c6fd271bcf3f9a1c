// C-ABI entry points that are not tied to one kernel family: context life
// cycle, kernel dispatch for the compare path, host-buffer variants.
#include <math_constants.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {
thread_local std::string g_last_error;

int grow(chess_ctx *ctx, void **ptr, size_t *have, size_t need, bool pinned) {
  if (*have >= need) return CHESS_OK;
  if (*ptr) {
    if (pinned) cudaFreeHost(*ptr);
    else cudaFree(*ptr);
    *ptr = nullptr;
    *have = 0;
  }
  size_t want = need + need / 2 + 4096;
  cudaError_t e = pinned ? cudaMallocHost(ptr, want) : cudaMalloc(ptr, want);
  if (e != cudaSuccess) {
    ctx->err = std::string("allocation of staging buffer failed: ") + cudaGetErrorString(e);
    return CHESS_ERR_NOMEM;
  }
  *have = want;
  return CHESS_OK;
}
}  // namespace

extern "C" int chess_abi_version(void) { return CHESS_B200_ABI_VERSION; }

extern "C" int chess_ctx_create(int device, chess_ctx **out) {
  if (!out) return CHESS_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_last_error = std::string("no CUDA device available: ") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                   " (chess_b200 has no CPU fallback)";
    return CHESS_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_last_error = "chess_ctx_create: device index out of range";
    return CHESS_ERR_ARG;
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_last_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
    return CHESS_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_last_error = std::string("chess_b200 is built for sm_100a only; device is ") + prop.name;
    return CHESS_ERR_UNSUPPORTED;
  }
  chess_ctx *ctx = new chess_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  *out = ctx;
  return CHESS_OK;
}

extern "C" void chess_ctx_destroy(chess_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->d_pairs) cudaFree(ctx->d_pairs);
  if (ctx->d_results) cudaFree(ctx->d_results);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->d_items) cudaFree(ctx->d_items);
  if (ctx->d_bg) cudaFree(ctx->d_bg);
  if (ctx->d_ingest) cudaFree(ctx->d_ingest);
  delete ctx;
}

extern "C" const char *chess_last_error(const chess_ctx *ctx) {
  return ctx ? ctx->err.c_str() : g_last_error.c_str();
}

extern "C" int64_t chess_launch_count(const chess_ctx *ctx) { return ctx ? ctx->launches : 0; }

static int validate(chess_ctx *ctx, const chess_params *p) {
  if (!p) { ctx->err = "params is NULL"; return CHESS_ERR_ARG; }
  if (p->abs_window < 0) { ctx->err = "abs_window must be >= 0"; return CHESS_ERR_ARG; }
  if (!(p->data_range >= 0.0)) { ctx->err = "data_range must be >= 0 (0 = from the data)"; return CHESS_ERR_ARG; }
  return CHESS_OK;
}

extern "C" int chess_compare_batch(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                                   const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                                   const chess_params *params, double *ssim, double *sn,
                                   int32_t *status, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  int rc = validate(ctx, params);
  if (rc) return rc;
  if (n_pairs < 0 || (n_pairs > 0 && (!ref_base || !qry_base || !pairs || !ssim || !sn || !status))) {
    ctx->err = "chess_compare_batch: NULL buffer";
    return CHESS_ERR_ARG;
  }
  if (n_pairs == 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (params->kernel != 1 && params->data_range == 0.0) {   // a fixed data range is the generic kernel's ...
    bool handled = false;
    rc = chess_launch_compare_fast(ctx, ref_base, qry_base, pairs, n_pairs, max_n, *params, ssim, sn,
                                   status, stream, &handled);
    if (rc) return rc;
    if (handled) return CHESS_OK;
  }
  if (params->kernel == 0 && params->data_range > 0.0 && (params->flags & CHESS_FLAG_SYMMETRIC) &&
      (params->flags & CHESS_FLAG_EQUAL_SIZES) && !(params->mappability_cutoff > 0)) {
    // ... except for symmetric, equal-sized matrices without masks (the dense one-vs-many batches of
    // utils/run_synthetic_test.py): they are views like any band view, and the triangle / block kernels need
    // nothing of a sample but its base pointer then
    chess_sample rs = {}, qs = {};
    rs.band = ref_base; qs.band = qry_base;
    bool handled = false;
    rc = chess_launch_compare_items(ctx, &rs, &qs, pairs, n_pairs, max_n, *params, ssim, sn, status, stream,
                                    &handled);
    if (rc) return rc;
    if (handled) return CHESS_OK;
  }
  return chess_launch_compare_generic(ctx, ref_base, qry_base, pairs, n_pairs, max_n, *params, ssim,
                                      sn, status, stream);
}

extern "C" int chess_compare_band_batch(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                                        const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                                        const chess_params *params, double *ssim, double *sn,
                                        int32_t *status, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  int rc = validate(ctx, params);
  if (rc) return rc;
  if (!ref || !qry || !ref->band || !qry->band ||
      n_pairs < 0 || (n_pairs > 0 && (!pairs || !ssim || !sn || !status))) {
    ctx->err = "chess_compare_band_batch: NULL buffer";
    return CHESS_ERR_ARG;
  }
  if (n_pairs == 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  chess_params p = *params;
  p.flags |= CHESS_FLAG_SYMMETRIC;  // a band is symmetric by construction
  if ((p.kernel == 0 || p.kernel == 3) && p.data_range == 0.0) {
    bool handled = false;
    rc = chess_launch_compare_items(ctx, ref, qry, pairs, n_pairs, max_n, p, ssim, sn, status, stream, &handled);
    if (rc) return rc;
    if (handled) return CHESS_OK;
  }
  return chess_compare_batch(ctx, ref->band, qry->band, pairs, n_pairs, max_n, &p, ssim, sn, status, stream_);
}

// Page-locked host memory the device can DMA from / into without staging (cudaHostAlloc, cudaHostRegister,
// torch's pinned allocator).  Pageable memory reports cudaMemoryTypeUnregistered.
static bool is_pinned(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// Shared body of the *_host entry points: stage the pair table, run `launch`, (optionally) the run-level
// z-score on the device, copy the results back in one transfer.
template <typename Launch>
static int host_roundtrip(chess_ctx *ctx, const chess_pair *pairs_host, int64_t n_pairs, double *ssim_host,
                          double *sn_host, int32_t *status_host, double *z_host, cudaStream_t stream,
                          Launch launch) {
  const size_t pair_bytes = (size_t)n_pairs * sizeof(chess_pair);
  const size_t res_bytes = (size_t)n_pairs * ((z_host ? 3 : 2) * sizeof(double) + sizeof(int32_t));
  // Small batches (a chromosome's worth of pairs) are latency-bound end to end: the kernels read the pair
  // table straight from the pinned staging buffer (unified addressing: 40 B per work item over PCIe instead
  // of a DMA that must finish before the launch) and write the results into it (no D2H copy behind the last
  // kernel): 144 -> 136 us per call on configs[1].  CHESS_ZEROCOPY=0 switches both off, =1 the table only (A/B).
  static const int zc_env = getenv("CHESS_ZEROCOPY") ? atoi(getenv("CHESS_ZEROCOPY")) : 2;
  const int zc = n_pairs <= 8192 ? zc_env : 0;
  const size_t res_off = (pair_bytes + 255) & ~size_t(255);
  int rc;
  if ((rc = grow(ctx, &ctx->d_pairs, &ctx->d_pairs_bytes, pair_bytes, false))) return rc;
  if ((rc = grow(ctx, &ctx->d_results, &ctx->d_results_bytes, res_bytes, false))) return rc;
  if ((rc = grow(ctx, &ctx->h_pinned, &ctx->h_pinned_bytes, res_off + res_bytes, true))) return rc;
  char *hp = reinterpret_cast<char *>(ctx->h_pinned);
  // A caller that already holds page-locked buffers (torch pinned tensors, cudaHostAlloc) is served by DMA
  // straight from / into them; pageable buffers go through the context's pinned staging area (one memcpy on
  // the host each way: 0.5 ms of the 13.7 ms call on the genome-wide table).
  const bool direct = zc == 0 && is_pinned(pairs_host) && is_pinned(ssim_host) && is_pinned(sn_host) &&
                      is_pinned(status_host) && (!z_host || is_pinned(z_host));
  if (!direct) std::memcpy(hp, pairs_host, pair_bytes);
  const chess_pair *k_pairs = reinterpret_cast<const chess_pair *>(hp);
  if (zc == 0) {
    CHESS_CUDA(ctx, cudaMemcpyAsync(ctx->d_pairs, direct ? (const void *)pairs_host : (const void *)hp, pair_bytes,
                                    cudaMemcpyHostToDevice, stream));
    k_pairs = reinterpret_cast<const chess_pair *>(ctx->d_pairs);
  }
  double *d_ssim = reinterpret_cast<double *>(zc == 2 ? (void *)(hp + res_off) : ctx->d_results);
  double *d_sn = d_ssim + n_pairs;
  double *d_z = d_sn + n_pairs;                      // only if z_host
  int32_t *d_status = reinterpret_cast<int32_t *>(d_sn + (z_host ? 2 : 1) * n_pairs);
  rc = launch(k_pairs, d_ssim, d_sn, d_status);
  if (rc) return rc;
  if (z_host && (rc = chess_zscore(ctx, d_ssim, n_pairs, d_z, stream))) return rc;
  const size_t col = (size_t)n_pairs * sizeof(double);
  if (direct) {
    CHESS_CUDA(ctx, cudaMemcpyAsync(ssim_host, d_ssim, col, cudaMemcpyDeviceToHost, stream));
    CHESS_CUDA(ctx, cudaMemcpyAsync(sn_host, d_sn, col, cudaMemcpyDeviceToHost, stream));
    if (z_host) CHESS_CUDA(ctx, cudaMemcpyAsync(z_host, d_z, col, cudaMemcpyDeviceToHost, stream));
    CHESS_CUDA(ctx, cudaMemcpyAsync(status_host, d_status, (size_t)n_pairs * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, stream));
    CHESS_CUDA(ctx, cudaStreamSynchronize(stream));
    return CHESS_OK;
  }
  if (zc != 2)
    CHESS_CUDA(ctx, cudaMemcpyAsync(hp + res_off, ctx->d_results, res_bytes, cudaMemcpyDeviceToHost, stream));
  CHESS_CUDA(ctx, cudaStreamSynchronize(stream));
  const char *h = hp + res_off;
  std::memcpy(ssim_host, h, col);
  std::memcpy(sn_host, h + col, col);
  if (z_host) std::memcpy(z_host, h + 2 * col, col);
  std::memcpy(status_host, h + (z_host ? 3 : 2) * col, (size_t)n_pairs * sizeof(int32_t));
  return CHESS_OK;
}

extern "C" int chess_compare_batch_host(chess_ctx *ctx, const double *ref_base,
                                        const double *qry_base, const chess_pair *pairs_host,
                                        int64_t n_pairs, int32_t max_n, const chess_params *params,
                                        double *ssim_host, double *sn_host, int32_t *status_host,
                                        void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_pairs < 0 || (n_pairs > 0 && (!pairs_host || !ssim_host || !sn_host || !status_host))) {
    ctx->err = "chess_compare_batch_host: NULL buffer";
    return CHESS_ERR_ARG;
  }
  if (n_pairs == 0) return CHESS_OK;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  return host_roundtrip(ctx, pairs_host, n_pairs, ssim_host, sn_host, status_host, nullptr, (cudaStream_t)stream_,
                        [&](const chess_pair *d_pairs, double *d_ssim, double *d_sn, int32_t *d_status) {
                          return chess_compare_batch(ctx, ref_base, qry_base, d_pairs, n_pairs, max_n, params,
                                                     d_ssim, d_sn, d_status, stream_);
                        });
}

extern "C" int chess_sim_band_batch_host(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                                        const chess_pair *pairs_host, int64_t n_pairs, int32_t max_n,
                                        const chess_params *params, double *ssim_host, double *sn_host,
                                        int32_t *status_host, double *z_host, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_pairs < 0 || (n_pairs > 0 && (!pairs_host || !ssim_host || !sn_host || !status_host))) {
    ctx->err = "chess_sim_band_batch_host: NULL buffer";
    return CHESS_ERR_ARG;
  }
  if (n_pairs == 0) return CHESS_OK;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  return host_roundtrip(ctx, pairs_host, n_pairs, ssim_host, sn_host, status_host, z_host, (cudaStream_t)stream_,
                        [&](const chess_pair *d_pairs, double *d_ssim, double *d_sn, int32_t *d_status) {
                          return chess_compare_band_batch(ctx, ref, qry, d_pairs, n_pairs, max_n, params,
                                                          d_ssim, d_sn, d_status, stream_);
                        });
}

extern "C" int chess_compare_band_batch_host(chess_ctx *ctx, const chess_sample *ref,
                                             const chess_sample *qry, const chess_pair *pairs_host,
                                             int64_t n_pairs, int32_t max_n, const chess_params *params,
                                             double *ssim_host, double *sn_host, int32_t *status_host,
                                             void *stream_) {
  return chess_sim_band_batch_host(ctx, ref, qry, pairs_host, n_pairs, max_n, params, ssim_host, sn_host,
                                   status_host, nullptr, stream_);
}

namespace {
__global__ void __launch_bounds__(1024) zscore_kernel(const double *__restrict__ s, long long n,
                                                      double *__restrict__ z) {
  // One CTA, fixed reduction order (deterministic).  Count, sum and sum of squared deviations from
  // a pivot (the first finite value) in a single sweep: with the pivot inside the data the
  // shifted one-pass variance is as accurate as numpy's two passes for these magnitudes (|ssim| <= 1).
  __shared__ double red[3][33];
  __shared__ double pivot_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cudaGridDependencySynchronize();  // launched early (programmatic serialisation): wait for the producer
  if (threadIdx.x == 0) {
    double pv = 0.0;
    for (long long i = 0; i < n; ++i) {
      const double v = s[i];
      if (v == v) { pv = v; break; }
    }
    pivot_s = pv;
  }
  __syncthreads();
  const double pivot = pivot_s;
  // one CTA is latency-bound: 16 loads per thread in flight (one at a time cost 46 us for the 101 354 values of
  // the genome-wide list, 2.7 % of an 8-GPU step); the order of the additions per thread is unchanged
  constexpr int U = 16;
  double cnt = 0.0, sum = 0.0, ssq = 0.0;
  long long i = threadIdx.x;
  for (; i + (long long)(U - 1) * 1024 < n; i += (long long)U * 1024) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = s[i + (long long)u * 1024];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (v[u] == v[u]) { const double d = v[u] - pivot; cnt += 1.0; sum += d; ssq = fma(d, d, ssq); }
  }
  for (; i < n; i += 1024) {
    const double v = s[i];
    if (v == v) { const double d = v - pivot; cnt += 1.0; sum += d; ssq = fma(d, d, ssq); }
  }
  cnt = warp_sum(cnt); sum = warp_sum(sum); ssq = warp_sum(ssq);
  if (lane == 0) { red[0][warp] = cnt; red[1][warp] = sum; red[2][warp] = ssq; }
  __syncthreads();
  if (warp == 0) {
    cnt = warp_sum(red[0][lane]); sum = warp_sum(red[1][lane]); ssq = warp_sum(red[2][lane]);
    if (lane == 0) { red[0][32] = cnt; red[1][32] = sum; red[2][32] = ssq; }
  }
  __syncthreads();
  cnt = red[0][32];
  const double dmean = red[1][32] / cnt;                 // mean - pivot
  const double var = red[2][32] / cnt - dmean * dmean;   // population variance (ddof 0)
  const double sd = sqrt(var > 0.0 ? var : 0.0);
  i = threadIdx.x;
  for (; i + (long long)(U - 1) * 1024 < n; i += (long long)U * 1024) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = s[i + (long long)u * 1024];
#pragma unroll
    for (int u = 0; u < U; ++u) z[i + (long long)u * 1024] = isfinite(v[u]) ? ((v[u] - pivot) - dmean) / sd : CUDART_NAN;
  }
  for (; i < n; i += 1024) {
    const double v = s[i];
    z[i] = isfinite(v) ? ((v - pivot) - dmean) / sd : CUDART_NAN;
  }
}
}  // namespace

extern "C" int chess_zscore(chess_ctx *ctx, const double *ssim, int64_t n, double *z, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n < 0 || (n > 0 && (!ssim || !z))) { ctx->err = "chess_zscore: bad arguments"; return CHESS_ERR_ARG; }
  if (n == 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(1024);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CHESS_CUDA(ctx, cudaLaunchKernelEx(&cfg, zscore_kernel, ssim, (long long)n, z));
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  return CHESS_OK;
}

extern "C" int chess_sn_dense(chess_ctx *ctx, const double *m1, const double *m2, int32_t n,
                              double *out, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n < 3 || !m1 || !m2 || !out) {
    ctx->err = "chess_sn_dense: need n >= 3 and non-NULL buffers";
    return CHESS_ERR_ARG;
  }
  // One dense pair through the generic kernel with masking off; the SSIM it also
  // computes is discarded (device scratch from the results staging buffer).
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = grow(ctx, &ctx->d_pairs, &ctx->d_pairs_bytes, sizeof(chess_pair), false))) return rc;
  if ((rc = grow(ctx, &ctx->d_results, &ctx->d_results_bytes, 64, false))) return rc;
  chess_pair pr;
  std::memset(&pr, 0, sizeof(pr));
  pr.ref_pitch = n; pr.qry_pitch = n; pr.ref_n = n; pr.qry_n = n;
  CHESS_CUDA(ctx, cudaMemcpyAsync(ctx->d_pairs, &pr, sizeof(pr), cudaMemcpyHostToDevice, stream));
  CHESS_CUDA(ctx, cudaStreamSynchronize(stream));  // pr lives on this stack frame
  chess_params p;
  std::memset(&p, 0, sizeof(p));
  p.min_bins = 0; p.keep_unmappable = 1; p.abs_window = 3; p.compute_sn = 1;
  p.rel_window = 1.0; p.mappability_cutoff = 0.0; p.default_value = 1.0; p.kernel = 1;
  double *d_ssim = reinterpret_cast<double *>(ctx->d_results);
  int32_t *d_status = reinterpret_cast<int32_t *>(d_ssim + 1);
  return chess_launch_compare_generic(ctx, m1, m2, reinterpret_cast<const chess_pair *>(ctx->d_pairs), 1, n, p,
                                      d_ssim, out, d_status, stream);
}
