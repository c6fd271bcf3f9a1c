// Shared declarations for the chess_b200 CUDA translation units (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/chess_b200.h"

struct chess_ctx {
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  std::string err;
  int64_t launches = 0;
  // staging for the *_host entry points (grown on demand, freed in destroy)
  void *d_pairs = nullptr;
  size_t d_pairs_bytes = 0;
  void *d_results = nullptr;
  size_t d_results_bytes = 0;
  void *h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  // workspace of the item kernel: work counters, retry flags, arrival counters, band records
  void *d_ingest = nullptr;   // sparse-text parser: tile counts / offsets (+ 256 spare bytes behind d_ingest_bytes)
  size_t d_ingest_bytes = 0;
  void *d_bg = nullptr;       // background path: pair descriptors + results of one chunk
  size_t d_bg_bytes = 0;
  void *d_items = nullptr;
  size_t d_items_bytes = 0;
  size_t items_done_cap = 0;
  int items_parity = 0;
};

#define CHESS_CUDA(ctx, call)                                                          \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                \
      return CHESS_ERR_CUDA;                                                           \
    }                                                                                  \
  } while (0)

// Internal per-pair status: the tuned kernel could not take this pair, the
// generic kernel must (never visible to callers).
#define CHESS_PAIR_RETRY 100

// Arguments of the work-item kernel (compare_kernels.cuh).
struct chess_item_args {
  const double *ref_band, *qry_band;
  const int *ref_lo, *ref_hi, *qry_lo, *qry_hi;
  const double *ref_pmax, *ref_pmin, *qry_pmax, *qry_pmin;  // running extrema, may be NULL
  const double *ref_pre, *qry_pre;                          // band-row prefix sums (triangle kernel only)
  const chess_pair *pairs;
  long long n_pairs;
  chess_params p;
  double *ssim, *sn;
  int *status;
  int ncol_pad;               // columns reserved per pair in maps and records (max n, rounded to 32)
  int max_n;                  // largest matrix of the batch, in bins
  int blocks;                 // block-triangle kernel: blocks per matrix side
  long long pair_items;       // triangle -r 1 kernel: leading pairs walked whole by one warp (set by its launcher)
  double *scratch;            // n_pairs * kBands * (kItemHdr + 2*ncol_pad)
  int *done;                  // n_pairs, zero on entry, zero on exit
  unsigned long long *counter_cur, *counter_next;
  int *retry_cur, *retry_next;
};

// Work-item kernel launchers, one translation unit per window half-width H (0 = -r 1 regime).
int chess_items_launch_h0(chess_ctx *ctx, chess_item_args &a, int G, bool multi, cudaStream_t stream);
int chess_items_launch_h1(chess_ctx *ctx, chess_item_args &a, int G, bool multi, cudaStream_t stream);
int chess_items_launch_h2(chess_ctx *ctx, chess_item_args &a, int G, bool multi, cudaStream_t stream);
int chess_items_launch_h3(chess_ctx *ctx, chess_item_args &a, int G, bool multi, cudaStream_t stream);
// -r 1 over the upper triangle (compare_tri.cuh)
int chess_items_launch_tri(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream);
// -r 1, matrices wider than one warp: square blocks on and above the diagonal (compare_blk.cuh)
int chess_items_launch_blk(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream);
int chess_items_launch_blkw(chess_ctx *ctx, chess_item_args &a, int G, int H, cudaStream_t stream);  // windows 3..11
// -a 7 over the upper triangle (compare_tri7.cuh)
int chess_items_launch_tri7(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream);
// -a 3 / -a 5 over the upper triangle (compare_triw.cuh)
int chess_items_launch_triw(chess_ctx *ctx, chess_item_args &a, int G, int H, cudaStream_t stream);

// Internal launchers (one per translation unit).
int chess_launch_compare_generic(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                                 const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                                 const chess_params &p, double *ssim, double *sn, int32_t *status,
                                 cudaStream_t stream, bool retry_only = false,
                                 const int *retry_flag = nullptr);
int chess_launch_compare_items(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                               const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                               const chess_params &p, double *ssim, double *sn, int32_t *status,
                               cudaStream_t stream, bool *handled);
int chess_launch_compare_fast(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                              const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                              const chess_params &p, double *ssim, double *sn, int32_t *status,
                              cudaStream_t stream, bool *handled);

// ---- small device helpers -------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Source index of output index o for an order-0 resize n_in -> n_out: the
// sampling rule of scipy.ndimage.zoom(order=0, grid_mode=True), which is what
// skimage.transform.resize(order=0) calls (chess/sim.py:323-329).  Plain
// round-to-nearest double operations, no FMA contraction, so the index agrees
// with the CPU for every o.
__device__ __forceinline__ int resize_src_index(int o, int n_in, int n_out) {
  double zoom = __ddiv_rn((double)n_in, (double)n_out);
  double cc = __dsub_rn(__dmul_rn(__dadd_rn((double)o, 0.5), zoom), 0.5);
  int idx = (int)floor(__dadd_rn(cc, 0.5));
  return min(max(idx, 0), n_in - 1);
}

// Window rule of chess/sim.py:356-365.
__host__ __device__ __forceinline__ int window_rule(int n, int abs_window, double rel_window) {
  int w = abs_window ? abs_window : (int)((double)n * rel_window);
  if ((w & 1) == 0) w -= 1;
  return w < 3 ? 3 : w;
}
