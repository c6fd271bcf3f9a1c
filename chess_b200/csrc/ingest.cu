// K8: sparse-matrix text -> COO on the device (chess/helpers.py:309-339).
//
// The reference reads "source<TAB>sink<TAB>weight" lines with a Python loop (int(), int(), float()
// per line).  Here the file bytes are copied to HBM once and parsed there:
//   1. newline count per 4 KB tile -> exclusive scan -> start offset of every line;
//   2. one thread per line: rstrip, split at tabs, int / int / float with parse_number.cuh
//      (bit-identical to Python or flagged for the host), canonical order source <= sink;
//   3. stable compaction of the data lines (file order is kept: a later duplicate must win,
//      chess/helpers.py:353-356).
// Lines the device cannot decide (unusual number spellings, > 19 significant digits, subnormals,
// malformed lines) are listed for the host, which parses them with Python itself and patches the
// per-line arrays before step 3.
#include "common.cuh"
#include "parse_number.cuh"

namespace {

__device__ const uint64_t d_pow5[] = CHESS_POW5_TABLE;
__device__ const short d_exp5[] = CHESS_POW5_EXP;
__device__ const double d_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                       1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

constexpr int kTile = 4096;       // bytes per tile
constexpr int kTileThreads = 256; // 16 bytes per thread
constexpr int kLineTile = 1024;   // lines per compaction tile

__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *sh /* [33] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int nw = (int)(blockDim.x >> 5);
    int w = lane < nw ? sh[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nw) sh[lane] = winc - w;
    if (lane == 31) sh[32] = winc;
  }
  __syncthreads();
  *total = sh[32];
  return inc - v + sh[warp];
}

// newlines (and carriage returns that are not followed by a newline) per tile
__global__ void __launch_bounds__(kTileThreads)
nl_count_kernel(const char *__restrict__ text, long long n, int *__restrict__ tile_nl, int *__restrict__ tile_cr) {
  __shared__ int sh[33];
  const long long n_tiles = (n + kTile - 1) / kTile;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * kTile + (long long)threadIdx.x * 16;
    int nl = 0, cr = 0;
    if (base + 16 <= n) {
      const uint4 v = *reinterpret_cast<const uint4 *>(text + base);
      const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const unsigned c = (wds[k >> 2] >> ((k & 3) * 8)) & 0xffu;
        nl += c == '\n';
        if (c == '\r') {
          const unsigned nx = k < 15 ? (wds[(k + 1) >> 2] >> (((k + 1) & 3) * 8)) & 0xffu
                                     : (base + 16 < n ? (unsigned)(unsigned char)text[base + 16] : 0u);
          cr += nx != '\n';
        }
      }
    } else {
      for (long long i = base; i < n; ++i) {
        const char c = text[i];
        nl += c == '\n';
        if (c == '\r') cr += !(i + 1 < n && text[i + 1] == '\n');
      }
    }
    int tot_nl, tot_cr;
    block_exclusive_scan(nl, &tot_nl, sh);
    block_exclusive_scan(cr, &tot_cr, sh);
    if (threadIdx.x == 0) { tile_nl[tile] = tot_nl; tile_cr[tile] = tot_cr; }
  }
}

// exclusive scan of int counts into 64-bit offsets, one CTA, fixed order; out[n] = total
__global__ void __launch_bounds__(1024)
scan_kernel(const int *__restrict__ counts, long long n, long long *__restrict__ out) {
  __shared__ long long sh[1024];
  const long long per = (n + 1023) / 1024;
  const long long lo = per * threadIdx.x, hi = lo + per < n ? lo + per : n;
  long long s = 0;
  for (long long i = lo; i < hi; ++i) s += counts[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int t = 0; t < 1024; ++t) { const long long v = sh[t]; sh[t] = run; run += v; }
    out[n] = run;
  }
  __syncthreads();
  long long run = sh[threadIdx.x];
  for (long long i = lo; i < hi; ++i) { out[i] = run; run += counts[i]; }
}

// line_start[k] = offset of the first byte of line k; line k ends (exclusive, without its newline)
// at line_start[k + 1] - 1
__global__ void __launch_bounds__(kTileThreads)
line_start_kernel(const char *__restrict__ text, long long n, const long long *__restrict__ tile_off,
                  long long n_lines, long long *__restrict__ line_start) {
  __shared__ int sh[33];
  const long long n_tiles = (n + kTile - 1) / kTile;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    line_start[0] = 0;
    // a last line without a newline ends at n: pretend its newline sits at offset n
    if (n_lines > tile_off[n_tiles]) line_start[n_lines] = n + 1;
  }
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * kTile + (long long)threadIdx.x * 16;
    unsigned mask = 0;
    for (int k = 0; k < 16; ++k) {
      const long long i = base + k;
      if (i < n && text[i] == '\n') mask |= 1u << k;
    }
    int tot;
    int rank = block_exclusive_scan(__popc(mask), &tot, sh);
    const long long first = tile_off[tile] + 1;
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      line_start[first + rank] = base + k + 1;
      ++rank;
    }
  }
}

__global__ void __launch_bounds__(128)
parse_lines_kernel(const char *__restrict__ text, const long long *__restrict__ line_start, long long n_lines,
                   int32_t *__restrict__ src, int32_t *__restrict__ sink, double *__restrict__ weight,
                   unsigned char *__restrict__ kind, long long *__restrict__ fb_lines, long long fb_cap,
                   unsigned long long *__restrict__ fb_count) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n_lines;
       k += (long long)gridDim.x * blockDim.x) {
    const long long b = line_start[k];
    long long e = line_start[k + 1] - 1;
    const char *s = text + b;
    int len = (int)(e - b);
    int kd = PARSE_FALLBACK;
    int32_t a = 0, c = 0;
    double w = 0.0;
    if (e - b > 4096) {
      kd = PARSE_FALLBACK;                       // absurdly long line: host
    } else if (len > 0 && s[0] == '#') {
      kd = 1;                                    // comment (startswith('#'), before rstrip)
    } else {
      while (len > 0 && chess_is_space(s[len - 1])) --len;   // line.rstrip()
      if (len == 0) {
        kd = 1;                                  // blank line
      } else {
        int t1 = -1, t2 = -1, t3 = len;
        for (int i = 0; i < len; ++i) {
          if (s[i] == '\t') {
            if (t1 < 0) t1 = i;
            else if (t2 < 0) t2 = i;
            else { t3 = i; break; }
          }
        }
        if (t1 >= 0 && t2 >= 0) {
          const int r0 = chess_parse_index(s, t1, &a);
          const int r1 = chess_parse_index(s + t1 + 1, t2 - t1 - 1, &c);
          const int r2 = chess_parse_double(s + t2 + 1, t3 - t2 - 1, d_pow5, d_exp5, d_pow10, &w);
          if (r0 == PARSE_OK && r1 == PARSE_OK && r2 != PARSE_FALLBACK) kd = r2;
        }
      }
    }
    if (kd == PARSE_FALLBACK) {
      const unsigned long long at = atomicAdd(fb_count, 1ull);
      if ((long long)at < fb_cap) fb_lines[at] = k;
    } else if (kd == PARSE_OK) {
      src[k] = a <= c ? a : c;                   // chess/helpers.py:332-335
      sink[k] = a <= c ? c : a;
      weight[k] = w;
    }
    kind[k] = (unsigned char)kd;
  }
}

__global__ void __launch_bounds__(256)
kind_count_kernel(const unsigned char *__restrict__ kind, long long n_lines, int *__restrict__ tile_keep,
                  unsigned long long *__restrict__ totals /* [nonfinite, fallback] */) {
  __shared__ int sh[33];
  const long long n_tiles = (n_lines + kLineTile - 1) / kLineTile;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int keep = 0, nonfinite = 0, fb = 0;
    for (int j = threadIdx.x; j < kLineTile; j += 256) {
      const long long k = tile * kLineTile + j;
      if (k < n_lines) {
        const int kd = kind[k];
        keep += kd == PARSE_OK;
        nonfinite += kd == PARSE_NONFINITE;
        fb += kd == PARSE_FALLBACK;
      }
    }
    int t0, t1, t2;
    block_exclusive_scan(keep, &t0, sh);
    block_exclusive_scan(nonfinite, &t1, sh);
    block_exclusive_scan(fb, &t2, sh);
    if (threadIdx.x == 0) {
      tile_keep[tile] = t0;
      if (t1) atomicAdd(&totals[0], (unsigned long long)t1);
      if (t2) atomicAdd(&totals[1], (unsigned long long)t2);
    }
  }
}

__global__ void __launch_bounds__(256)
compact_kernel(const unsigned char *__restrict__ kind, long long n_lines, const long long *__restrict__ tile_off,
               const int32_t *__restrict__ src, const int32_t *__restrict__ sink,
               const double *__restrict__ weight, int32_t *__restrict__ out_src, int32_t *__restrict__ out_sink,
               double *__restrict__ out_weight) {
  __shared__ int sh[33];
  const long long n_tiles = (n_lines + kLineTile - 1) / kLineTile;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // thread t owns the 4 consecutive lines 4t .. 4t+3 of the tile: ranks stay in file order
    const long long k0 = tile * kLineTile + (long long)threadIdx.x * 4;
    int flags = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (k0 + j < n_lines && kind[k0 + j] == PARSE_OK) flags |= 1 << j;
    int tot;
    long long at = tile_off[tile] + block_exclusive_scan(__popc(flags), &tot, sh);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (flags & (1 << j)) {
        out_src[at] = src[k0 + j];
        out_sink[at] = sink[k0 + j];
        out_weight[at] = weight[k0 + j];
        ++at;
      }
    }
  }
}

int ingest_workspace(chess_ctx *ctx, size_t need) {
  if (ctx->d_ingest_bytes >= need) return CHESS_OK;
  if (ctx->d_ingest) cudaFree(ctx->d_ingest);
  ctx->d_ingest = nullptr;
  ctx->d_ingest_bytes = 0;
  const size_t usable = ((need + need / 8) + 255) & ~(size_t)255;
  if (cudaMalloc(&ctx->d_ingest, usable + 256) != cudaSuccess) {  // 256 spare bytes: the fallback counter
    ctx->err = "allocation of the ingest workspace failed";
    return CHESS_ERR_NOMEM;
  }
  ctx->d_ingest_bytes = usable;
  return CHESS_OK;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// count + scan of the newline tiles; leaves tile offsets (n_tiles + 1 int64) at the start of the workspace
int index_tiles(chess_ctx *ctx, const char *text, int64_t n_bytes, cudaStream_t stream, int64_t *n_newlines,
                int64_t *n_lone_cr, char *last_byte) {
  const long long n_tiles = (n_bytes + kTile - 1) / kTile;
  const size_t off_bytes = align256((size_t)(n_tiles + 1) * 8);
  const size_t cnt_bytes = align256((size_t)n_tiles * 4);
  int rc = ingest_workspace(ctx, 2 * off_bytes + 2 * cnt_bytes);
  if (rc) return rc;
  char *base = reinterpret_cast<char *>(ctx->d_ingest);
  long long *nl_off = reinterpret_cast<long long *>(base);
  long long *cr_off = reinterpret_cast<long long *>(base + off_bytes);
  int *nl_cnt = reinterpret_cast<int *>(base + 2 * off_bytes);
  int *cr_cnt = reinterpret_cast<int *>(base + 2 * off_bytes + cnt_bytes);
  long long grid = n_tiles < 148 * 8 ? n_tiles : 148 * 8;
  nl_count_kernel<<<(unsigned)grid, kTileThreads, 0, stream>>>(text, n_bytes, nl_cnt, cr_cnt);
  scan_kernel<<<1, 1024, 0, stream>>>(nl_cnt, n_tiles, nl_off);
  scan_kernel<<<1, 1024, 0, stream>>>(cr_cnt, n_tiles, cr_off);
  ctx->launches += 3;
  CHESS_CUDA(ctx, cudaGetLastError());
  long long h_nl = 0, h_cr = 0;
  CHESS_CUDA(ctx, cudaMemcpyAsync(&h_nl, nl_off + n_tiles, 8, cudaMemcpyDeviceToHost, stream));
  CHESS_CUDA(ctx, cudaMemcpyAsync(&h_cr, cr_off + n_tiles, 8, cudaMemcpyDeviceToHost, stream));
  CHESS_CUDA(ctx, cudaMemcpyAsync(last_byte, text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, stream));
  CHESS_CUDA(ctx, cudaStreamSynchronize(stream));
  *n_newlines = h_nl;
  *n_lone_cr = h_cr;
  return CHESS_OK;
}

}  // namespace

extern "C" int chess_sparse_text_lines(chess_ctx *ctx, const char *text, int64_t n_bytes, int64_t *n_lines,
                                       int64_t *n_lone_cr, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_bytes < 0 || (n_bytes > 0 && !text) || !n_lines || !n_lone_cr) {
    ctx->err = "chess_sparse_text_lines: bad arguments";
    return CHESS_ERR_ARG;
  }
  *n_lines = 0;
  *n_lone_cr = 0;
  if (n_bytes == 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t nl = 0, cr = 0;
  char last = 0;
  int rc = index_tiles(ctx, text, n_bytes, stream, &nl, &cr, &last);
  if (rc) return rc;
  *n_lines = nl + (last != '\n' ? 1 : 0);
  *n_lone_cr = cr;
  return CHESS_OK;
}

extern "C" int chess_sparse_text_parse(chess_ctx *ctx, const char *text, int64_t n_bytes, int64_t n_lines,
                                       int32_t *src, int32_t *sink, double *weight, uint8_t *kind,
                                       int64_t *line_start, int64_t *fb_lines, int64_t fb_cap,
                                       int64_t *n_fallback, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_bytes < 0 || n_lines < 0 || !n_fallback ||
      (n_lines > 0 && (!text || !src || !sink || !weight || !kind || !line_start)) || (fb_cap > 0 && !fb_lines)) {
    ctx->err = "chess_sparse_text_parse: bad arguments";
    return CHESS_ERR_ARG;
  }
  *n_fallback = 0;
  if (n_lines == 0 || n_bytes == 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t nl = 0, cr = 0;
  char last = 0;
  int rc = index_tiles(ctx, text, n_bytes, stream, &nl, &cr, &last);
  if (rc) return rc;
  if (n_lines != nl + (last != '\n' ? 1 : 0)) {
    ctx->err = "chess_sparse_text_parse: n_lines does not match the text (call chess_sparse_text_lines first)";
    return CHESS_ERR_ARG;
  }
  const long long n_tiles = (n_bytes + kTile - 1) / kTile;
  const long long *nl_off = reinterpret_cast<const long long *>(ctx->d_ingest);
  long long grid = n_tiles < 148 * 8 ? n_tiles : 148 * 8;
  line_start_kernel<<<(unsigned)grid, kTileThreads, 0, stream>>>(text, n_bytes, nl_off, n_lines,
                                                                reinterpret_cast<long long *>(line_start));
  // the fallback counter lives in the spare bytes behind the workspace
  unsigned long long *fb_count = reinterpret_cast<unsigned long long *>(
      reinterpret_cast<char *>(ctx->d_ingest) + ctx->d_ingest_bytes);
  CHESS_CUDA(ctx, cudaMemsetAsync(fb_count, 0, 8, stream));
  long long pgrid = (n_lines + 127) / 128;
  if (pgrid > 148 * 32) pgrid = 148 * 32;
  parse_lines_kernel<<<(unsigned)pgrid, 128, 0, stream>>>(text, reinterpret_cast<const long long *>(line_start),
                                                          n_lines, src, sink, weight, kind,
                                                          reinterpret_cast<long long *>(fb_lines), fb_cap, fb_count);
  ctx->launches += 2;
  CHESS_CUDA(ctx, cudaGetLastError());
  unsigned long long h_fb = 0;
  CHESS_CUDA(ctx, cudaMemcpyAsync(&h_fb, fb_count, 8, cudaMemcpyDeviceToHost, stream));
  CHESS_CUDA(ctx, cudaStreamSynchronize(stream));
  *n_fallback = (int64_t)h_fb;
  return CHESS_OK;
}

extern "C" int chess_sparse_text_compact(chess_ctx *ctx, int64_t n_lines, const int32_t *src, const int32_t *sink,
                                         const double *weight, const uint8_t *kind, int32_t *out_src,
                                         int32_t *out_sink, double *out_weight, int64_t out_capacity,
                                         int64_t *counts, void *stream_) {
  if (!ctx) return CHESS_ERR_ARG;
  if (n_lines < 0 || !counts || (n_lines > 0 && (!src || !sink || !weight || !kind))) {
    ctx->err = "chess_sparse_text_compact: bad arguments";
    return CHESS_ERR_ARG;
  }
  counts[0] = counts[1] = counts[2] = 0;
  if (n_lines == 0) return CHESS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  CHESS_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long n_tiles = (n_lines + kLineTile - 1) / kLineTile;
  const size_t off_bytes = align256((size_t)(n_tiles + 1) * 8);
  const size_t cnt_bytes = align256((size_t)n_tiles * 4);
  int rc = ingest_workspace(ctx, off_bytes + cnt_bytes + 256);
  if (rc) return rc;
  char *base = reinterpret_cast<char *>(ctx->d_ingest);
  long long *off = reinterpret_cast<long long *>(base);
  int *cnt = reinterpret_cast<int *>(base + off_bytes);
  unsigned long long *totals = reinterpret_cast<unsigned long long *>(base + off_bytes + cnt_bytes);
  CHESS_CUDA(ctx, cudaMemsetAsync(totals, 0, 16, stream));
  long long grid = n_tiles < 148 * 8 ? n_tiles : 148 * 8;
  kind_count_kernel<<<(unsigned)grid, 256, 0, stream>>>(kind, n_lines, cnt, totals);
  scan_kernel<<<1, 1024, 0, stream>>>(cnt, n_tiles, off);
  ctx->launches += 2;
  CHESS_CUDA(ctx, cudaGetLastError());
  long long kept = 0;
  unsigned long long h_tot[2] = {0, 0};
  CHESS_CUDA(ctx, cudaMemcpyAsync(&kept, off + n_tiles, 8, cudaMemcpyDeviceToHost, stream));
  CHESS_CUDA(ctx, cudaMemcpyAsync(h_tot, totals, 16, cudaMemcpyDeviceToHost, stream));
  CHESS_CUDA(ctx, cudaStreamSynchronize(stream));
  counts[0] = kept;
  counts[1] = kept + (int64_t)h_tot[0];   // data lines = finite + non-finite (the 5 % warning, helpers.py:337-339)
  counts[2] = (int64_t)h_tot[0];
  if (h_tot[1] != 0) {
    ctx->err = "chess_sparse_text_compact: lines flagged for the host are still unresolved";
    return CHESS_ERR_ARG;
  }
  if (out_src == nullptr && out_sink == nullptr && out_weight == nullptr) return CHESS_OK;  // counting pass
  if (!out_src || !out_sink || !out_weight || out_capacity < kept) {
    ctx->err = "chess_sparse_text_compact: output buffers missing or too small";
    return CHESS_ERR_ARG;
  }
  compact_kernel<<<(unsigned)grid, 256, 0, stream>>>(kind, n_lines, off, src, sink, weight, out_src, out_sink,
                                                    out_weight);
  ctx->launches += 1;
  CHESS_CUDA(ctx, cudaGetLastError());
  CHESS_CUDA(ctx, cudaStreamSynchronize(stream));
  return CHESS_OK;
}
