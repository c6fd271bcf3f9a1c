/* chess_b200.h -- C ABI of the B200-native `chess sim` comparison hot path.
 *
 * The reference (vaquerizaslab/chess 0.3.8) is pure Python and has no FFI: the
 * boundary of its hot path is a set of Python call signatures.  Every entry
 * point below names the reference interface it stands behind (paths relative
 * to the reference checkout).  The Python package `chess_b200` binds this
 * library with ctypes and keeps the reference's names and signatures
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - plain C, opaque context, `int` status (0 = ok, <0 = error; message via
 *    chess_last_error()).
 *  - The caller owns every buffer.  Pointers marked DEVICE are device
 *    pointers of the context's GPU (e.g. torch.Tensor.data_ptr()); the
 *    library never frees caller memory.  The `_host` entry points take HOST
 *    pointers and perform the host<->device copies themselves: by DMA
 *    straight from / into the caller's arrays when ALL of them are
 *    page-locked (cudaHostAlloc, cudaHostRegister, torch pinned tensors),
 *    through context-owned pinned staging buffers otherwise (pageable
 *    memory: one extra host copy each way).  Tables of up to 8192 pairs are
 *    read from, and their results written to, the staging area directly
 *    by the kernels (no DMA on the path of a latency-bound call).
 *  - All launches go to the caller-supplied stream (a cudaStream_t passed as
 *    void*; NULL = legacy default stream).  Device entry points do not
 *    synchronise; `_host` entry points return after the results are in host
 *    memory.
 *  - One context per GPU; a context is not thread-safe.
 *  - There is no CPU fallback: without a CUDA device every compute entry
 *    point fails.
 */
#ifndef CHESS_B200_H
#define CHESS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHESS_B200_ABI_VERSION 4

typedef struct chess_ctx chess_ctx;

/* ---- status codes of the API calls ------------------------------------ */
#define CHESS_OK 0
#define CHESS_ERR_CUDA (-1)     /* a CUDA runtime call failed                */
#define CHESS_ERR_ARG (-2)      /* invalid argument                          */
#define CHESS_ERR_UNSUPPORTED (-3) /* shape outside what the kernels cover   */
#define CHESS_ERR_NOMEM (-4)

/* ---- per-pair status written by the compare kernels -------------------- */
#define CHESS_PAIR_OK 0
#define CHESS_PAIR_NOT_FOUND 1   /* region has no bins: ValueError in the gather
                                    -> (nan, nan), chess/sim.py:198-200        */
#define CHESS_PAIR_TOO_SMALL 2   /* < min_bins, chess/sim.py:319-320           */
#define CHESS_PAIR_UNMAPPABLE 3  /* masked fraction > cutoff, sim.py:344-347   */
#define CHESS_PAIR_WINDOW_EXCEEDS 4 /* window larger than the compacted matrix:
                                    skimage raises ValueError, which aborts the
                                    reference run (sim.py:217-220)             */

/* One comparison.  A submatrix is the square [0,n) x [0,n) of the 2-D view
 *     element(r, c) = base[off + r * pitch + c]
 * which covers both storages the path uses:
 *   - symmetric band (chess_oe_band_build): off = first_bin*(2W-1) + (W-1),
 *     pitch = 2W-2;
 *   - packed dense matrices (chess/sim.py:306 compare_matrix_pair callers):
 *     off = offset of the matrix, pitch = n.
 * Stands behind the (pair_ix, reference_region, query_region) tuple consumed
 * by chunk_comparison_worker, chess/sim.py:147-222, after the region -> bin
 * lookup of chess/helpers.py:222-268.                                        */
typedef struct chess_pair {
  int64_t ref_off;   /* element offset of (0,0) of the reference submatrix    */
  int64_t qry_off;
  int32_t ref_pitch; /* row pitch in elements                                  */
  int32_t qry_pitch;
  int32_t ref_n;     /* bins; 0 or negative = region not found                */
  int32_t qry_n;
  int32_t flip;      /* 1 if the strands differ (sim.py:332-333)               */
  int32_t reserved;
} chess_pair;

/* Keyword arguments of compare_matrix_pair, chess/sim.py:306-317.            */
typedef struct chess_params {
  int32_t min_bins;          /* 20 from the CLI (bin/chess:204)                */
  int32_t keep_unmappable;   /* --keep-unmappable-bins                         */
  int32_t abs_window;        /* -a; 0 = None                                   */
  int32_t compute_sn;        /* False for background comparisons (bin/chess:306) */
  double rel_window;         /* -r                                             */
  double mappability_cutoff; /* --mappability-cutoff                           */
  double default_value;      /* fill / masking value, 1 from the CLI           */
  int32_t kernel;            /* 0 = auto, 1 = force the generic kernel,
                                2 = skip the work-item kernels, 3 = work-item
                                kernel without the upper-triangle form (A/B)   */
  int32_t flags;             /* CHESS_FLAG_* promises about the batch          */
  double data_range;         /* 0 = max - min of the compacted pair, as chess sim
                                passes it (sim.py:368-374); > 0 = this value, e.g.
                                2.0, what skimage's legacy compare_ssim assumes for
                                float images when utils/run_synthetic_test.py:306
                                calls it without data_range.  Such batches run on
                                the generic kernel.  (ABI 3)                      */
} chess_params;

/* Promises the caller makes about a batch; they only select faster kernels and
 * never change results.  SYMMETRIC: every submatrix view is symmetric (true for
 * the band, chess/helpers.py:301-302 mirrors every edge).  EQUAL_SIZES: every
 * pair whose two regions were found has ref_n == qry_n (no resize step,
 * chess/sim.py:323-329).                                                      */
#define CHESS_FLAG_SYMMETRIC 1
#define CHESS_FLAG_EQUAL_SIZES 2
/* MOSTLY_EQUAL_SIZES: nearly all pairs have ref_n == qry_n (e.g. background regions of one length in bp
 * whose last bin is shorter at a chromosome end): the work-item kernels are worth launching, the few unequal
 * pairs take the resize path in the retry pass.  A hint like the others: results do not depend on it.       */
#define CHESS_FLAG_MOSTLY_EQUAL_SIZES 4

/* ---- context ------------------------------------------------------------ */
int chess_abi_version(void);
int chess_ctx_create(int device, chess_ctx **out);
void chess_ctx_destroy(chess_ctx *ctx);
const char *chess_last_error(const chess_ctx *ctx); /* ctx may be NULL */
/* Number of kernels this context has launched (bench.py's gpu_launches).     */
int64_t chess_launch_count(const chess_ctx *ctx);

/* ---- K1: per-chromosome per-diagonal expected values -------------------- *
 * Replaces expected_values + possible_contacts, chess/oe.py:17-146.
 * COO triples are canonical (src <= sink), non-finite weights already dropped
 * (chess/helpers.py:328-335).  bin_chrom[i] is the chromosome id of bin i,
 * chrom_start[c] the first bin of chromosome c (n_chrom+1 entries, bins of a
 * chromosome contiguous).  Outputs (DEVICE):
 *   expected[chrom_start[c] + d]  expected value of chromosome c at distance d
 *   inter_expected[0]             inter-chromosomal expected value
 *   mappable[i]                   1 if the marginal of bin i is > 0
 *   possible[chrom_start[c] + d]  #(i, i+d) pairs with both bins mappable
 *   diag_sum[chrom_start[c] + d]  (optional, may be NULL) the diagonal sums
 * Sums are accumulated in 128-bit fixed point, so the result does not depend
 * on the order of the atomics (bit-identical across runs and GPU counts).
 * `workspace` (DEVICE) must hold chess_expected_workspace_bytes(n_bins).      */
size_t chess_expected_workspace_bytes(int64_t n_bins);
int chess_expected_from_coo(chess_ctx *ctx, const int32_t *src, const int32_t *sink,
                            const double *weight, int64_t nnz, const int32_t *bin_chrom,
                            const int64_t *chrom_start, int32_t n_chrom, int64_t n_bins,
                            double *expected, double *inter_expected, uint8_t *mappable,
                            int64_t *possible, double *diag_sum, void *workspace,
                            size_t workspace_bytes, void *stream);
/* possible_contacts alone, chess/oe.py:17-75 (intra-chromosomal counts).      */
int chess_possible_contacts(chess_ctx *ctx, const uint8_t *mappable, const int64_t *chrom_start,
                            int32_t n_chrom, int64_t n_bins, int64_t *possible, void *stream);

/* ---- K2: O/E values and the symmetric band ------------------------------ *
 * chess_oe_values replaces _oe_generator, chess/oe.py:149-166: out[k] =
 * weight[k] / expected (1 when the expected value is exactly 0); apply_log2
 * is the optional fused transform (default off: `chess sim` applies none).   */
int chess_oe_values(chess_ctx *ctx, const int32_t *src, const int32_t *sink, const double *weight,
                    int64_t nnz, const int32_t *bin_chrom, const int64_t *chrom_start,
                    const double *expected, const double *inter_expected, int32_t apply_log2,
                    double *out, void *stream);
/* chess_band_build replaces the dict-of-dict edges store (helpers.py:342-356)
 * and the per-cell fill of sub_matrix_from_edges_dict (helpers.py:296-304):
 * band[i*(2W-1) + (j-i) + W-1] for |j-i| < W, pre-filled with `fill`, then
 * both mirror cells of every triple are written.  band holds
 * n_bins*(2W-1) doubles (DEVICE).                                            */
int chess_band_build(chess_ctx *ctx, const int32_t *src, const int32_t *sink, const double *value,
                     int64_t nnz, int64_t n_bins, int32_t W, double fill, double *band,
                     void *stream);
/* The same for bins [bin_lo, bin_lo + n_range) only (ABI 4): the band of the stretch of the genome one GPU's
 * share of the pair list touches (the reference forks the whole dict into every worker, bin/chess:196-208;
 * here a GPU holds what its contiguous range of pairs reads).  band holds n_range*(2W-1) doubles, row 0 = bin
 * bin_lo; triples with a bin outside the range are skipped.  Pair offsets into it use first_bin - bin_lo.      */
int chess_band_build_range(chess_ctx *ctx, const int32_t *src, const int32_t *sink, const double *value,
                           int64_t nnz, int64_t bin_lo, int64_t n_range, int32_t W, double fill, double *band,
                           void *stream);
/* Facts about a device-resident COO list (ABI 4): out4 (DEVICE, 4 x int64) = smallest index, largest index,
 * number of triples with src > sink, number of triples that do not come strictly after their predecessor in
 * (src, sink) order.  0 in the last slot = sorted and free of duplicates, so that "a later line overwrites an
 * earlier one" (edges_d[source][sink] = weight, chess/helpers.py:353-356) has nothing to decide and contiguous
 * bin ranges are contiguous slices.                                                                           */
int chess_coo_check(chess_ctx *ctx, const int32_t *src, const int32_t *sink, int64_t nnz, int64_t *out4,
                    void *stream);
/* out[0] (DEVICE int32) = 1 if any of `count` dense n x n matrices packed back to back at `base` differs from its
 * transpose, else 0 (ABI 4).  What CHESS_FLAG_SYMMETRIC asserts for the dense batches of
 * utils/run_synthetic_test.py:97-121, checked where the matrices already are.                                */
int chess_symmetric_check(chess_ctx *ctx, const double *base, int32_t n, int64_t count, int32_t *out, void *stream);

/* Dense copy of one submatrix view, out[r*n + c] = base[off + r*pitch + c]:
 * the return value of sub_matrix_from_edges_dict, chess/helpers.py:286-306.   */
int chess_gather_submatrix(chess_ctx *ctx, const double *base, int64_t off, int32_t pitch,
                           int32_t n, double *out, void *stream);

/* ---- K3-K5: gather + masks + SSIM + SN ---------------------------------- *
 * Replaces, per pair, sub_matrix_from_edges_dict (helpers.py:286-306) and
 * compare_matrix_pair (sim.py:306-380) including skimage's
 * structural_similarity and SN (sim.py:15-71).  All pointers DEVICE.
 * ssim/sn are NaN and status != 0 where the reference returns (nan, nan);
 * sn is NaN when compute_sn == 0 (the reference returns None).
 * max_n = largest ref_n/qry_n in the batch (sizes shared memory).            */
int chess_compare_batch(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                        const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                        const chess_params *params, double *ssim, double *sn, int32_t *status,
                        void *stream);
/* Same with HOST pair descriptors and HOST result arrays; the copies are part
 * of the call (this is the entry point bench.py times as `e2e`).             */
int chess_compare_batch_host(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                             const chess_pair *pairs_host, int64_t n_pairs, int32_t max_n,
                             const chess_params *params, double *ssim_host, double *sn_host,
                             int32_t *status_host, void *stream);

/* ---- band-resident samples: the fastest form of the compare call ------- *
 * A sample = its band plus a per-bin index built once by chess_band_index_build:
 * nd_lo[i] / nd_hi[i] = nearest bin k <= i / k >= i (within the band) whose entry
 * band(i,k) differs from the default value (-1 / INT32_MAX if none).  With it a
 * column of a submatrix is known to be all-default -- i.e. an unmappable bin,
 * chess/sim.py:74-87 -- without summing it, so the tuned kernels need no separate
 * column-sum pass; the prediction is verified against the column sums the kernel
 * accumulates anyway, and the rare mismatch is redone by the generic kernel.
 * Pairs must be band views (off = first_bin*(2W-1)+W-1, pitch = 2W-2).          */
typedef struct chess_sample {
  const double *band;    /* DEVICE, n_bins*(2W-1) doubles (chess_band_build)   */
  const int32_t *nd_lo;  /* DEVICE, n_bins, may be NULL (then no prediction)    */
  const int32_t *nd_hi;  /* DEVICE, n_bins, may be NULL                         */
  const double *pmax;    /* DEVICE, n_bins*W: pmax[i*W+d] = max band(i, i..i+d); may be NULL */
  const double *pmin;    /* DEVICE, n_bins*W: running minimum likewise; may be NULL */
  int64_t n_bins;
  int32_t W;
  int32_t reserved;
  const double *rowpre;  /* DEVICE, n_bins*2W: prefix sums along every band row (chess_band_prefix_build);
                            may be NULL (then the -r 1 kernel walks the full square instead of the triangle) */
} chess_sample;
int chess_band_index_build(chess_ctx *ctx, const double *band, int64_t n_bins, int32_t W,
                           double default_value, int32_t *nd_lo, int32_t *nd_hi, void *stream);
/* Running extrema along every upper band row.  The data range of a submatrix
 * (chess/sim.py:368-370) starting at bin i0 with n bins is then
 * max_r pmax[(i0+r)*W + n-1-r] - min_r pmin[...]: n lookups instead of n*n
 * compares (the matrix is symmetric, so its upper triangle holds every value). */
/* Prefix sums along every band row: rowpre[i*2W + k] = sum of band(i, .) over the first k entries
 * of row i (k = 0 .. 2W-1).  The sum of a submatrix column -- the quantity find_masked_rows tests,
 * chess/sim.py:74-87 -- is then one subtraction; the -r 1 kernel uses it to screen the columns it
 * did NOT predict as masked for a coincidental hit of n*default (confirmed exactly, in the
 * reference's own order, before the pair is handed to the generic kernel).                     */
int chess_band_prefix_build(chess_ctx *ctx, const double *band, int64_t n_bins, int32_t W,
                            double *rowpre, void *stream);
int chess_band_extrema_build(chess_ctx *ctx, const double *band, int64_t n_bins, int32_t W,
                             double *pmax, double *pmin, void *stream);
int chess_compare_band_batch(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                             const chess_pair *pairs, int64_t n_pairs, int32_t max_n,
                             const chess_params *params, double *ssim, double *sn, int32_t *status,
                             void *stream);
int chess_compare_band_batch_host(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                                  const chess_pair *pairs_host, int64_t n_pairs, int32_t max_n,
                                  const chess_params *params, double *ssim_host, double *sn_host,
                                  int32_t *status_host, void *stream);
/* The same plus the run-level z-score (K6, bin/chess:216-233) of this batch, computed on the device and
 * returned in the same transfer; z_host may be NULL.  One call = what `chess sim` reports per pair.  */
int chess_sim_band_batch_host(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                              const chess_pair *pairs_host, int64_t n_pairs, int32_t max_n,
                              const chess_params *params, double *ssim_host, double *sn_host,
                              int32_t *status_host, double *z_host, void *stream);

/* K6: run-level z-score, bin/chess:216-233: z = (ssim - nanmean) / nanstd
 * (ddof 0) over the non-NaN entries; entries that are not finite keep NaN.
 * Fixed-order reduction in one CTA, so the result is deterministic. DEVICE.   */
int chess_zscore(chess_ctx *ctx, const double *ssim, int64_t n, double *z, void *stream);

/* K7: background statistics of `chess sim --background-query` / `--background-regions`,
 * bin/chess:259-332.  The REFERENCE region of every pair is compared with every background
 * region of the query sample (compute_sn = False, bin/chess:300-307); per pair
 *     z_bg = (ssim - nanmean(bg)) / nanstd(bg)            (ddof 0, bin/chess:323-326)
 *     p_bg = (1 + #{ssim <= bg}) / len(bg)                (NaNs counted in len, bin/chess:327-329)
 * and NaN / NaN where the pair's own ssim is NaN (bin/chess:319-321).  Regions are bin spans of
 * band-resident samples (first bin, bin count; count <= 0 = region not found); strands are small
 * integer codes, a pair is flipped iff the two codes differ (chess/sim.py:332-333; background BEDs
 * load with strand None, which differs from '+' and '-').  own_first / own_n / own_strand (nullable
 * together) give every pair's own QUERY region: a background region with the same bins and the
 * same orientation repeats the pair's own comparison, where the reference obtains the bit-identical
 * value (it counts in `ssim <= bg`); the pair's ssim is used for it, so the tie does not depend on
 * which kernel produced either number.  The n_pairs x n_bg pair descriptors,
 * their results and the reductions never leave the device.  n_exceeds (nullable) receives the
 * number of comparisons whose window exceeds the compacted matrix -- where skimage raises and the
 * reference run aborts.  params->compute_sn is ignored.  All pointers DEVICE.                  */
/* Background regions of `chess sim --background-query`, bin/chess:272-298, made on the device (ABI 4): for every
 * query bin (bin_begin, bin_end, chromosome id; bins of a chromosome contiguous
 * and sorted, chrom_first[c] = its first bin, n_chrom + 1 entries; chrom_end_bp[c] = largest bin end) the region
 * [start, start + size] on both strands, kept where start + size <= chrom_end_bp (and, limit_chrom >= 0, on that
 * chromosome only), each looked up like helpers.py:212-242 does (half-open overlap with [start - 1, start + size)).
 * bin_begin[i] is the tree interval's begin, i.e. region.start - 1 where region.start is the BED start as
 * load_regions keeps it.  Output in the reference's order (bin order, '+' then '-'): bg_first / bg_n / bg_strand
 * (codes 0 / 1) hold up to 2 * n_bins entries, n_regions[0] the number written.  workspace: 13 * n_bins bytes.
 * All pointers DEVICE.                                                                                        */
int chess_background_regions(chess_ctx *ctx, const int64_t *bin_begin, const int64_t *bin_end,
                             const int32_t *bin_chrom, const int64_t *chrom_first, const int64_t *chrom_end_bp,
                             int64_t n_bins, int64_t size, int32_t limit_chrom, int64_t *bg_first, int32_t *bg_n,
                             int32_t *bg_strand, int64_t *n_regions, void *workspace, void *stream);

int chess_background_batch(chess_ctx *ctx, const chess_sample *ref, const chess_sample *qry,
                           const int64_t *ref_first, const int32_t *ref_n, const int32_t *ref_strand,
                           const int64_t *own_first, const int32_t *own_n, const int32_t *own_strand,
                           const double *pair_ssim, int64_t n_pairs, const int64_t *bg_first,
                           const int32_t *bg_n, const int32_t *bg_strand, int64_t n_bg, int32_t max_n,
                           const chess_params *params, double *z_bg, double *p_bg, int32_t *n_exceeds,
                           void *stream);

/* K8: sparse-matrix text ("source<TAB>sink<TAB>weight" lines) -> COO on the device; replaces the
 * Python line loop of edges_from_sparse_matrix, chess/helpers.py:309-339.  `text` is the file
 * content in DEVICE memory (16-byte aligned).  Three synchronous steps:
 *   chess_sparse_text_lines    number of lines, and of carriage returns not followed by a newline
 *                              (Python's universal newlines would split there: use the host parser);
 *   chess_sparse_text_parse    per line k: kind[k] = 0 data (src[k] <= sink[k], weight[k] set),
 *                              1 comment ('#' first) or blank, 2 non-finite weight (dropped by the
 *                              reference, helpers.py:328-330), 3 = the device could not decide
 *                              (spelling outside the plain int / float grammar, > 19 significant
 *                              digits, subnormal, malformed).  int() / float() results are
 *                              bit-identical to Python's.  Kind-3 line numbers go to fb_lines (up to
 *                              fb_cap; *n_fallback is the full count); line_start[n_lines + 1]
 *                              receives the byte offset of every line (line k ends at
 *                              line_start[k+1] - 1), so the host can re-parse those lines with
 *                              Python itself and patch src / sink / weight / kind;
 *   chess_sparse_text_compact  stable compaction of the kind-0 lines (file order kept: a later
 *                              duplicate wins, helpers.py:353-356).  counts[0] = kept, counts[1] =
 *                              data lines (kept + non-finite), counts[2] = non-finite.  With all
 *                              three outputs NULL only the counts are produced.  Fails if a kind-3
 *                              line is left.                                                        */
int chess_sparse_text_lines(chess_ctx *ctx, const char *text, int64_t n_bytes, int64_t *n_lines,
                            int64_t *n_lone_cr, void *stream);
int chess_sparse_text_parse(chess_ctx *ctx, const char *text, int64_t n_bytes, int64_t n_lines,
                            int32_t *src, int32_t *sink, double *weight, uint8_t *kind,
                            int64_t *line_start, int64_t *fb_lines, int64_t fb_cap, int64_t *n_fallback,
                            void *stream);
int chess_sparse_text_compact(chess_ctx *ctx, int64_t n_lines, const int32_t *src, const int32_t *sink,
                              const double *weight, const uint8_t *kind, int32_t *out_src,
                              int32_t *out_sink, double *out_weight, int64_t out_capacity,
                              int64_t *counts, void *stream);

/* K9: first stage of `chess extract`, chess/get_structures.py:99-121.  ref_base / qry_base are
 * bands (or dense tiles) built with default value 0 (sub_matrix_from_edges_dict(..., default_weight=0.)).
 * Per pair k with equal sizes n: OR = log(query / reference), NaN and +-inf -> 0; std = np.std(OR);
 * positive[k*max_n*max_n + r*n + c] = OR > 0.5 std ? OR : 0, negative = OR < -0.5 std ? -OR : 0
 * (n*n doubles, row-major, at a stride of max_n*max_n per pair); stats[4k..] = {std, mean(positive),
 * mean(negative), n}.  status: 0 ok, 1 region not found, 100 the two regions differ in size (numpy
 * cannot divide them: the reference raises).  No strand handling: the reference has none here.
 * All pointers DEVICE.                                                                          */
int chess_log_ratio_batch(chess_ctx *ctx, const double *ref_base, const double *qry_base,
                          const chess_pair *pairs, int64_t n_pairs, int32_t max_n, double *positive,
                          double *negative, double *stats, int32_t *status, void *stream);

/* SN(M1, M2, size=7), chess/sim.py:15-30, on n x n DEVICE matrices.          */
int chess_sn_dense(chess_ctx *ctx, const double *m1, const double *m2, int32_t n, double *out,
                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CHESS_B200_H */
