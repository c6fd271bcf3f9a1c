"""CPU oracle for the `chess sim` comparison hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy *restatement* of the reference algorithm
(vaquerizaslab/chess 0.3.8).  It is the checker for the CUDA path: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it.  Nothing under `chess_b200/` imports it, and the product path never
falls back to it.

Pinning status
--------------
* Everything that is the reference's *own* numpy code (SN, masked-row
  detection, row removal, the window rule, expected values, O/E, the
  dict-of-dict gather, the z-score / background statistics, the TSV writer) is
  pinned: `tests/golden/make_golden.py` executes the reference's functions
  straight from `/root/reference/chess/*.py` (missing third-party imports
  shimmed) and `tests/test_oracle_golden.py` checks this restatement against
  those outputs bit for bit.
* `structural_similarity` and `resize` are arithmetic of scikit-image, an
  un-vendored and unpinned dependency (`setup.py:45`) that is absent from this
  image.  They are restated from the published algorithm
  (`skimage.metrics.structural_similarity`, `skimage.transform.resize`,
  scikit-image >= 0.19) on top of `scipy.ndimage`, and cross-checked against an
  independent brute-force formulation.  **Parity unpinned** for those two
  functions: no scikit-image output could be generated here.

Each function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import bisect
import gzip
import io
import math
from collections import defaultdict

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view
from scipy import ndimage as ndi

ORACLE_VERSION = "chess-oracle-1 (numpy %s)" % np.__version__


# --------------------------------------------------------------------------
# regions, pairs, sparse matrices  (chess/helpers.py)
# --------------------------------------------------------------------------
class Region:
    """Genomic interval; equality ignores strand (chess/helpers.py:12-150)."""

    __slots__ = ("chromosome", "start", "end", "ix", "strand")

    def __init__(self, start, end, chromosome=None, ix=None, strand=None):
        self.chromosome = chromosome
        self.start = start
        self.end = end
        self.ix = ix
        self.strand = strand

    def __eq__(self, other):  # helpers.py:126-136
        return (self.chromosome == other.chromosome and self.start == other.start
                and self.end == other.end)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return "%s:%s-%s:%s" % (self.chromosome, self.start, self.end, self.strand)


def _open_text(path):
    """helpers.py:153-162."""
    if path.endswith(".gz") or path.endswith(".gzip"):
        return io.TextIOWrapper(gzip.open(path, mode="r"))
    return open(path, "r")


def load_regions(path, sep=None, ignore_ix=False):
    """BED -> (regions, name->ix or None, ix->[chrom,start,end]).

    helpers.py:165-206.  Note the BED start is kept 0-based (`:186`), the
    bin index is the line number (`:188`), and column 4 (unless '.') names the
    bin for HiC-Pro style sparse files (`:190-198`).
    """
    regions, ix2reg, names = [], {}, None
    with _open_text(path) as fh:
        for lineno, raw in enumerate(fh):
            cols = raw.rstrip().split(sep)
            if len(cols) < 3:
                continue
            chrom, start, end = cols[0], int(cols[1]), int(cols[2])
            ix2reg[lineno] = [chrom, start, end]
            if not ignore_ix and len(cols) > 3 and cols[3] != ".":
                if names is None:
                    names = {}
                if cols[3] in names:
                    raise ValueError("name column in region BED must only contain "
                                     "unique values! ({})".format(cols[3]))
                names[cols[3]] = lineno
            regions.append(Region(start, end, chrom, ix=lineno))
    if names is not None:  # helpers.py:203-205
        back = {str(v): int(k) for k, v in names.items()}
        ix2reg = {back[str(k)]: v for k, v in ix2reg.items()}
    return regions, names, ix2reg


def iter_sparse(path, ix_converter=None, sep="\t"):
    """Yield canonical (src<=sink, weight) triples; helpers.py:309-339."""
    with _open_text(path) as fh:
        for raw in fh:
            if raw.startswith("#"):
                continue
            cols = raw.rstrip().split(sep)
            if ix_converter is None:
                a, b = int(cols[0]), int(cols[1])
            else:
                a, b = ix_converter[cols[0]], ix_converter[cols[1]]
            w = float(cols[2])
            if not np.isfinite(w):
                continue
            yield (a, b, w) if a <= b else (b, a, w)


def edges_dict(triples):
    """{src: {sink: w}}, later duplicates win; helpers.py:342-356."""
    out = defaultdict(dict)
    for a, b, w in triples:
        out[a][b] = w
    return out


def load_pairs_for_chroms(path, chromosomes, sep=None):
    """10-column BEDPE -> ([(ID, ref, qry)], dropped); helpers.py:483-527.

    start = bedpe_start + 1, end = bedpe_end, strands kept as strings.
    Duplicate IDs raise; pairs on unknown chromosomes are counted as dropped.
    """
    seen, out, dropped = set(), [], 0
    with _open_text(path) as fh:
        for raw in fh:
            if raw.startswith("#"):
                continue
            cols = raw.rstrip().split(sep)
            if len(cols) != 10:
                raise ValueError("10 columns expected but {} found in bedpe input".format(len(cols)))
            c1, s1, e1, c2, s2, e2, pid, _, st1, st2 = raw.split(sep)
            if pid in seen:
                raise ValueError("Pair ID {} not unique in bedpe input!")
            seen.add(pid)
            if c1 not in chromosomes or c2 not in chromosomes:
                dropped += 1
                continue
            out.append((pid,
                        Region(int(s1) + 1, int(e1), str(c1), strand=str(st1)),
                        Region(int(s2) + 1, int(e2), str(c2), strand=str(st2))))
    return out, dropped


def chunks(seq, p):
    """helpers.py:530-534."""
    size = int(np.ceil(len(seq) / p))
    for lo in range(0, len(seq), size):
        yield seq[lo:lo + size]


class RegionTree:
    """Stand-in for `intervaltree.IntervalTree` as used by helpers.py:209-241.

    Intervals are half open [begin, end); `overlap(a, b)` returns the data of
    every interval with begin < b and end > a, which is intervaltree's slice
    semantics (intervaltree 3.x, un-vendored dependency).  Implemented as a
    linear scan between two bisections over the sorted begins; regions of one
    chromosome in a BED file are (near-)sorted and short, so the scan bound
    `max_len` keeps this exact for arbitrary inputs.
    """

    def __init__(self, items):
        # items: iterable of (begin, end, data)
        self.items = sorted(items, key=lambda t: (t[0], t[1]))
        self.begins = [t[0] for t in self.items]
        self.max_len = max((t[1] - t[0] for t in self.items), default=0)

    def overlap(self, a, b):
        lo = bisect.bisect_right(self.begins, a - self.max_len)
        hi = bisect.bisect_left(self.begins, b)
        return [d for (s, e, d) in self.items[max(lo - 1, 0):hi] if s < b and e > a]


def region_trees(regions):
    """{chrom: RegionTree} with intervals [start-1, end); helpers.py:209-219."""
    per = defaultdict(list)
    for r in regions:
        per[r.chromosome].append((r.start - 1, r.end, r))
    return {c: RegionTree(v) for c, v in per.items()}


def sub_regions(trees, region):
    """Bins overlapping `region`; helpers.py:222-268 (interval-tree branch).

    Query slice is [region.start-1, region.end).  No hit -> ValueError
    (`min([])` at :241).  Unknown chromosome -> KeyError (uncaught, :233).
    """
    tree = trees[region.chromosome]
    hits = tree.overlap(region.start - 1, region.end)
    ixs = [h.ix for h in hits]
    lo, hi = min(ixs), max(ixs)  # ValueError on empty, like the reference
    return hits, lo, hi


def sub_matrix_from_edges_dict(edges, trees, region, default_weight=0.0):
    """Dense symmetric submatrix, absent entries = default; helpers.py:286-306.

    Deliberately keeps the reference's cost profile (upper-triangle loop over a
    dict of dicts) because this function is part of the timed CPU baseline.
    """
    hits, lo, hi = sub_regions(trees, region)
    n = hi - lo + 1
    m = np.full((n, n), default_weight, dtype="float64")
    for i in range(lo, hi + 1):
        for j in range(i, hi + 1):
            try:
                w = edges[i][j]
            except KeyError:
                continue
            m[i - lo, j - lo] = w
            m[j - lo, i - lo] = w
    return m, hits


# --------------------------------------------------------------------------
# observed / expected  (chess/oe.py)
# --------------------------------------------------------------------------
def chromosome_bins(regions):
    """{chrom: [first_ix, last_ix+1]}; oe.py:7-14."""
    out = {}
    for r in regions:
        if r.chromosome not in out:
            out[r.chromosome] = [r.ix, r.ix + 1]
        else:
            out[r.chromosome][1] = r.ix + 1
    return out


def possible_contacts(regions, mappable):
    """Number of (i, i+d) pairs with both bins mappable, per chromosome and d,
    and the inter-chromosomal mappable pair count; oe.py:17-75.

    Restated as a direct count (the reference subtracts unmappable rows and
    columns from `count - d`; the golden test shows both give the same ints).
    Returns (genome_wide[d], {chrom: [count_d]}, inter_total).
    """
    cb = chromosome_bins(regions)
    order = []
    for r in regions:
        if not order or order[-1] != r.chromosome:
            order.append(r.chromosome)
    mappable = np.asarray(mappable, dtype=bool)
    max_d = max(hi - lo for lo, hi in cb.values())
    genome = np.zeros(max_d, dtype=np.int64)
    per_chrom, n_map = {}, {}
    for chrom in order:
        lo, hi = cb[chrom]
        m = mappable[lo:hi].astype(np.int64)
        n = hi - lo
        # autocorrelation of the mappable indicator = #pairs at each distance
        full = np.correlate(m, m, mode="full")[n - 1:] if n < 4096 else _autocorr(m)
        counts = np.rint(full).astype(np.int64)
        per_chrom[chrom] = [int(v) for v in counts]
        genome[:n] += counts
        n_map[chrom] = int(m.sum())
    inter = 0
    for a in range(len(order)):
        for b in range(a + 1, len(order)):
            inter += n_map[order[a]] * n_map[order[b]]
    return [int(v) for v in genome], per_chrom, inter


def _autocorr(m):
    n = len(m)
    size = 1 << int(math.ceil(math.log2(2 * n)))
    f = np.fft.rfft(m.astype(np.float64), size)
    return np.fft.irfft(f * np.conj(f), size)[:n]


def expected_values(regions, triples):
    """(genome-wide expected[d], {chrom: expected[d]}, inter expected); oe.py:78-146.

    Sums run in input order with Python floats, exactly like the reference, so
    the results are bit-identical to it.
    """
    cb = chromosome_bins(regions)
    chrom_of = {}
    max_d = 0
    for chrom, (lo, hi) in cb.items():
        max_d = max(max_d, hi - lo)
        for i in range(lo, hi):
            chrom_of[i] = chrom
    sums = {c: [0.0] * (hi - lo) for c, (lo, hi) in cb.items()}
    marg = [0.0] * len(regions)
    genome_sums = [0.0] * max_d
    inter_sum = 0.0
    for a, b, w in triples:
        if not np.isfinite(w):
            continue
        ca, cbb = chrom_of[a], chrom_of[b]
        marg[a] += w
        marg[b] += w
        if ca != cbb:
            inter_sum += w
        else:
            genome_sums[b - a] += w
            sums[ca][b - a] += w
    mappable = [v > 0 for v in marg]  # oe.py:119
    genome_tot, chrom_tot, inter_tot = possible_contacts(regions, mappable)
    inter_exp = 0 if inter_tot == 0 else inter_sum / inter_tot  # oe.py:123
    genome_exp = [0.0] * max_d
    for d in range(max_d):
        if genome_tot[d] > 0:
            genome_exp[d] = genome_sums[d] / genome_tot[d]
    chrom_exp = {}
    for c, (lo, hi) in cb.items():
        e = [0.0] * (hi - lo)
        for d in range(hi - lo):
            if chrom_tot[c][d] > 0:
                e[d] = sums[c][d] / chrom_tot[c][d]
        chrom_exp[c] = e
    return genome_exp, chrom_exp, inter_exp, mappable


def observed_expected(regions, triples):
    """[(src, sink, w / expected)], per-chromosome expected; oe.py:149-181.

    Division by an expected of exactly 0 yields 1 (the reference catches
    ZeroDivisionError, oe.py:158-161).  No log transform (`log=False`).
    """
    triples = list(triples)
    _, chrom_exp, inter_exp, _ = expected_values(regions, triples)
    chrom_of = {i: r.chromosome for i, r in enumerate(regions)}
    out = []
    for a, b, w in triples:
        if chrom_of[a] == chrom_of[b]:
            e = chrom_exp[chrom_of[a]][b - a]
        else:
            e = inter_exp
        try:
            oe = w / e
        except ZeroDivisionError:
            oe = 1
        out.append((a, b, oe))
    return out


# --------------------------------------------------------------------------
# scikit-image restatements (un-vendored dependency -- PARITY UNPINNED)
# --------------------------------------------------------------------------
def structural_similarity(im1, im2, win_size, data_range):
    """Mean SSIM of two 2-D float64 images.

    Restates `skimage.metrics.structural_similarity` (scikit-image >= 0.19) for
    the call at chess/sim.py:371-374: uniform window, K1=0.01, K2=0.03,
    sample covariance, `scipy.ndimage.uniform_filter` (mode='reflect'), crop
    (win-1)//2 each side, float64 mean.
    """
    im1 = np.asarray(im1, dtype=np.float64)
    im2 = np.asarray(im2, dtype=np.float64)
    if np.any((np.asarray(im1.shape) - win_size) < 0):
        raise ValueError("win_size exceeds image extent. Either ensure that your images are "
                         "at least 7x7; or pass win_size explicitly in the function call, "
                         "with an odd value less than or equal to the smaller side of your "
                         "images. If your images are multichannel (with color channels), set "
                         "channel_axis to the axis number corresponding to the channels.")
    if win_size % 2 != 1:
        raise ValueError("Window size must be odd.")
    npix = win_size ** im1.ndim
    cov_norm = npix / (npix - 1)
    ux = ndi.uniform_filter(im1, size=win_size)
    uy = ndi.uniform_filter(im2, size=win_size)
    uxx = ndi.uniform_filter(im1 * im1, size=win_size)
    uyy = ndi.uniform_filter(im2 * im2, size=win_size)
    uxy = ndi.uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    a1, a2 = 2 * ux * uy + c1, 2 * vxy + c2
    b1, b2 = ux ** 2 + uy ** 2 + c1, vx + vy + c2
    with np.errstate(invalid="ignore", divide="ignore"):
        s = (a1 * a2) / (b1 * b2)
    pad = (win_size - 1) // 2
    core = s[tuple(slice(pad, dim - pad) for dim in s.shape)]
    return core.mean(dtype=np.float64)


def structural_similarity_bruteforce(im1, im2, win_size, data_range):
    """Same quantity from explicit valid windows (no filter, no boundary mode).

    Independent formulation used to cross-check `structural_similarity`: after
    the crop only windows fully inside the image survive, so boundary handling
    cannot matter.
    """
    x = np.asarray(im1, dtype=np.float64)
    y = np.asarray(im2, dtype=np.float64)
    npix = win_size * win_size
    cov_norm = npix / (npix - 1)
    wx = sliding_window_view(x, (win_size, win_size))
    wy = sliding_window_view(y, (win_size, win_size))
    ux = wx.mean(axis=(2, 3))
    uy = wy.mean(axis=(2, 3))
    uxx = (wx * wx).mean(axis=(2, 3))
    uyy = (wy * wy).mean(axis=(2, 3))
    uxy = (wx * wy).mean(axis=(2, 3))
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    with np.errstate(invalid="ignore", divide="ignore"):
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    return s.mean(dtype=np.float64)


def synthetic_window_size(n_bins, rel_window):
    """Window of `compare_matrices`, utils/run_synthetic_test.py:302-305."""
    win = int(n_bins * rel_window)
    if win % 2 == 0:
        win -= 1
    return max(win, 7)


def synthetic_compare_matrices(a, b, rel_window, data_range=2.0):
    """`compare_matrices` of utils/run_synthetic_test.py:279-308 (BASELINE configs[0]), scaling mode 1:
    the smaller matrix is resized up (order 0), then `compare_ssim(a, b, win_size=...)` -- the legacy
    skimage call without data_range, which for float images takes the dtype's range -1 .. 1 = 2.0
    (un-vendored dependency, PARITY UNPINNED; the value is stated and is a parameter)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape[0] > b.shape[0]:
        big, small = a, resize_order0(b, a.shape)
        a, b = big, small
    elif b.shape[0] > a.shape[0]:
        a, b = b, resize_order0(a, b.shape)
    win = synthetic_window_size(a.shape[0], rel_window)
    return structural_similarity(a, b, win_size=win, data_range=data_range)


def synthetic_truth_and_controls(reference, query, pool, rel_window, data_range=2.0):
    """Truth SSIM, control SSIMs, p and z of utils/run_synthetic_test.py:105-115."""
    truth = synthetic_compare_matrices(reference, query, rel_window, data_range)
    control = np.array([synthetic_compare_matrices(reference, t, rel_window, data_range) for t in pool])
    p = float(np.sum(control >= truth)) / max(len(pool), 1)
    allv = np.concatenate([control, [truth]])
    z = float((truth - allv.mean()) / allv.std())
    return truth, control, p, z


def resize_nearest_index(n_in, n_out):
    """Source index for every output index of an order-0 resize n_in -> n_out.

    `skimage.transform.resize(order=0, mode='reflect')` (>= 0.19) calls
    `scipy.ndimage.zoom(img, n_out/n_in, order=0, mode='mirror', grid_mode=True)`,
    whose C kernel samples at cc = (o + 0.5) * (n_in / n_out) - 0.5 and takes
    floor(cc + 0.5).  Same double arithmetic here (and in the CUDA gather).
    """
    zoom = float(n_in) / float(n_out)
    o = np.arange(n_out, dtype=np.float64)
    cc = (o + 0.5) * zoom - 0.5
    idx = np.floor(cc + 0.5).astype(np.int64)
    return np.clip(idx, 0, n_in - 1)


def resize_order0(img, out_shape):
    """Order-0 upscale of a square matrix via scipy.ndimage.zoom itself."""
    img = np.asarray(img, dtype=np.float64)
    factors = [o / i for o, i in zip(out_shape, img.shape)]
    return ndi.zoom(img, factors, order=0, mode="mirror", grid_mode=True).astype(np.float64)


# --------------------------------------------------------------------------
# the per-pair comparison  (chess/sim.py)
# --------------------------------------------------------------------------
def nan_window_filter(m, size, func):
    """func over the size x size window at every pixel, NaN-padded border.

    chess/sim.py:33-71.  Materialises the (n, m, size*size) copy like the
    reference does (this is the dominant CPU cost and is kept on purpose).
    """
    assert size % 2 == 1 and m.dtype == np.float64
    half = (size - 1) // 2
    padded = np.full((m.shape[0] + size - 1, m.shape[1] + size - 1), np.nan)
    padded[half:half + m.shape[0], half:half + m.shape[1]] = m
    win = sliding_window_view(padded, (size, size))
    flat = win.copy().reshape(m.shape[0], m.shape[1], size * size)
    return func(flat, axis=2)


def signal_to_noise(m1, m2, size=7):
    """mean over pixels of |window mean| / window variance of m1 - m2; sim.py:15-30."""
    d = m1 - m2
    mu = nan_window_filter(d, size, np.nanmean)
    var = nan_window_filter(d, size, np.nanvar)
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(np.nanmean(np.abs(mu) / var))


def find_masked_rows(m, masking_value=0):
    """Columns whose float64 sum equals n * masking_value exactly; sim.py:74-87."""
    return np.where(np.sum(m, 0) == m.shape[0] * masking_value)[0]


def remove_rows(m, rows):
    """Delete rows and columns `rows`; sim.py:90-102."""
    return np.delete(np.delete(m, rows, 0), rows, 1)


def window_rule(n, absolute_window_size=None, relative_window_size=1.0):
    """sim.py:356-365: absolute overrides relative, forced odd, floor 3."""
    if absolute_window_size:
        w = absolute_window_size
    else:
        w = int(n * relative_window_size)
    if w % 2 == 0:
        w -= 1
    return max(w, 3)


def compare_matrix_pair(reference, reference_strand, query, query_strand, pair_ix=None,
                        min_bins=20, keep_unmappable_bins=False, absolute_window_size=None,
                        relative_window_size=1.0, mappability_cutoff=0.1, default_value=1,
                        compute_sn=True):
    """(pair_ix, ssim, sn) for one pair of dense matrices; sim.py:306-380."""
    if query.shape[0] < min_bins or reference.shape[0] < min_bins:
        return pair_ix, np.nan, np.nan
    ratio = len(reference) / len(query)
    if ratio < 1:
        reference = resize_order0(reference, query.shape)
    elif ratio > 1:
        query = resize_order0(query, reference.shape)
    if reference_strand != query_strand:
        query = query[::-1, ::-1]
    masked, ok = [], True
    if mappability_cutoff > 0:
        for m in (reference, query):
            idx = find_masked_rows(np.array(m), masking_value=default_value)
            masked.append(idx)
            if len(idx) / m.shape[0] > mappability_cutoff:
                ok = False
    if not ok:
        return pair_ix, np.nan, np.nan
    if not keep_unmappable_bins:
        drop = np.union1d(*masked)  # TypeError when cutoff <= 0, like the reference
        reference = remove_rows(reference, drop)
        query = remove_rows(query, drop)
    win = window_rule(reference.shape[0], absolute_window_size, relative_window_size)
    drange = max(reference.max(), query.max()) - min(reference.min(), query.min())
    ssim = structural_similarity(reference, query, win_size=win, data_range=drange)
    sn = signal_to_noise(np.ascontiguousarray(reference), np.ascontiguousarray(query)) \
        if compute_sn else None
    return pair_ix, ssim, sn


def compare_pair(pair, ref_edges, ref_trees, qry_edges, qry_trees, compute_sn=True,
                 default_value=1, **kw):
    """Gather both submatrices, then compare; sim.py:225-303."""
    pid, rr, qr = pair
    try:
        ref, _ = sub_matrix_from_edges_dict(ref_edges, ref_trees, rr, default_weight=default_value)
        qry, _ = sub_matrix_from_edges_dict(qry_edges, qry_trees, qr, default_weight=default_value)
    except ValueError:
        return pid, np.nan, np.nan
    return compare_matrix_pair(ref, rr.strand, qry, qr.strand, pair_ix=pid,
                               default_value=default_value, compute_sn=compute_sn, **kw)


def compare_chunk(chunk, ref_edges, ref_trees, qry_edges, qry_trees, compute_sn=True,
                  default_value=1, **kw):
    """One worker's chunk with the 1-deep region cache; sim.py:160-216."""
    prev_r, prev_q = [None, None], [None, None]
    out = []
    for pid, rr, qr in chunk:
        try:
            if prev_r[0] is not None and rr == prev_r[0]:
                ref = prev_r[1]
            else:
                ref, _ = sub_matrix_from_edges_dict(ref_edges, ref_trees, rr, default_weight=default_value)
                prev_r = [rr, ref]
            if prev_q[0] is not None and qr == prev_q[0]:
                qry = prev_q[1]
            else:
                qry, _ = sub_matrix_from_edges_dict(qry_edges, qry_trees, qr, default_weight=default_value)
                prev_q = [qr, qry]
        except ValueError:
            out.append([pid, np.nan, np.nan])
            continue
        _, s, sn = compare_matrix_pair(ref, rr.strand, qry, qr.strand, pair_ix=pid,
                                       default_value=default_value, compute_sn=compute_sn, **kw)
        out.append([pid, s, sn])
    return out


# --------------------------------------------------------------------------
# run-level statistics and the output table  (bin/chess:216-353)
# --------------------------------------------------------------------------
def run_zscores(results):
    """results: {id: [ssim, sn]} -> {id: [ssim, sn, z]}; bin/chess:216-233."""
    vals = [v[0] for v in results.values() if not np.isnan(v[0])]
    with np.errstate(invalid="ignore"), _quiet():
        mean, sd = np.nanmean(vals), np.nanstd(vals)
    out = {}
    for k, (s, sn) in results.items():
        z = np.nan
        if np.isfinite(s):
            with np.errstate(invalid="ignore", divide="ignore"):
                z = (s - mean) / sd
        out[k] = [s, sn, z]
    return out


class _quiet:
    def __enter__(self):
        import warnings
        self._cm = warnings.catch_warnings()
        self._cm.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *a):
        return self._cm.__exit__(*a)


def background_stats(original, bg_values):
    """(z_bg, p_bg) of one pair against its background ssims; bin/chess:317-332."""
    if np.isnan(original):
        return np.nan, np.nan
    bg = list(bg_values)
    with np.errstate(invalid="ignore", divide="ignore"), _quiet():
        mean, sd = np.nanmean(bg), np.nanstd(bg)
        z = (original - mean) / sd
        p = (1 + np.sum(original <= bg)) / len(bg)
    return z, p


def format_table(pairs, results, bg=None):
    """TSV text in pairs-file order; bin/chess:236-256, 334-353."""
    lines = ["ID\tSN\tssim\tz_ssim\n" if bg is None else "ID\tSN\tssim\tz_ssim\tz_bg\tp_bg\n"]
    for pid, _, _ in pairs:
        if pid in results:
            s, sn, z = results[pid]
        else:
            s, sn, z = np.nan, np.nan, np.nan
        if bg is None:
            lines.append("{}\t{}\t{}\t{}\n".format(pid, sn, s, z))
        else:
            zb, pb = bg.get(pid, (np.nan, np.nan)) if pid in results else (np.nan, np.nan)
            lines.append("{}\t{}\t{}\t{}\t{}\t{}\n".format(pid, sn, s, z, zb, pb))
    return "".join(lines)


def background_regions_from_query(query_regions, query_region, limit_background=False):
    """Every query bin start, both strands, with the span of `query_region`; kept if it ends
    inside its chromosome, optionally only on the query region's chromosome (bin/chess:272-298)."""
    ends = defaultdict(int)
    for r in query_regions:
        ends[r.chromosome] = max(ends[r.chromosome], r.end)
    span = query_region.end - query_region.start
    out = []
    for r in query_regions:
        if limit_background and r.chromosome != query_region.chromosome:
            continue
        for strand in ("+", "-"):
            if r.start + span <= ends[r.chromosome]:
                out.append(Region(r.start, r.start + span, r.chromosome, strand=strand))
    return out


def log_ratio_stage(ref_edges, ref_trees, qry_edges, qry_trees, pair):
    """Head of `extract_structures`, chess/get_structures.py:99-113 (numpy code restated verbatim):
    gather with default weight 0, OR = log(query / reference) with NaN / +-inf -> 0, 0.5-std thresholds.
    Returns (positive, negative, std, mean(positive), mean(negative))."""
    _, rr, qr = pair
    reference, _ = sub_matrix_from_edges_dict(ref_edges, ref_trees, rr, default_weight=0.)
    query, _ = sub_matrix_from_edges_dict(qry_edges, qry_trees, qr, default_weight=0.)
    with np.errstate(divide="ignore", invalid="ignore"):
        or_matrix = np.log(np.divide(query, reference))
    or_matrix[np.isnan(or_matrix)] = 0.
    or_matrix[or_matrix == -np.inf] = 0.
    or_matrix[or_matrix == np.inf] = 0.
    std = np.std(or_matrix)
    positive = np.where(or_matrix > (0.5 * std), or_matrix, 0.)
    negative = np.abs(np.where(or_matrix < -(0.5 * std), or_matrix, 0.))
    return positive, negative, std, np.mean(positive), np.mean(negative)


def sim(ref_edges, ref_regions, qry_edges, qry_regions, pairs, background_regions=None,
        background_query=False, limit_background=False, **kw):
    """Serial equivalent of `Chess.sim` after loading (bin/chess:189-353).

    Returns (results {id: [ssim, sn, z]}, bg {id: (z_bg, p_bg)} or None).
    """
    rt, qt = region_trees(ref_regions), region_trees(qry_regions)
    raw = {}
    for pid, s, sn in compare_chunk(pairs, ref_edges, rt, qry_edges, qt, True, **kw):
        raw[pid] = [s, sn]
    results = run_zscores(raw)
    if background_regions is None and not background_query:
        return results, None
    bg = {}
    for pid, rr, qr in pairs:
        if background_query:
            background_regions = background_regions_from_query(qry_regions, qr, limit_background)
        bpairs = [(k, rr, b) for k, b in enumerate(background_regions)]
        vals = [s for _, s, _ in compare_chunk(bpairs, ref_edges, rt, qry_edges, qt, False, **kw)]
        bg[pid] = background_stats(results[pid][0], vals)
    return results, bg
