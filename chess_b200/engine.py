"""Device-side state and the batched comparison engine.

PyTorch is used for device memory, streams and (in bench.py) torch.distributed
only; all arithmetic runs in the hand-written kernels behind the C ABI
(include/chess_b200.h).  Nothing here falls back to the CPU.

Data layout in HBM (one per sample, per GPU)
    COO      src int32[nnz], sink int32[nnz], value float64[nnz]   (src <= sink)
    band     float64[n_bins, 2W-1], band[i, (j-i)+W-1] = value(i, j) for |j-i| < W,
             absent cells = default value (1 for O/E).  Row r of the submatrix
             that starts at bin i0 is the contiguous run
             band.flat[i0*(2W-1) + (W-1) + r*(2W-2) : ... + n], i.e. every
             submatrix is a rectangular 2-D view with pitch 2W-2 -- rows load
             coalesced and no per-pair copy is ever made.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _native
from ._native import PAIR_DTYPE, lib, check
from .distributed import shard_bounds

_contexts = {}
MAX_BINS = 1024   # largest region, in bins, the comparison kernels take (csrc/compare_generic.cu)


def get_context(device=0):
    """The (cached) library context of one GPU."""
    device = int(device)
    if device not in _contexts:
        _contexts[device] = _native.Context(device)
    return _contexts[device]


def visible_devices():
    """GPU indices to use: $CHESS_B200_DEVICES ("0,1,2") or every visible GPU."""
    env = os.environ.get("CHESS_B200_DEVICES")
    if env:
        return [int(v) for v in env.split(",") if v.strip() != ""]
    if not torch.cuda.is_available():
        raise _native.ChessNativeError("chess_b200 needs a CUDA device (B200); none is visible "
                                       "and there is no CPU fallback")
    return list(range(torch.cuda.device_count()))


def plan_devices(devices, pixel_pairs, per_gpu_rate=2.5e11, context_cost_s=1.0):
    """How many of ``devices`` a run of ``pixel_pairs`` (comparisons x bins^2) should use.  A CUDA context costs
    about a second to create (and contexts are created one after the other by the driver), a GPU compares
    roughly 2.5e11 pixel pairs per second: a further GPU pays off once it saves more than it costs.  At least
    two when there are two (the two samples of a run are loaded side by side)."""
    devices = list(devices)
    if len(devices) <= 2:
        return devices
    seconds = float(pixel_pairs) / per_gpu_rate
    n = 2
    while n < len(devices) and seconds / n - seconds / (n + 1) > context_cost_s:
        n += 1
    return devices[:n]


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class DeviceTriples(object):
    """Canonical COO triples that live on a GPU (what the device text parser returns): int32 src / sink and
    float64 weight tensors on ``device``.  Lets the loaders hand a matrix from the parser to the O/E kernels and
    on to the contact store without a round trip through host memory."""

    def __init__(self, src, sink, value, device):
        self.src, self.sink, self.value, self.device = src, sink, value, int(device)

    def __len__(self):
        return int(self.value.numel())

    def host(self):
        return (self.src.cpu().numpy().astype(np.int64), self.sink.cpu().numpy().astype(np.int64),
                self.value.cpu().numpy())


def coo_facts(d_src, d_sink, device):
    """(min index, max index, #src > sink, #not strictly after the predecessor) of a device COO list."""
    device = int(device)
    ctx = get_context(device)
    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        out = torch.empty(4, dtype=torch.int64, device=dev)
        check(lib.chess_coo_check(ctx.handle, _ptr(d_src), _ptr(d_sink), int(d_src.numel()), _ptr(out),
                                  _stream_ptr(dev)), ctx.handle, "chess_coo_check")
        return [int(v) for v in out.cpu().tolist()]


class DeviceContacts(object):
    """The edges store of one sample: COO triples plus the HBM-resident band.

    Replaces the ``{source: {sink: weight}}`` dict built by
    chess/helpers.py:342-391.  ``store[source][sink]`` still works (host-side
    CSR view) so code written against the dict keeps running; the comparison
    kernels read the band.

    The triples live in host arrays, or (``from_device``) on the GPU that parsed
    and normalised them -- then the host copy is made only if something asks for
    it.  Other GPUs receive what they need from there: the whole list, or, for a
    sorted list, the contiguous slice of the bin range their share of the pair
    list touches (peer copy), and build the band of that range only.
    """

    def __init__(self, src, sink, value, n_bins=None, fill=1.0):
        src = np.ascontiguousarray(src, dtype=np.int64)
        sink = np.ascontiguousarray(sink, dtype=np.int64)
        value = np.ascontiguousarray(value, dtype=np.float64)
        swap = src > sink
        if swap.any():
            src, sink = np.where(swap, sink, src), np.where(swap, src, sink)
        if len(src) and int(src.min()) < 0:
            raise ValueError("negative contact index")
        self.sorted_unique = True
        if len(src) > 1:
            # a (source, sink) listed more than once: the reference's dict of dicts keeps the LAST occurrence
            # in input order (chess/helpers.py:353-356); keep exactly that one, so that the band scatter has
            # one writer per cell (deterministic, and both mirror cells agree)
            key = src * (int(sink.max()) + 1) + sink
            if not bool(np.all(key[1:] > key[:-1])):   # sorted input (the usual dump order): nothing repeats
                self.sorted_unique = False
                order = np.argsort(key, kind="stable")
                ks = key[order]
                dup = ks[1:] == ks[:-1]
                if dup.any():
                    last = np.ones(len(ks), dtype=bool)
                    last[:-1] = ~dup
                    keep = np.sort(order[last])
                    src, sink, value = src[keep], sink[keep], value[keep]
        self._host = (src, sink, value)
        self._home = None
        hi = int(sink.max()) + 1 if len(sink) else 0
        self.n_bins = int(n_bins) if n_bins is not None else hi
        if self.n_bins < hi:
            raise ValueError("contact indices exceed the number of regions")
        self._init_caches(fill)

    def _init_caches(self, fill):
        self.fill = float(fill)
        self._bands = {}   # device -> (W, fill, bin_lo, bin_hi, tensor)
        self._index = {}   # device -> (nd_lo, nd_hi, pmax, pmin, rowpre) tensors for the current band
        self._coo = {}     # device -> (src32, sink32, value) tensors: the whole list
        self._coo_part = {}  # device -> (bin_lo, bin_hi, (src32, sink32, value)): a slice of a sorted list
        self._csr = None

    @classmethod
    def from_device(cls, triples, n_bins, fill=1.0):
        """Store around triples that already live on a GPU (DeviceTriples).  A list that is not sorted and
        duplicate-free goes through the host constructor (which keeps the last of repeated entries)."""
        lo, hi, noncanon, unordered = coo_facts(triples.src, triples.sink, triples.device) if len(triples) else (0, -1, 0, 0)
        if noncanon or unordered:
            return cls(*triples.host(), n_bins=n_bins, fill=fill)
        if len(triples) and (lo < 0 or hi >= int(n_bins)):
            raise ValueError("contact indices exceed the number of regions")
        self = cls.__new__(cls)
        self._host = None
        self._home = triples.device
        self.sorted_unique = True
        self.n_bins = int(n_bins)
        self._init_caches(fill)
        self._coo[triples.device] = (triples.src, triples.sink, triples.value)
        return self

    # -- host view of the triples (made on demand when they live on a GPU)
    def _host_arrays(self):
        if self._host is None:
            s, t, v = self._coo[self._home]
            self._host = (s.cpu().numpy().astype(np.int64), t.cpu().numpy().astype(np.int64), v.cpu().numpy())
        return self._host

    src = property(lambda self: self._host_arrays()[0])
    sink = property(lambda self: self._host_arrays()[1])
    value = property(lambda self: self._host_arrays()[2])

    def __len__(self):
        starts, _, _ = self._rows()
        return int(np.count_nonzero(np.diff(starts)))

    def n_contacts(self):
        return len(self._host[2]) if self._host is not None else int(self._coo[self._home][2].numel())

    def ensure_bins(self, n_bins):
        """Make the store cover at least ``n_bins`` bins.  Bins after the last one with a contact
        (chrY, chrM, telomeric or unplaced bins are routinely empty) hold the fill value like every
        absent cell (chess/helpers.py:297); a store sized by its largest contact index alone would
        leave regions that reach into them outside every per-sample buffer."""
        n_bins = int(n_bins)
        if n_bins > self.n_bins:
            self.n_bins = n_bins
            self._bands.clear()
            self._index.clear()
            self._csr = None

    def check_span(self, first, n, what="region"):
        """Raise unless every span [first, first + n) with n > 0 lies inside the store."""
        first = np.asarray(first, dtype=np.int64)
        n = np.asarray(n, dtype=np.int64)
        live = n > 0
        if live.any() and (int(first[live].min()) < 0 or int((first[live] + n[live]).max()) > self.n_bins):
            raise ValueError("{} bins [{}, {}) lie outside the contact store of {} bins: the regions file "
                             "does not match the contact matrix".format(
                                 what, int(first[live].min()), int((first[live] + n[live]).max()), self.n_bins))

    # -- dict-like view (compatibility with code indexing the reference's dict)
    def _rows(self):
        if self._csr is None:
            src, sink, value = self._host_arrays()
            order = np.lexsort((sink, src))
            s, t, w = src[order], sink[order], value[order]
            starts = np.searchsorted(s, np.arange(self.n_bins + 1))
            self._csr = (starts, t, w)
        return self._csr

    def __getitem__(self, source):
        starts, t, w = self._rows()
        if source < 0 or source >= self.n_bins or starts[source] == starts[source + 1]:
            raise KeyError(source)
        lo, hi = starts[source], starts[source + 1]
        return dict(zip(t[lo:hi].tolist(), w[lo:hi].tolist()))

    def __contains__(self, source):
        starts, _, _ = self._rows()
        return 0 <= source < self.n_bins and starts[source] != starts[source + 1]

    def triples(self):
        return self._host_arrays()

    # -- device residency
    def coo_on(self, device, bin_range=None):
        """(src32, sink32, value) tensors on ``device``: the whole list, or -- for a sorted list and a bin range
        -- the contiguous slice with src in [lo, hi) (all a band of those bins can use).  Comes from the host
        arrays, or by peer copy from the GPU that holds the list."""
        device = int(device)
        if device in self._coo:
            return self._coo[device]
        dev = torch.device("cuda", device)
        if bin_range is not None and self.sorted_unique:
            lo, hi = int(bin_range[0]), int(bin_range[1])
            have = self._coo_part.get(device)
            if have is not None and have[0] <= lo and have[1] >= hi:
                return have[2]
            if self._host is not None:
                i0, i1 = np.searchsorted(self._host[0], [lo, hi])
                part = tuple(torch.from_numpy(np.ascontiguousarray(a[i0:i1], dtype=dt)).to(dev) for a, dt in
                             zip(self._host, (np.int32, np.int32, np.float64)))
            else:
                s, t, v = self._coo[self._home]
                cut = torch.searchsorted(s, torch.tensor([lo, hi], dtype=s.dtype, device=s.device)).cpu().tolist()
                part = tuple(a[cut[0]:cut[1]].to(dev, non_blocking=True) for a in (s, t, v))
            self._coo_part[device] = (lo, hi, part)
            return part
        if self._host is not None:
            self._coo[device] = tuple(torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev) for a, dt in
                                      zip(self._host, (np.int32, np.int32, np.float64)))
        else:
            self._coo[device] = tuple(a.to(dev, non_blocking=True) for a in self._coo[self._home])
        return self._coo[device]

    def band_on(self, device, min_w, fill=None, bin_range=None):
        """(W, band tensor, bin_lo): the band on ``device`` with half-width >= min_w covering ``bin_range``
        (default: every bin); rebuilt if the one there is too narrow or covers less."""
        device = int(device)
        fill = self.fill if fill is None else float(fill)
        lo, hi = (0, self.n_bins) if bin_range is None else (max(int(bin_range[0]), 0), min(int(bin_range[1]), self.n_bins))
        if hi <= lo:
            lo, hi = 0, max(min(self.n_bins, 1), 1)
        have = self._bands.get(device)
        if have is not None and have[0] >= min_w and have[1] == fill and have[2] <= lo and have[3] >= hi:
            return have[0], have[4], have[2]
        W = max(int(min_w), 1)
        W = min(((W + 15) // 16) * 16, max(self.n_bins, 1))
        W = max(W, int(min_w))
        ctx = get_context(device)
        dev = torch.device("cuda", device)
        if have is not None:
            del self._bands[device]
            del have
        band = torch.empty((hi - lo) * (2 * W - 1), dtype=torch.float64, device=dev)
        whole = lo == 0 and hi == self.n_bins
        s, t, v = self.coo_on(device, None if whole else (lo, hi))
        with torch.cuda.device(dev):
            check(lib.chess_band_build_range(ctx.handle, _ptr(s), _ptr(t), _ptr(v), int(v.numel()), lo, hi - lo, W,
                                             fill, _ptr(band), _stream_ptr(dev)),
                  ctx.handle, "chess_band_build_range")
        self._bands[device] = (W, fill, lo, hi, band)
        self._index.pop(device, None)
        return W, band, lo

    def sample_on(self, device, min_w, fill=None, bin_range=None, with_origin=False):
        """(W, band tensor, chess_sample) on ``device``: the band plus the per-bin index of the
        nearest non-default entries (built once per band, lets the kernels predict masked bins).
        With ``bin_range`` the band covers those bins only; ``with_origin`` appends the first bin of the
        band (pair offsets into it count from there)."""
        device = int(device)
        W, band, bin_lo = self.band_on(device, min_w, fill, bin_range)
        _, fill_v, lo, hi, _ = self._bands[device]
        n_local = hi - lo
        have = self._index.get(device)
        if have is None:
            ctx = get_context(device)
            dev = torch.device("cuda", device)
            nd_lo = torch.empty(n_local, dtype=torch.int32, device=dev)
            nd_hi = torch.empty(n_local, dtype=torch.int32, device=dev)
            pmax = torch.empty(n_local * W, dtype=torch.float64, device=dev)
            pmin = torch.empty(n_local * W, dtype=torch.float64, device=dev)
            rowpre = torch.empty(n_local * 2 * W, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                stream = _stream_ptr(dev)
                check(lib.chess_band_index_build(ctx.handle, _ptr(band), n_local, W, fill_v, _ptr(nd_lo),
                                                 _ptr(nd_hi), stream), ctx.handle, "chess_band_index_build")
                check(lib.chess_band_extrema_build(ctx.handle, _ptr(band), n_local, W, _ptr(pmax), _ptr(pmin),
                                                   stream), ctx.handle, "chess_band_extrema_build")
                check(lib.chess_band_prefix_build(ctx.handle, _ptr(band), n_local, W, _ptr(rowpre), stream),
                      ctx.handle, "chess_band_prefix_build")
            have = (nd_lo, nd_hi, pmax, pmin, rowpre)
            self._index[device] = have
        sample = _native.chess_sample(band.data_ptr(), have[0].data_ptr(), have[1].data_ptr(),
                                      have[2].data_ptr(), have[3].data_ptr(), n_local, W, 0,
                                      have[4].data_ptr())
        if with_origin:
            return W, band, sample, bin_lo
        if bin_lo != 0:
            raise _native.ChessNativeError("internal error: a partial band was handed to a caller that indexes it "
                                           "with absolute bins")
        return W, band, sample

    def drop_device_copies(self):
        self._host_arrays()          # the triples must survive somewhere
        self._home = None
        self._bands.clear()
        self._index.clear()
        self._coo.clear()
        self._coo_part.clear()

    def submatrix(self, start_ix, n, default_weight=None, device=None):
        """Dense n x n numpy tile starting at bin ``start_ix`` (device gather + D2H)."""
        device = visible_devices()[0] if device is None else int(device)
        self.check_span([start_ix], [n])
        W, band, bin_lo = self.band_on(device, n, default_weight)
        ctx = get_context(device)
        dev = torch.device("cuda", device)
        out = torch.empty((n, n), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib.chess_gather_submatrix(ctx.handle, _ptr(band), (int(start_ix) - bin_lo) * (2 * W - 1) + W - 1,
                                             2 * W - 2, int(n), _ptr(out), _stream_ptr(dev)),
                  ctx.handle, "chess_gather_submatrix")
        return out.cpu().numpy()


def index_bins(region_index):
    """Number of bins a region index addresses (largest ``ix`` + 1); None when it cannot tell."""
    try:
        if isinstance(region_index, dict):
            tops = [int(ix.ixs.max()) for ix in region_index.values() if len(ix)]
        else:
            tops = [int(r.ix) for r in region_index if r.ix is not None]
        return max(tops) + 1 if tops else None
    except (AttributeError, TypeError, ValueError):
        return None


def as_device_contacts(edges, n_bins=None):
    """Accept a DeviceContacts or a ``{src: {sink: w}}`` mapping (uploaded as is).
    ``n_bins``: the number of bins of the regions the store goes with (bins after the
    last contact exist and hold the default value)."""
    if isinstance(edges, DeviceContacts):
        if n_bins is not None:
            edges.ensure_bins(n_bins)
        return edges
    cached = getattr(edges, "_chess_b200_contacts", None)
    if cached is not None:
        if n_bins is not None:
            cached.ensure_bins(n_bins)
        return cached
    src, sink, val = [], [], []
    for a, row in edges.items():
        for b, w in row.items():
            src.append(a)
            sink.append(b)
            val.append(w)
    contacts = DeviceContacts(np.asarray(src, dtype=np.int64), np.asarray(sink, dtype=np.int64),
                              np.asarray(val, dtype=np.float64))
    if n_bins is not None:
        contacts.ensure_bins(n_bins)
    try:
        edges._chess_b200_contacts = contacts
    except AttributeError:
        pass
    return contacts


def region_spans(region_index, regions):
    """(first_bin, n_bins) arrays for a list of regions; n = 0 where not found.

    Batched form of the lookup in sub_regions (chess/helpers.py:233-242).  A
    chromosome missing from the index raises KeyError like the reference.
    """
    n = len(regions)
    first = np.full(n, -1, dtype=np.int64)
    count = np.zeros(n, dtype=np.int64)
    by_chrom = {}
    for k, r in enumerate(regions):
        by_chrom.setdefault(r.chromosome, []).append(k)
    for chrom, ks in by_chrom.items():
        index = region_index[chrom]
        ks = np.asarray(ks)
        qb = np.array([regions[k].start - 1 for k in ks], dtype=np.int64)
        qe = np.array([regions[k].end for k in ks], dtype=np.int64)
        lo, hi = index.span(qb, qe)
        ok = lo >= 0
        first[ks[ok]] = lo[ok]
        count[ks[ok]] = hi[ok] - lo[ok] + 1
    return first, count


def region_spans_columns(region_index, chroms, starts, ends):
    """region_spans for regions given as columns (chromosome names, 1-based starts, ends): no per-region Python."""
    import pandas as pd
    n = len(starts)
    first = np.full(n, -1, dtype=np.int64)
    count = np.zeros(n, dtype=np.int64)
    if n == 0:
        return first, count
    starts = np.asarray(starts, dtype=np.int64)
    ends = np.asarray(ends, dtype=np.int64)
    codes, names = pd.factorize(np.asarray(chroms, dtype=object), sort=False)
    for c, chrom in enumerate(names):
        index = region_index[chrom]          # KeyError for an unknown chromosome, like the reference
        ks = np.flatnonzero(codes == c)
        lo, hi = index.span(starts[ks] - 1, ends[ks])
        ok = lo >= 0
        first[ks[ok]] = lo[ok]
        count[ks[ok]] = hi[ok] - lo[ok] + 1
    return first, count


def build_pair_table(ref_first, ref_n, ref_w, qry_first, qry_n, qry_w, flip):
    """chess_pair records for band-resident samples."""
    n = len(ref_first)
    tab = np.zeros(n, dtype=PAIR_DTYPE)
    tab["ref_off"] = np.where(ref_n > 0, ref_first * (2 * ref_w - 1) + ref_w - 1, 0)   # first = bin - band origin
    tab["qry_off"] = np.where(qry_n > 0, qry_first * (2 * qry_w - 1) + qry_w - 1, 0)
    tab["ref_pitch"] = 2 * ref_w - 2
    tab["qry_pitch"] = 2 * qry_w - 2
    tab["ref_n"] = ref_n
    tab["qry_n"] = qry_n
    tab["flip"] = flip
    return tab


class QueryBackgroundSet(object):
    """The background regions `chess sim --background-query` derives from the query bins (bin/chess:272-298) for
    one region size [and chromosome], enumerated and looked up ON THE DEVICE (chess_background_regions): per GPU
    that asks, (first_bin, n_bins, strand code) tensors in the reference's order.  Needs a RegionTable whose bins
    are sorted and contiguous per chromosome (any BED that `chess pairs` / a matrix goes with)."""

    def __init__(self, table, size, limit_chromosome=None):
        self.table, self.size = table, int(size)
        codes = np.asarray(table.chrom_codes)
        change = np.flatnonzero(np.diff(codes)) + 1 if len(codes) else np.zeros(0, dtype=np.int64)
        firsts = np.concatenate([[0], change]).astype(np.int64) if len(codes) else np.zeros(0, dtype=np.int64)
        runs = codes[firsts] if len(codes) else codes
        self.ok = len(set(runs.tolist())) == len(runs) and bool(np.array_equal(table.ixs, np.arange(len(table)))) \
            and all(bool(np.all(np.diff(table.starts[a:b]) > 0) and np.all(np.diff(table.ends[a:b]) >= 0))
                    for a, b in zip(firsts, list(firsts[1:]) + [len(codes)]))
        self.run_of_code = {int(c): k for k, c in enumerate(runs)}
        self.chrom_first = np.concatenate([firsts, [len(codes)]]).astype(np.int64)
        self.bin_run = np.repeat(np.arange(len(firsts), dtype=np.int32), np.diff(self.chrom_first))
        ends = np.zeros(len(firsts), dtype=np.int64)
        if len(codes):
            np.maximum.at(ends, self.bin_run, table.ends)
        self.chrom_end = ends
        if limit_chromosome is None:
            self.limit = -1
        elif limit_chromosome in table.chrom_names and table.chrom_names.index(limit_chromosome) in self.run_of_code:
            self.limit = self.run_of_code[table.chrom_names.index(limit_chromosome)]
        else:
            self.limit = len(firsts)          # a chromosome the query does not have: no region at all
        self._dev = {}

    def on(self, device):
        """(d_first int64, d_n int32, d_strand int32) on ``device``."""
        device = int(device)
        if device not in self._dev:
            t = self.table
            ctx = get_context(device)
            dev = torch.device("cuda", device)
            n = len(t)
            with torch.cuda.device(dev):
                up = [torch.from_numpy(np.array(a)).to(dev) for a in
                      (t.starts - 1, t.ends, self.bin_run, self.chrom_first, self.chrom_end)]
                d_first = torch.empty(2 * n, dtype=torch.int64, device=dev)
                d_n = torch.empty(2 * n, dtype=torch.int32, device=dev)
                d_strand = torch.empty(2 * n, dtype=torch.int32, device=dev)
                d_count = torch.zeros(1, dtype=torch.int64, device=dev)
                ws = torch.empty(13 * max(n, 1), dtype=torch.uint8, device=dev)
                check(lib.chess_background_regions(ctx.handle, _ptr(up[0]), _ptr(up[1]), _ptr(up[2]), _ptr(up[3]),
                                                   _ptr(up[4]), n, self.size, self.limit, _ptr(d_first), _ptr(d_n),
                                                   _ptr(d_strand), _ptr(d_count), _ptr(ws), _stream_ptr(dev)),
                      ctx.handle, "chess_background_regions")
                m = int(d_count.cpu()[0])
            self._dev[device] = (d_first[:m], d_n[:m], d_strand[:m])
        return self._dev[device]

    def host(self, device):
        return tuple(a.cpu().numpy() for a in self.on(device))


class CompareEngine(object):
    """Batched replacement for the worker pool of bin/chess:189-225.

    Holds the two samples resident on one or more GPUs and compares arbitrary
    lists of (ID, reference_region, query_region) pairs.  Pairs are sharded in
    contiguous ranges over the GPUs (no collective; results are concatenated
    on the host in input order).
    """

    def __init__(self, reference_edges, reference_regions, query_edges, query_regions, devices=None,
                 placement="auto"):
        """``placement``: "auto" = with several GPUs each one holds the band of the bin range its share of the pair
        list reads, with one GPU the whole band; "range" / "whole" force either (tests, A/B runs)."""
        self.placement = placement
        self.ref = as_device_contacts(reference_edges, index_bins(reference_regions))
        self.qry = as_device_contacts(query_edges, index_bins(query_regions))
        self.ref_index = reference_regions
        self.qry_index = query_regions
        self.devices = list(devices) if devices is not None else visible_devices()[:1]
        if not self.devices:
            raise _native.ChessNativeError("no CUDA device selected")

    def compare_spans(self, ref_first, ref_n, qry_first, qry_n, flip, default_value=1, **params):
        """Compare pairs given as bin spans -> (ssim, sn, status) numpy arrays."""
        ref_first = np.asarray(ref_first, dtype=np.int64)
        ref_n = np.asarray(ref_n, dtype=np.int64)
        qry_first = np.asarray(qry_first, dtype=np.int64)
        qry_n = np.asarray(qry_n, dtype=np.int64)
        flip = np.asarray(flip, dtype=np.int32)
        total = len(ref_first)
        ssim = np.full(total, np.nan)
        sn = np.full(total, np.nan)
        status = np.zeros(total, dtype=np.int32)
        if total == 0:
            return ssim, sn, status
        self.ref.check_span(ref_first, ref_n, "reference region")
        self.qry.check_span(qry_first, qry_n, "query region")
        max_n = int(max(ref_n.max(), qry_n.max(), 1))
        # the band is symmetric by construction; equal sizes let the kernels skip the resize path
        flags = _native.FLAG_SYMMETRIC
        equal = (ref_n == qry_n) | (ref_n <= 0) | (qry_n <= 0)
        if bool(np.all(equal)):
            flags |= _native.FLAG_EQUAL_SIZES
        elif equal.mean() >= 0.9:
            flags |= _native.FLAG_MOSTLY_EQUAL_SIZES
        p = _native.make_params(default_value=default_value, flags=flags, **params)
        ndev = min(len(self.devices), total)
        bounds = shard_bounds(total, ndev)
        pending = []

        def span_range(first, n):
            live = n > 0
            if not live.any():
                return (0, 1)
            return int(first[live].min()), int((first[live] + n[live]).max())

        for k in range(ndev):
            device = self.devices[k]
            lo, hi = int(bounds[k]), int(bounds[k + 1])
            if hi == lo:
                continue
            ctx = get_context(device)
            dev = torch.device("cuda", device)
            with torch.cuda.device(dev):
                # one GPU: the whole band (later calls reuse it).  Several: every GPU holds the stretch of the
                # genome its contiguous range of pairs reads (pairs files run chromosome by chromosome)
                ranged = self.placement == "range" or (self.placement == "auto" and ndev > 1)
                r_rng = span_range(ref_first[lo:hi], ref_n[lo:hi]) if ranged else None
                q_rng = span_range(qry_first[lo:hi], qry_n[lo:hi]) if ranged else None
                rw, rband, rsample, r0 = self.ref.sample_on(device, max_n, default_value, r_rng, with_origin=True)
                qw, qband, qsample, q0 = self.qry.sample_on(device, max_n, default_value, q_rng, with_origin=True)
                tab = build_pair_table(ref_first[lo:hi] - r0, ref_n[lo:hi], rw, qry_first[lo:hi] - q0, qry_n[lo:hi], qw,
                                       flip[lo:hi])
                h_tab = torch.from_numpy(tab.view(np.uint8)).pin_memory()
                d_tab = h_tab.to(dev, non_blocking=True)
                d_ssim = torch.empty(hi - lo, dtype=torch.float64, device=dev)
                d_sn = torch.empty(hi - lo, dtype=torch.float64, device=dev)
                d_status = torch.empty(hi - lo, dtype=torch.int32, device=dev)
                check(lib.chess_compare_band_batch(ctx.handle, C.byref(rsample), C.byref(qsample), _ptr(d_tab),
                                                   hi - lo, max_n, C.byref(p), _ptr(d_ssim), _ptr(d_sn),
                                                   _ptr(d_status), _stream_ptr(dev)),
                      ctx.handle, "chess_compare_band_batch")
                h_ssim = torch.empty(hi - lo, dtype=torch.float64).pin_memory()
                h_sn = torch.empty(hi - lo, dtype=torch.float64).pin_memory()
                h_status = torch.empty(hi - lo, dtype=torch.int32).pin_memory()
                h_ssim.copy_(d_ssim, non_blocking=True)
                h_sn.copy_(d_sn, non_blocking=True)
                h_status.copy_(d_status, non_blocking=True)
            pending.append((dev, lo, hi, h_ssim, h_sn, h_status, (h_tab, d_tab, d_ssim, d_sn, d_status)))
        for dev, lo, hi, h_ssim, h_sn, h_status, _keep in pending:
            torch.cuda.synchronize(dev)
            ssim[lo:hi] = h_ssim.numpy()
            sn[lo:hi] = h_sn.numpy()
            status[lo:hi] = h_status.numpy()
        if (status >= 100).any():
            raise _native.ChessNativeError("internal error: a pair was left unprocessed by the tuned kernel")
        return ssim, sn, status

    def background(self, ref_first, ref_n, ref_strand, pair_ssim, bg_first, bg_n, bg_strand, own=None,
                   default_value=1, **params):
        """z_bg / p_bg of bin/chess:300-329 for pairs given as reference bin spans.

        Every pair's reference region is compared with every background region
        (bin spans of the query sample, strands as integer codes: a comparison is
        flipped iff the codes differ).  Descriptors, SSIM values and the reductions
        stay on the device (`chess_background_batch`); pairs are sharded over the
        GPUs in contiguous ranges.  `own` = (first, n, strand) of every pair's own
        query region: a background region identical to it takes the pair's ssim (the
        reference computes the identical value twice; keeps `ssim <= bg` exact).
        Returns (z_bg, p_bg) numpy arrays.
        """
        ref_first = np.ascontiguousarray(ref_first, dtype=np.int64)
        ref_n = np.ascontiguousarray(ref_n, dtype=np.int32)
        ref_strand = np.ascontiguousarray(ref_strand, dtype=np.int32)
        pair_ssim = np.ascontiguousarray(pair_ssim, dtype=np.float64)
        bg_set = bg_first if isinstance(bg_first, QueryBackgroundSet) else None
        if bg_set is not None:          # regions enumerated on the device; the host only needs their bin counts
            bg_n = bg_set.on(self.devices[0])[1].cpu().numpy()
            bg_first = np.zeros(0, dtype=np.int64)
            bg_strand = np.zeros(0, dtype=np.int32)
        bg_first = np.ascontiguousarray(bg_first, dtype=np.int64)
        bg_n = np.ascontiguousarray(bg_n, dtype=np.int32)
        bg_strand = np.ascontiguousarray(bg_strand, dtype=np.int32)
        if own is not None:
            own = (np.ascontiguousarray(own[0], dtype=np.int64), np.ascontiguousarray(own[1], dtype=np.int32),
                   np.ascontiguousarray(own[2], dtype=np.int32))
        total = len(ref_first)
        z_bg = np.full(total, np.nan)
        p_bg = np.full(total, np.nan)
        if total == 0:
            return z_bg, p_bg
        self.ref.check_span(ref_first, ref_n, "reference region")
        if bg_set is None:
            self.qry.check_span(bg_first, bg_n, "background region")
        if own is not None:
            self.qry.check_span(own[0], own[1], "query region")
        max_n = int(max(ref_n.max(), bg_n.max() if len(bg_n) else 1, 1))
        flags = _native.FLAG_SYMMETRIC
        sizes = np.unique(np.concatenate([ref_n[ref_n > 0], bg_n[bg_n > 0]]))
        if len(sizes) <= 1:
            flags |= _native.FLAG_EQUAL_SIZES
        elif len(bg_n):
            # fraction of the pair x region grid with equal sizes (regions of one length in bp differ by a bin
            # where a chromosome ends): the work-item kernels take those, the retry pass the rest
            rs, rc = np.unique(ref_n, return_counts=True)
            bs, bc = np.unique(bg_n, return_counts=True)
            same = sum(int(c) * int(bc[np.searchsorted(bs, v)]) for v, c in zip(rs, rc)
                       if v > 0 and v in bs)
            if same >= 0.9 * len(ref_n) * len(bg_n):
                flags |= _native.FLAG_MOSTLY_EQUAL_SIZES
        params.pop("compute_sn", None)
        p = _native.make_params(default_value=default_value, flags=flags, compute_sn=False, **params)
        ndev = min(len(self.devices), total)
        bounds = shard_bounds(total, ndev)
        pending = []
        for k in range(ndev):
            device = self.devices[k]
            lo, hi = int(bounds[k]), int(bounds[k + 1])
            if hi == lo:
                continue
            ctx = get_context(device)
            dev = torch.device("cuda", device)
            with torch.cuda.device(dev):
                rw, rband, rsample = self.ref.sample_on(device, max_n, default_value)
                qw, qband, qsample = self.qry.sample_on(device, max_n, default_value)
                up = [torch.from_numpy(a).to(dev) for a in (ref_first[lo:hi], ref_n[lo:hi], ref_strand[lo:hi],
                                                            pair_ssim[lo:hi])]
                up += list(bg_set.on(device)) if bg_set is not None else \
                    [torch.from_numpy(a).to(dev) for a in (bg_first, bg_n, bg_strand)]
                d_z = torch.empty(hi - lo, dtype=torch.float64, device=dev)
                d_p = torch.empty(hi - lo, dtype=torch.float64, device=dev)
                d_exc = torch.zeros(1, dtype=torch.int32, device=dev)
                if own is not None:
                    up += [torch.from_numpy(a[lo:hi]).to(dev) for a in own]
                    o0, o1, o2 = _ptr(up[7]), _ptr(up[8]), _ptr(up[9])
                else:
                    o0 = o1 = o2 = None
                check(lib.chess_background_batch(ctx.handle, C.byref(rsample), C.byref(qsample), _ptr(up[0]),
                                                 _ptr(up[1]), _ptr(up[2]), o0, o1, o2, _ptr(up[3]), hi - lo, _ptr(up[4]),
                                                 _ptr(up[5]), _ptr(up[6]), len(bg_n), max_n, C.byref(p),
                                                 _ptr(d_z), _ptr(d_p), _ptr(d_exc), _stream_ptr(dev)),
                      ctx.handle, "chess_background_batch")
            pending.append((dev, lo, hi, d_z, d_p, d_exc, up))
        exceeds = 0
        for dev, lo, hi, d_z, d_p, d_exc, _keep in pending:
            torch.cuda.synchronize(dev)
            z_bg[lo:hi] = d_z.cpu().numpy()
            p_bg[lo:hi] = d_p.cpu().numpy()
            exceeds += int(d_exc.cpu()[0])
        if exceeds:
            raise ValueError("win_size exceeds image extent. Either ensure that your images are "
                             "at least 7x7; or pass win_size explicitly in the function call, "
                             "with an odd value less than or equal to the smaller side of your images.")
        return z_bg, p_bg

    def compare(self, pairs, compute_sn=True, default_value=1, **params):
        """Compare (ID, reference_region, query_region) tuples.

        Returns (ids, ssim, sn, status).  Where the reference yields
        ``(nan, nan)`` the arrays hold NaN and ``status`` says why
        (``_native.PAIR_*``).  A window larger than a compacted matrix raises
        ValueError -- in the reference skimage raises it and the run aborts
        (chess/sim.py:217-220, bin/chess:220-221).
        """
        from .helpers import PairTable
        if params.get("mappability_cutoff", 0.1) <= 0 and not params.get("keep_unmappable_bins", False):
            # np.union1d(*[]) in chess/sim.py:351
            raise TypeError("union1d() missing 2 required positional arguments: 'ar1' and 'ar2'")
        table = pairs if isinstance(pairs, PairTable) else None
        if table is not None:           # columns: spans and flips without a Python loop over the pairs
            ids = table.ids
            rf, rn = region_spans_columns(self.ref_index, table.ref[0], table.ref[1], table.ref[2])
            qf, qn = region_spans_columns(self.qry_index, table.qry[0], table.qry[1], table.qry[2])
        else:
            pairs = list(pairs)
            ids = [p[0] for p in pairs]
            rf, rn = region_spans(self.ref_index, [p[1] for p in pairs])
            qf, qn = region_spans(self.qry_index, [p[2] for p in pairs])
        big = np.nonzero((rn > MAX_BINS) | (qn > MAX_BINS))[0]
        if len(big):
            k = int(big[0])
            raise ValueError("region pair {} ({} / {}) spans {} bins; the comparison kernels handle at most {} bins "
                             "per region ({} such pair(s) in this run): use smaller regions or a larger bin size"
                             .format(ids[k], pairs[k][1], pairs[k][2], int(max(rn[k], qn[k])), MAX_BINS, len(big)))
        flip = table.flips() if table is not None else \
            np.array([1 if p[1].strand != p[2].strand else 0 for p in pairs], dtype=np.int32)
        self.last_spans = (rf, rn, qf, qn)      # for callers that go on to the background statistics
        ssim, sn, status = self.compare_spans(rf, rn, qf, qn, flip, default_value=default_value,
                                              compute_sn=compute_sn, **params)
        if (status == _native.PAIR_WINDOW_EXCEEDS).any():
            raise ValueError("win_size exceeds image extent. Either ensure that your images are "
                             "at least 7x7; or pass win_size explicitly in the function call, "
                             "with an odd value less than or equal to the smaller side of your images.")
        return ids, ssim, sn, status


def compare_dense_pairs(references, queries, flips, device=None, default_value=1, symmetric=False, **params):
    """Compare already-dense matrix pairs (lists of square float64 arrays).

    The matrices are packed back to back in one device buffer per side and
    described to the kernel as pitch-n views (the dense branch of chess_pair).
    An array object that occurs several times in a list is uploaded once: the
    one-vs-many shape of utils/run_synthetic_test.py:105-110 (one reference
    against the truth and 1000 decoys) costs one copy of the reference.
    `data_range=2.0` reproduces what that script's legacy `compare_ssim` call
    assumes for float images (run_synthetic_test.py:306); the default takes the
    range from the data as `chess sim` does (sim.py:368-374).
    """
    device = visible_devices()[0] if device is None else int(device)
    ctx = get_context(device)
    dev = torch.device("cuda", device)
    n_pairs = len(references)
    tab = np.zeros(n_pairs, dtype=PAIR_DTYPE)

    def pack(mats):
        offs, seen, parts, total = [], {}, [], 0
        for m in mats:
            if m.ndim != 2 or m.shape[0] != m.shape[1]:
                raise ValueError("compare_dense_pairs expects square matrices")
            if id(m) not in seen:
                seen[id(m)] = total
                parts.append(np.ascontiguousarray(m, dtype=np.float64).ravel())
                total += m.size
            offs.append(seen[id(m)])
        return offs, (np.concatenate(parts) if parts else np.zeros(1))

    r_offs, h_ref = pack(references)
    q_offs, h_qry = pack(queries)
    for k, (r, q) in enumerate(zip(references, queries)):
        tab[k] = (r_offs[k], q_offs[k], r.shape[0], q.shape[0], r.shape[0], q.shape[0], 1 if flips[k] else 0, 0)
    max_n = int(max([r.shape[0] for r in references] + [q.shape[0] for q in queries] + [1]))
    if symmetric:       # the caller vouches for symmetric matrices: the upper-triangle kernels may take the batch
        params["flags"] = params.get("flags", 0) | _native.FLAG_SYMMETRIC | (
            _native.FLAG_EQUAL_SIZES if len({m.shape[0] for m in list(references) + list(queries)}) == 1 else 0)
    p = _native.make_params(default_value=default_value, **params)
    ssim = np.empty(n_pairs)
    sn = np.empty(n_pairs)
    status = np.empty(n_pairs, dtype=np.int32)
    with torch.cuda.device(dev):
        d_ref = torch.from_numpy(h_ref).to(dev)
        d_qry = torch.from_numpy(h_qry).to(dev)
        check(lib.chess_compare_batch_host(ctx.handle, _ptr(d_ref), _ptr(d_qry), tab.ctypes.data_as(C.c_void_p),
                                           n_pairs, max_n, C.byref(p), ssim.ctypes.data_as(C.c_void_p),
                                           sn.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p),
                                           _stream_ptr(dev)), ctx.handle, "chess_compare_batch_host")
    return ssim, sn, status
