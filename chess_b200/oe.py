"""Observed / expected normalisation on the device.

Mirror of chess/oe.py (reference 0.3.8): same public names and return shapes.
The per-diagonal sums, the mappable-pair counts and the O/E division run in
the K1 / K2 kernels (chess_b200/csrc/oe.cu); the Python below only lays out
the chromosome table and converts results to the reference's return types.
"""
import numpy as np
import torch

from . import _native
from ._native import lib, check
from .engine import DeviceContacts, DeviceTriples, coo_facts, get_context, visible_devices, _ptr, _stream_ptr
from .helpers import load_regions, _parse_sparse, string_types


def _chromosome_bins(regions):
    """{chromosome: [first_ix, last_ix + 1]}; chess/oe.py:7-14."""
    bins = {}
    for r in regions:
        if bins.get(r.chromosome) is None:
            bins[r.chromosome] = [r.ix, r.ix + 1]
        else:
            bins[r.chromosome][1] = r.ix + 1
    return bins


class _Layout(object):
    """Chromosome table of a regions list in the form the kernels take."""

    def __init__(self, regions):
        self.n_bins = len(regions)
        table = getattr(regions, "chrom_codes", None)
        if table is not None:          # helpers.RegionTable: columns instead of 300 000 objects
            codes = np.asarray(regions.chrom_codes)
            change = np.flatnonzero(np.diff(codes)) + 1 if len(codes) else np.zeros(0, dtype=np.int64)
            firsts = np.concatenate([[0], change]).astype(np.int64) if len(codes) else np.zeros(0, dtype=np.int64)
            run_codes = codes[firsts] if len(codes) else codes
            if len(set(run_codes.tolist())) != len(run_codes):
                name = regions.chrom_names[int(run_codes[np.argmax(np.bincount(run_codes))])]
                raise ValueError("regions of chromosome {} are not contiguous".format(name))
            self.names = [regions.chrom_names[int(c)] for c in run_codes]
            self.chrom_start = np.concatenate([firsts, [self.n_bins]]).astype(np.int64)
            self.bin_chrom = np.repeat(np.arange(len(firsts), dtype=np.int32), np.diff(self.chrom_start))
            return
        self.names = []
        starts = []
        bin_chrom = np.empty(self.n_bins, dtype=np.int32)
        seen = {}
        for i, r in enumerate(regions):
            if not self.names or self.names[-1] != r.chromosome:
                if r.chromosome in seen:
                    raise ValueError("regions of chromosome {} are not contiguous".format(r.chromosome))
                seen[r.chromosome] = len(self.names)
                self.names.append(r.chromosome)
                starts.append(i)
            bin_chrom[i] = len(self.names) - 1
        starts.append(self.n_bins)
        self.chrom_start = np.asarray(starts, dtype=np.int64)
        self.bin_chrom = bin_chrom


def _as_triples(sparse_matrix):
    if isinstance(sparse_matrix, DeviceContacts):
        return sparse_matrix.triples()
    if isinstance(sparse_matrix, tuple) and len(sparse_matrix) == 3 and isinstance(sparse_matrix[0], np.ndarray):
        a, b, w = sparse_matrix
    else:
        rows = list(sparse_matrix)
        if rows:
            a, b, w = zip(*rows)
        else:
            a, b, w = (), (), ()
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    finite = np.isfinite(w)  # oe.py:103-104
    if not finite.all():
        a, b, w = a[finite], b[finite], w[finite]
    return a, b, w


def _device_expected(regions, triples, device=None, want_oe=False, apply_log2=False, keep_oe_on_device=False):
    """Run K1 (and optionally the O/E division of K2) on one GPU.  ``triples``: host arrays (a, b, w) or a
    DeviceTriples (then nothing is uploaded, and with ``keep_oe_on_device`` the O/E values stay there too)."""
    layout = _Layout(regions)
    n = layout.n_bins
    if n >= 2 ** 31:
        raise ValueError("more than 2^31 - 1 regions are not supported")
    on_device = isinstance(triples, DeviceTriples)
    if on_device:
        device = triples.device
        nnz = len(triples)
        lo, hi, _, _ = coo_facts(triples.src, triples.sink, device) if nnz else (0, -1, 0, 0)
    else:
        device = visible_devices()[0] if device is None else int(device)
        a, b, w = triples
        nnz = len(w)
        lo, hi = (int(min(a.min(), b.min())), int(max(a.max(), b.max()))) if nnz else (0, -1)
    # the kernels index marginals, chromosome table and expected values with these: a matrix that does not
    # go with the regions file must fail here, as it does in the reference (marginals[source], oe.py:107)
    if nnz and (lo < 0 or hi >= n):
        raise IndexError("matrix index {} is out of bounds for {} regions: the sparse matrix does not match "
                         "the regions file".format(hi if lo >= 0 else lo, n))
    ctx = get_context(device)
    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        stream = _stream_ptr(dev)
        if on_device:
            d_a, d_b, d_w = triples.src, triples.sink, triples.value
        else:
            d_a = torch.from_numpy(a.astype(np.int32)).to(dev)
            d_b = torch.from_numpy(b.astype(np.int32)).to(dev)
            d_w = torch.from_numpy(np.ascontiguousarray(w)).to(dev)
        d_chrom = torch.from_numpy(layout.bin_chrom).to(dev)
        d_start = torch.from_numpy(layout.chrom_start).to(dev)
        d_exp = torch.empty(n, dtype=torch.float64, device=dev)
        d_sum = torch.empty(n, dtype=torch.float64, device=dev)
        d_inter = torch.empty(1, dtype=torch.float64, device=dev)
        d_map = torch.empty(n, dtype=torch.uint8, device=dev)
        d_poss = torch.empty(n, dtype=torch.int64, device=dev)
        ws_bytes = int(lib.chess_expected_workspace_bytes(n))
        d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib.chess_expected_from_coo(ctx.handle, _ptr(d_a), _ptr(d_b), _ptr(d_w), nnz, _ptr(d_chrom),
                                          _ptr(d_start), len(layout.names), n, _ptr(d_exp), _ptr(d_inter),
                                          _ptr(d_map), _ptr(d_poss), _ptr(d_sum), _ptr(d_ws), ws_bytes, stream),
              ctx.handle, "chess_expected_from_coo")
        out = {"layout": layout, "device": device}
        if not keep_oe_on_device:
            out.update(expected=d_exp.cpu().numpy(), inter=float(d_inter.cpu()[0]),
                       mappable=d_map.cpu().numpy().astype(bool), possible=d_poss.cpu().numpy(),
                       diag_sum=d_sum.cpu().numpy())
        if want_oe:
            d_oe = torch.empty(nnz, dtype=torch.float64, device=dev)
            check(lib.chess_oe_values(ctx.handle, _ptr(d_a), _ptr(d_b), _ptr(d_w), nnz, _ptr(d_chrom),
                                      _ptr(d_start), _ptr(d_exp), _ptr(d_inter), 1 if apply_log2 else 0,
                                      _ptr(d_oe), stream), ctx.handle, "chess_oe_values")
            if keep_oe_on_device:
                out["oe_triples"] = DeviceTriples(d_a, d_b, d_oe, device)
            else:
                out["oe"] = d_oe.cpu().numpy()
    return out


def possible_contacts(regions, mappability):
    """(intra_total, {chromosome: totals}, inter_total); chess/oe.py:17-75."""
    device = visible_devices()[0]
    layout = _Layout(regions)
    ctx = get_context(device)
    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        d_map = torch.from_numpy(np.asarray(mappability, dtype=bool).astype(np.uint8)).to(dev)
        d_start = torch.from_numpy(layout.chrom_start).to(dev)
        d_poss = torch.empty(layout.n_bins, dtype=torch.int64, device=dev)
        check(lib.chess_possible_contacts(ctx.handle, _ptr(d_map), _ptr(d_start), len(layout.names),
                                          layout.n_bins, _ptr(d_poss), _stream_ptr(dev)),
              ctx.handle, "chess_possible_contacts")
        poss = d_poss.cpu().numpy()
    return _totals(layout, poss, np.asarray(mappability, dtype=bool))


def _totals(layout, poss, mappable):
    cs = layout.chrom_start
    sizes = np.diff(cs)
    intra = np.zeros(int(sizes.max()) if len(sizes) else 0, dtype=np.int64)
    per_chrom, n_map = {}, []
    for c, name in enumerate(layout.names):
        seg = poss[cs[c]:cs[c + 1]]
        per_chrom[name] = [int(v) for v in seg]
        intra[:len(seg)] += seg
        n_map.append(int(mappable[cs[c]:cs[c + 1]].sum()))
    inter, seen = 0, 0
    for v in n_map:
        inter += seen * v
        seen += v
    return [int(v) for v in intra], per_chrom, inter


def expected_values(regions, sparse_matrix, selected_chromosome=None):
    """(intra_expected, {chromosome: expected}, inter_expected); chess/oe.py:78-146."""
    res = _device_expected(regions, _as_triples(sparse_matrix))
    layout, cs = res["layout"], res["layout"].chrom_start
    chrom_expected = {name: [float(v) for v in res["expected"][cs[c]:cs[c + 1]]]
                      for c, name in enumerate(layout.names)}
    if selected_chromosome is not None:
        return chrom_expected[selected_chromosome]
    sizes = np.diff(cs)
    max_d = int(sizes.max()) if len(sizes) else 0
    sums = np.zeros(max_d)
    counts = np.zeros(max_d, dtype=np.int64)
    for c in range(len(layout.names)):
        sums[:sizes[c]] += res["diag_sum"][cs[c]:cs[c + 1]]
        counts[:sizes[c]] += res["possible"][cs[c]:cs[c + 1]]
    intra = [float(sums[d] / counts[d]) if counts[d] > 0 else 0.0 for d in range(max_d)]
    return intra, chrom_expected, res["inter"]


def _oe_generator(sparse_matrix, ix_to_chromosome, intra_expected, inter_expected, log=False):
    """Per-edge O/E from given expected tables; chess/oe.py:149-166 (host glue,
    kept for callers that bring their own expected values)."""
    for source, sink, weight in sparse_matrix:
        if ix_to_chromosome[source] == ix_to_chromosome[sink]:
            e = intra_expected[ix_to_chromosome[source]][sink - source]
        else:
            e = inter_expected
        try:
            oe = weight / e
        except ZeroDivisionError:
            oe = 1
        if log:
            # NB deliberate bug-compatibility: chess/oe.py:163-166 has no `else`, so log=True yields every edge
            # twice (log2 value, then the plain value).  No caller sets it (oe.py:181); the device path
            # (`observed_expected_arrays(apply_log2=True)`) yields each edge once, log2 only.
            yield source, sink, np.log2(oe)
        yield source, sink, oe


def observed_expected_arrays(regions, triples, apply_log2=False, device=None):
    """O/E of canonical COO triples -> DeviceContacts (array form used by the CLI).  Triples that already
    live on a GPU (DeviceTriples, from the device text parser) are normalised there and stay there."""
    if isinstance(triples, DeviceTriples):
        res = _device_expected(regions, triples, want_oe=True, apply_log2=apply_log2, keep_oe_on_device=True)
        return DeviceContacts.from_device(res["oe_triples"], n_bins=len(regions))
    a, b, w = _as_triples(triples)
    res = _device_expected(regions, (a, b, w), device=device, want_oe=True, apply_log2=apply_log2)
    return DeviceContacts(a, b, res["oe"], n_bins=len(regions))


def observed_expected(regions, sparse_matrix):
    """(regions, [(source, sink, oe), ...]); chess/oe.py:169-181."""
    ix_converter = None
    if isinstance(regions, string_types):
        regions, ix_converter, _ = load_regions(regions)
    if isinstance(sparse_matrix, string_types):
        sparse_matrix = _parse_sparse(sparse_matrix, ix_converter)
    a, b, w = _as_triples(sparse_matrix)
    res = _device_expected(regions, (a, b, w), want_oe=True)
    return regions, list(zip(a.tolist(), b.tolist(), res["oe"].tolist()))
