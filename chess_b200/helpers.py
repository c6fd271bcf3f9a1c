"""Regions, pairs, sparse matrices and the submatrix gather.

Host-side mirror of chess/helpers.py (reference 0.3.8): same public names,
argument meaning and error behaviour.  What changes underneath:

* the dict-of-dict ``edges`` store becomes :class:`DeviceContacts`, a COO copy
  plus a symmetric O/E band resident in HBM (``engine.py``);
* ``region_interval_trees`` returns per-chromosome :class:`RegionIndex` objects
  (sorted arrays + bisection) that answer the same half-open overlap query as
  ``intervaltree`` and also answer it for whole arrays of regions at once;
* ``sub_matrix_from_edges_dict`` gathers on the device.

File parsing stays on the host (pandas' C parser for the matrix file).
"""
import gzip
import io
import logging
import os
from collections import defaultdict

import numpy as np

logger = logging.getLogger(__name__)

string_types = (str,)


class GenomicRegion(object):
    """A genomic interval (chess/helpers.py:12-150).

    ``start``/``end`` in base pairs, ``ix`` the bin index of the region in its
    matrix, ``strand`` whatever the input file said ('+', '-', None).
    Equality compares chromosome, start and end only (helpers.py:126-136).
    """

    def __init__(self, start, end, chromosome=None, ix=None, strand=None):
        self.start = start
        self.end = end
        self.chromosome = chromosome
        self.ix = ix
        self.strand = strand

    @classmethod
    def from_string(cls, region_string):
        """``<chromosome>[:<start>-<end>[:<strand>]]`` -> region (helpers.py:43-94)."""
        text = "".join(region_string.split())
        fields = text.split(":")
        if len(fields) > 3:
            raise ValueError("Genomic range string must be of the form "
                             "<chromosome>[:<start>-<end>:[<strand>]]")
        chromosome = fields[0] if fields else None
        start = end = None
        if len(fields) > 1 and fields[1] != "":
            bounds = fields[1].split("-")
            try:
                start = int(bounds[0])
            except ValueError:
                raise ValueError("Start of genomic range must be integer")
            if len(bounds) > 1:
                try:
                    end = int(bounds[1])
                except ValueError:
                    raise ValueError("End of genomic range must be integer")
                if not end > start:
                    raise ValueError("The end coordinate must be bigger than the start.")
        return cls(start=start, end=end, chromosome=chromosome)

    def overlaps(self, region):
        if isinstance(region, str):
            region = GenomicRegion.from_string(region)
        return (region.chromosome == self.chromosome and region.start <= self.end
                and region.end >= self.start)

    def contains(self, region):
        if isinstance(region, str):
            region = GenomicRegion.from_string(region)
        return (region.chromosome == self.chromosome and region.start >= self.start
                and region.end <= self.end)

    def _equals(self, region):
        return (region.chromosome == self.chromosome and region.start == self.start
                and region.end == self.end)

    def __eq__(self, other):
        return self._equals(other)

    def __ne__(self, other):
        return not self._equals(other)

    __hash__ = None

    def __str__(self):
        return "{}:{}-{}:{}".format(self.chromosome, self.start, self.end, self.strand)

    def copy(self):
        return GenomicRegion(chromosome=self.chromosome, start=self.start, end=self.end)


def is_gzipped(file_name):
    return file_name.endswith(".gz") or file_name.endswith(".gzip")


def _open(file_name, mode="r"):
    if is_gzipped(file_name):
        return io.TextIOWrapper(gzip.open(file_name, mode=mode))
    return open(file_name, mode=mode)


def load_regions(file_name, sep=None, ignore_ix=False):
    """BED -> (regions, ix_converter, ix2reg); chess/helpers.py:165-206.

    The BED start is kept as is (0-based), the bin index is the line number,
    and a 4th column other than '.' names the bin for HiC-Pro style matrices
    (must be unique, and must parse as an integer -- helpers.py:204).
    """
    regions, ix2reg, ix_converter = [], {}, None
    with _open(file_name, "r") as fh:
        for i, line in enumerate(fh):
            fields = line.rstrip().split(sep)
            if len(fields) <= 2:
                continue
            chromosome, start, end = fields[0], int(fields[1]), int(fields[2])
            ix2reg[i] = [chromosome, start, end]
            if not ignore_ix and len(fields) > 3 and fields[3] != ".":
                if ix_converter is None:
                    ix_converter = dict()
                if fields[3] in ix_converter:
                    raise ValueError("name column in region BED must only contain unique "
                                     "values! ({})".format(fields[3]))
                ix_converter[fields[3]] = i
            regions.append(GenomicRegion(chromosome=chromosome, start=start, end=end, ix=i))
    if ix_converter is not None:
        rev = {str(v): int(k) for k, v in ix_converter.items()}
        ix2reg = {rev[str(k)]: v for k, v in ix2reg.items()}
    return regions, ix_converter, ix2reg


class RegionTable(object):
    """The regions of a BED file as columns (chromosome codes, starts, ends, bin indices) that behave like the
    list of GenomicRegion objects ``load_regions`` returns: ``len``, indexing and iteration make the objects on
    demand.  What the batched engine needs -- overlap queries for whole pair lists, the chromosome table of
    the O/E kernels -- works on the columns, so a 300 000-bin genome costs no Python loop."""

    def __init__(self, chrom_names, chrom_codes, starts, ends, ixs=None):
        self.chrom_names = list(chrom_names)
        self.chrom_codes = np.ascontiguousarray(chrom_codes, dtype=np.int32)
        self.starts = np.ascontiguousarray(starts, dtype=np.int64)
        self.ends = np.ascontiguousarray(ends, dtype=np.int64)
        self.ixs = np.arange(len(self.starts), dtype=np.int64) if ixs is None else np.ascontiguousarray(ixs, dtype=np.int64)

    def __len__(self):
        return len(self.starts)

    def _make(self, k):
        return GenomicRegion(chromosome=self.chrom_names[int(self.chrom_codes[k])], start=int(self.starts[k]),
                             end=int(self.ends[k]), ix=int(self.ixs[k]))

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._make(i) for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        return self._make(k)

    def __iter__(self):
        for k in range(len(self)):
            yield self._make(k)

    def max_bin_size(self):
        """max(r.end - r.start + 1) over the regions (bin/chess:140-143)."""
        return int((self.ends - self.starts + 1).max()) if len(self) else 0

    def chromosome_ends(self):
        """{chromosome: largest end}."""
        out = {}
        for c, name in enumerate(self.chrom_names):
            sel = self.chrom_codes == c
            if sel.any():
                out[name] = int(self.ends[sel].max())
        return out


class _TableRows(object):
    """Rows of a RegionTable as a sequence of region objects (made on demand)."""

    def __init__(self, table, rows):
        self.table, self.rows = table, rows

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, k):
        return self.table._make(int(self.rows[k]))


def _arrow_table(file_name, string_cols=()):
    """A plain tab-separated table through pyarrow's multithreaded reader (101 354 BEDPE lines: 7 ms against 200 ms
    in pandas' C parser), or None unless the file is exactly that: no comment or blank lines, no spaces, quotes or
    carriage returns, the same number of fields on every line.  Whatever is not plain is left to the slower
    readers, which end in the reference's own loops."""
    try:
        import pyarrow as pa
        import pyarrow.csv as pc
    except ImportError:
        return None
    try:
        if is_gzipped(file_name):
            return None
        with open(file_name, "rb") as fh:
            data = fh.read()
        if not data or data[:1] == b"\n" or b"\n\n" in data or any(c in data for c in (b" ", b"#", b'"', b"\r")):
            return None
        table = pc.read_csv(
            pa.BufferReader(data), read_options=pc.ReadOptions(autogenerate_column_names=True),
            parse_options=pc.ParseOptions(delimiter="\t", quote_char=False, ignore_empty_lines=False),
            convert_options=pc.ConvertOptions(column_types={"f%d" % k: pa.string() for k in string_cols},
                                              strings_can_be_null=False))
        for k, col in enumerate(table.columns):
            if col.null_count:
                return None
            if pa.types.is_string(col.type) and k in string_cols:
                import pyarrow.compute as pcomp
                if pcomp.any(pcomp.equal(pcomp.utf8_length(col), 0)).as_py():
                    return None
        return table
    except Exception:
        return None


def _arrow_codes(col):
    """(names in order of first appearance, int codes) of a string column."""
    enc = col.combine_chunks().dictionary_encode()
    return enc.dictionary.to_numpy(zero_copy_only=False).tolist(), enc.indices.to_numpy(zero_copy_only=False).astype(np.int64)


def _arrow_int64(col):
    import pyarrow as pa
    if not pa.types.is_int64(col.type):
        raise ValueError("not an integer column")
    return col.to_numpy()


def load_regions_table(file_name, ignore_ix=False):
    """``load_regions`` for the batched engine: (RegionTable, ix_converter).  The BED is read by the pandas C
    parser; a file it cannot take as a plain table (ragged lines, blank lines in the middle, a name column) goes
    through ``load_regions`` itself, so behaviour and errors are the reference's in every case."""
    import pandas as pd
    try:
        table = _arrow_table(file_name, string_cols=(0,))
        if table is not None and table.num_columns >= 3:
            import pyarrow as pa
            plain = True
            if table.num_columns > 3 and not ignore_ix:          # a name column: only "." everywhere is plain
                col = table.column(3)
                plain = pa.types.is_string(col.type) and col.unique().to_pylist() == ["."]
            if plain:
                names, codes = _arrow_codes(table.column(0))
                return RegionTable(names, codes, _arrow_int64(table.column(1)), _arrow_int64(table.column(2))), None
    except Exception:
        pass
    try:
        with _open(file_name, "r") as fh:
            frame = pd.read_csv(fh, sep=r"\s+", header=None, comment=None, skip_blank_lines=False, engine="c",
                                dtype={0: str}, keep_default_na=False, quoting=3)     # 3 = csv.QUOTE_NONE
        plain = frame.shape[1] >= 3 and not frame.isnull().any().any() and not (frame.astype(str) == "").any().any()
        if plain and frame.shape[1] > 3 and not ignore_ix:
            plain = bool((frame[3].astype(str) == ".").all())
        if plain:
            codes, names = pd.factorize(frame[0], sort=False)
            return RegionTable(list(names), codes, frame[1].to_numpy(dtype=np.int64),
                               frame[2].to_numpy(dtype=np.int64)), None
    except Exception:
        pass
    regions, ix_converter, _ = load_regions(file_name, ignore_ix=ignore_ix)
    names, codes = [], []
    seen = {}
    for r in regions:
        if r.chromosome not in seen:
            seen[r.chromosome] = len(names)
            names.append(r.chromosome)
        codes.append(seen[r.chromosome])
    return RegionTable(names, codes, [r.start for r in regions], [r.end for r in regions],
                       [r.ix for r in regions]), ix_converter


class PairTable(object):
    """A pair list as columns.  Behaves like the list of (ID, reference_region, query_region) tuples of
    ``load_pairs_for_chroms`` (len, indexing, iteration make the tuples on demand); the engine reads the columns."""

    def __init__(self, ids, ref_chrom, ref_start, ref_end, ref_strand, qry_chrom, qry_start, qry_end, qry_strand):
        self.ids = list(ids)
        self.ref = (np.asarray(ref_chrom, dtype=object), np.asarray(ref_start, dtype=np.int64),
                    np.asarray(ref_end, dtype=np.int64), np.asarray(ref_strand, dtype=object))
        self.qry = (np.asarray(qry_chrom, dtype=object), np.asarray(qry_start, dtype=np.int64),
                    np.asarray(qry_end, dtype=np.int64), np.asarray(qry_strand, dtype=object))

    def __len__(self):
        return len(self.ids)

    def _make(self, k):
        r, q = self.ref, self.qry
        return (self.ids[k],
                GenomicRegion(chromosome=r[0][k], start=int(r[1][k]), end=int(r[2][k]), strand=r[3][k]),
                GenomicRegion(chromosome=q[0][k], start=int(q[1][k]), end=int(q[2][k]), strand=q[3][k]))

    def __getitem__(self, k):
        if isinstance(k, slice):
            idx = range(*k.indices(len(self)))
            return PairTable([self.ids[i] for i in idx], *[c[k] for c in self.ref], *[c[k] for c in self.qry])
        if k < 0:
            k += len(self)
        return self._make(int(k))

    def __iter__(self):
        for k in range(len(self)):
            yield self._make(k)

    def flips(self):
        """1 where the strands of a pair differ (the query is rotated by 180 degrees, chess/sim.py:332-333)."""
        return (self.ref[3] != self.qry[3]).astype(np.int32)

    def max_spans_bp(self):
        """(max reference span, max query span) in bp, as bin/chess:160-167 computes them."""
        return (int((self.ref[2] - self.ref[1] + 1).max()), int((self.qry[2] - self.qry[1] + 1).max()))


def load_pairs_table(file_name, chromosomes, sep=None):
    """``load_pairs_for_chroms`` (chess/helpers.py:483-527) as a PairTable: (pairs, dropped).  Read by the pandas C
    parser; anything that is not a plain 10-column table (comment lines, ragged lines) goes through
    ``load_pairs_for_chroms`` itself, so errors and corner cases are the reference's."""
    import pandas as pd
    try:
        table = _arrow_table(file_name, string_cols=(0, 3, 6, 7, 8, 9)) if sep is None else None
        if table is not None and table.num_columns == 10:
            ids = table.column(6).to_numpy(zero_copy_only=False).tolist()   # (to_pylist is 10 x slower)
            if len(set(ids)) == len(ids):                        # a duplicate ID: the reference's error, below
                known = set(chromosomes)
                (n1, c1), (n2, c2) = _arrow_codes(table.column(0)), _arrow_codes(table.column(3))
                (s1, k1), (s2, k2) = _arrow_codes(table.column(8)), _arrow_codes(table.column(9))
                keep = np.array([n in known for n in n1], dtype=bool)[c1] & np.array([n in known for n in n2], dtype=bool)[c2]
                dropped = int((~keep).sum())
                obj = lambda names, codes: np.asarray(names, dtype=object)[codes[keep]]  # noqa: E731
                ints = [_arrow_int64(table.column(k))[keep] for k in (1, 2, 4, 5)]
                if not keep.all():
                    ids = [v for v, k in zip(ids, keep.tolist()) if k]
                return PairTable(ids, obj(n1, c1), ints[0] + 1, ints[1], obj(s1, k1),
                                 obj(n2, c2), ints[2] + 1, ints[3], obj(s2, k2)), dropped
    except Exception:
        pass
    try:
        with _open(file_name, "r") as fh:
            text = fh.read()
        if "#" in text or sep is not None:
            raise ValueError("not a plain table")
        frame = pd.read_csv(io.StringIO(text), sep=r"\s+", header=None, engine="c", dtype=str, keep_default_na=False,
                            skip_blank_lines=False, quoting=3)                        # 3 = csv.QUOTE_NONE
        if frame.shape[1] != 10 or frame.isnull().any().any() or (frame == "").any().any():
            raise ValueError("not a plain table")
        ids = frame[6]
        if ids.duplicated().any():
            raise ValueError("duplicate")
        keep = (frame[0].isin(chromosomes) & frame[3].isin(chromosomes)).to_numpy()
        dropped = int((~keep).sum())
        f = frame[keep]
        table = PairTable(f[6].tolist(), f[0].to_numpy(), f[1].to_numpy(dtype=np.int64) + 1, f[2].to_numpy(dtype=np.int64),
                          f[8].to_numpy(), f[3].to_numpy(), f[4].to_numpy(dtype=np.int64) + 1,
                          f[5].to_numpy(dtype=np.int64), f[9].to_numpy())
        return table, dropped
    except Exception:
        pass
    res, dropped = load_pairs_for_chroms(file_name, chromosomes, sep=sep)
    if not res:
        return PairTable([], [], [], [], [], [], [], [], []), dropped
    ids = [p[0] for p in res]
    cols = lambda side, attr: [getattr(p[side], attr) for p in res]  # noqa: E731
    return PairTable(ids, cols(1, "chromosome"), cols(1, "start"), cols(1, "end"), cols(1, "strand"),
                     cols(2, "chromosome"), cols(2, "start"), cols(2, "end"), cols(2, "strand")), dropped


class _Hit(object):
    """What iterating an intervaltree slice yields: an object with ``.data``."""
    __slots__ = ("begin", "end", "data")

    def __init__(self, begin, end, data):
        self.begin, self.end, self.data = begin, end, data


class RegionIndex(object):
    """All bins of one chromosome, queryable by half-open overlap.

    Stands in for ``intervaltree.IntervalTree`` as built by
    chess/helpers.py:209-219: every region r contributes [r.start-1, r.end).
    ``index[a:b]`` yields the hits like the tree does (helpers.py:237);
    :meth:`span` answers min/max bin index for arrays of queries at once, which
    is what the batched engine uses instead of one tree query per pair.
    """

    def __init__(self, regions, arrays=None):
        """``regions``: the chromosome's region objects -- or, with ``arrays`` = (starts, ends, ixs), any
        sequence that yields them on demand (RegionTable rows), so that nothing is materialised up front."""
        if arrays is not None:
            self.regions = regions
            begins = np.asarray(arrays[0], dtype=np.int64) - 1
            ends = np.asarray(arrays[1], dtype=np.int64)
            ixs = np.asarray(arrays[2], dtype=np.int64)
        else:
            self.regions = list(regions)
            begins = np.array([r.start - 1 for r in self.regions], dtype=np.int64)
            ends = np.array([r.end for r in self.regions], dtype=np.int64)
            ixs = np.array([r.ix for r in self.regions], dtype=np.int64)
        order = np.argsort(begins, kind="stable")
        self.order = order
        self.begins, self.ends, self.ixs = begins[order], ends[order], ixs[order]
        # the bisection shortcut needs begins, ends and bin indices all increasing
        self.monotone = bool(len(order) < 2 or (np.all(np.diff(self.ends) >= 0)
                                                and np.all(np.diff(self.ixs) > 0)))

    def __len__(self):
        return len(self.begins)

    def __getitem__(self, sl):
        a, b = sl.start, sl.stop
        hit = np.nonzero((self.begins < b) & (self.ends > a))[0]
        return set(_Hit(int(self.begins[k]), int(self.ends[k]), self.regions[self.order[k]]) for k in hit)

    def span(self, q_begin, q_end):
        """(first_ix, last_ix) of the bins overlapping [q_begin, q_end); -1 where none.

        Vectorised restatement of chess/helpers.py:233-242 (hits, then min and
        max of their ``ix``).
        """
        q_begin = np.asarray(q_begin, dtype=np.int64)
        q_end = np.asarray(q_end, dtype=np.int64)
        lo = np.full(q_begin.shape, -1, dtype=np.int64)
        hi = np.full(q_begin.shape, -1, dtype=np.int64)
        if len(self.begins) == 0:
            return lo, hi
        if self.monotone:
            first = np.searchsorted(self.ends, q_begin, side="right")   # first end > q_begin
            last = np.searchsorted(self.begins, q_end, side="left") - 1  # last begin < q_end
            ok = first <= last
            lo[ok] = self.ixs[first[ok]]
            hi[ok] = self.ixs[last[ok]]
            return lo, hi
        for k in range(q_begin.size):  # irregular BED (overlapping / unsorted bins): exact scan
            hit = (self.begins < q_end.flat[k]) & (self.ends > q_begin.flat[k])
            if hit.any():
                lo.flat[k] = self.ixs[hit].min()
                hi.flat[k] = self.ixs[hit].max()
        return lo, hi


def region_interval_trees(regions):
    """{chromosome: RegionIndex}; chess/helpers.py:209-219."""
    if isinstance(regions, RegionTable):
        out = {}
        codes = regions.chrom_codes
        order = np.argsort(codes, kind="stable")
        cuts = np.searchsorted(codes[order], np.arange(len(regions.chrom_names) + 1))
        for c, name in enumerate(regions.chrom_names):
            rows = order[cuts[c]:cuts[c + 1]]
            if len(rows):
                out[name] = RegionIndex(_TableRows(regions, rows), (regions.starts[rows], regions.ends[rows],
                                                                    regions.ixs[rows]))
        return out
    per_chrom = defaultdict(list)
    for region in regions:
        per_chrom[region.chromosome].append(region)
    return {chrom: RegionIndex(rs) for chrom, rs in per_chrom.items()}


def sub_regions(regions, region):
    """Bins of ``regions`` overlapping ``region`` -> (hits, start_ix, end_ix).

    chess/helpers.py:222-268.  ``regions`` is the dict from
    :func:`region_interval_trees` (half-open overlap, KeyError for an unknown
    chromosome, ValueError when nothing overlaps) or a plain list of regions
    (closed overlap, helpers.py:249-251).
    """
    if isinstance(region, string_types):
        region = GenomicRegion.from_string(region)
    if isinstance(regions, dict):
        index = regions[region.chromosome]
        hits = [h.data for h in index[region.start - 1: region.end]]
        if not hits:
            raise ValueError("Region not found in dataset! {}:{}-{}".format(
                region.chromosome, region.start, region.end))
        ixs = [h.ix for h in hits]
        return hits, min(ixs), max(ixs)
    sr, start_ix, end_ix = [], None, None
    for i, r in enumerate(regions):
        if r.chromosome == region.chromosome and r.start <= region.end and r.end >= region.start:
            if start_ix is None:
                start_ix = i
            end_ix = i
            sr.append(r.copy())
        elif end_ix is not None:
            break
    if len(sr) == 0:
        raise ValueError("Region not found in dataset! {}:{}-{}".format(
            region.chromosome, region.start, region.end))
    return sr, start_ix, end_ix


def sub_matrix_regions(hic_matrix, regions, region):
    """chess/helpers.py:271-283."""
    sr, start_ix, end_ix = sub_regions(regions, region)
    if start_ix is None:
        return np.empty((0, 0)), sr
    return np.copy(hic_matrix[start_ix:end_ix + 1, start_ix:end_ix + 1]), sr


def sub_matrix_from_edges_dict(edges, regions, region, default_weight=0.0):
    """Square submatrix overlapping ``region`` -> (ndarray, hits).

    chess/helpers.py:286-306.  ``edges`` is a :class:`DeviceContacts` (or any
    ``{src: {sink: w}}`` mapping, which is uploaded first); the dense tile is
    produced by a gather kernel from the device-resident band whose absent
    cells hold ``default_weight``.
    """
    from .engine import as_device_contacts, index_bins
    sr, start_ix, end_ix = sub_regions(regions, region)
    contacts = as_device_contacts(edges, index_bins(regions))
    return contacts.submatrix(start_ix, end_ix - start_ix + 1, default_weight), sr


def _cuda_visible():
    """File parsing alone (no comparison) also works on a host without a GPU: then pandas parses."""
    import torch
    return torch.cuda.is_available()


def _parse_sparse(file_name, ix_converter=None, sep="\t", want_device=False, device=None):
    """(src, sink, weight) arrays, canonical and finite; the array form of
    edges_from_sparse_matrix (chess/helpers.py:309-339).  ``want_device``: when the text was parsed on a GPU,
    return the triples where they are (engine.DeviceTriples) instead of copying them to host arrays."""
    if ix_converter is None and sep == "\t" and _cuda_visible():
        # index-addressed matrix: the text is parsed on the GPU (csrc/ingest.cu); None = this file
        # needs the host parser below (lone carriage returns / mostly unusual number spellings)
        from .ingest import parse_sparse_text
        parsed = parse_sparse_text(os.path.expanduser(file_name), device=device, return_device=want_device)
        if parsed is not None:
            lo, hi, w, total, n_bad = parsed
            if total and n_bad / total > 0.05:
                logger.warning("Could not import more than 5% of matrix entries ({}/{}), "
                               "as they were not finite".format(n_bad, total))
            if want_device:
                from .engine import DeviceTriples, visible_devices
                return DeviceTriples(lo, hi, w, visible_devices()[0] if device is None else device)
            return lo, hi, w
    import pandas as pd
    with _open(file_name, "r") as fh:
        try:
            frame = pd.read_csv(fh, sep=sep, comment="#", header=None, usecols=[0, 1, 2],
                                dtype={2: np.float64} if ix_converter is None else {0: str, 1: str, 2: np.float64},
                                engine="c", na_filter=True, skip_blank_lines=True,
                                float_precision="round_trip")  # bit-identical to float(str)
        except pd.errors.EmptyDataError:
            return (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64))
    total = len(frame)
    if ix_converter is None:
        a = frame[0].to_numpy(dtype=np.int64)
        b = frame[1].to_numpy(dtype=np.int64)
    else:
        a = frame[0].map(ix_converter).to_numpy()
        b = frame[1].map(ix_converter).to_numpy()
        if np.isnan(a.astype(np.float64)).any() or np.isnan(b.astype(np.float64)).any():
            bad = frame[0][np.isnan(a.astype(np.float64))]
            raise KeyError(bad.iloc[0] if len(bad) else frame[1][np.isnan(b.astype(np.float64))].iloc[0])
        a, b = a.astype(np.int64), b.astype(np.int64)
    w = frame[2].to_numpy(dtype=np.float64)
    finite = np.isfinite(w)
    n_bad = int(total - finite.sum())
    if total and n_bad / total > 0.05:
        logger.warning("Could not import more than 5% of matrix entries ({}/{}), "
                       "as they were not finite".format(n_bad, total))
    a, b, w = a[finite], b[finite], w[finite]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    return lo, hi, w


def edges_from_sparse_matrix(file_name, ix_converter=None, sep="\t"):
    """Generator of canonical (source, sink, weight); chess/helpers.py:309-339."""
    a, b, w = _parse_sparse(file_name, ix_converter, sep)
    for k in range(len(w)):
        yield int(a[k]), int(b[k]), float(w[k])


def load_sparse_matrix(file_name, ix_converter=None, sep="\t"):
    """chess/helpers.py:394-395."""
    return list(edges_from_sparse_matrix(file_name, ix_converter, sep))


def edges_dict_from_sparse(edges, n_bins=None):
    """Edges store from (source, sink, weight) triples; chess/helpers.py:342-356.

    Returns a :class:`DeviceContacts` (device-resident), which answers
    ``store[source][sink]`` like the reference's dict of dicts.  ``n_bins`` (not in
    the reference signature) = number of regions the matrix goes with: bins after
    the last contact exist and read as the default value.
    """
    from .engine import DeviceContacts, DeviceTriples
    if isinstance(edges, DeviceTriples):
        if n_bins is None:
            return DeviceContacts(*edges.host())
        return DeviceContacts.from_device(edges, n_bins=n_bins)
    if isinstance(edges, tuple) and len(edges) == 3 and isinstance(edges[0], np.ndarray):
        return DeviceContacts(*edges, n_bins=n_bins)
    triples = list(edges)
    if triples:
        a, b, w = zip(*triples)
    else:
        a, b, w = (), (), ()
    return DeviceContacts(np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64),
                          np.asarray(w, dtype=np.float64), n_bins=n_bins)


def _fanc_bins(hic):
    try:
        return len(hic.regions)
    except TypeError:
        return sum(1 for _ in hic.regions)


def edges_dict_from_fanc(hic):
    """chess/helpers.py:359-373 (needs the third-party fanc object)."""
    return edges_dict_from_sparse(((e.source, e.sink, e.weight) for e in hic.edges(lazy=True)),
                                  n_bins=_fanc_bins(hic))


def oe_edges_dict_from_fanc(hic):
    """chess/helpers.py:376-391 (O/E computed inside fanc)."""
    return edges_dict_from_sparse(((e.source, e.sink, e.weight) for e in hic.edges(oe=True, lazy=True)),
                                  n_bins=_fanc_bins(hic))


def load_matrix(file_name, size=None, sep=None, ix_converter=None, is_oe=False):
    """Dense numpy matrix from a sparse file; chess/helpers.py:398-423."""
    if size is None:
        raise ValueError("Must provide matrix size!")
    m = np.ones((size, size)) if is_oe else np.zeros((size, size))
    a, b, w = _parse_sparse(file_name, ix_converter, sep if sep is not None else "\t")
    m[a, b] = w
    m[b, a] = w
    return m


def _pair_fields(line, sep):
    fields = line.rstrip().split(sep)
    if len(fields) != 10:
        raise ValueError("10 columns expected but {} found in bedpe input".format(len(fields)))
    return line.split(sep)


def load_pairs_iter(file_name, sep=None):
    """Yield (ID, reference_region, query_region); chess/helpers.py:445-480."""
    with _open(file_name, "r") as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            c1, s1, e1, c2, s2, e2, ID, _, str1, str2 = _pair_fields(line, sep)
            yield (ID,
                   GenomicRegion(chromosome=str(c1), start=int(s1) + 1, end=int(e1), strand=str(str1)),
                   GenomicRegion(chromosome=str(c2), start=int(s2) + 1, end=int(e2), strand=str(str2)))


def load_pairs(file_name, sep=None):
    """{ID: (reference_region, query_region)}; chess/helpers.py:426-442."""
    return {ID: (r, q) for ID, r, q in load_pairs_iter(file_name, sep=sep)}


def load_pairs_for_chroms(file_name, chromosomes, sep=None):
    """([(ID, ref, qry)], dropped) restricted to known chromosomes.

    chess/helpers.py:483-527: start = BEDPE start + 1, duplicate IDs raise,
    pairs touching an unknown chromosome are counted and skipped.
    """
    seen, res, dropped = set(), [], 0
    with _open(file_name, "r") as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            c1, s1, e1, c2, s2, e2, ID, _, str1, str2 = _pair_fields(line, sep)
            if ID in seen:
                raise ValueError("Pair ID {} not unique in bedpe input!")
            seen.add(ID)
            if c1 not in chromosomes or c2 not in chromosomes:
                dropped += 1
                continue
            res.append((ID,
                        GenomicRegion(chromosome=str(c1), start=int(s1) + 1, end=int(e1), strand=str(str1)),
                        GenomicRegion(chromosome=str(c2), start=int(s2) + 1, end=int(e2), strand=str(str2))))
    return res, dropped


def chunks(l, p):
    """Successive ceil(len/p)-sized chunks; chess/helpers.py:530-534."""
    n = int(np.ceil(len(l) / p))
    for i in range(0, len(l), n):
        yield l[i:i + n]


def read_chromosome_sizes(file_name):
    """chess/helpers.py:537-545."""
    sizes = {}
    with open(os.path.expanduser(file_name), "r") as fh:
        for line in fh:
            line = line.rstrip()
            if line != "":
                chromosome, length = line.split("\t")
                sizes[chromosome] = int(length)
    return sizes


def _try_fanc(matrix_file):
    """fanc.load if fanc is installed, else the ValueError that sends the
    reference down its sparse-format branch (chess/helpers.py:553-560)."""
    try:
        import fanc  # third-party, optional
    except ImportError:
        raise ValueError("{}: fanc is not installed, so only the sparse format "
                         "(matrix + regions BED) can be read".format(matrix_file))
    return fanc.load(matrix_file)


def load_contacts(matrix_file, regions_file=None, device=None):
    """Balanced contacts -> (O/E edges store, region index, regions).

    chess/helpers.py:548-573.  Sparse-format input is normalised on the device
    (``oe.observed_expected_arrays``); fanc-readable files use fanc's own O/E.
    ``device`` (not in the reference): the GPU that parses and normalises.  The text
    goes to that GPU once; parsed triples, expected values and O/E values never
    leave it, and the regions come back as a RegionTable (columns, objects on demand).
    """
    from .oe import observed_expected_arrays
    try:
        loaded = _try_fanc(matrix_file)
        edges = oe_edges_dict_from_fanc(loaded)
        regions = loaded.regions
        region_trees = region_interval_trees(regions)
    except ValueError as initial_error:
        print(initial_error)
        try:
            assert regions_file is not None, ("Regions file needs to be"
                                              "specified for sparse input.")
            regions, ix_converter = load_regions_table(regions_file)
            region_trees = region_interval_trees(regions)
            triples = _parse_sparse(matrix_file, ix_converter, want_device=True, device=device)
            edges = observed_expected_arrays(regions, triples, device=device)
        except AssertionError as error:
            raise ValueError(error)
    return edges, region_trees, regions


def load_oe_contacts(matrix_file, regions_file=None, device=None):
    """Already O/E-transformed contacts; chess/helpers.py:576-598."""
    try:
        loaded = _try_fanc(matrix_file)
        edges = edges_dict_from_fanc(loaded)
        regions = loaded.regions
        region_trees = region_interval_trees(regions)
    except ValueError:
        try:
            assert regions_file is not None
            regions, ix_converter = load_regions_table(regions_file)
            region_trees = region_interval_trees(regions)
            edges = edges_dict_from_sparse(_parse_sparse(matrix_file, ix_converter, want_device=True, device=device),
                                           n_bins=len(regions))
        except AssertionError:
            raise ValueError
    return edges, region_trees, regions
