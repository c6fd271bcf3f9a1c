"""CPU-only tests: host-side mirror of the reference interface, C-ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import chess_oracle as O
from tests.conftest import ROOT


def test_library_exports_every_declared_symbol():
    from chess_b200 import _native
    header = open(os.path.join(ROOT, "include", "chess_b200.h")).read()
    declared = set(re.findall(r"\b(chess_[a-z0-9_]+)\s*\(", header))
    declared -= {"chess_oe_band_build"}  # mentioned in a comment only
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    from chess_b200 import _native
    assert lib.chess_abi_version() == _native.ABI_VERSION == 4
    assert ctypes.sizeof(_native.chess_pair) == 40 and ctypes.sizeof(_native.chess_params) == 56


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from chess_b200 import _native
    with pytest.raises(_native.ChessNativeError, match="no CPU fallback"):
        _native.Context(0)


def test_region_index_matches_oracle_lookup(toy):
    from chess_b200.helpers import load_regions, region_interval_trees, GenomicRegion, sub_regions
    from chess_b200.engine import region_spans
    regions, _, _ = load_regions(toy("ref.bed"))
    trees = region_interval_trees(regions)
    oregions, _, _ = O.load_regions(toy("ref.bed"))
    otrees = O.region_trees(oregions)
    rng = np.random.default_rng(0)
    queries = []
    for _ in range(300):
        chrom = ["chrA", "chrB", "chrC"][rng.integers(0, 3)]
        s = int(rng.integers(1, 70000))
        e = s + int(rng.integers(1, 40000))
        queries.append(GenomicRegion(chromosome=chrom, start=s, end=e))
    first, count = region_spans(trees, queries)
    for q, f, n in zip(queries, first, count):
        try:
            _, lo, hi = O.sub_regions(otrees, O.Region(q.start, q.end, q.chromosome))
            assert (f, n) == (lo, hi - lo + 1)
            hits, lo2, hi2 = sub_regions(trees, q)
            assert (lo2, hi2) == (lo, hi) and len(hits) == hi - lo + 1
        except ValueError:
            assert n == 0
            with pytest.raises(ValueError):
                sub_regions(trees, q)
    with pytest.raises(KeyError):
        region_spans(trees, [GenomicRegion(chromosome="chrZ", start=1, end=10)])


def test_region_index_irregular_bins():
    from chess_b200.helpers import GenomicRegion, RegionIndex
    regs = [GenomicRegion(start=0, end=100, chromosome="c", ix=3), GenomicRegion(start=50, end=400, chromosome="c", ix=0),
            GenomicRegion(start=100, end=200, chromosome="c", ix=7), GenomicRegion(start=300, end=350, chromosome="c", ix=1)]
    idx = RegionIndex(regs)
    assert not idx.monotone
    lo, hi = idx.span(np.array([120, 0, 1000]), np.array([130, 10, 1100]))
    assert list(lo) == [0, 3, -1] and list(hi) == [7, 3, -1]


def test_pairs_and_regions_parsers_match_oracle(toy):
    from chess_b200.helpers import load_pairs_for_chroms, load_regions, load_pairs, chunks
    pairs, dropped = load_pairs_for_chroms(toy("pairs.bedpe"), {"chrA", "chrB", "chrC"})
    opairs, odropped = O.load_pairs_for_chroms(toy("pairs.bedpe"), {"chrA", "chrB", "chrC"})
    assert dropped == odropped == 1
    assert [(p, str(r), str(q)) for p, r, q in pairs] == \
        [(p, "%s:%s-%s:%s" % (r.chromosome, r.start, r.end, r.strand),
          "%s:%s-%s:%s" % (q.chromosome, q.start, q.end, q.strand)) for p, r, q in opairs]
    assert len(load_pairs(toy("pairs.bedpe"))) == 21
    regions, conv, ix2reg = load_regions(toy("named.bed"))
    oregions, oconv, oix2reg = O.load_regions(toy("named.bed"))
    assert conv == oconv and ix2reg == oix2reg and len(regions) == len(oregions)
    assert [len(c) for c in chunks(list(range(10)), 3)] == [4, 4, 2]


def test_sparse_parser_matches_oracle(toy):
    from chess_b200.helpers import _parse_sparse, load_regions, load_sparse_matrix
    for tag in ("ref", "named"):
        _, conv, _ = load_regions(toy(tag + ".bed"))
        a, b, w = _parse_sparse(toy(tag + ".sparse"), conv)
        want = list(O.iter_sparse(toy(tag + ".sparse"), conv))
        assert len(want) == len(w)
        assert [(int(x), int(y), float(z)) for x, y, z in zip(a, b, w)] == want
    assert load_sparse_matrix(toy("ref.sparse"))[:50] == list(O.iter_sparse(toy("ref.sparse")))[:50]


def test_sim_parser_contract():
    from chess_b200.commands import sim_parser
    args = sim_parser().parse_args(["r.m", "q.m", "p.bedpe", "out.tsv"])
    assert (args.threads, args.keep_unmappable_bins, args.mappability_cutoff, args.relative_windowsize,
            args.absolute_windowsize, args.oe_input, args.limit_background, args.background_query,
            args.background_regions, args.reference_regions, args.query_regions) == \
        (1, False, 0.1, 1, None, False, False, False, None, None, None)
    args = sim_parser().parse_args(["r", "q", "p", "o", "--reference-regions", "a.bed", "--query-regions", "b.bed",
                                    "-p", "4", "-a", "7", "-r", "0.5", "--oe-input", "--keep-unmappable-bins",
                                    "--mappability-cutoff", "0.2", "--background-query", "--limit-background",
                                    "--background-regions", "bg.bed"])
    assert args.threads == 4 and args.absolute_windowsize == 7 and args.relative_windowsize == 0.5
    assert args.oe_input and args.keep_unmappable_bins and args.background_query and args.limit_background


def test_synthetic_generator_is_seeded():
    from chess_b200 import synth
    g = synth.Genome([("chr1", 300), ("chr2", 200)], 10000)
    a = synth.synthetic_sample(g, 60, seed=1, structure_seed=5)
    b = synth.synthetic_sample(g, 60, seed=1, structure_seed=5)
    c = synth.synthetic_sample(g, 60, seed=2, structure_seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert len(c[2]) != len(a[2]) or not np.array_equal(a[2], c[2])
    assert (a[0] <= a[1]).all() and a[1].max() < g.n_bins and (a[2] > 0).all()
    pairs = synth.sliding_pairs(g, 400000, 100000)
    assert len(pairs) == (3000000 - 400000) // 100000 + 1 + (2000000 - 400000) // 100000 + 1


def test_contact_store_covers_every_region_bin(toy, tmp_path):
    """Bins after the last bin with a contact (chrY, chrM, telomeres) exist: the store built on the
    --oe-input path must span the whole regions file, not just max(sink) + 1; a region outside it is an
    error, not an out-of-bounds read."""
    from chess_b200.engine import DeviceContacts, as_device_contacts, index_bins
    from chess_b200.helpers import GenomicRegion, edges_dict_from_sparse, region_interval_trees
    regions = [GenomicRegion(chromosome="c", start=k * 10, end=k * 10 + 10, ix=k) for k in range(100)]
    trees = region_interval_trees(regions)
    assert index_bins(trees) == 100 and index_bins(regions) == 100
    a = np.arange(0, 59)
    store = edges_dict_from_sparse((a, a + 1, np.ones(59)), n_bins=len(regions))
    assert store.n_bins == 100
    short = DeviceContacts(a, a + 1, np.ones(59))
    assert short.n_bins == 60
    with pytest.raises(ValueError, match="outside the contact store"):
        short.check_span([50], [41])
    assert as_device_contacts(short, index_bins(trees)).n_bins == 100
    short.check_span([50], [41])
    short.check_span([-1, 95], [0, 5])          # n = 0 marks "region not found": ignored
    with pytest.raises(ValueError):
        short.check_span([95], [6])
    d = as_device_contacts({0: {1: 2.0}, 1: {1: 3.0}}, 7)
    assert d.n_bins == 7 and d[0] == {1: 2.0}
    # load_oe_contacts passes the number of regions
    bed = tmp_path / "r.bed"
    bed.write_text("".join("c\t%d\t%d\n" % (k * 10, k * 10 + 10) for k in range(30)))
    mat = tmp_path / "m.sparse"
    mat.write_text("0\t1\t1.5\n3\t4\t0.5\n")
    import torch
    if not torch.cuda.is_available():      # file parsing alone works without a GPU
        from chess_b200.helpers import load_oe_contacts
        edges, _, regs = load_oe_contacts(str(mat), str(bed))
        assert edges.n_bins == len(regs) == 30


def test_duplicate_contacts_keep_the_last_occurrence():
    """chess/helpers.py:353-356: edges_d[source][sink] = weight -- a later line overwrites an earlier
    one, and (i, j) / (j, i) are the same cell after canonicalisation."""
    from chess_b200.engine import DeviceContacts
    src = np.array([0, 1, 2, 1, 0, 5, 2])
    snk = np.array([1, 2, 1, 2, 1, 5, 2])
    val = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0])
    d = DeviceContacts(src, snk, val)
    ref = {}
    for a, b, w in zip(src, snk, val):
        ref.setdefault(min(a, b), {})[max(a, b)] = w
    got = {int(a): d[int(a)] for a in set(d.src.tolist())}
    assert got == {int(k): {int(j): float(w) for j, w in row.items()} for k, row in ref.items()}
    assert len(d.value) == sum(len(r) for r in ref.values())
    # sorted, duplicate-free input is untouched (and takes the O(n) path)
    e = DeviceContacts(np.array([0, 0, 1]), np.array([0, 2, 1]), np.array([1.0, 2.0, 3.0]))
    assert e.value.tolist() == [1.0, 2.0, 3.0]
    # oracle agrees (edges_from_sparse_matrix canonicalises before the dict is built, helpers.py:332-335)
    lo, hi = np.minimum(src, snk), np.maximum(src, snk)
    assert O.edges_dict(zip(lo.tolist(), hi.tolist(), val.tolist())) == \
        {int(k): {int(j): float(w) for j, w in row.items()} for k, row in ref.items()}


def test_oversized_and_mismatched_inputs_fail_with_a_clear_message():
    import torch
    if torch.cuda.is_available():
        pytest.skip("host-side validation is exercised on the CPU box")
    from chess_b200.helpers import GenomicRegion
    from chess_b200 import oe
    regions = [GenomicRegion(chromosome="c", start=k * 10, end=k * 10 + 10, ix=k) for k in range(10)]
    # index validation happens before any device work
    import chess_b200.oe as oemod
    orig = oemod.visible_devices
    oemod.visible_devices = lambda: [0]
    try:
        with pytest.raises(IndexError, match="does not match the regions file"):
            oe._device_expected(regions, (np.array([0, 3]), np.array([1, 12]), np.array([1.0, 2.0])))
        with pytest.raises(IndexError):
            oe._device_expected(regions, (np.array([-1]), np.array([1]), np.array([1.0])))
    finally:
        oemod.visible_devices = orig


def test_region_and_pair_tables_match_the_object_loaders(toy, tmp_path):
    """The columnar loaders of the batched engine (RegionTable, PairTable) against the reference-style loaders:
    same regions, indices, pairs, dropped counts and errors -- including files the fast path must hand over."""
    from chess_b200 import helpers as H
    for bed in ("ref.bed", "qry.bed", "named.bed", "background.bed"):
        table, conv = H.load_regions_table(toy(bed))
        regions, conv0, _ = H.load_regions(toy(bed))
        assert conv == conv0 and len(table) == len(regions)
        assert all(a.__dict__ == b.__dict__ for a, b in zip(table, regions))
        assert table.max_bin_size() == max(r.end - r.start + 1 for r in regions)
        trees, trees0 = H.region_interval_trees(table), H.region_interval_trees(regions)
        assert set(trees) == set(trees0)
        for chrom in trees:
            assert np.array_equal(trees[chrom].begins, trees0[chrom].begins)
            assert np.array_equal(trees[chrom].ixs, trees0[chrom].ixs)
            hits = sorted(h.data.ix for h in trees[chrom][5000:12000])
            assert hits == sorted(h.data.ix for h in trees0[chrom][5000:12000])
    # a BED with a blank line in the middle: the bin index is the LINE number (helpers.py:176-201)
    odd = tmp_path / "odd.bed"
    odd.write_text("c\t0\t10\nc\t10\t20\n\nc\t20\t30\n")
    table, _ = H.load_regions_table(str(odd))
    regions, _, _ = H.load_regions(str(odd))
    assert [r.ix for r in table] == [r.ix for r in regions] == [0, 1, 3]
    trees = H.region_interval_trees(H.load_regions_table(toy("ref.bed"))[0])
    chroms = set(trees)
    pt, dropped = H.load_pairs_table(toy("pairs.bedpe"), chroms)
    pl, dropped0 = H.load_pairs_for_chroms(toy("pairs.bedpe"), chroms)
    assert dropped == dropped0 and len(pt) == len(pl)
    for a, b in zip(pt, pl):
        assert a[0] == b[0] and str(a[1]) == str(b[1]) and str(a[2]) == str(b[2])
    assert pt.flips().tolist() == [1 if p[1].strand != p[2].strand else 0 for p in pl]
    assert pt.max_spans_bp() == (max(p[1].end - p[1].start + 1 for p in pl), max(p[2].end - p[2].start + 1 for p in pl))
    assert [p[0] for p in pt[3:7]] == [p[0] for p in pl[3:7]]
    # comment lines, duplicate IDs and ragged lines: the reference loader decides
    text = open(toy("pairs.bedpe")).read()
    (tmp_path / "c.bedpe").write_text("# comment\n" + text)
    assert len(H.load_pairs_table(str(tmp_path / "c.bedpe"), chroms)[0]) == len(pl)
    (tmp_path / "d.bedpe").write_text(text + text.splitlines()[0] + "\n")
    with pytest.raises(ValueError, match="not unique"):
        H.load_pairs_table(str(tmp_path / "d.bedpe"), chroms)
    (tmp_path / "r.bedpe").write_text(text + "chrA\t1\t2\n")
    with pytest.raises(ValueError, match="10 columns expected"):
        H.load_pairs_table(str(tmp_path / "r.bedpe"), chroms)


def test_background_spans_and_table_text_match_the_loops(toy):
    """Vectorised pieces of cli.run_sim against the loops they replace: background regions generated from the
    query bins (bin/chess:272-298) and the number formatting of the output table."""
    from chess_b200 import cli, helpers as H
    from chess_b200.engine import region_spans
    table, _ = H.load_regions_table(toy("qry.bed"))
    regions, _, _ = H.load_regions(toy("qry.bed"))
    trees = H.region_interval_trees(table)
    for q, limit in ((H.GenomicRegion(chromosome="chrA", start=4001, end=28000, strand="+"), True),
                     (H.GenomicRegion(chromosome="chrB", start=1, end=20000, strand="-"), False),
                     (H.GenomicRegion(chromosome="chrQ", start=1, end=20000), True)):
        regs = cli._background_regions_from_query(regions, q, limit)
        bf0, bn0 = region_spans(H.region_interval_trees(regions), regs) if regs else (np.zeros(0), np.zeros(0))
        bs0 = [0 if r.strand == '+' else 1 for r in regs]
        bf, bn, bs = cli._background_spans_from_query(table, trees, q, limit)
        assert np.array_equal(bf, bf0) and np.array_equal(bn, bn0) and bs.tolist() == bs0
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(0, 1, 2000) * 10.0 ** rng.integers(-20, 20, 2000),
                           [np.nan, 0.0, -0.0, 1.0, 1e16, 1e-5, 123456789.125, np.inf, -np.inf, 5e-324, 1.7976931348623157e308]])
    assert cli._float_column(vals) == ["{}".format(np.float64(v)) for v in vals]


def test_gpus_are_planned_by_the_work():
    """A CUDA context costs about a second, so `chess sim` adds GPUs only when the comparisons are worth it."""
    from chess_b200.engine import plan_devices
    eight = list(range(8))
    assert plan_devices([0], 1e15) == [0] and plan_devices([0, 1], 1.0) == [0, 1]
    assert plan_devices(eight, 101354 * 201 ** 2) == [0, 1]                 # configs[2]: 40 ms of work
    assert plan_devices(eight, 101354 * 26000.0 * 201 ** 2) == eight         # the same list with --background-query
    n = [len(plan_devices(eight, w)) for w in (1e9, 1e11, 1e12, 3e12, 1e13, 1e14)]
    assert n == sorted(n) and n[0] == 2 and n[-1] == 8


def test_run_level_zscores_keep_nan_and_match_the_reference_rule():
    """bin/chess:228-233: nanmean / nanstd (ddof 0) over the non-NaN ssim values; a NaN ssim keeps z = NaN."""
    from chess_b200.distributed import run_zscores
    s = np.array([0.5, np.nan, 0.25, 0.75, np.nan, 0.5])
    z = run_zscores(s)
    ok = ~np.isnan(s)
    assert np.all(np.isnan(z[~ok]))
    np.testing.assert_allclose(z[ok], (s[ok] - np.nanmean(s)) / np.nanstd(s), rtol=1e-15)
    assert np.all(np.isnan(run_zscores(np.array([np.nan, np.nan])))) and len(run_zscores(np.zeros(0))) == 0


def test_work_estimate_from_file_heads(tmp_path):
    """cli._estimate_work: comparisons x bins^2 from the first lines and the line counts, host only."""
    import argparse
    from chess_b200 import cli
    bed = tmp_path / "r.bed"
    bed.write_text("".join("chr1\t%d\t%d\n" % (i * 10000, (i + 1) * 10000) for i in range(500)))
    pairs = tmp_path / "p.bedpe"
    pairs.write_text("".join("chr1\t%d\t%d\tchr1\t%d\t%d\t%d\t.\t+\t+\n" % (s, s + 2000000, s, s + 2000000, k)
                             for k, s in enumerate(range(0, 3000000, 100000))))
    args = argparse.Namespace(pairs=str(pairs), reference_regions=str(bed), query_regions=str(bed),
                              background_query=False, limit_background=False, background_regions=None)
    assert cli._estimate_work(args) == 30 * 201.0 * 201
    args.background_query = True
    assert cli._estimate_work(args) == 30 * (1 + 2 * 500.0) * 201 * 201
    args.background_query, args.background_regions = False, str(bed)
    assert cli._estimate_work(args) == 30 * (1 + 500.0) * 201 * 201
    args.reference_regions = None                       # a matrix format with its own regions: no estimate
    assert cli._estimate_work(args) is None
    # the user's choice of GPUs is never overridden
    args.reference_regions = str(bed)
    before = os.environ.get("CUDA_VISIBLE_DEVICES")
    cli._limit_visible_gpus(args)                       # programmatic calls leave the environment alone
    assert os.environ.get("CUDA_VISIBLE_DEVICES") == before
    cli.LIMIT_GPUS = True
    try:
        os.environ["CHESS_B200_DEVICES"] = "0"          # GPUs named explicitly: nothing is hidden
        cli._limit_visible_gpus(args)
        assert os.environ.get("CUDA_VISIBLE_DEVICES") == before
        del os.environ["CHESS_B200_DEVICES"]
        os.environ["CUDA_VISIBLE_DEVICES"] = "3,5,6,7"  # a list is narrowed to the GPUs the work is worth
        cli._limit_visible_gpus(args)
        assert os.environ["CUDA_VISIBLE_DEVICES"] == "3,5"
    finally:
        cli.LIMIT_GPUS = False
        os.environ.pop("CHESS_B200_DEVICES", None)
        if before is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = before


def _same_outcome(fast, slow):
    """Both callables give the same value, or raise the same exception class."""
    try:
        want = slow()
    except Exception as exc:        # noqa: BLE001  (the class is the contract)
        with pytest.raises(type(exc)):
            fast()
        return None
    got = fast()
    return got, want


def test_table_loaders_hand_over_whatever_is_not_a_plain_table(tmp_path):
    """The pyarrow fast path of load_regions_table / load_pairs_table takes plain tab-separated files only; every
    other spelling the reference's loops accept (or reject) must come out as those loops decide."""
    from chess_b200 import helpers as H
    row = "chr1\t%d\t%d\tchr2\t%d\t%d\t%s\t.\t+\t-\n"
    plain = "".join(row % (k * 100, k * 100 + 500, k * 50, k * 50 + 500, "id%d" % k) for k in range(40))
    variants = {
        "plain": plain,
        "no_final_newline": plain.rstrip("\n"),
        "spaces": plain.replace("\t", " "),
        "mixed_whitespace": plain.replace("\t.\t", " \t . \t"),
        "crlf": plain.replace("\n", "\r\n"),
        "comment": "#header\n" + plain,
        "blank_middle": plain.replace("id7\t", "id7\t", 1).replace("\n", "\n\n", 1),
        "blank_end": plain + "\n",
        "trailing_tab": plain.replace("-\n", "-\t\n", 1),
        "ragged": plain + "chr1\t1\t2\n",
        "eleven": plain + (row % (1, 2, 3, 4, "x")).rstrip("\n") + "\textra\n",
        "duplicate_id": plain + row % (1, 2, 3, 4, "id3"),
        "plus_sign": plain + "chr1\t+5\t900\tchr2\t7\t800\tplus\t.\t+\t+\n",
        "float_start": plain + "chr1\t5.0\t900\tchr2\t7\t800\tfl\t.\t+\t+\n",
        "empty_field": plain + "chr1\t\t900\tchr2\t7\t800\tempty\t.\t+\t+\n",
        "na_names": plain + "NA\t1\t900\tnull\t7\t800\tNaN\t.\t+\t+\n",
        "numeric_chroms": plain.replace("chr1", "1").replace("chr2", "2"),
        "quotes": plain.replace("id5", '"id5"'),
        "unknown_chrom": plain + "chrZ\t1\t900\tchr2\t7\t800\tz\t.\t+\t+\n",
        "empty": "",
    }
    chroms = {"chr1", "chr2", "NA", "null", "1", "2"}
    for name, text in variants.items():
        path = tmp_path / (name + ".bedpe")
        path.write_text(text)
        out = _same_outcome(lambda: H.load_pairs_table(str(path), chroms),
                            lambda: H.load_pairs_for_chroms(str(path), chroms))
        if out is None:
            continue
        (table, dropped), (want, dropped0) = out
        assert dropped == dropped0 and len(table) == len(want), name
        for a, b in zip(table, want):
            assert a[0] == b[0] and a[1].__dict__ == b[1].__dict__ and a[2].__dict__ == b[2].__dict__, name
        assert [type(v) for v in table.ids] == [str] * len(table), name
    bed = "".join("chr%d\t%d\t%d\n" % (1 + k // 30, (k % 30) * 1000, (k % 30 + 1) * 1000) for k in range(60))
    beds = {
        "plain": bed,
        "dots": bed.replace("\n", "\t.\n"),
        "names": "".join(line + "\tbin%d\n" % k for k, line in enumerate(bed.splitlines())),
        "int_names": "".join(line + "\t%d\n" % (k + 1) for k, line in enumerate(bed.splitlines())),
        "one_name": bed.replace("\n", "\t.\n").replace("\t.\n", "\tx\n", 1),
        "spaces": bed.replace("\t", "  "),
        "crlf": bed.replace("\n", "\r\n"),
        "blank_middle": bed.replace("\n", "\n\n", 1),
        "short_line": bed + "chr3\t5\n" + "chr3\t0\t1000\n",
        "five_columns": bed.replace("\n", "\t.\t0\n"),
        "bad_int": bed + "chr3\tx\t1000\n",
        "plus_sign": bed + "chr3\t+0\t1000\n",
        "numeric_chroms": bed.replace("chr", ""),
        "empty": "",
    }
    for name, text in beds.items():
        path = tmp_path / (name + ".bed")
        path.write_text(text)
        for ignore_ix in (False, True):
            out = _same_outcome(lambda: H.load_regions_table(str(path), ignore_ix=ignore_ix),
                                lambda: H.load_regions(str(path), ignore_ix=ignore_ix))
            if out is None:
                continue
            (table, conv), (want, conv0, _) = out
            assert conv == conv0 and len(table) == len(want), (name, ignore_ix)
            assert all(a.__dict__ == b.__dict__ for a, b in zip(table, want)), (name, ignore_ix)
            assert all(type(r.chromosome) is str for r in table), name
