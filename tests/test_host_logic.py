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
    assert lib.chess_abi_version() == _native.ABI_VERSION == 3
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
