"""The oracle restatement against outputs of the reference's own code.

Fixtures come from tests/golden/make_golden.py (reference executed in the build
container).  Everything here is CPU-only and bit-exact: the oracle performs the
same numpy operations in the same order as the reference.
"""
import os

import numpy as np
import pytest

from oracle import chess_oracle as O
from tests.conftest import unhex, unhex_arr, same_float, close, GOLDEN_DIR, ROOT


def test_sn_matches_reference(golden):
    for case in golden["SN"]:
        n = case["n"]
        a, b = unhex_arr(case["a"], (n, n)), unhex_arr(case["b"], (n, n))
        assert same_float(O.signal_to_noise(a, b), unhex(case["sn"]))


def test_masked_rows(golden):
    g = golden["masked"]
    m = unhex_arr(g["m"], (g["n"], g["n"]))
    idx = O.find_masked_rows(m, masking_value=1)
    assert list(map(int, idx)) == g["idx"]
    assert np.array_equal(O.remove_rows(m, idx).ravel(), unhex_arr(g["removed"]))


@pytest.mark.parametrize("tag", ["ref", "qry", "named"])
def test_expected_and_oe(golden, toy, tag):
    g = golden["oe"][tag]
    regions, conv, _ = O.load_regions(toy(tag + ".bed"))
    triples = list(O.iter_sparse(toy(tag + ".sparse"), conv))
    assert len(triples) == g["n_edges"]
    genome, chrom, inter, mappable = O.expected_values(regions, triples)
    assert mappable == g["mappable"]
    assert np.array_equal(np.array(genome), unhex_arr(g["intra_expected"]))
    assert same_float(inter, unhex(g["inter_expected"]))
    for c, v in g["chrom_expected"].items():
        assert np.array_equal(np.array(chrom[c]), unhex_arr(v)), c
    tot, ctot, itot = O.possible_contacts(regions, mappable)
    assert tot == g["possible_intra"] and itot == g["possible_inter"]
    for c, v in g["possible_chrom"].items():
        assert ctot[c] == v
    oe = O.observed_expected(regions, triples)
    if g["oe"] is not None:
        want = g["oe"]
    else:
        want = [[a, b, float(w).hex()] for a, b, w in O.iter_sparse(toy(tag + ".oe.sparse"))]
    assert len(oe) == len(want)
    for (a, b, w), (ga, gb, gw) in zip(oe, want):
        assert (a, b) == (ga, gb) and same_float(w, unhex(gw))


def _load(toy, tag):
    regions, conv, _ = O.load_regions(toy(tag + ".bed"))
    triples = list(O.iter_sparse(toy(tag + ".sparse"), conv))
    edges = O.edges_dict(O.observed_expected(regions, triples))
    return edges, O.region_trees(regions), regions


def test_gather(golden, toy):
    edges, trees, _ = _load(toy, "ref")
    for case in golden["gather"]["cases"]:
        reg = O.Region(case["start"], case["end"], case["chrom"])
        m, hits = O.sub_matrix_from_edges_dict(edges, trees, reg, default_weight=1)
        assert m.shape[0] == case["n"]
        assert min(h.ix for h in hits) == case["first_ix"]
        assert np.array_equal(m.ravel(), unhex_arr(case["m"]))
    with pytest.raises(ValueError):
        O.sub_matrix_from_edges_dict(edges, trees, O.Region(90001, 120000, "chrA"), 1)
    assert golden["gather"]["out_of_range"] == "ValueError"


def test_pairs(golden, toy):
    _, rt, _ = _load(toy, "ref")
    pairs, dropped = O.load_pairs_for_chroms(toy("pairs.bedpe"), set(rt.keys()))
    assert dropped == golden["pairs"]["dropped"]
    rows = [[pid, r.chromosome, r.start, r.end, r.strand, q.chromosome, q.start, q.end, q.strand]
            for pid, r, q in pairs]
    assert rows == golden["pairs"]["rows"]


def test_worker_all_parameter_sets(golden, toy):
    re_, rt, _ = _load(toy, "ref")
    qe, qt, _ = _load(toy, "qry")
    pairs, _ = O.load_pairs_for_chroms(toy("pairs.bedpe"), set(rt.keys()) | set(qt.keys()))
    for tag, g in golden["worker"].items():
        rows = O.compare_chunk(pairs, re_, rt, qe, qt, True, **g["params"])
        assert len(rows) == len(g["rows"])
        for (pid, s, sn), (gpid, gs, gsn) in zip(rows, g["rows"]):
            assert pid == gpid, tag
            assert same_float(s, unhex(gs)), (tag, pid, s, unhex(gs))
            assert same_float(sn, unhex(gsn)), (tag, pid, sn, unhex(gsn))


def test_compare_pair_and_structures(golden, toy):
    re_, rt, _ = _load(toy, "ref")
    qe, qt, _ = _load(toy, "qry")
    pairs, _ = O.load_pairs_for_chroms(toy("pairs.bedpe"), set(rt.keys()) | set(qt.keys()))
    for pair, (gpid, gs, gsn) in zip(pairs[:4], golden["compare_pair_nosn"]):
        pid, s, sn = O.compare_pair(pair, re_, rt, qe, qt, compute_sn=False)
        assert pid == gpid and same_float(s, unhex(gs))
        assert (sn is None) if gsn is None else (np.isnan(gsn) and np.isnan(sn))
    for pair in pairs[:6]:
        pid, s, sn = O.compare_pair(pair, re_, rt, qe, qt, absolute_window_size=7)
        gs, gsn = golden["compare_structures_abs7"][pid]
        assert same_float(s, unhex(gs))
        if not np.isnan(s):
            assert same_float(sn, unhex(gsn))


def _cli_table(toy, background=None, oe_input=False, background_query=False, limit_background=False, **kw):
    if oe_input:
        def load(tag):
            regions, conv, _ = O.load_regions(toy(tag + ".bed"))
            return O.edges_dict(O.iter_sparse(toy(tag + ".oe.sparse"), conv)), regions
        (re_, rr), (qe, qr) = load("ref"), load("qry")
    else:
        re_, _, rr = _load(toy, "ref")
        qe, _, qr = _load(toy, "qry")
    chroms = {r.chromosome for r in rr} | {r.chromosome for r in qr}
    pairs, _ = O.load_pairs_for_chroms(toy("pairs.bedpe"), chroms)
    bgr = None
    if background:
        bgr, _, _ = O.load_regions(toy(background), ignore_ix=True)
    res, bg = O.sim(re_, rr, qe, qr, pairs, background_regions=bgr, background_query=background_query,
                    limit_background=limit_background, **kw)
    return O.format_table(pairs, res, bg)


def test_cli_tables_byte_identical(toy):
    assert _cli_table(toy) == open(toy("cli_default.tsv")).read()
    assert _cli_table(toy, background="background.bed", absolute_window_size=7) == \
        open(toy("cli_abs7_bg.tsv")).read()
    assert _cli_table(toy, oe_input=True, relative_window_size=0.4, keep_unmappable_bins=True) == \
        open(toy("cli_oe_input.tsv")).read()
    assert _cli_table(toy, background_query=True, limit_background=True, absolute_window_size=5) == \
        open(toy("cli_bgquery_limit.tsv")).read()
    assert _cli_table(toy, background_query=True, mappability_cutoff=0.3) == \
        open(toy("cli_bgquery_all.tsv")).read()


def test_ssim_restatement_equals_bruteforce():
    rng = np.random.default_rng(5)
    for n, win in ((20, 3), (31, 7), (40, 39), (41, 41), (64, 11), (50, 25)):
        a = rng.lognormal(0, 0.5, (n, n))
        b = a * rng.lognormal(0, 0.3, (n, n))
        dr = max(a.max(), b.max()) - min(a.min(), b.min())
        s1 = O.structural_similarity(a, b, win, dr)
        s2 = O.structural_similarity_bruteforce(a, b, win, dr)
        assert abs(s1 - s2) <= 1e-12 * abs(s2)
    with pytest.raises(ValueError):
        O.structural_similarity(a, b, 51, dr)
    with pytest.raises(ValueError):
        O.structural_similarity(a, b, 8, dr)


def test_resize_index_rule_equals_scipy_zoom():
    for n_in, n_out in ((20, 21), (26, 41), (29, 31), (33, 100), (20, 20), (64, 65)):
        src = np.arange(n_in * n_in, dtype=np.float64).reshape(n_in, n_in)
        got = O.resize_order0(src, (n_out, n_out))
        idx = O.resize_nearest_index(n_in, n_out)
        assert np.array_equal(got, src[np.ix_(idx, idx)])


def test_window_rule():
    assert O.window_rule(120) == 119 and O.window_rule(121) == 121
    assert O.window_rule(80, absolute_window_size=8) == 7
    assert O.window_rule(80, absolute_window_size=1) == 3
    assert O.window_rule(30, relative_window_size=0.01) == 3


def test_extract_head_oracle_matches_reference_golden(toy):
    """oracle.log_ratio_stage against what the reference's own extract_structures handed to its first
    denoise_bilateral calls (tests/golden/make_golden_extract.py): bit for bit (same numpy)."""
    import os
    from tests.conftest import GOLDEN_DIR
    gold = np.load(os.path.join(GOLDEN_DIR, "extract_golden.npz"))
    trees, edges = [], []
    for tag in ("ref", "qry"):
        regions, conv, _ = O.load_regions(toy(tag + ".bed"))
        trees.append(O.region_trees(regions))
        edges.append(O.edges_dict(O.iter_sparse(toy(tag + ".oe.sparse"), conv)))
    pairs, _ = O.load_pairs_for_chroms(toy("pairs.bedpe"), set(trees[0]) | set(trees[1]))
    assert [p[0] for p in pairs] == gold["all_ids"].tolist()
    have = set(gold["ids"].tolist())
    for pair in pairs:
        pid = pair[0]
        if pid not in have:
            with pytest.raises(ValueError):
                O.log_ratio_stage(edges[0], trees[0], edges[1], trees[1], pair)
            continue
        pos, neg, std, mp, mn = O.log_ratio_stage(edges[0], trees[0], edges[1], trees[1], pair)
        assert np.array_equal(pos, gold["pos_" + pid]) and np.array_equal(neg, gold["neg_" + pid])
        assert [std, mp, mn] == gold["stats_" + pid].tolist()


# ---------------------------------------------------------------- north-star shape: 201 bins
def _golden_201_inputs():
    """The sample pair of tests/golden/make_golden_201.py, regenerated from its seeds; skipped (loudly) if the
    generator no longer reproduces the contacts the reference's code was run on."""
    import hashlib
    import importlib.util
    g = np.load(os.path.join(GOLDEN_DIR, "golden_201.npz"))
    spec = importlib.util.spec_from_file_location("synth_for_golden", os.path.join(ROOT, "chess_b200", "synth.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    genome = S.Genome([("chr1", int(g["n_bins"]))], int(g["bin"]))
    s = [int(v) for v in g["seeds"]]
    a = S.synthetic_sample(genome, 215, seed=s[0], structure_seed=s[1])
    b = S.synthetic_sample(genome, 215, seed=s[2], structure_seed=s[3], perturb=0.1)
    h = hashlib.sha256()
    for smp in (a, b):
        h.update(np.asarray(smp[0], dtype=np.int64).tobytes())
        h.update(np.asarray(smp[1], dtype=np.int64).tobytes())
        h.update(np.asarray(smp[2], dtype=np.float64).tobytes())
    if h.hexdigest() != str(g["inputs_sha256"]):
        pytest.skip("chess_b200/synth.py no longer reproduces the inputs of golden_201.npz: regenerate the golden")
    return g, genome, a, b


GOLDEN_201_PARAMS = {"default": dict(), "abs7": dict(absolute_window_size=7), "abs3": dict(absolute_window_size=3),
                     "abs11": dict(absolute_window_size=11), "keep": dict(keep_unmappable_bins=True),
                     "rel05": dict(relative_window_size=0.5)}


def test_oracle_matches_reference_at_201_bins():
    """The oracle against results of the reference's own compare_pair (chess/sim.py:225-380, chess/oe.py) at the
    shape BASELINE configs[2] is quoted on: 10 kb bins, 2 Mb regions = 201 bins, both strands, six parameter sets."""
    g, genome, a, b = _golden_201_inputs()
    regions = genome.regions(O.Region)
    trees = O.region_trees(regions)
    edges = []
    for smp in (a, b):
        triples = list(zip(smp[0].tolist(), smp[1].tolist(), smp[2].tolist()))
        edges.append(O.edges_dict(O.observed_expected(regions, triples)))
    bin_size = int(g["bin"])
    pairs = []
    for k, st in enumerate(int(v) for v in g["starts"]):
        for strand in ("+", "-"):
            r = O.Region(st * bin_size + 1, st * bin_size + 2000000, "chr1", strand="+")
            q = O.Region(st * bin_size + 1, st * bin_size + 2000000, "chr1", strand=strand)
            pairs.append(("p%d%s" % (k, strand), r, q))
    m, _ = O.sub_matrix_from_edges_dict(edges[0], trees, pairs[0][1], 1.0)
    assert m.shape[0] == int(g["bins_first_pair"]) == 201
    assert np.array_equal(O.find_masked_rows(m, masking_value=1), g["masked_first_pair_ref"])
    for tag, kw in GOLDEN_201_PARAMS.items():
        for k, pair in enumerate(pairs):
            _, s, sn = O.compare_pair(pair, edges[0], trees, edges[1], trees, compute_sn=True, **kw)
            assert close(s, g["ssim_" + tag][k], 1e-12), (tag, k, s, g["ssim_" + tag][k])
            assert close(sn, g["sn_" + tag][k], 1e-12), (tag, k, sn, g["sn_" + tag][k])
