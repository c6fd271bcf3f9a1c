"""GPU parity, round 2: the kernels that carry the north-star numbers checked DIRECTLY against the CPU
oracle at the shapes those numbers are quoted on (201 bins: the triangle kernels of configs[2]; 401 / 601
bins: the block kernels of configs[3]), the fused log2 transform, trailing contact-less bins, and the head
of `chess extract` against a golden produced by the reference's own code.  Everything goes through the
C ABI (ctypes); the oracle is the checker only."""
import os

import numpy as np
import pytest

from oracle import chess_oracle as O
from tests.conftest import close, GOLDEN_DIR

pytestmark = pytest.mark.gpu
RTOL = 1e-6   # north_star: 1e-6 relative in fp64 mode (observed ~1e-13)


@pytest.fixture(scope="module")
def cb():
    import chess_b200
    return chess_b200


def _oracle_side(regions, *contacts):
    oregions = [O.Region(r.start, r.end, r.chromosome, ix=r.ix) for r in regions]
    otrees = O.region_trees(oregions)
    return otrees, [O.edges_dict(zip(*[x.tolist() for x in c.triples()])) for c in contacts]


def _opair(pid, rr, qr):
    return (pid, O.Region(rr.start, rr.end, rr.chromosome, strand=rr.strand),
            O.Region(qr.start, qr.end, qr.chromosome, strand=qr.strand))


def _check_against_oracle(cb, eng, plist, otrees, ore, oqe, kw, worst, kernel=0):
    ids, s, sn, st = eng.compare(plist, kernel=kernel, **kw)
    n_ok = 0
    for k, (pid, rr, qr) in enumerate(plist):
        _, os_, osn = O.compare_pair(_opair(pid, rr, qr), ore, otrees, oqe, otrees, **kw)
        assert close(s[k], os_, RTOL), (kw, pid, s[k], os_)
        if kw.get("compute_sn", True):
            assert close(sn[k], osn, RTOL), (kw, pid, sn[k], osn)
        if not np.isnan(os_):
            n_ok += 1
            worst[0] = max(worst[0], abs(s[k] - os_) / abs(os_))
            if kw.get("compute_sn", True) and np.isfinite(osn) and osn != 0:
                worst[1] = max(worst[1], abs(sn[k] - osn) / abs(osn))
    return n_ok


@pytest.fixture(scope="module")
def hg38_like(cb):
    """10 kb bins, 2 Mb regions = 201 bins: the shape of BASELINE configs[2]; the query sample has its own
    unmappable runs (structure seed differs) so that bins masked in one sample only occur."""
    from chess_b200 import synth
    genome = synth.Genome([("chr1", 1500), ("chr2", 900)], 10000)
    ref = synth.synthetic_sample(genome, 215, seed=11, structure_seed=7)
    qry = synth.synthetic_sample(genome, 215, seed=12, structure_seed=7, perturb=0.1)
    qry2 = synth.synthetic_sample(genome, 215, seed=13, structure_seed=9, perturb=0.1)
    regions = genome.regions()
    trees = cb.region_interval_trees(regions)
    re_ = cb.observed_expected_arrays(regions, ref)
    qe = cb.observed_expected_arrays(regions, qry)
    qe2 = cb.observed_expected_arrays(regions, qry2)
    pairs = synth.sliding_pairs(genome, 2000000, 1370000)
    return genome, regions, trees, re_, qe, qe2, pairs


@pytest.mark.parametrize("kw", [dict(), dict(absolute_window_size=3), dict(absolute_window_size=5),
                                dict(absolute_window_size=7), dict(absolute_window_size=9),
                                dict(absolute_window_size=11), dict(compute_sn=False),
                                dict(absolute_window_size=7, mappability_cutoff=0.3)],
                         ids=["r1", "a3", "a5", "a7", "a9", "a11", "r1_nosn", "a7_cut03"])
def test_201_bin_kernels_against_oracle(cb, hg38_like, kw):
    """compare_r1_tri_kernel<7,.>, compare_win7_tri_kernel<7>, compare_win_tri_kernel<7,H> (the kernels of the
    genome-wide bench line) against the oracle itself, pair by pair: aligned masks, unaligned masks, flipped."""
    genome, regions, trees, re_, qe, qe2, pairs = hg38_like
    from chess_b200.engine import region_spans
    assert set(region_spans(trees, [p[1] for p in pairs])[1].tolist()) == {201}      # 2 Mb through the BED lookup
    worst = [0.0, 0.0]
    checked = 0
    for contacts in (qe, qe2):
        eng = cb.CompareEngine(re_, trees, contacts, trees)
        otrees, (ore, oqe) = _oracle_side(regions, re_, contacts)
        plist = pairs[:5] if contacts is qe else pairs[5:9]
        flipped = [(pid + "f", r, cb.GenomicRegion(chromosome=q.chromosome, start=q.start, end=q.end, strand="-"))
                   for pid, r, q in plist[:2]]
        checked += _check_against_oracle(cb, eng, plist + flipped, otrees, ore, oqe, kw, worst)
    assert checked >= 6, checked
    assert worst[0] < 1e-9 and worst[1] < 1e-6, worst


@pytest.mark.parametrize("n_bins,kw", [(401, dict()), (401, dict(absolute_window_size=7)),
                                       (601, dict()), (601, dict(absolute_window_size=5)),
                                       (601, dict(absolute_window_size=9, compute_sn=False))],
                         ids=["401_r1", "401_a7", "601_r1", "601_a5", "601_a9_nosn"])
def test_block_kernels_against_oracle(cb, n_bins, kw):
    """compare_blk* (222 ... 1024 bins; 3 Mb at 5 kb = 601 bins is the DLBCL shape) against the oracle."""
    from chess_b200 import synth
    genome = synth.Genome([("chr2", 1900)], 5000)
    ref = synth.synthetic_sample(genome, n_bins + 14, seed=21, structure_seed=5)
    qry = synth.synthetic_sample(genome, n_bins + 14, seed=22, structure_seed=6, perturb=0.1)
    regions = genome.regions()
    trees = cb.region_interval_trees(regions)
    re_ = cb.observed_expected_arrays(regions, ref)
    qe = cb.observed_expected_arrays(regions, qry)
    eng = cb.CompareEngine(re_, trees, qe, trees)
    otrees, (ore, oqe) = _oracle_side(regions, re_, qe)
    plist = synth.sliding_pairs(genome, (n_bins - 1) * 5000, 2900000)[:2]
    plist.append(("f", plist[0][1], cb.GenomicRegion(chromosome="chr2", start=plist[1][2].start,
                                                     end=plist[1][2].end, strand="-")))
    worst = [0.0, 0.0]
    assert _check_against_oracle(cb, eng, plist, otrees, ore, oqe, kw, worst) >= 2
    assert worst[0] < 1e-9 and worst[1] < 1e-6, worst


def test_background_set_at_201_bins_against_oracle(cb, hg38_like):
    """One --background-regions style set at 201 bins: every comparison the device makes for z_bg / p_bg
    (SN off, strand None -> always flipped, bin/chess:300-307) against the oracle, then the reduction."""
    from chess_b200.engine import region_spans
    genome, regions, trees, re_, qe, qe2, pairs = hg38_like
    eng = cb.CompareEngine(re_, trees, qe2, trees)
    otrees, (ore, oqe) = _oracle_side(regions, re_, qe2)
    pick = [pairs[1], pairs[6]]
    ids, s, sn, st = eng.compare(pick, absolute_window_size=7)
    bg = [cb.GenomicRegion(chromosome=c, start=b + 1, end=b + 2000000, strand=None)
          for c, b in (("chr1", 0), ("chr1", 1230000), ("chr1", 5550000), ("chr1", 12990000),
                       ("chr2", 70000), ("chr2", 3330000), ("chr2", 6990000))]
    rf, rn = region_spans(trees, [p[1] for p in pick])
    bf, bn = region_spans(trees, bg)
    assert (bn == 201).all()
    z, pv = eng.background(rf, rn, np.zeros(2, dtype=np.int32), s, bf, bn, np.full(len(bg), 2, dtype=np.int32),
                           absolute_window_size=7)
    for k, (pid, rr, qr) in enumerate(pick):
        vals = []
        for b in bg:
            _, v, _ = O.compare_pair(("bg", O.Region(rr.start, rr.end, rr.chromosome, strand=rr.strand),
                                      O.Region(b.start, b.end, b.chromosome, strand=None)),
                                     ore, otrees, oqe, otrees, compute_sn=False, absolute_window_size=7)
            vals.append(v)
        _, os_, _ = O.compare_pair(_opair(pid, rr, qr), ore, otrees, oqe, otrees, absolute_window_size=7)
        assert close(s[k], os_, RTOL)
        oz, op = O.background_stats(os_, np.array(vals))
        assert close(z[k], oz, RTOL), (k, z[k], oz)
        assert pv[k] == op, (k, pv[k], op)


def test_fused_log2_transform(cb, hg38_like):
    """north_star's "O/E normalisation and the log transform": apply_log2=True fuses log2 into the O/E kernel
    (default OFF -- `chess sim` itself applies none, chess/oe.py:149,181).  Checked against np.log2 of the
    oracle's O/E values, and a comparison on the log2 data (absent cells = log2(1) = 0) against the oracle
    run on log2'ed edges (the shape of utils/run_synthetic_test.py:34-35, 92-94)."""
    from chess_b200 import synth
    genome, regions, trees, re_, qe, qe2, pairs = hg38_like
    sub = synth.Genome([("chrL", 400)], 10000)
    ref = synth.synthetic_sample(sub, 120, seed=31, structure_seed=3)
    qry = synth.synthetic_sample(sub, 120, seed=32, structure_seed=3, perturb=0.2)
    sregions = sub.regions()
    strees = cb.region_interval_trees(sregions)
    oregions = [O.Region(r.start, r.end, r.chromosome, ix=r.ix) for r in sregions]
    log_contacts, oracle_edges = [], []
    for smp in (ref, qry):
        got = cb.observed_expected_arrays(sregions, smp, apply_log2=True)
        plain = cb.observed_expected_arrays(sregions, smp)
        want = O.observed_expected(oregions, list(zip(*[x.tolist() for x in smp])))
        w = np.array([t[2] for t in want])
        a, b, v = got.triples()
        assert np.array_equal(a, [t[0] for t in want]) and np.array_equal(b, [t[1] for t in want])
        np.testing.assert_allclose(plain.triples()[2], w, rtol=1e-12, atol=0)
        np.testing.assert_allclose(v, np.log2(w), rtol=1e-12, atol=1e-14)   # values near log2(1) = 0
        assert (v < 0).any() and (v > 0).any()
        log_contacts.append(got)
        oracle_edges.append(O.edges_dict(zip(a.tolist(), b.tolist(), np.log2(w).tolist())))
    otrees = O.region_trees(oregions)
    eng = cb.CompareEngine(log_contacts[0], strees, log_contacts[1], strees)
    plist = synth.sliding_pairs(sub, 1000000, 700000)
    ids, s, sn, st = eng.compare(plist, default_value=0, absolute_window_size=7)
    checked = 0
    for k, (pid, rr, qr) in enumerate(plist):
        _, os_, osn = O.compare_pair(_opair(pid, rr, qr), oracle_edges[0], otrees, oracle_edges[1], otrees,
                                     default_value=0, absolute_window_size=7)
        assert close(s[k], os_, RTOL), (pid, s[k], os_)
        assert close(sn[k], osn, RTOL), (pid, sn[k], osn)
        checked += int(not np.isnan(os_))
    assert checked >= 3


def test_regions_reaching_into_trailing_empty_bins(cb):
    """A matrix whose last bins carry no contact at all (chrY / chrM / telomeres): on the --oe-input path the
    store must still span every bin of the regions file; cells there read as the default value, as the
    reference's dict lookup does (chess/helpers.py:297-304)."""
    from chess_b200 import synth
    from chess_b200.helpers import edges_dict_from_sparse
    genome = synth.Genome([("chrT", 300)], 10000)
    a, b, w = synth.synthetic_sample(genome, 60, seed=41, structure_seed=2)
    keep = b < 180                       # nothing at all from bin 180 on
    rng = np.random.default_rng(0)
    w = np.exp(rng.normal(0, 0.3, keep.sum()))       # O/E-like values
    regions = genome.regions()
    trees = cb.region_interval_trees(regions)
    store_r = edges_dict_from_sparse((a[keep], b[keep], w), n_bins=len(regions))
    store_q = {}
    w2 = w * np.exp(rng.normal(0, 0.1, len(w)))
    for i, j, v in zip(a[keep].tolist(), b[keep].tolist(), w2.tolist()):   # the dict form: size unknown up front
        store_q.setdefault(i, {})[j] = v
    eng = cb.CompareEngine(store_r, trees, store_q, trees)
    assert eng.ref.n_bins == eng.qry.n_bins == 300
    plist = [("in", cb.GenomicRegion(chromosome="chrT", start=1000001, end=1400000, strand="+"),
              cb.GenomicRegion(chromosome="chrT", start=1000001, end=1400000, strand="+")),
             ("edge", cb.GenomicRegion(chromosome="chrT", start=1400001, end=1900000, strand="+"),
              cb.GenomicRegion(chromosome="chrT", start=1400001, end=1900000, strand="+")),
             ("mostly_empty", cb.GenomicRegion(chromosome="chrT", start=1700001, end=2100000, strand="+"),
              cb.GenomicRegion(chromosome="chrT", start=1700001, end=2100000, strand="+")),
             ("empty", cb.GenomicRegion(chromosome="chrT", start=2500001, end=2990000, strand="+"),
              cb.GenomicRegion(chromosome="chrT", start=2500001, end=2990000, strand="+"))]
    oregions = [O.Region(r.start, r.end, r.chromosome, ix=r.ix) for r in regions]
    otrees = O.region_trees(oregions)
    ore = O.edges_dict(zip(a[keep].tolist(), b[keep].tolist(), w.tolist()))
    oqe = O.edges_dict(zip(a[keep].tolist(), b[keep].tolist(), w2.tolist()))
    for kw in (dict(), dict(absolute_window_size=7), dict(mappability_cutoff=0.5),
               dict(keep_unmappable_bins=True, absolute_window_size=7)):
        for kernel in (0, 1):
            ids, s, sn, st = eng.compare(plist, kernel=kernel, **kw)
            for k, (pid, rr, qr) in enumerate(plist):
                _, os_, osn = O.compare_pair(_opair(pid, rr, qr), ore, otrees, oqe, otrees, **kw)
                assert close(s[k], os_, RTOL), (kw, kernel, pid, s[k], os_)
                assert close(sn[k], osn, RTOL), (kw, kernel, pid, sn[k], osn)
    m = cb.sub_matrix_from_edges_dict(store_r, trees, plist[1][1], default_weight=1.0)[0]
    om = O.sub_matrix_from_edges_dict(ore, otrees, _opair(*plist[1])[1], default_weight=1.0)[0]
    assert np.array_equal(m, om) and (m[-5:, :] == 1.0).all()


def test_extract_head_against_reference_golden(cb, toy):
    """K9 against tests/golden/extract_golden.npz: `positive`, `negative`, np.std and the two means exactly as
    the reference's own extract_structures computed them (get_structures.py:99-121, captured at its
    denoise_bilateral call by tests/golden/make_golden_extract.py)."""
    from chess_b200.extract import log_ratio_stage
    from chess_b200.helpers import load_oe_contacts, load_pairs_for_chroms
    gold = np.load(os.path.join(GOLDEN_DIR, "extract_golden.npz"))
    redges, rtrees, _ = load_oe_contacts(toy("ref.oe.sparse"), toy("ref.bed"))
    qedges, qtrees, _ = load_oe_contacts(toy("qry.oe.sparse"), toy("qry.bed"))
    pairs, _ = load_pairs_for_chroms(toy("pairs.bedpe"), set(rtrees.keys()) | set(qtrees.keys()))
    assert [p[0] for p in pairs] == gold["all_ids"].tolist()
    got = log_ratio_stage(redges, rtrees, qedges, qtrees, pairs)
    have = set(gold["ids"].tolist())
    checked = 0
    for (pid, rr, qr), g in zip(pairs, got):
        if pid not in have:
            assert g is None, pid                        # the reference raises for these pairs
            continue
        pos, neg, (std, mp, mn) = gold["pos_" + pid], gold["neg_" + pid], gold["stats_" + pid]
        assert g is not None and g[0].shape == pos.shape, pid
        assert close(g[2], std, 1e-12), (pid, g[2], std)
        # log is not correctly rounded on either side: a value within rounding of a threshold may fall either way
        edge = (np.abs(np.maximum(pos, g[0]) - 0.5 * std) < 1e-12) | (np.abs(np.maximum(neg, g[1]) - 0.5 * std) < 1e-12)
        assert edge.sum() <= 2
        np.testing.assert_allclose(g[0][~edge], pos[~edge], rtol=1e-12, atol=0)
        np.testing.assert_allclose(g[1][~edge], neg[~edge], rtol=1e-12, atol=0)
        if not edge.any():
            assert close(g[3], mp, 1e-11) and close(g[4], mn, 1e-11), pid
        checked += 1
    assert checked == len(have) >= 10


def test_mismatched_matrix_is_rejected_before_any_launch(cb):
    """A sparse matrix that does not go with the regions file: IndexError as in the reference (oe.py:107),
    not device memory corruption."""
    from chess_b200 import synth
    genome = synth.Genome([("c", 50)], 10000)
    regions = genome.regions()
    with pytest.raises(IndexError):
        cb.observed_expected_arrays(regions, (np.array([0, 3]), np.array([1, 60]), np.array([1.0, 2.0])))
    trees = cb.region_interval_trees(regions)
    big = synth.Genome([("c", 1500)], 10000)
    good = cb.observed_expected_arrays(big.regions(), synth.synthetic_sample(big, 30, seed=1))
    eng = cb.CompareEngine(good, cb.region_interval_trees(big.regions()), good, cb.region_interval_trees(big.regions()))
    huge = cb.GenomicRegion(chromosome="c", start=1, end=11000000, strand="+")
    with pytest.raises(ValueError, match="at most 1024 bins"):
        eng.compare([("huge", huge, huge)])


def test_full_size_hg38_properties(cb):
    """BASELINE configs[2] at full size (hg38 @ 10 kb, 101 354 pairs of 201 bins): size-independent properties
    of the kernels the north-star number is quoted on."""
    import bench
    w = bench.workload_spec(bench.parse_args([]))
    assert w["name"] == "hg38_10kb" and w["pairs"] == 101354
    genome, ref, qry, pairs = bench.make_workload(w)
    assert len(pairs) == w["pairs"] and genome.n_bins == w["n_bins"]
    regions = genome.regions()
    trees = cb.region_interval_trees(regions)
    re_ = cb.observed_expected_arrays(regions, ref)
    qe = cb.observed_expected_arrays(regions, qry)
    fwd = cb.CompareEngine(re_, trees, qe, trees)
    rev = cb.CompareEngine(qe, trees, re_, trees)
    for kw in (dict(), dict(absolute_window_size=7)):
        ids, s, sn, st = fwd.compare(pairs, **kw)
        assert ids == [p[0] for p in pairs] and (st == 0).sum() > 95000
        # SSIM and SN are symmetric in (reference, query)
        _, s_r, sn_r, st_r = rev.compare(pairs, **kw)
        assert np.array_equal(st, st_r)
        np.testing.assert_allclose(s, s_r, rtol=1e-10, equal_nan=True)
        np.testing.assert_allclose(sn, sn_r, rtol=1e-9, equal_nan=True)
        # results do not depend on how the pair list is cut into launches (1 / 2 / 4 / 8 GPUs cut it like this)
        from chess_b200.distributed import shard_bounds
        for world in (2, 8):
            b = shard_bounds(len(pairs), world)
            cut = [fwd.compare(pairs[b[k]:b[k + 1]], **kw) for k in range(world)]
            assert np.concatenate([c[1] for c in cut]).tobytes() == s.tobytes()
            assert np.concatenate([c[2] for c in cut]).tobytes() == sn.tobytes()
        # SSIM is bounded, SN non-negative, failures are NaN in both columns
        ok = st == 0
        assert np.all(np.abs(s[ok]) <= 1.0 + 1e-12) and np.all(sn[ok] >= 0)
        assert np.all(np.isnan(s[~ok])) and np.all(np.isnan(sn[~ok]))
        # a sample against the generic kernel (a different implementation of the same function)
        pick = pairs[::977]
        a = fwd.compare(pick, **kw)
        g = fwd.compare(pick, kernel=1, **kw)
        assert np.array_equal(a[3], g[3])
        np.testing.assert_allclose(a[1], g[1], rtol=1e-9, equal_nan=True)
        np.testing.assert_allclose(a[2], g[2], rtol=1e-8, equal_nan=True)
        assert np.array_equal(a[1], s[::977], equal_nan=True)      # and a pair's bits do not depend on its batch
    # a matrix against itself
    ids, s, sn, st = cb.CompareEngine(re_, trees, re_, trees).compare(pairs[::50])
    ok = st == 0
    assert np.all(np.abs(s[ok] - 1.0) < 1e-12) and np.all(np.isnan(sn[ok]))
    re_.drop_device_copies()
    qe.drop_device_copies()


def test_ranged_bands_and_device_resident_triples(cb, hg38_like, tmp_path):
    """Per-GPU placement: a GPU holds the band of the bin range its share of the pair list reads
    (chess_band_build_range), fed by a slice of the sorted COO list; triples that were parsed on the GPU are
    normalised and kept there (DeviceTriples -> DeviceContacts.from_device).  Same bytes as the whole band."""
    import torch
    from chess_b200 import synth
    from chess_b200.engine import DeviceTriples, DeviceContacts, coo_facts
    from chess_b200.helpers import load_contacts
    genome, regions, trees, re_, qe, qe2, pairs = hg38_like
    whole = cb.CompareEngine(re_, trees, qe2, trees, devices=[0]).compare(pairs, absolute_window_size=7)
    # three shards on one GPU, each with the band of its own bin range
    for devs in ([0, 0, 0], [0] * 7):
        part = cb.CompareEngine(re_, trees, qe2, trees, devices=devs, placement="range").compare(
            pairs, absolute_window_size=7)
        assert part[1].tobytes() == whole[1].tobytes() and part[2].tobytes() == whole[2].tobytes()
        assert np.array_equal(part[3], whole[3])
    # the same samples through files: parsed on the device, O/E on the device, never on the host
    files = synth.write_dataset(str(tmp_path), genome, synth.synthetic_sample(genome, 215, seed=11, structure_seed=7),
                                synth.synthetic_sample(genome, 215, seed=13, structure_seed=9, perturb=0.1), pairs)
    r_edges, r_trees, r_regions = load_contacts(files["reference"], files["regions"], device=0)
    q_edges, q_trees, q_regions = load_contacts(files["query"], files["regions"], device=0)
    assert r_edges._host is None and r_edges._home == 0 and r_edges.sorted_unique      # still only on the GPU
    assert len(r_regions) == genome.n_bins and r_regions[7].ix == 7
    for devs, placement in (([0], "auto"), ([0, 0, 0], "range")):
        got = cb.CompareEngine(r_edges, r_trees, q_edges, q_trees, devices=devs, placement=placement).compare(
            pairs, absolute_window_size=7)
        # %.17g text round-trips every weight, so the file path gives the very same numbers
        assert got[1].tobytes() == whole[1].tobytes() and got[2].tobytes() == whole[2].tobytes()
    assert np.array_equal(r_edges.triples()[2], re_.triples()[2])      # host view on demand
    # an unsorted / repeated list on the device goes through the host constructor (last entry wins)
    dev = torch.device("cuda", 0)
    t = DeviceTriples(torch.tensor([3, 0, 3, 1], dtype=torch.int32, device=dev),
                      torch.tensor([4, 1, 4, 1], dtype=torch.int32, device=dev),
                      torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64, device=dev), 0)
    assert coo_facts(t.src, t.sink, 0) == [0, 4, 0, 2]
    dc = DeviceContacts.from_device(t, n_bins=6)
    assert dc[3] == {4: 3.0} and dc[0] == {1: 2.0} and dc[1] == {1: 4.0} and dc.n_bins == 6


def test_background_regions_enumerated_on_the_device(cb, toy, hg38_like):
    """chess_background_regions (bin/chess:272-298 on the GPU) against the host columns and, through them
    (tests/test_host_logic.py), against the reference's loop: same regions, same order, same bin spans."""
    from chess_b200 import cli, helpers as H
    from chess_b200.engine import QueryBackgroundSet
    table, _ = H.load_regions_table(toy("qry.bed"))
    trees = H.region_interval_trees(table)
    cases = [(table, trees, H.GenomicRegion(chromosome="chrA", start=4001, end=28000, strand="+"), True),
             (table, trees, H.GenomicRegion(chromosome="chrB", start=1, end=20000, strand="-"), False),
             (table, trees, H.GenomicRegion(chromosome="chrQ", start=1, end=20000), True)]
    genome = hg38_like[0]
    big = H.RegionTable(genome.names, np.repeat(np.arange(len(genome.names)), genome.sizes),
                        [r.start for r in hg38_like[1]], [r.end for r in hg38_like[1]])
    big_trees = H.region_interval_trees(big)
    cases += [(big, big_trees, H.GenomicRegion(chromosome="chr2", start=1, end=2000000), True),
              (big, big_trees, H.GenomicRegion(chromosome="chr1", start=10001, end=1230000), False)]
    for t, tr, q, limit in cases:
        want = cli._background_spans_from_query(t, tr, q, limit)
        bg = QueryBackgroundSet(t, q.end - q.start, q.chromosome if limit else None)
        assert bg.ok
        got = bg.host(0)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), (q.chromosome, limit, len(a), len(b))
    # an irregular BED (bins out of order) is left to the host path
    odd = H.RegionTable(["c"], [0, 0, 0], [20, 0, 10], [30, 10, 20])
    assert not QueryBackgroundSet(odd, 10).ok


def test_host_call_serves_pinned_and_pageable_arrays_alike(cb, hg38_like):
    """chess_sim_band_batch_host copies straight from / into page-locked caller arrays and stages pageable
    ones through its own pinned area: the same bytes either way, for a table above and one below the 8 192-pair
    limit of the direct-read path."""
    import ctypes as C
    import torch
    from chess_b200 import _native
    from chess_b200._native import lib, check
    from chess_b200.engine import get_context, region_spans, build_pair_table
    genome, regions, trees, re_, qe, qe2, pairs = hg38_like
    rf, rn = region_spans(trees, [p[1] for p in pairs])
    qf, qn = region_spans(trees, [p[2] for p in pairs])
    max_n = int(max(rn.max(), qn.max()))
    rw, rband, rsample, r0 = re_.sample_on(0, max_n, None, None, with_origin=True)
    qw, qband, qsample, q0 = qe.sample_on(0, max_n, None, None, with_origin=True)
    flip = np.zeros(len(pairs), dtype=np.int32)
    small = build_pair_table(rf - r0, rn, rw, qf - q0, qn, qw, flip)
    ctx = get_context(0)
    sp = C.c_void_p(torch.cuda.current_stream(torch.device("cuda", 0)).cuda_stream)
    p = _native.make_params(compute_sn=True, flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES)
    for table in (small, np.tile(small, (9000 // len(small)) + 1)):
        n = len(table)
        got = {}
        for kind in ("pageable", "pinned"):
            if kind == "pinned":
                h_tab = torch.from_numpy(table.view(np.uint8).copy()).pin_memory().numpy()
                outs = [torch.empty(n, dtype=dt).pin_memory().numpy()
                        for dt in (torch.float64, torch.float64, torch.int32, torch.float64)]
            else:
                h_tab = table.view(np.uint8).copy()
                outs = [np.empty(n), np.empty(n), np.empty(n, dtype=np.int32), np.empty(n)]
            ptr = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
            check(lib.chess_sim_band_batch_host(ctx.handle, C.byref(rsample), C.byref(qsample), ptr(h_tab), n, max_n,
                                                C.byref(p), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]),
                                                sp), ctx.handle, kind)
            got[kind] = [o.copy() for o in outs]
        for a, b in zip(got["pageable"], got["pinned"]):
            assert a.tobytes() == b.tobytes()
        assert (got["pinned"][2] == 0).sum() > 0.8 * n
        # and the engine (device-resident path) gives the same values for the pairs of the list
        ids, s, sn, st = cb.CompareEngine(re_, trees, qe, trees).compare(pairs)
        np.testing.assert_allclose(got["pinned"][0][:len(pairs)], s, rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(got["pinned"][1][:len(pairs)], sn, rtol=1e-10, equal_nan=True)
        # z_ssim of the call = bin/chess:228-233 over its own ssim column
        ss = got["pinned"][0]
        valid = ss[~np.isnan(ss)]
        np.testing.assert_allclose(got["pinned"][3], (ss - valid.mean()) / valid.std(), rtol=1e-9, atol=1e-12,
                                   equal_nan=True)


def test_kernels_match_reference_golden_at_201_bins(cb):
    """The CUDA path against results of the REFERENCE's own compare_pair at the north-star shape (10 kb bins,
    2 Mb regions = 201 bins; tests/golden/make_golden_201.py), both strands, six parameter sets."""
    from tests.test_oracle_golden import _golden_201_inputs, GOLDEN_201_PARAMS
    g, genome, a, b = _golden_201_inputs()
    regions = genome.regions(cb.GenomicRegion)      # (the generator module was loaded by file path: no package)
    trees = cb.region_interval_trees(regions)
    re_ = cb.observed_expected_arrays(regions, a)
    qe = cb.observed_expected_arrays(regions, b)
    bin_size = int(g["bin"])
    pairs = []
    for k, st in enumerate(int(v) for v in g["starts"]):
        for strand in ("+", "-"):
            r = cb.GenomicRegion(chromosome="chr1", start=st * bin_size + 1, end=st * bin_size + 2000000, strand="+")
            q = cb.GenomicRegion(chromosome="chr1", start=st * bin_size + 1, end=st * bin_size + 2000000, strand=strand)
            pairs.append(("p%d%s" % (k, strand), r, q))
    eng = cb.CompareEngine(re_, trees, qe, trees)
    for tag, kw in GOLDEN_201_PARAMS.items():
        for kernel in (0, 1):
            ids, s, sn, st = eng.compare(pairs, kernel=kernel, **kw)
            for k in range(len(pairs)):
                assert close(s[k], g["ssim_" + tag][k], RTOL), (tag, kernel, k, s[k], g["ssim_" + tag][k])
                assert close(sn[k], g["sn_" + tag][k], RTOL), (tag, kernel, k, sn[k], g["sn_" + tag][k])
