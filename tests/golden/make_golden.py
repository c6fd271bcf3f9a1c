#!/usr/bin/env python
"""Generate golden vectors by executing the REFERENCE's own code.

Run in the build container only (it reads /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

The reference (chess-hic 0.3.8) cannot be imported as a package here because
`future`, `intervaltree`, `scikit-image` and `fanc` are not installed.  Its
modules are therefore loaded file by file with four shims:

* `future.utils.string_types`            -> (str,)
* `intervaltree.Interval/IntervalTree`   -> a half-open linear-scan stand-in
* `skimage.metrics.structural_similarity`, `skimage.transform.resize`
                                         -> the oracle's restatements
                                            (PARITY UNPINNED for these two,
                                            see oracle/chess_oracle.py)
* `fanc.load`                            -> raises ValueError, which is the
                                            reference's own cue to take the
                                            sparse-format path
                                            (chess/helpers.py:558-570)

Everything else that runs is the reference's code, unmodified, including
`Chess.sim` from bin/chess (multiprocessing pool and all).  Outputs are
written next to this script as small JSON / text fixtures; floats are stored
with `float.hex()` so they round-trip exactly.
"""
import importlib.util
import json
import os
import queue
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import chess_oracle as oracle  # noqa: E402


# ----------------------------------------------------------------- shims
def install_shims():
    fut = types.ModuleType("future")
    fut_utils = types.ModuleType("future.utils")
    fut_utils.string_types = (str,)
    fut.utils = fut_utils
    sys.modules["future"] = fut
    sys.modules["future.utils"] = fut_utils

    it = types.ModuleType("intervaltree")

    class Interval(object):
        __slots__ = ("begin", "end", "data")

        def __init__(self, begin, end, data=None):
            self.begin, self.end, self.data = begin, end, data

    class IntervalTree(object):
        def __init__(self, intervals=()):
            self.ivs = list(intervals)

        def __getitem__(self, sl):
            a, b = sl.start, sl.stop
            return set(iv for iv in self.ivs if iv.begin < b and iv.end > a)

    it.Interval, it.IntervalTree = Interval, IntervalTree
    sys.modules["intervaltree"] = it

    sk = types.ModuleType("skimage")
    skt = types.ModuleType("skimage.transform")
    skm = types.ModuleType("skimage.metrics")

    def resize(image, output_shape, order=0, mode="reflect"):
        assert order == 0 and mode == "reflect"
        return oracle.resize_order0(image, output_shape)

    def structural_similarity(im1, im2, win_size=None, data_range=None):
        return oracle.structural_similarity(im1, im2, win_size=win_size, data_range=data_range)

    skt.resize = resize
    skm.structural_similarity = structural_similarity
    sk.transform, sk.metrics = skt, skm
    sys.modules["skimage"] = sk
    sys.modules["skimage.transform"] = skt
    sys.modules["skimage.metrics"] = skm

    fanc = types.ModuleType("fanc")
    fanc.__version__ = "shim"

    def load(path, *a, **k):
        raise ValueError("fanc shim: {} is not a fanc/cooler/juicer file".format(path))

    fanc.load = load
    sys.modules["fanc"] = fanc


def load_reference():
    """Load chess.{version,helpers,oe,sim,commands} without chess/__init__."""
    pkg = types.ModuleType("chess")
    pkg.__path__ = [os.path.join(REF, "chess")]
    sys.modules["chess"] = pkg
    mods = {}
    for name in ("version", "helpers", "oe", "sim", "commands"):
        spec = importlib.util.spec_from_file_location(
            "chess." + name, os.path.join(REF, "chess", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["chess." + name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        mods[name] = mod
    pkg.__version__ = mods["version"].__version__
    return mods


def load_cli_class():
    src = open(os.path.join(REF, "bin", "chess")).read()
    glb = {"__name__": "chess_cli_ref"}
    exec(compile(src, os.path.join(REF, "bin", "chess"), "exec"), glb)
    return glb["Chess"]


# ----------------------------------------------------------------- helpers
def hx(v):
    if v is None:
        return None
    v = float(v)
    if np.isnan(v):
        return "nan"
    return v.hex()


def hx_arr(a):
    return [hx(v) for v in np.asarray(a, dtype=np.float64).ravel()]


# ----------------------------------------------------------------- toy genome
CHROMS = [("chrA", 64), ("chrB", 48), ("chrC", 30)]
BIN = 1000


def make_toy(seed, unmappable, explicit_zero_frac=0.01, named=False):
    """(bed_text, sparse_text) of a small banded 3-chromosome sample."""
    rng = np.random.default_rng(seed)
    bed, off, starts = [], 0, {}
    for c, n in CHROMS:
        starts[c] = off
        for k in range(n):
            name = "\t%d" % (off + k + 1) if named else ""  # HiC-Pro: 1-based integer names
            bed.append("%s\t%d\t%d%s\n" % (c, k * BIN, (k + 1) * BIN, name))
        off += n
    total = off
    lines = ["# toy sparse matrix seed=%d\n" % seed]
    dead = set(unmappable)
    for c, n in CHROMS:
        o = starts[c]
        tad = rng.integers(4, 12)
        for i in range(n):
            for j in range(i, min(n, i + 40)):
                gi, gj = o + i, o + j
                if gi in dead or gj in dead:
                    continue
                d = j - i
                if d > 6 and rng.random() < 0.15 + 0.01 * d:
                    continue  # absent entry
                base = 10.0 * (d + 1) ** -0.85
                if (i // tad) == (j // tad):
                    base *= 1.8
                w = base * rng.lognormal(0, 0.35)
                if rng.random() < explicit_zero_frac:
                    w = 0.0
                a, b = (gi, gj) if rng.random() < 0.8 else (gj, gi)  # some lower-triangle lines
                if named:
                    lines.append("%d\t%d\t%r\n" % (a + 1, b + 1, w))
                else:
                    lines.append("%d\t%d\t%r\n" % (a, b, w))
    # a few inter-chromosomal entries and one non-finite value
    for _ in range(12):
        a = int(rng.integers(0, CHROMS[0][1]))
        b = int(rng.integers(CHROMS[0][1], total))
        if a in dead or b in dead:
            continue
        if named:
            lines.append("%d\t%d\t%r\n" % (a + 1, b + 1, float(rng.random())))
        else:
            lines.append("%d\t%d\t%r\n" % (a, b, float(rng.random())))
    if not named:
        lines.append("3\t5\tnan\n")
    return "".join(bed), "".join(lines)


def make_pairs():
    """BEDPE text exercising strands, sizes, quirks and failure modes."""
    rows = []
    pid = 0

    def add(c1, s1, e1, c2, s2, e2, st1="+", st2="+"):
        nonlocal pid
        rows.append("%s\t%d\t%d\t%s\t%d\t%d\t%d\t.\t%s\t%s\n" % (c1, s1, e1, c2, s2, e2, pid, st1, st2))
        pid += 1

    for s in range(0, 36, 4):           # chrA sliding 24-kb windows, same region both sides
        add("chrA", s * BIN, (s + 24) * BIN, "chrA", s * BIN, (s + 24) * BIN)
    for s in range(0, 20, 5):           # chrB, minus-strand query -> flipped
        add("chrB", s * BIN, (s + 26) * BIN, "chrB", s * BIN, (s + 26) * BIN, "+", "-")
    add("chrA", 2000, 30000, "chrB", 1000, 31000)            # 29 vs 31 bins -> resize ref
    add("chrA", 5000, 45000, "chrB", 3000, 28000, "-", "+")  # 41 vs 26 -> resize qry + flip
    add("chrA", 10000, 22000, "chrA", 10000, 22000)          # 13 bins < min_bins -> nan
    add("chrC", 0, 29999, "chrC", 0, 29999)                  # whole chrC (30 bins)
    add("chrA", 40000, 80000, "chrA", 40000, 80000)          # runs past chrA end (64 bins)
    add("chrA", 90000, 120000, "chrA", 90000, 120000)        # beyond the chromosome -> not found -> nan
    add("chrZ", 0, 30000, "chrA", 0, 30000)                  # unknown chromosome -> dropped
    add("chrB", 1500, 26500, "chrB", 20500, 45500)           # unaligned coordinates, different loci
    return "".join(rows)


def main():
    install_shims()
    ref = load_reference()
    H, OE, SIM = ref["helpers"], ref["oe"], ref["sim"]
    out = {"generator": "tests/golden/make_golden.py", "reference": "chess-hic " + ref["version"].__version__,
           "numpy": np.__version__, "oracle": oracle.ORACLE_VERSION,
           "unpinned": ["skimage.metrics.structural_similarity", "skimage.transform.resize",
                        "intervaltree slice semantics", "fanc"]}

    # ---- files
    files = {}
    files["ref.bed"], files["ref.sparse"] = make_toy(1, unmappable=[7, 8, 30, 70, 71, 72, 100])
    files["qry.bed"], files["qry.sparse"] = make_toy(2, unmappable=[8, 9, 31, 85, 120])
    files["named.bed"], files["named.sparse"] = make_toy(3, unmappable=[5], named=True)
    files["pairs.bedpe"] = make_pairs()
    bg_rows = []
    for c, n in CHROMS[:2]:
        for s in range(0, n - 24, 6):
            bg_rows.append("%s\t%d\t%d\n" % (c, s * BIN, (s + 24) * BIN))
    files["background.bed"] = "".join(bg_rows)
    data_dir = os.path.join(HERE, "toy")
    os.makedirs(data_dir, exist_ok=True)
    for name, text in files.items():
        with open(os.path.join(data_dir, name), "w") as fh:
            fh.write(text)
    P = lambda n: os.path.join(data_dir, n)  # noqa: E731

    # ---- SN / masks on seeded matrices  (chess/sim.py:15-102)
    rng = np.random.default_rng(11)
    sn_cases = []
    for n in (20, 23, 37, 64):
        a = rng.lognormal(0, 0.4, (n, n))
        a = (a + a.T) / 2
        b = a * rng.lognormal(0, 0.2, (n, n))
        b = (b + b.T) / 2
        if n == 37:           # identical block -> zero-variance windows -> NaN skipped
            b[:12, :12] = a[:12, :12]
        sn_cases.append({"n": n, "a": hx_arr(a), "b": hx_arr(b), "sn": hx(SIM.SN(a, b))})
    out["SN"] = sn_cases

    m = rng.lognormal(0, 0.3, (12, 12))
    m = (m + m.T) / 2
    for k in (2, 7):
        m[k, :] = 1.0
        m[:, k] = 1.0
    m[0, 1] = m[1, 0] = 0.5  # column 1 unaffected
    idx = SIM.find_masked_rows(m, masking_value=1)
    out["masked"] = {"m": hx_arr(m), "n": 12, "idx": [int(v) for v in idx],
                     "removed": hx_arr(SIM.remove_rows(m, idx))}

    # ---- O/E  (chess/oe.py)
    oe_out, oe_lists = {}, {}
    for tag in ("ref", "qry", "named"):
        regions, conv, _ = H.load_regions(P(tag + ".bed"))
        sparse = H.load_sparse_matrix(P(tag + ".sparse"), ix_converter=conv)
        intra, chrom_intra, inter = OE.expected_values(regions, sparse)
        marg = [0.0] * len(regions)
        for s, t, w in sparse:
            marg[s] += w
            marg[t] += w
        mappable = [v > 0 for v in marg]
        tot, chrom_tot, inter_tot = OE.possible_contacts(regions, mappable)
        _, oe_edges = OE.observed_expected(P(tag + ".bed"), P(tag + ".sparse"))
        oe_out[tag] = {
            "n_edges": len(sparse),
            "intra_expected": hx_arr(intra),
            "chrom_expected": {c: hx_arr(v) for c, v in chrom_intra.items()},
            "inter_expected": hx(inter),
            "possible_intra": [int(v) for v in tot],
            "possible_chrom": {c: [int(x) for x in v] for c, v in chrom_tot.items()},
            "possible_inter": int(inter_tot),
            "mappable": [bool(v) for v in mappable],
            # ref/qry O/E edges are kept as toy/<tag>.oe.sparse (repr floats round-trip exactly)
            "oe": [[int(s), int(t), hx(w)] for s, t, w in oe_edges] if tag == "named" else None,
        }
        oe_lists[tag] = oe_edges
    out["oe"] = oe_out

    # ---- gather  (chess/helpers.py:222-306)
    ref_edges, ref_trees, ref_regions = H.load_contacts(P("ref.sparse"), P("ref.bed"))
    qry_edges, qry_trees, qry_regions = H.load_contacts(P("qry.sparse"), P("qry.bed"))
    gather = []
    for chrom, start, end in (("chrA", 1, 24000), ("chrA", 4001, 28000), ("chrB", 1001, 30500),
                              ("chrC", 1, 30000), ("chrA", 50001, 90000)):
        reg = H.GenomicRegion(chromosome=chrom, start=start, end=end)
        mat, sr = H.sub_matrix_from_edges_dict(ref_edges, ref_trees, reg, default_weight=1)
        gather.append({"chrom": chrom, "start": start, "end": end, "n": int(mat.shape[0]),
                       "first_ix": int(min(r.ix for r in sr)), "m": hx_arr(mat)})
    try:
        H.sub_matrix_from_edges_dict(ref_edges, ref_trees,
                                     H.GenomicRegion(chromosome="chrA", start=90001, end=120000), 1)
        raised = None
    except ValueError as e:
        raised = "ValueError"
    out["gather"] = {"cases": gather, "out_of_range": raised}

    # ---- pairs loading  (chess/helpers.py:483-527)
    chroms = set(ref_trees.keys()) | set(qry_trees.keys())
    pairs, dropped = H.load_pairs_for_chroms(P("pairs.bedpe"), chroms)
    out["pairs"] = {"dropped": dropped,
                    "rows": [[pid, r.chromosome, r.start, r.end, r.strand,
                              q.chromosome, q.start, q.end, q.strand] for pid, r, q in pairs]}

    # ---- chunk worker under several parameter sets  (chess/sim.py:147-222, 306-380)
    param_sets = {
        "default": dict(),
        "abs7": dict(absolute_window_size=7),
        "abs5_keep": dict(absolute_window_size=5, keep_unmappable_bins=True),
        "rel05": dict(relative_window_size=0.5),
        "cutoff03": dict(mappability_cutoff=0.3, absolute_window_size=9),
        "abs_even": dict(absolute_window_size=8),
    }
    worker = {}
    for tag, kw in param_sets.items():
        qi, qo = queue.Queue(), queue.Queue()
        qi.put([pairs, True])
        qi.put(None)
        SIM.chunk_comparison_worker(qi, qo, ref_edges, ref_trees, qry_edges, qry_trees,
                                    20, kw.get("keep_unmappable_bins", False),
                                    kw.get("absolute_window_size"), kw.get("relative_window_size", 1.),
                                    kw.get("mappability_cutoff", 0.1))
        res = qo.get()
        if isinstance(res, Exception):
            raise res
        worker[tag] = {"params": kw, "rows": [[pid, hx(s), hx(sn)] for pid, s, sn in res]}
    out["worker"] = worker

    # compare_pair without SN, and compare_structures (sim.py:225-303, 383-459)
    out["compare_pair_nosn"] = []
    for pair in pairs[:4]:
        pid, s, sn = SIM.compare_pair(pair, ref_edges, ref_trees, qry_edges, qry_trees, compute_sn=False)
        out["compare_pair_nosn"].append([pid, hx(s), sn])
    cs = SIM.compare_structures(ref_edges, ref_trees, qry_edges, qry_trees, pairs[:6], "w0",
                                absolute_window_size=7)
    out["compare_structures_abs7"] = {k: [hx(v[0]), hx(v[1])] for k, v in cs.items()}

    # ---- the CLI itself  (bin/chess:80-359), two runs
    Chess = load_cli_class()
    for tag, extra in (("cli_default", []),
                       ("cli_abs7_bg", ["-a", "7", "--background-regions", P("background.bed")]),
                       ("cli_bgquery_limit", ["-a", "5", "--background-query", "--limit-background"]),
                       ("cli_bgquery_all", ["--background-query", "--mappability-cutoff", "0.3"])):
        out_path = P(tag + ".tsv")
        argv = ["chess", "sim", P("ref.sparse"), P("qry.sparse"), P("pairs.bedpe"), out_path,
                "--reference-regions", P("ref.bed"), "--query-regions", P("qry.bed"), "-p", "1"] + extra  # p=1: with p>1 the z-score sum order follows queue arrival order
        obj = object.__new__(Chess)
        obj.sim(argv)
    # the --oe-input path: feed the reference's own O/E output back in
    for tag in ("ref", "qry"):
        with open(P(tag + ".oe.sparse"), "w") as fh:
            for s, t, w in oe_lists[tag]:
                fh.write("%d\t%d\t%r\n" % (s, t, float(w)))
    obj = object.__new__(Chess)
    obj.sim(["chess", "sim", P("ref.oe.sparse"), P("qry.oe.sparse"), P("pairs.bedpe"), P("cli_oe_input.tsv"),
             "--reference-regions", P("ref.bed"), "--query-regions", P("qry.bed"), "--oe-input",
             "-r", "0.4", "--keep-unmappable-bins"])

    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print("wrote", os.path.join(HERE, "golden.json"), os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
