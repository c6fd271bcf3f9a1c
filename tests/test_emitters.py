"""`chess pairs` / `chess background` (bin/chess:382-518) against files written by the reference's own
CLI (tests/golden/make_golden_emitters.py), and the array generators behind them."""
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TOY = os.path.join(HERE, "golden", "toy")
P = lambda name: os.path.join(TOY, name)  # noqa: E731


def _run(argv):
    from chess_b200.cli import Chess
    Chess(["chess"] + argv)


@pytest.mark.parametrize("golden,argv", [
    ("emit_pairs.bedpe", ["pairs", "{sizes}", "20000", "7000", "{out}", "--file-input"]),
    ("emit_pairs.bedpe", ["pairs", "{sizes}", "20000", "7000", "{out}"]),   # no pybedtools: falls back to the file
    ("emit_pairs_chrB.bedpe", ["pairs", "{sizes}", "9000", "4000", "{out}", "--chromosome", "chrB"]),
    ("emit_background.bed", ["background", "{sizes}", "20000", "9000", "{out}"]),
    ("emit_background_minus.bed", ["background", "{sizes}", "8000", "5000", "{out}", "--strand", "-"]),
    ("emit_background_region.bed", ["background", "chrA:1001-40000", "10000", "6000", "{out}", "--strand", "+"]),
])
def test_emitters_match_reference_files(tmp_path, golden, argv):
    out = str(tmp_path / "out.txt")
    _run([a.format(sizes=P("chrom.sizes"), out=out) for a in argv])
    assert open(out).read() == open(P(golden)).read()


def test_emitter_errors(tmp_path):
    out = str(tmp_path / "o.bed")
    with pytest.raises(ValueError):
        _run(["pairs", P("chrom.sizes"), "1000", "500", out, "--chromosome", "chrZ"])
    with pytest.raises(ValueError):
        _run(["background", P("chrom.sizes"), "1000", "500", out, "--strand", "x"])
    with pytest.raises(RuntimeError):
        _run(["pairs", P("chrom.sizes"), "1000", "500", str(tmp_path / "missing" / "o.bedpe")])


def test_window_pairs_feed_the_engine_like_the_file(tmp_path):
    """The tuples built from arrays are the tuples `chess sim` loads from the BEDPE file."""
    from chess_b200 import emit
    from chess_b200.helpers import load_pairs_iter
    sizes = emit.load_chromosome_sizes(P("chrom.sizes"))
    assert list(sizes.items()) == [("chrA", 60000), ("chrB", 25500), ("chrC", 9000)]
    path = str(tmp_path / "p.bedpe")
    n = emit.write_pairs(path, sizes, 9000, 4000)
    direct = emit.window_pair_tuples(sizes, 9000, 4000)
    loaded = list(load_pairs_iter(path))
    assert n == len(direct) == len(loaded) > 10
    for (i1, a1, b1), (i2, a2, b2) in zip(direct, loaded):
        assert str(i1) == str(i2)
        for x, y in ((a1, a2), (b1, b2)):
            assert (x.chromosome, x.start, x.end, x.strand) == (y.chromosome, y.start, y.end, y.strand)
    names, cix, starts, ends, ids = emit.window_pairs(sizes, 9000, 4000, chromosome="chrC")
    assert names == ["chrC"] and starts.tolist() == [0] and ends.tolist() == [9000]
    assert len(emit.window_pairs({"tiny": 10}, 100, 10)[2]) == 0
    rows = emit.background_windows([("c", 1, 50)], 10, 20)
    assert rows == [("c", 1, 11, 0, "+"), ("c", 1, 11, 1, "-"), ("c", 21, 31, 2, "+"), ("c", 21, 31, 3, "-")]
