"""Wall-clock of the whole `chess sim` command on synthetic chr2 @ 25 kb files (development tool).

    python tools/cli_e2e.py [--background]

Writes ref/qry sparse matrices, the region BED and the pair BEDPE to a temp dir, runs the CLI in-process and
prints the time of each stage (from the log timestamps) and the total.
"""
import argparse
import logging
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--background", action="store_true")
    ap.add_argument("--window", default=None)
    args = ap.parse_args()
    import bench
    from chess_b200.cli import Chess
    genome, ref, qry, pairs = bench.make_workload(bench.workload_spec(bench.parse_args(["--workload", "chr2_25kb"])))
    tmp = tempfile.mkdtemp()
    t0 = time.perf_counter()
    regions = genome.regions()
    with open(os.path.join(tmp, "regions.bed"), "w") as fh:
        for r in regions:
            fh.write("%s\t%d\t%d\n" % (r.chromosome, r.start, r.end))
    for name, s in (("ref", ref), ("qry", qry)):
        with open(os.path.join(tmp, name + ".sparse"), "w") as fh:
            fh.write("".join("%d\t%d\t%r\n" % t for t in zip(s[0].tolist(), s[1].tolist(), s[2].tolist())))
    with open(os.path.join(tmp, "pairs.bedpe"), "w") as fh:
        for pid, a, b in pairs:
            fh.write("%s\t%d\t%d\t%s\t%d\t%d\t%s\t.\t%s\t%s\n" % (a.chromosome, a.start - 1, a.end, b.chromosome,
                                                                  b.start - 1, b.end, pid, a.strand, b.strand))
    print("wrote inputs in %.1f s: %d contacts per sample, %d pairs" % (time.perf_counter() - t0, len(ref[2]),
                                                                        len(pairs)))
    stamps = []

    class H(logging.Handler):
        def emit(self, record):
            stamps.append((time.perf_counter(), record.getMessage()[:60]))

    logging.getLogger('').addHandler(H())
    argv = ["chess", "sim", os.path.join(tmp, "ref.sparse"), os.path.join(tmp, "qry.sparse"),
            os.path.join(tmp, "pairs.bedpe"), os.path.join(tmp, "out.tsv"), "--reference-regions",
            os.path.join(tmp, "regions.bed"), "--query-regions", os.path.join(tmp, "regions.bed")]
    if args.background:
        argv += ["--background-query", "--limit-background"]
    if args.window:
        argv += ["-a", args.window]
    for rep in range(2):
        del stamps[:]
        t0 = time.perf_counter()
        Chess(argv)
        total = time.perf_counter() - t0
        print("run %d: total %.3f s" % (rep, total))
        prev = t0
        for t, msg in stamps:
            print("   +%.3f s  %s" % (t - prev, msg))
            prev = t
    print(open(os.path.join(tmp, "out.tsv")).read()[:300])


if __name__ == "__main__":
    main()
