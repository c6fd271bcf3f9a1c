"""Wall clock of the whole `chess sim` command on synthetic input FILES, with the time of every stage.

    python tools/cli_e2e.py [--workload chr2_25kb|hg38_10kb] [--window 7] [--background] [--runs 2] [--devices 0,1,..]

Writes the reference / query sparse matrices, the region BED and the pair BEDPE of the bench workload to a
temp dir (in parallel: the hg38 matrices are 47 M lines each), runs the CLI in-process `--runs` times and prints,
per run, the stage table that cli.run_sim records (cli.LAST_STAGES) and the total.  One JSON line at the end.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _write_part(args):
    path, a, b, w = args
    import pyarrow as pa
    import pyarrow.csv as pc
    pc.write_csv(pa.table({"a": a, "b": b, "w": w}), path,
                 write_options=pc.WriteOptions(include_header=False, delimiter="\t"))
    return path


def write_sparse_parallel(path, sample, pool, parts):
    a, b, w = sample
    cuts = np.linspace(0, len(w), parts + 1).astype(np.int64)
    names = pool.map(_write_part, [(path + ".part%d" % k, a[cuts[k]:cuts[k + 1]], b[cuts[k]:cuts[k + 1]],
                                    w[cuts[k]:cuts[k + 1]]) for k in range(parts)])
    with open(path, "wb") as out:
        for name in names:
            with open(name, "rb") as fh:
                while True:
                    buf = fh.read(64 << 20)
                    if not buf:
                        break
                    out.write(buf)
            os.remove(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="chr2_25kb", choices=["chr2_25kb", "hg38_10kb"])
    ap.add_argument("--background", action="store_true")
    ap.add_argument("--window", default=None)
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--devices", default=None, help="comma-separated GPU indices (default: all visible)")
    ap.add_argument("--keep", default=None, help="directory to write the input files to (kept)")
    args = ap.parse_args()
    if args.devices:
        os.environ["CHESS_B200_DEVICES"] = args.devices
    visible_at_start = os.environ.get("CUDA_VISIBLE_DEVICES")
    import bench
    t0 = time.perf_counter()
    genome, ref, qry, pairs = bench.make_workload(bench.workload_spec(bench.parse_args(["--workload", args.workload])))
    t_gen = time.perf_counter() - t0
    tmp = args.keep or tempfile.mkdtemp()
    os.makedirs(tmp, exist_ok=True)
    t0 = time.perf_counter()
    genome.write_bed(os.path.join(tmp, "regions.bed"))
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(min(cores, 16)) as pool:
        for name, s in (("ref", ref), ("qry", qry)):
            write_sparse_parallel(os.path.join(tmp, name + ".sparse"), s, pool, min(cores, 16))
    with open(os.path.join(tmp, "pairs.bedpe"), "w") as fh:
        fh.write("".join("%s\t%d\t%d\t%s\t%d\t%d\t%s\t.\t%s\t%s\n" % (a.chromosome, a.start - 1, a.end, b.chromosome,
                                                                      b.start - 1, b.end, pid, a.strand, b.strand)
                         for pid, a, b in pairs))
    sizes = {f: os.path.getsize(os.path.join(tmp, f)) for f in ("ref.sparse", "qry.sparse", "regions.bed", "pairs.bedpe")}
    print("generated in %.1f s, wrote inputs in %.1f s: %d contacts per sample, %d pairs, files %s" % (
        t_gen, time.perf_counter() - t0, len(ref[2]), len(pairs), sizes), flush=True)
    n_pairs, n_contacts = len(pairs), int(len(ref[2]))
    del ref, qry, pairs

    from chess_b200 import cli
    cli.LIMIT_GPUS = True          # as bin/chess does: the run may hide GPUs it does not need from the driver
    argv = ["chess", "sim", os.path.join(tmp, "ref.sparse"), os.path.join(tmp, "qry.sparse"),
            os.path.join(tmp, "pairs.bedpe"), os.path.join(tmp, "out.tsv"), "--reference-regions",
            os.path.join(tmp, "regions.bed"), "--query-regions", os.path.join(tmp, "regions.bed")]
    if args.background:
        argv += ["--background-query", "--limit-background"]
    if args.window:
        argv += ["-a", args.window]
    runs = []
    for rep in range(args.runs):
        t0 = time.perf_counter()
        cli.Chess(argv)
        total = time.perf_counter() - t0
        stages = dict(cli.LAST_STAGES)
        runs.append({"total_s": total, "stages": stages})
        print("run %d: total %.3f s" % (rep, total))
        for name, secs in stages.items():
            print("   %-34s %8.3f s" % (name, secs))
        print("   %-34s %8.3f s" % ("(unaccounted)", total - sum(stages.values())), flush=True)
    print(open(os.path.join(tmp, "out.tsv")).read()[:240])
    import torch
    print(json.dumps({"workload": args.workload, "pairs": n_pairs, "contacts_per_sample": n_contacts,
                      "devices": os.environ.get("CHESS_B200_DEVICES") or "all (%d)" % torch.cuda.device_count(),
                      "cuda_visible_devices": {"at_start": visible_at_start, "after": os.environ.get("CUDA_VISIBLE_DEVICES")},
                      "argv": argv[2:], "input_bytes": sizes, "runs": runs}))


if __name__ == "__main__":
    main()
