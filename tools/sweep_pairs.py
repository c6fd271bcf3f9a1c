"""Time the compare launch as a function of the pair count (wave quantisation of the work-item kernel).

    python tools/sweep_pairs.py [--window r1|a7] [--counts 592,1184,...]

Development tool; prints one line per count: pairs, microseconds per launch (median of 30, L2 flushed), ns per pair.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--window", default="r1")
    ap.add_argument("--counts", default="148,296,592,1184,1776,2368,2402,2960,3552,4736")
    ap.add_argument("--no-flush", action="store_true", help="leave L2 warm between launches")
    ap.add_argument("--bin-size", type=int, default=25000)
    ap.add_argument("--window-bp", type=int, default=2000000)
    args = ap.parse_args()
    import torch
    import bench
    import chess_b200 as cb
    from chess_b200 import _native
    from chess_b200._native import lib, check
    from chess_b200.engine import get_context, region_spans, build_pair_table

    dev = torch.device("cuda", 0)
    genome, ref, qry, pairs = bench.make_workload(bench.workload_spec(bench.parse_args(
        ["--workload", "chr2_25kb", "--bin-size", str(args.bin_size), "--window-bp", str(args.window_bp)])))
    regions = genome.regions()
    trees = cb.region_interval_trees(regions)
    p = _native.make_params(compute_sn=True, kernel=0, flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES,
                            **bench.window_params(args.window))
    ctx = get_context(0)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rf, rn = region_spans(trees, [q[1] for q in pairs])
    qf, qn = region_spans(trees, [q[2] for q in pairs])
    flip = np.zeros(len(pairs), dtype=np.int32)
    max_n = int(max(rn.max(), qn.max()))
    re_ = cb.observed_expected_arrays(regions, ref, device=0)
    qe = cb.observed_expected_arrays(regions, qry, device=0)
    rw, rband, rsample = re_.sample_on(0, max_n)
    qw, qband, qsample = qe.sample_on(0, max_n)
    table = build_pair_table(rf, rn, rw, qf, qn, qw, flip)
    counts = [int(c) for c in args.counts.split(",")]
    reps = (max(counts) + len(table) - 1) // len(table)
    big = np.concatenate([table] * reps)
    d_table = torch.from_numpy(big.view(np.uint8)).to(dev)
    nmax = len(big)
    d_ssim = torch.empty(nmax, dtype=torch.float64, device=dev)
    d_sn = torch.empty(nmax, dtype=torch.float64, device=dev)
    d_status = torch.empty(nmax, dtype=torch.int32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for cnt in counts:
        ts = []
        for k in range(35):
            if not args.no_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib.chess_compare_band_batch(ctx.handle, C.byref(rsample), C.byref(qsample), ptr(d_table), cnt,
                                               max_n, C.byref(p), ptr(d_ssim), ptr(d_sn), ptr(d_status), sp),
                  ctx.handle, "chess_compare_band_batch")
            e1.record()
            torch.cuda.synchronize(dev)
            if k >= 5:
                ts.append(e0.elapsed_time(e1) * 1e3)
        med = float(np.median(ts))
        print("pairs %6d  n %d  %8.1f us  %7.2f ns/pair  min %.1f" % (cnt, max_n, med, med * 1e3 / cnt, min(ts)),
              flush=True)


if __name__ == "__main__":
    main()
