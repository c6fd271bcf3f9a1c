"""BASELINE configs[3] (DLBCL-style): chr2 @ 5 kb, windows of 1 Mb (200 bins) and 3 Mb (600 bins), every pair
against 1000 background regions on both strands (compute_sn off), reduced to z_bg / p_bg on the device.
Development / evidence tool:  python tools/bench_config4.py [--pairs 64]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    args = ap.parse_args()
    import torch
    import bench
    import chess_b200 as cb
    from chess_b200.engine import region_spans
    for window_bp in (1000000, 3000000):
        genome, ref, qry, pairs = bench.make_workload(0, "chr2_25kb", 1, 5000, window_bp)
        regions = genome.regions()
        trees = cb.region_interval_trees(regions)
        re_ = cb.observed_expected_arrays(regions, ref, device=0)
        qe = cb.observed_expected_arrays(regions, qry, device=0)
        eng = cb.CompareEngine(re_, trees, qe, trees, devices=[0])
        n = window_bp // 5000 + 1
        step = max(1, len(pairs) // args.pairs)
        pick = pairs[::step][:args.pairs]
        ids, s, sn, st = eng.compare(pick)
        rf, rn = region_spans(trees, [p[1] for p in pick])
        rng = np.random.default_rng(1)
        starts = rng.integers(0, len(regions) - n, 1000)
        bf = np.concatenate([starts, starts])
        bn = np.full(2000, n)
        bstrand = np.concatenate([np.zeros(1000), np.ones(1000)]).astype(np.int32)
        zeros = np.zeros(len(pick), dtype=np.int32)
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            z, p = eng.background(rf, rn, zeros, s, bf, bn, bstrand)
            dt = time.perf_counter() - t0
        ncmp = len(pick) * len(bf)
        alg = ncmp * (2.0 * n * n * 8 + 24)
        print("n=%d: %d pairs x %d background regions = %d comparisons in %.1f ms: %.3g comparisons/s, "
              "algorithmic %.0f GB/s (%.2f of 6525.9), finite z %d" % (n, len(pick), len(bf), ncmp, dt * 1e3,
                                                                       ncmp / dt, alg / dt / 1e9, alg / dt / 6525.9e9,
                                                                       int(np.isfinite(z).sum())), flush=True)
        re_.drop_device_copies()
        qe.drop_device_copies()


if __name__ == "__main__":
    main()
