"""BASELINE configs[0]: one reference against its truth and a pool of decoys, dense matrices, SSIM only
(utils/run_synthetic_test.py:97-121), per relative window of the script.  The pool is uploaded once; every
window is one chess_compare_batch launch on device-resident data, timed with CUDA events; the CPU column is
the oracle's restatement of `compare_matrices` on one core over a few decoys.
Development / evidence tool:  python tools/bench_config0.py [--n 100] [--pool 1000]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--pool", type=int, default=1000)
    ap.add_argument("--cpu-sample", type=int, default=8)
    args = ap.parse_args()
    import torch
    from chess_b200 import synth, _native
    from chess_b200 import synthetic_test as S
    from chess_b200.engine import get_context, PAIR_DTYPE, _ptr, _stream_ptr, check
    from oracle import chess_oracle as O
    n, P = args.n, args.pool + 1
    rng = np.random.default_rng(0)
    base = synth.dense_decay(n)
    feat = synth.dense_features(n, rng)
    ref = synth.dense_oe(synth.dense_noisy(base + feat, rng))
    dev = torch.device("cuda", 0)
    ctx = get_context(0)
    d_ref = torch.from_numpy(ref.ravel().copy()).to(dev)
    d_pool = torch.empty(P * n * n, dtype=torch.float64, device=dev)
    keep = []
    for k in range(P):   # entry 0 is the truth (same features), the rest are decoys
        f = feat if k == 0 else synth.dense_features(n, rng)
        q = synth.dense_oe(synth.dense_noisy(base + f, rng))
        d_pool[k * n * n:(k + 1) * n * n] = torch.from_numpy(q.ravel()).to(dev)
        if k < args.cpu_sample:
            keep.append(q)
    tab = np.zeros(P, dtype=PAIR_DTYPE)
    for k in range(P):
        tab[k] = (0, k * n * n, n, n, n, n, 0, 0)
    d_tab = torch.from_numpy(tab.view(np.uint8)).to(dev)
    ssim = torch.empty(P, dtype=torch.float64, device=dev)
    sn = torch.empty(P, dtype=torch.float64, device=dev)
    status = torch.empty(P, dtype=torch.int32, device=dev)
    stream = _stream_ptr(dev)
    alg = P * (2.0 * n * n * 8 + 24)
    for rel in S.distinct_windows(n):
        win = S.window_size(n, rel)
        p = _native.make_params(absolute_window_size=win, min_bins=0, keep_unmappable_bins=True,
                                mappability_cutoff=0.0, compute_sn=False, data_range=2.0)
        times = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(_native.lib.chess_compare_batch(ctx.handle, _ptr(d_ref), _ptr(d_pool), _ptr(d_tab), P, n,
                                                  C.byref(p), _ptr(ssim), _ptr(sn), _ptr(status), stream),
                  ctx.handle, "chess_compare_batch")
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
        dt = min(times[1:])
        got = ssim[:len(keep)].cpu().numpy()
        t0 = time.perf_counter()
        want = np.array([O.synthetic_compare_matrices(ref, q, rel) for q in keep])
        cpu = (time.perf_counter() - t0) / len(keep)
        dev_rel = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)))
        print("n=%d win=%3d: %d pairs in %8.3f ms = %.3g pairs/s, algorithmic %.0f GB/s (%.3f of 6525.9); "
              "CPU oracle 1 core %.3g pairs/s; max rel dev %.1e" %
              (n, win, P, dt * 1e3, P / dt, alg / dt / 1e9, alg / dt / 6525.9e9, 1.0 / cpu, dev_rel), flush=True)


if __name__ == "__main__":
    main()
