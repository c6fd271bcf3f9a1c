"""One-vs-many comparison of dense synthetic matrices on the device.

Host-side mirror of the comparison loop of the reference's synthetic benchmark
(BASELINE configs[0]): `utils/run_synthetic_test.py:97-121` compares every
reference matrix with its true counterpart and with a pool of 1000 decoys, per
matrix type {balanced, O/E, log2(O/E)} and per relative window, one
`compare_ssim` call per pair (`compare_matrices`, :279-308).  Here the truth
and the whole pool are ONE batched launch against a single device copy of the
reference (`engine.compare_dense_pairs`): dense, pre-packed, no masks, SSIM
only.

`data_range`: the script calls the legacy `skimage.measure.compare_ssim`
without `data_range`; for float images that function assumes the range of the
dtype, -1 .. 1, i.e. 2.0.  That is the default here; `data_range=None` takes
max - min of the pair as `chess sim` does (chess/sim.py:368-374).
"""
import numpy as np

from .engine import compare_dense_pairs

REL_WINDOWS = (0.01, 0.05, 0.1, 0.2, 0.4, 0.5, 0.6, 0.8, 1)   # run_synthetic_test.py:97


def window_size(n_bins, rel_window):
    """Odd window of `compare_matrices` (run_synthetic_test.py:302-305)."""
    win = int(n_bins * rel_window)
    if win % 2 == 0:
        win -= 1
    return max(win, 7)


def distinct_windows(n_bins, rel_windows=REL_WINDOWS):
    """The relative windows `run_test` actually runs: 7-bin windows once (:98-103)."""
    out, ran7 = [], False
    for rel in rel_windows:
        if max(int(n_bins * rel), 7) == 7:
            if ran7:
                continue
            ran7 = True
        out.append(rel)
    return out


def _params(n, rel_window, data_range):
    return dict(absolute_window_size=window_size(n, rel_window), relative_window_size=1.0, min_bins=0,
                keep_unmappable_bins=True, mappability_cutoff=0.0, compute_sn=False, data_range=data_range)


def compare_matrices(a, b, rel_window, data_range=2.0, device=None):
    """SSIM of two dense matrices; the smaller one is resized up first
    (scaling mode 1, run_synthetic_test.py:289-293: order 0, as sim.py:323-329)."""
    n = max(a.shape[0], b.shape[0])
    ssim, _, _ = compare_dense_pairs([a], [b], [0], device=device, **_params(n, rel_window, data_range))
    return float(ssim[0])


def truth_and_controls(reference, query, pool, rel_window, data_range=2.0, device=None):
    """(truth_ssim, control_ssim[len(pool)], p, z) of run_synthetic_test.py:105-115.

    p = #{control >= truth} / len(pool); z = last entry of the population
    z-score of controls + [truth].
    """
    n = max([reference.shape[0], query.shape[0]] + [t.shape[0] for t in pool])
    queries = [query] + list(pool)
    ssim, _, _ = compare_dense_pairs([reference] * len(queries), queries, [0] * len(queries), device=device,
                                     **_params(n, rel_window, data_range))
    truth, control = float(ssim[0]), ssim[1:]
    p = float(np.sum(control >= truth)) / max(len(pool), 1)
    allv = np.concatenate([control, [truth]])
    z = float((truth - allv.mean()) / allv.std())
    return truth, control, p, z
