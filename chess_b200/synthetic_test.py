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
import ctypes as C

import numpy as np
import torch

from . import _native
from ._native import lib, check
from .engine import compare_dense_pairs, get_context, visible_devices, PAIR_DTYPE, _ptr, _stream_ptr

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
    # symmetric matrices of one size (what the script generates) may take the upper-triangle kernels
    symmetric = all(m.shape == (n, n) and np.array_equal(m, m.T) for m in [reference] + queries)
    ssim, _, _ = compare_dense_pairs([reference] * len(queries), queries, [0] * len(queries), device=device,
                                     symmetric=symmetric, **_params(n, rel_window, data_range))
    truth, control = float(ssim[0]), ssim[1:]
    p = float(np.sum(control >= truth)) / max(len(pool), 1)
    allv = np.concatenate([control, [truth]])
    z = float((truth - allv.mean()) / allv.std())
    return truth, control, p, z


class OneVsMany(object):
    """One reference against its truth and a pool of decoys, resident on the device.

    `run_test` (run_synthetic_test.py:97-121) compares the same reference with the same pool once per relative
    window; here the matrices are uploaded ONCE and every window is one launch on the resident batch
    (equal sizes only; `truth_and_controls` handles the general case).
    """

    def __init__(self, reference, query, pool, device=None):
        self.device = visible_devices()[0] if device is None else int(device)
        self.n = int(reference.shape[0])
        mats = [query] + list(pool)
        if any(m.shape != (self.n, self.n) for m in [reference] + mats):
            raise ValueError("OneVsMany expects square matrices of one size")
        self.n_pool = len(pool)
        dev = torch.device("cuda", self.device)
        self.ctx = get_context(self.device)
        n2 = self.n * self.n
        with torch.cuda.device(dev):
            self.d_ref = torch.from_numpy(np.ascontiguousarray(reference, dtype=np.float64).ravel()).to(dev)
            # one packed host buffer, one copy (a thousand 80 KB copies cost more than the 80 MB they move)
            self.d_pool = torch.from_numpy(np.ascontiguousarray(np.stack(mats), dtype=np.float64).reshape(-1)).to(dev)
            # Hi-C matrices are symmetric; when every matrix of the batch is (checked where the matrices already
            # are, not assumed), the batch may take the upper-triangle kernels (CHESS_FLAG_SYMMETRIC)
            flags = torch.zeros(2, dtype=torch.int32, device=dev)
            for k, (buf, cnt) in enumerate(((self.d_pool, len(mats)), (self.d_ref, 1))):
                check(lib.chess_symmetric_check(self.ctx.handle, _ptr(buf), self.n, cnt,
                                                C.c_void_p(flags.data_ptr() + 4 * k), _stream_ptr(dev)),
                      self.ctx.handle, "chess_symmetric_check")
            self.symmetric = not bool(flags.cpu().any())
            tab = np.zeros(len(mats), dtype=PAIR_DTYPE)
            for k in range(len(mats)):
                tab[k] = (0, k * n2, self.n, self.n, self.n, self.n, 0, 0)
            self.d_tab = torch.from_numpy(tab.view(np.uint8)).to(dev)
            self.d_res = torch.empty(len(mats) * 5, dtype=torch.float32, device=dev)   # ssim f64, sn f64, status i32
        self.P = len(mats)

    def run(self, rel_window, data_range=2.0):
        """(truth_ssim, control_ssim, p, z) for one relative window (run_synthetic_test.py:105-115)."""
        dev = torch.device("cuda", self.device)
        P = self.P
        flags = (_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES) if self.symmetric else 0
        p = _native.make_params(flags=flags, **_params(self.n, rel_window, data_range))
        base = self.d_res.data_ptr()
        with torch.cuda.device(dev):
            check(lib.chess_compare_batch(self.ctx.handle, _ptr(self.d_ref), _ptr(self.d_pool), _ptr(self.d_tab), P,
                                          self.n, C.byref(p), C.c_void_p(base), C.c_void_p(base + 8 * P),
                                          C.c_void_p(base + 16 * P), _stream_ptr(dev)),
                  self.ctx.handle, "chess_compare_batch")
            ssim = self.d_res[:2 * P].view(torch.float64).cpu().numpy()
        truth, control = float(ssim[0]), ssim[1:]
        pval = float(np.sum(control >= truth)) / max(self.n_pool, 1)
        allv = np.concatenate([control, [truth]])
        z = float((truth - allv.mean()) / allv.std())
        return truth, control, pval, z
