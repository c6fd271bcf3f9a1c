"""BASELINE configs[0]: the one-vs-many comparison of dense synthetic matrices
(utils/make_synthetic_references.py + utils/run_synthetic_test.py).

CPU part: the oracle's restatement of `compare_matrices` against the brute-force
valid-window SSIM and the script's window rule.  GPU part: one batched launch
per (matrix type, window) through the C ABI against the oracle, <= 1e-6 relative
on every SSIM, identical p, z within 1e-6.  data_range = 2.0 (what the script's
legacy `compare_ssim` call assumes for float images) and the `chess sim` rule
(max - min of the pair) are both covered.
"""
import numpy as np
import pytest

from oracle import chess_oracle as O
from chess_b200 import synth
from tests.conftest import close

RTOL = 1e-6


def _types(m):
    oe = synth.dense_oe(m)
    return [m, oe, np.log2(oe)]          # run_synthetic_test.py:34, 87-95


def _case(n, n_pool, seed):
    rng = np.random.default_rng(seed)
    base = synth.dense_decay(n)
    feat = synth.dense_features(n, rng)
    ref = _types(synth.dense_noisy(base + feat, rng))
    truth = _types(synth.dense_noisy(base + feat, rng))
    pool = [_types(synth.dense_noisy(base + synth.dense_features(n, rng), rng)) for _ in range(n_pool)]
    return ref, truth, pool


def test_window_rule_of_the_script():
    assert [O.synthetic_window_size(1000, r) for r in (0.01, 0.05, 0.1, 0.2, 0.4, 0.5, 0.6, 0.8, 1)] == \
        [9, 49, 99, 199, 399, 499, 599, 799, 999]
    assert [O.synthetic_window_size(100, r) for r in (0.01, 0.05, 0.1, 0.2, 0.5, 1)] == [7, 7, 9, 19, 49, 99]
    from chess_b200.synthetic_test import window_size, distinct_windows
    for n in (40, 100, 257, 1000):
        for r in (0.01, 0.05, 0.1, 0.2, 0.4, 0.5, 0.6, 0.8, 1):
            assert window_size(n, r) == O.synthetic_window_size(n, r)
    assert distinct_windows(100) == [0.01, 0.1, 0.2, 0.4, 0.5, 0.6, 0.8, 1]   # 7 bins run once (:98-103)
    assert distinct_windows(1000) == [0.01, 0.05, 0.1, 0.2, 0.4, 0.5, 0.6, 0.8, 1]


def test_oracle_compare_matrices_equals_bruteforce():
    ref, truth, pool = _case(60, 2, 3)
    for t in range(3):
        for rel in (0.05, 0.2, 0.5, 1):
            win = O.synthetic_window_size(60, rel)
            a = O.synthetic_compare_matrices(ref[t], truth[t], rel)
            b = O.structural_similarity_bruteforce(ref[t], truth[t], win, 2.0)
            assert close(a, b, 1e-11), (t, rel, a, b)
    # unequal sizes: the smaller matrix is resized up, whichever side it is on
    small = ref[1][:45, :45]
    assert O.synthetic_compare_matrices(small, truth[1], 0.2) == \
        O.synthetic_compare_matrices(O.resize_order0(small, (60, 60)), truth[1], 0.2)
    tr, co, p, z = O.synthetic_truth_and_controls(ref[1], truth[1], [q[1] for q in pool], 0.2)
    assert co.shape == (2,) and 0.0 <= p <= 1.0 and np.isfinite(z)


@pytest.mark.gpu
@pytest.mark.parametrize("n,n_pool,rels", [
    (100, 24, (0.01, 0.1, 0.2, 0.4, 0.5, 0.6, 0.8, 1)),
    (257, 6, (0.05, 0.5, 1)),
    (1000, 2, (0.01, 0.1, 0.5, 1)),
])
def test_one_vs_many_against_oracle(n, n_pool, rels):
    from chess_b200 import synthetic_test as S
    ref, truth, pool = _case(n, n_pool, n)
    for t in range(3):
        queries = [q[t] for q in pool]
        for rel in rels:
            tr, co, p, z = S.truth_and_controls(ref[t], truth[t], queries, rel)
            otr, oco, op, oz = O.synthetic_truth_and_controls(ref[t], truth[t], queries, rel)
            assert close(tr, otr, RTOL), (n, t, rel, tr, otr)
            for a, b in zip(co, oco):
                assert close(a, b, RTOL), (n, t, rel, a, b)
            assert p == op and close(z, oz, 1e-5), (n, t, rel, p, op, z, oz)
    # the resident form: one upload, one launch per window, identical numbers
    batch = S.OneVsMany(ref[2], truth[2], [q[2] for q in pool])
    for rel in rels[:2]:
        tr, co, p, z = batch.run(rel)
        tr2, co2, p2, z2 = S.truth_and_controls(ref[2], truth[2], [q[2] for q in pool], rel)
        assert tr == tr2 and np.array_equal(co, co2) and p == p2 and z == z2
    # the chess sim rule for the data range, and a single pair of unequal sizes
    tr, co, _, _ = S.truth_and_controls(ref[1], truth[1], [q[1] for q in pool[:2]], 0.2, data_range=None)
    lo = min(ref[1].min(), truth[1].min())
    hi = max(ref[1].max(), truth[1].max())
    win = O.synthetic_window_size(n, 0.2)
    assert close(tr, O.structural_similarity(ref[1], truth[1], win, hi - lo), RTOL)
    if n <= 257:
        small = np.ascontiguousarray(ref[1][:n - 17, :n - 17])
        assert close(S.compare_matrices(small, truth[1], 0.2), O.synthetic_compare_matrices(small, truth[1], 0.2), RTOL)


@pytest.mark.gpu
def test_fixed_data_range_is_rejected_when_negative():
    import chess_b200
    from chess_b200 import synthetic_test as S
    a = synth.dense_decay(30)
    with pytest.raises(Exception):
        S.compare_matrices(a, a, 0.5, data_range=-1.0)
    assert S.compare_matrices(a, a, 0.5) == 1.0


@pytest.mark.gpu
def test_large_batch_uses_the_same_arithmetic():
    """More than a wave of pairs switches the one-warp kernel to its 12-warps-per-SM build: same code, so the
    numbers of a pair do not depend on the batch it is in; a few are checked against the oracle."""
    from chess_b200 import synthetic_test as S
    n = 40
    rng = np.random.default_rng(5)
    base = synth.dense_decay(n)
    ref = synth.dense_oe(synth.dense_noisy(base + synth.dense_features(n, rng), rng))
    pool = [synth.dense_oe(synth.dense_noisy(base + synth.dense_features(n, rng), rng)) for _ in range(1500)]
    for rel in (0.2, 0.5):
        _, big, _, _ = S.truth_and_controls(ref, pool[0], pool, rel)
        _, small, _, _ = S.truth_and_controls(ref, pool[0], pool[:20], rel)
        assert np.array_equal(big[:20], small)
        for k in (0, 7, 1499):
            assert close(big[k], O.synthetic_compare_matrices(ref, pool[k], rel), RTOL)
