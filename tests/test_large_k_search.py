"""Host logic of the k > 240 route (robot_ebert_b200.catalog.large_k_search) against brute force, with a simulated fast
pass whose scores differ from the exact ones by up to eps — including mass ties at the cut, fewer allowed rows than k, and
a sweep buffer that overflows."""
import numpy as np
import pytest

from robot_ebert_b200.catalog import large_k_search


def _setup(n, eps, seed, ties=False):
    rng = np.random.default_rng(seed)
    exact = rng.uniform(-1.0, 1.0, size=n)
    if ties:
        exact = np.round(exact, 2)                                  # thousands of exact ties, many right at the cut
    fast = (exact + rng.uniform(-eps, eps, size=n)).astype(np.float32).astype(np.float64)
    rows = np.arange(n, dtype=np.int64)
    calls = {"count": 0, "sweep": 0}

    def count(thr):
        calls["count"] += 1
        return int(np.count_nonzero(fast >= np.float32(thr)))

    def sweep(thr):
        calls["sweep"] += 1
        sel = fast >= np.float32(thr)
        return rows[sel], exact[sel]

    return exact, rows, count, sweep, calls


@pytest.mark.parametrize("ties", [False, True])
@pytest.mark.parametrize("n,k", [(5000, 241), (5000, 1000), (300, 500), (2000, 2000), (40_000, 4096)])
def test_matches_brute_force(n, k, ties):
    eps = 3e-4
    exact, rows, count, sweep, calls = _setup(n, eps + 6e-8, seed=n + k, ties=ties)   # + fp32 rounding of the simulated fast score
    got_rows, got_scores = large_k_search(count, sweep, k, 2 * eps)
    order = np.lexsort((rows, -exact))[:k]
    np.testing.assert_array_equal(got_rows, rows[order])
    np.testing.assert_array_equal(got_scores, exact[order])
    assert calls["count"] <= 42 and calls["sweep"] <= 4


def test_sweep_overflow_is_an_error_not_a_wrong_answer():
    _, _, count, _, _ = _setup(1000, 1e-4, seed=1)
    with pytest.raises(RuntimeError, match="overflow"):
        large_k_search(count, lambda thr: None, 300, 1e-4)
