"""The threshold hints of gemv_topk_kernel are LOWER BOUNDS of the kc-th best allowed score; rows strictly below a hint are
dropped before the insert path, so an invalid hint would silently lose results.  This restates the two bounds in numpy over
a simulated CTA (8 warps with private top-kc lists, rows dealt to warps in the kernel's order) and checks the invariant
at every step, including with excluded rows and heavy ties."""
import numpy as np
import pytest

WARPS = 8


def _simulate(scores, allowed, kc):
    n = len(scores)
    lists = [[] for _ in range(WARPS)]                       # per-warp kept scores, descending
    seen_allowed = []
    for start in range(0, n, WARPS):                         # one step: every warp takes one row (the kernel's row order)
        for w in range(WARPS):
            r = start + w
            if r >= n:
                break
            # the bounds the warps can publish BEFORE this row is looked at
            full = [l[kc - 1] for l in lists if len(l) >= kc]
            warp_bound = max(full) if full else -np.inf                                  # a full list's threshold
            q = kc // 8
            cta_bound = min(l[q - 1] for l in lists) if all(len(l) >= q for l in lists) else -np.inf   # min of (kc/8)-th bests
            hint = max(warp_bound, cta_bound)
            if len(seen_allowed) >= kc:
                truth = np.sort(seen_allowed)[::-1][kc - 1]                              # kc-th best allowed score seen so far
                assert hint <= truth, (r, hint, truth)
            else:
                assert hint == -np.inf or hint <= min(seen_allowed)
            if not allowed[r]:
                continue
            seen_allowed.append(scores[r])
            l = lists[w]
            thr = l[kc - 1] if len(l) >= kc else -np.inf
            if scores[r] > thr and scores[r] >= hint:        # the kernel's test: strictly above the own threshold, >= the hint
                l.append(scores[r])
                l.sort(reverse=True)
                del l[kc:]
    return lists, seen_allowed


@pytest.mark.parametrize("kc", [32, 64, 128])
@pytest.mark.parametrize("kind", ["gauss", "ties", "ascending", "descending"])
def test_hints_never_exceed_the_kth_best_and_no_result_is_lost(kc, kind):
    rng = np.random.default_rng(kc + len(kind))
    n = 3000
    s = rng.standard_normal(n)
    if kind == "ties":
        s = np.round(s, 1)
    if kind == "ascending":
        s = np.sort(s)
    if kind == "descending":
        s = np.sort(s)[::-1]
    allowed = rng.random(n) > 0.1
    lists, seen = _simulate(s.astype(np.float32).astype(np.float64), allowed, kc)
    # nothing that belongs to the top-kc (by score; ties at the boundary aside) was dropped by a hint
    kept = np.sort(np.concatenate([np.array(l) for l in lists]))[::-1][:kc]
    want = np.sort(np.array(seen))[::-1][:kc]
    np.testing.assert_array_equal(kept, want)
