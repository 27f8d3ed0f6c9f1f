"""Property-based tests (hypothesis).  CPU: the two oracle forms agree; GPU: CUDA path == oracle on random inputs.

The committed settings are derandomised (same examples every run, so the suite cannot flake at review time); during
development the GPU property was explored with several runs of 600 random examples, which is how the d = 1 / scaled
one-hot exact-tie cases and the k-boundary near-tie case were found."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import reference_scoring as ora
from robot_ebert_b200 import synth


@settings(max_examples=30, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(n=st.integers(3, 120), d=st.integers(1, 24), k=st.integers(1, 15), seed=st.integers(0, 10_000), nex=st.integers(0, 20),
       nliked=st.integers(1, 6))
def test_oracle_forms_agree_cpu(n, d, k, seed, nex, nliked):
    """DataFrame form (the reference's own expressions) == array form used at large N."""
    rng = np.random.default_rng(seed)
    m = synth.catalog_rows_f32(seed, 0, n, d, scale_rows=True).astype(np.float64)
    ids = synth.row_ids(n)
    emb = ora.catalog_frame(ids, m)
    excl = rng.choice(n, size=min(nex, n - 1), replace=False)
    liked = rng.choice(n, size=min(nliked, n), replace=False)
    q = rng.standard_normal(d)
    a = ora.single_query(emb, q, [ids[r] for r in excl], k)
    rows, scores = ora.query_rows(m, q, excl, k)
    assert [int(i) for i, _ in a] == rows.tolist()
    np.testing.assert_allclose([s for _, s in a], scores, atol=1e-14)
    import pandas as pd
    rated = np.union1d(liked, excl)
    ratings = pd.DataFrame({"tmdb_id": [ids[r] for r in rated], "rating": [5.0 if r in set(liked.tolist()) else 1.0 for r in rated]})
    b = ora.user_recs_ranked(emb, ratings, k)
    rows, scores = ora.recommend_rows(m, np.sort(liked), rated, k)
    assert [int(i) for i, _ in b] == rows.tolist()
    np.testing.assert_allclose([s for _, s in b], scores, atol=1e-14)


@pytest.mark.gpu
@settings(max_examples=200, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(n=st.integers(1, 6000), d=st.sampled_from([1, 3, 8, 32, 50, 64, 100, 256, 300, 768, 1536, 1600]),
       dtype=st.sampled_from(["fp32", "bf16"]), k=st.sampled_from([1, 5, 10, 17, 50, 100, 200]), seed=st.integers(0, 1000),
       excl_frac=st.sampled_from([0.0, 0.01, 0.3, 0.9]), dups=st.integers(0, 6), zero_rows=st.integers(0, 3), use_profile=st.booleans(),
       weighted=st.booleans())
def test_cuda_equals_oracle_on_random_inputs_gpu(n, d, dtype, k, seed, excl_frac, dups, zero_rows, use_profile, weighted):
    import torch
    from robot_ebert_b200 import CatalogStore
    rng = np.random.default_rng(seed)
    m = synth.catalog_rows_f32(seed, 0, n, d, scale_rows=True)
    for _ in range(min(dups, n - 1)):
        a, b = rng.integers(0, n, size=2)
        m[a] = m[b]                                   # exact duplicates -> exact score ties
    for _ in range(min(zero_rows, n)):
        m[rng.integers(0, n)] = 0.0                   # zero-norm rows score exactly 0
    store = CatalogStore.from_host(synth.row_ids(n), m, dtype)
    stored = store.rows[:n, :d].to(torch.float64).cpu().numpy()
    excl = rng.choice(n, size=int(excl_frac * n), replace=False) if excl_frac else None
    kx = k + 8                                        # a few extra oracle results to see ties that straddle the k boundary
    if use_profile:
        liked = np.sort(rng.choice(n, size=min(n, int(rng.integers(1, 40))), replace=False))
        w = rng.uniform(0.1, 2.0, size=len(liked)).astype(np.float32) if weighted else None
        rows, scores = store.recommend(liked_rows=liked, weights=w, exclude_rows=excl, k=k)
        want_rows, want_scores = ora.recommend_rows(stored, liked, excl, kx, weights=None if w is None else w.astype(np.float64))
    else:
        q = rng.standard_normal(d).astype(np.float32) if rng.random() > 0.1 else m[rng.integers(0, n)].copy()
        rows, scores = store.recommend(query=q, exclude_rows=excl, k=k)
        want_rows, want_scores = ora.query_rows(stored, q.astype(np.float64), excl, kx)
    assert len(rows) == min(k, len(want_rows))
    np.testing.assert_allclose(scores, want_scores[:k], rtol=1e-9, atol=1e-13)
    if not np.array_equal(rows, want_rows[:k]):
        # Ids may differ from the oracle ONLY where the oracle's own order rests on float64 rounding noise: two of its
        # scores within 1e-12 of each other at or around the differing positions (it orders mathematically tied rows by
        # the last bit of sums like x_hat . x_hat).  Exact ties are NOT excused: those must follow row order.
        diff = np.nonzero(rows != want_rows[:len(rows)])[0]
        ext = want_scores[:len(rows) + 8]
        for pos in diff:
            near = np.abs(ext - want_scores[pos])
            near = near[(near > 0) & (near < 1e-12)]
            assert near.size, (pos, rows, want_rows[:k], want_scores[:k + 2])
