"""Parity at BASELINE.json's full sizes, through size-independent properties (the oracle cannot run 10M x 1536 in seconds):
planted rows, scale invariance, idempotence, two independent kernels agreeing, exact fp64 re-scoring of the winners on the
host from regenerated rows."""
import numpy as np
import pytest
import torch

from robot_ebert_b200 import CatalogStore, synth

pytestmark = pytest.mark.gpu


def _need_memory(gib):
    free, _ = torch.cuda.mem_get_info()
    if free < gib * 2**30:
        pytest.skip(f"needs {gib} GiB of free HBM")


def _host_exact(seed, rows, d, dtype, qn):
    out = []
    for r in rows:
        row = synth.quantise(synth.catalog_rows_f32(seed, int(r), 1, d), dtype)[0]
        out.append(float(row @ qn / np.linalg.norm(row)))
    return np.array(out)


def test_ten_million_rows_single_query_properties():
    """BASELINE headline config: 10M x 1536 bf16, top-10 with a 133-row exclusion."""
    _need_memory(40)
    n, d, k = 10_000_000, 1536, 10
    store = CatalogStore.synthetic(0, n, d, "bf16")
    q = synth.query_f32(1, d)
    qn = q.astype(np.float64) / np.linalg.norm(q.astype(np.float64))
    excl = np.random.default_rng(1).choice(n, size=133, replace=False)
    rows, scores, info = store.recommend(query=q, exclude_rows=excl, k=k, return_info=True)
    assert info["proven_exact"] and len(rows) == k and not set(rows.tolist()) & set(excl.tolist())
    # (1) the winners' scores are the exact fp64 cosines of the regenerated rows; order is (score desc, row asc)
    np.testing.assert_allclose(scores, _host_exact(0, rows, d, "bf16", qn), rtol=1e-9)
    assert all(scores[i] > scores[i + 1] or (scores[i] == scores[i + 1] and rows[i] < rows[i + 1]) for i in range(k - 1))
    # (2) idempotence and scale invariance of the cosine
    r2, s2 = store.recommend(query=q, exclude_rows=excl, k=k)
    np.testing.assert_array_equal(rows, r2)
    np.testing.assert_array_equal(scores, s2)
    r3, s3 = store.recommend(query=(2.5 * q).astype(np.float32), exclude_rows=excl, k=k)
    np.testing.assert_array_equal(rows, r3)
    np.testing.assert_allclose(scores, s3, rtol=1e-6)
    # (3) an independent kernel (plain dense scores) + torch.topk finds the same rows
    q32 = torch.zeros((1, store.ld), dtype=torch.float32, device=store.device)
    q32[0, :d] = torch.from_numpy(qn.astype(np.float32)).to(store.device)
    dense = store.scores_dense(q32)[0]
    dense[torch.from_numpy(excl).to(store.device)] = -float("inf")
    cand = torch.topk(dense, k + 8).indices.cpu().numpy()
    exact = _host_exact(0, cand, d, "bf16", qn)
    order = np.lexsort((cand, -exact))[:k]
    np.testing.assert_array_equal(rows, cand[order])
    # (4) planted rows: copies of one catalog row must all surface, ties broken by ascending row, excluded ones skipped
    src = int(rows[3])
    planted = np.array([17, 4_000_000, 9_999_999, 123_456, 7_777_777], dtype=np.int64)
    idx = torch.from_numpy(planted).to(store.device)
    store.rows[idx] = store.rows[src].clone()
    store.inv_norm[idx] = store.inv_norm[src].clone()
    store.norm64[idx] = store.norm64[src].clone()
    probe = store.rows[src, :d].to(torch.float32).cpu().numpy()
    r4, s4 = store.recommend(query=probe, exclude_rows=np.array([123_456]), k=5)
    assert r4.tolist() == sorted([17, 4_000_000, 7_777_777, 9_999_999, src])
    assert np.all(np.abs(s4 - 1.0) < 1e-12)


def test_full_size_batch_agrees_with_single_query_kernel():
    """BASELINE config 3: 4096 users x 1M x 1536 bf16, top-100 — the tensor-core pipeline and the GEMV pipeline are
    independent implementations; on a sample of users they must return identical ids and scores."""
    _need_memory(12)
    n, d, b, k = 1_000_000, 1536, 4096, 100
    store = CatalogStore.synthetic(0, n, d, "bf16")
    q = synth.catalog_rows_f32(11, 0, b, d)
    rng = np.random.default_rng(2)
    ep = np.zeros(b + 1, dtype=np.int64)
    ec = []
    for u in range(b):
        c = np.unique(rng.integers(0, n, size=133))
        ec.append(c)
        ep[u + 1] = ep[u] + len(c)
    rows, scores, counts, info = store.recommend_batch(queries=q, excl_ptr=ep, excl_col=np.concatenate(ec), k=k, return_info=True)
    assert (info["status"] != 0).sum() <= 4 and (counts == k).all()
    for u in rng.choice(b, size=24, replace=False):
        r1, s1, i1 = store.recommend(query=q[u], exclude_rows=ec[u], k=k, return_info=True)
        assert i1["proven_exact"]
        np.testing.assert_array_equal(rows[u], r1, err_msg=f"user {u}")
        np.testing.assert_allclose(scores[u], s1, rtol=1e-12)
        assert not set(rows[u].tolist()) & set(ec[u].tolist())
