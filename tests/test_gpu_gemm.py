"""GPU parity of the batched tcgen05 path: the GEMM building block against the plain dense-score kernel, and
recommend_batch against the oracle (ids bit-exact, scores 1e-9)."""
import numpy as np
import pytest
import torch

from oracle import reference_scoring as ora
from robot_ebert_b200 import CatalogStore, synth

pytestmark = pytest.mark.gpu


def _stored_f64(store):
    return store.rows[:store.n, :store.d].to(torch.float64).cpu().numpy()


def _queries(b, d, seed=11):
    return synth.catalog_rows_f32(seed, 0, b, d)


@pytest.mark.parametrize("n,d,b", [(4096, 1536, 200), (5000, 1536, 128), (2048, 256, 1), (70_000, 64, 300)])
def test_gemm_scores_match_dense_kernel(n, d, b):
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    qn32, qn64, qbf = store.prepare_queries(_queries(b, d))
    nrows = (n + 255) // 256 * 256
    got = store.gemm_scores(qbf, 0, nrows)[:, :n]
    want = store.scores_dense(qbf.to(torch.float32))          # same bf16-rounded queries, fp32 CUDA-core dot
    torch.cuda.synchronize()
    err = (got - want).abs().max().item()
    assert err < 2e-6, err
    # and against float64 on the host for a few rows
    m = _stored_f64(store)
    qh = qbf.to(torch.float64).cpu().numpy()[:, :d]
    ref = (qh[:4] @ m[:512].T) / np.linalg.norm(m[:512], axis=1)
    np.testing.assert_allclose(got[:4, :512].cpu().numpy(), ref, atol=3e-6)


def test_gemm_scores_row_window():
    store = CatalogStore.synthetic(0, 8192, 128, "bf16")
    _, _, qbf = store.prepare_queries(_queries(130, 128))
    full = store.gemm_scores(qbf, 0, 8192)
    win = store.gemm_scores(qbf, 2048, 1024)
    torch.cuda.synchronize()
    assert torch.equal(full[:, 2048:3072], win)


@pytest.mark.parametrize("n,d,b,k", [(40_000, 1536, 300, 10), (100_000, 256, 1000, 10), (150_000, 128, 257, 100)])
def test_recommend_batch_queries_vs_oracle(n, d, b, k):
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    m = _stored_f64(store)
    q = _queries(b, d)
    rng = np.random.default_rng(5)
    ptr = [0]
    cols = []
    for u in range(b):
        c = np.sort(rng.choice(n, size=int(rng.integers(0, 200)), replace=False))
        cols.append(c)
        ptr.append(ptr[-1] + len(c))
    rows, scores, counts, info = store.recommend_batch(queries=q, excl_ptr=np.array(ptr), excl_col=np.concatenate(cols), k=k,
                                                       return_info=True)
    assert (info["status"] != 0).mean() < 0.05, info["status"]          # the fast path must carry almost everything
    unit = m / np.linalg.norm(m, axis=1, keepdims=True)
    qn = q.astype(np.float64)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    for u in range(b):
        want_rows, want_scores = ora.topk_rows(unit @ qn[u], k, cols[u])
        assert counts[u] == len(want_rows)
        np.testing.assert_array_equal(rows[u, :counts[u]], want_rows, err_msg=f"query {u} status {info['status'][u]}")
        np.testing.assert_allclose(scores[u, :counts[u]], want_scores, rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize("sel", ["wide", "narrow"])
def test_recommend_batch_with_genre_year_predicate(sel):
    """BASELINE config 5 at test size: CSR profiles + per-user exclusions + a genre/year predicate, top-50."""
    from robot_ebert_b200 import RowFilter
    n, d, b, k = 120_000, 256, 96, 50
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    g, y = synth.movie_metadata(3, 0, n)
    store.set_metadata(g, y)
    want_g, y0, y1 = (0b1011, 1960, 2000) if sel == "wide" else (0b1, 1990, 1999)
    keep = ((g & want_g) != 0) & (y >= y0) & (y <= y1)
    m = _stored_f64(store)
    users = synth.user_ratings(2, n, b)
    lp, lc, ep, ec = [0], [], [0], []
    for rated, rts in users:
        liked = rated[rts >= 3.5] if (rts >= 3.5).any() else rated[:1]
        lc.append(liked); lp.append(lp[-1] + len(liked)); ec.append(rated); ep.append(ep[-1] + len(rated))
    rows, scores, counts, info = store.recommend_batch(liked_ptr=np.array(lp), liked_col=np.concatenate(lc), excl_ptr=np.array(ep),
                                                       excl_col=np.concatenate(ec), k=k, return_info=True,
                                                       row_filter=RowFilter(genre_any=want_g, year_lo=y0, year_hi=y1))
    assert (info["status"] != 0).mean() < 0.2, (sel, info["status"])
    for u in range(b):
        want_rows, want_scores = ora.recommend_rows(m, lc[u], ec[u], k, keep_mask=keep)
        np.testing.assert_array_equal(rows[u, :counts[u]], want_rows, err_msg=f"user {u}")
        np.testing.assert_allclose(scores[u, :counts[u]], want_scores, rtol=1e-9, atol=1e-15)
        assert keep[rows[u, :counts[u]]].all()


def test_recommend_batch_large_sample_path():
    """n large enough that the threshold sample (> 48K rows) is radix-selected in place in global memory."""
    n, d, b, k = 3_200_000, 64, 48, 10
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    plan = store.gemm_plan(b, k)
    assert plan.sample_rows > 48 * 1024
    q = _queries(b, d)
    rows, scores, counts, info = store.recommend_batch(queries=q, k=k, return_info=True)
    assert (info["status"] != 0).sum() <= 2
    qn = q.astype(np.float64)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    best = np.full((b, k), -np.inf)
    best_rows = np.zeros((b, k), dtype=np.int64)
    for s0 in range(0, n, 400_000):                      # streamed oracle: keep the running top-k per query
        m = store.rows[s0:s0 + 400_000, :d].to(torch.float64).cpu().numpy()
        sc = (m / np.linalg.norm(m, axis=1, keepdims=True)) @ qn.T          # [chunk, b]
        for u in range(b):
            r, v = ora.topk_rows(sc[:, u], k)
            allv = np.concatenate([best[u], v]); allr = np.concatenate([best_rows[u], r + s0])
            order = np.lexsort((allr, -allv))[:k]
            best[u], best_rows[u] = allv[order], allr[order]
    np.testing.assert_array_equal(rows, best_rows)
    np.testing.assert_allclose(scores, best, rtol=1e-9)


def test_recommend_batch_profiles_vs_oracle():
    n, d, b, k = 60_000, 1536, 64, 10
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    m = _stored_f64(store)
    users = synth.user_ratings(2, n, b)
    lp, lc, ep, ec = [0], [], [0], []
    for rated, rts in users:
        liked = rated[rts >= 3.5]
        if len(liked) == 0:
            liked = rated[:1]
        lc.append(liked)
        lp.append(lp[-1] + len(liked))
        ec.append(rated)
        ep.append(ep[-1] + len(rated))
    rows, scores, counts = store.recommend_batch(liked_ptr=np.array(lp), liked_col=np.concatenate(lc), excl_ptr=np.array(ep),
                                                 excl_col=np.concatenate(ec), k=k)
    for u in range(b):
        want_rows, want_scores = ora.recommend_rows(m, lc[u], ec[u], k)
        np.testing.assert_array_equal(rows[u, :counts[u]], want_rows)
        np.testing.assert_allclose(scores[u, :counts[u]], want_scores, rtol=1e-9, atol=1e-15)


def test_small_or_fp32_catalogs_are_served_by_the_single_query_kernel():
    """Below the batched path's minimum size (or for fp32 storage) recommend_batch loops the fused GEMV: same answers."""
    for n, d, dtype in [(2000, 64, "bf16"), (30_000, 64, "fp32")]:
        store = CatalogStore.synthetic(0, n, d, dtype)
        m = _stored_f64(store)
        q = _queries(5, d)
        rows, scores, counts, info = store.recommend_batch(queries=q, k=10, return_info=True)
        assert info["plan"] is None
        for u in range(5):
            want_rows, want_scores = ora.query_rows(m, q[u].astype(np.float64), None, 10)
            np.testing.assert_array_equal(rows[u], want_rows)
            np.testing.assert_allclose(scores[u], want_scores, rtol=1e-9)
    store = CatalogStore.synthetic(0, 40_000, 64, "fp32")
    _, _, qbf = store.prepare_queries(_queries(4, 64))
    with pytest.raises(Exception):
        store.gemm_scores(qbf, 0, 256)          # the tensor-core building block itself is bf16-only


# ---------------------------------------------------------------------------------------------------------------------
# int8 operands (tcgen05 kind::i8 over the prefilter shadow): exact integer arithmetic in the GEMM, same results end to end
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,b", [(4096, 1536, 200), (5000, 256, 128), (2048, 128, 1), (70_000, 512, 300)])
def test_int8_gemm_scores_are_the_integer_dot_products(n, d, b):
    import ctypes as C
    from robot_ebert_b200 import _native as nat
    lib = nat.load()
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    store.enable_prefilter()
    assert store.batch_shadow_ok
    qn32, qn64, qbf = store.prepare_queries(_queries(b, d))
    q8, qscale, qeps = store.quantize_queries(qn32)
    nrows = (n + 255) // 256 * 256
    out = torch.empty((b, nrows), dtype=torch.float32, device=store.device)
    nat.check(lib.rebert_gemm_scores_i8(C.byref(store._c8), q8.data_ptr(), qscale.data_ptr(), b, 0, nrows, out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    rows8 = store._q8_rows[:n].cpu().numpy().astype(np.int64)
    factor = store._q8_factor[:n].cpu().numpy().astype(np.float64)
    acc = q8.cpu().numpy().astype(np.int64) @ rows8.T                               # exact integer dot products
    want = acc.astype(np.float64) * factor[None, :] * qscale.cpu().numpy().astype(np.float64)[:, None]
    got = out[:, :n].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, want, rtol=3e-7, atol=1e-9)                     # two fp32 multiplies on an exact integer
    # the quantised scores approximate the true cosines within the per-query bound the proof uses (6 sigma: generous)
    m = _stored_f64(store)
    true = (qn64[:8, :d].cpu().numpy() @ m[:2048].T) / np.linalg.norm(m[:2048], axis=1)
    assert np.abs(got[:8, :2048] - true).max() < float(qeps[:8].max().item())


@pytest.mark.parametrize("n,d,b,k", [(200_000, 1536, 300, 10), (150_000, 256, 257, 100), (300_000, 128, 64, 240)])
def test_int8_batched_path_equals_bf16_batched_path_and_oracle(n, d, b, k):
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    store.enable_prefilter()
    m = _stored_f64(store)
    q = _queries(b, d)
    rng = np.random.default_rng(5)
    ptr, cols = [0], []
    for u in range(b):
        c = np.sort(rng.choice(n, size=int(rng.integers(0, 200)), replace=False))
        cols.append(c)
        ptr.append(ptr[-1] + len(c))
    ptr, col = np.array(ptr), np.concatenate(cols)
    r8, s8, c8, i8 = store.recommend_batch(queries=q, excl_ptr=ptr, excl_col=col, k=k, return_info=True, prefilter=True)
    rb, sb, cb, ib = store.recommend_batch(queries=q, excl_ptr=ptr, excl_col=col, k=k, return_info=True, prefilter=False)
    assert i8["int8_operands"] and not ib["int8_operands"]
    assert (i8["status"] != 0).mean() < 0.10, (i8["status"] != 0).mean()            # the int8 filter must carry almost everything
    np.testing.assert_array_equal(r8, rb)
    np.testing.assert_array_equal(s8, sb)                                           # same exact pass on the same rows: same bits
    np.testing.assert_array_equal(c8, cb)
    unit = m / np.linalg.norm(m, axis=1, keepdims=True)
    qn = q.astype(np.float64)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    for u in range(0, b, 7):
        want_rows, want_scores = ora.topk_rows(unit @ qn[u], k, cols[u])
        np.testing.assert_array_equal(r8[u, :c8[u]], want_rows, err_msg=f"query {u} status {i8['status'][u]}")
        np.testing.assert_allclose(s8[u, :c8[u]], want_scores, rtol=1e-9, atol=1e-15)


def test_int8_batched_profiles_with_predicate_vs_oracle():
    """CSR profiles (shorter than unit vectors: the error bound scales with their length) + genre/year predicate."""
    from robot_ebert_b200 import RowFilter
    n, d, b, k = 160_000, 256, 96, 50
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    store.enable_prefilter()
    g, y = synth.movie_metadata(3, 0, n)
    store.set_metadata(g, y)
    m = _stored_f64(store)
    lp, lc, ep, ec = [0], [], [0], []
    users = synth.user_ratings(2, n, b)
    for rated, rts in users:
        liked = rated[rts >= 3.5] if (rts >= 3.5).any() else rated[:1]
        lc.append(liked); lp.append(lp[-1] + len(liked)); ec.append(rated); ep.append(ep[-1] + len(rated))
    rf = RowFilter(genre_any=0b1011, year_lo=1960, year_hi=2000)
    keep = ((g & 0b1011) != 0) & (y >= 1960) & (y <= 2000)
    rows, scores, counts, info = store.recommend_batch(liked_ptr=np.array(lp), liked_col=np.concatenate(lc), excl_ptr=np.array(ep),
                                                       excl_col=np.concatenate(ec), k=k, row_filter=rf, return_info=True, prefilter=True)
    assert info["int8_operands"]
    for u in range(0, b, 5):
        want_rows, want_scores = ora.recommend_rows(m, lc[u], ec[u], k, keep_mask=keep)
        np.testing.assert_array_equal(rows[u, :counts[u]], want_rows, err_msg=f"user {u} status {info['status'][u]}")
        np.testing.assert_allclose(scores[u, :counts[u]], want_scores, rtol=1e-9, atol=1e-15)


def test_int8_filter_with_saturated_accumulators():
    """Rows and queries of +-1 entries quantise to +-127 everywhere, so the int32 accumulators reach 1536 * 127^2 = 24.8M (beyond
    fp32's 24-bit integers: the int->float conversion of the epilogue rounds there) and the shadow is exact — the proof's error
    bound collapses to its floor.  The results must still equal the bf16 path's and the oracle's."""
    n, d, b, k = 70_000, 1536, 140, 10
    rng = np.random.default_rng(11)
    base = rng.integers(0, 2, size=(b, d)).astype(np.float32) * 2 - 1
    m = rng.integers(0, 2, size=(n, d)).astype(np.float32) * 2 - 1
    for u in range(b):                                         # plant near-copies of every query: high positive accumulators
        for j in range(12):
            r = (u * 12 + j) * 37 % n
            flips = rng.choice(d, size=8 * (j + 1), replace=False)
            m[r] = base[u]
            m[r, flips] *= -1
    m[5] = -base[0]                                            # and a strongly negative one
    store = CatalogStore.from_host(None, m, dtype="bf16")
    store.enable_prefilter()
    assert store.batch_shadow_ok
    r8, s8, c8, i8 = store.recommend_batch(queries=base, k=k, return_info=True, prefilter=True)
    rb, sb, cb, ib = store.recommend_batch(queries=base, k=k, return_info=True, prefilter=False)
    assert i8["int8_operands"]
    np.testing.assert_array_equal(r8, rb)
    np.testing.assert_array_equal(s8, sb)
    m64 = m.astype(np.float64)
    unit = m64 / np.linalg.norm(m64, axis=1, keepdims=True)
    for u in range(0, b, 9):
        qn = base[u].astype(np.float64) / np.linalg.norm(base[u].astype(np.float64))
        want_rows, want_scores = ora.topk_rows(unit @ qn, k, np.zeros(0, dtype=np.int64))
        np.testing.assert_array_equal(r8[u, :c8[u]], want_rows, err_msg=f"query {u} status {i8['status'][u]}")
        np.testing.assert_allclose(s8[u, :c8[u]], want_scores, rtol=1e-9, atol=1e-15)
