"""SURVEY.md §8(f) rows N2 (exact content retriever), N3 (catalog file format / loader / upsert), N4 (ratings batcher)."""
import numpy as np
import pytest

from oracle import reference_scoring as ora
from robot_ebert_b200 import synth
from robot_ebert_b200.batcher import RatingsBatcher, recommend_all
from robot_ebert_b200.retriever import ContentRetriever, ExactSearchEngine
from robot_ebert_b200.store_io import read_catalog_file, write_catalog_file
from tests.helpers import OracleCatalog, build_catalog_f32, build_catalog_f64


# ---------------------------------------------------------------- CPU: host logic ----------------------------
def test_catalog_file_roundtrip_cpu(tmp_path):
    m = synth.catalog_rows_f32(1, 0, 37, 64)
    for dtype, rows in [("fp32", m), ("bf16", synth.f32_to_bf16_bits(m))]:
        path = str(tmp_path / f"cat_{dtype}.rbc")
        inv, nrm = np.arange(37, dtype=np.float32), np.arange(37, dtype=np.float64) + 0.5
        write_catalog_file(path, synth.row_ids(37), rows, inv, nrm, d=64, dtype=dtype, row_base=5)
        h, r2, i2, n2 = read_catalog_file(path)
        assert (h["n"], h["d"], h["ld"], h["dtype"], h["row_base"]) == (37, 64, 64, dtype, 5) and h["ids"] == synth.row_ids(37)
        np.testing.assert_array_equal(np.asarray(r2), rows)
        np.testing.assert_array_equal(np.asarray(i2), inv)
        np.testing.assert_array_equal(np.asarray(n2), nrm)
    with pytest.raises(ValueError):
        write_catalog_file(str(tmp_path / "bad"), None, m, inv, nrm, d=64, dtype="bf16")
    (tmp_path / "junk").write_bytes(b"not a catalog")
    with pytest.raises(ValueError):
        read_catalog_file(str(tmp_path / "junk"))


def test_batcher_applies_reference_rules_cpu(golden):
    m = build_catalog_f64(golden["catalogs"]["collab32"])
    cat = OracleCatalog(synth.row_ids(m.shape[0]), m)
    b = RatingsBatcher(cat)
    assert not b.add_user("nobody", [])                                       # lib.py:39-40
    assert not b.add_user("grump", [("00000001", 1.0), ("00000002", 3.0)])     # rated, none liked
    assert b.add_user("fan", [("00000005", 5.0), ("00000007", 2.0), ("99999999", 5.0), ("00000003", 3.5)])
    assert b.add_user("dup", [("00000009", 4.0), ("00000009", 4.0)])
    lp, lc, lw, ep, ec = b.build()
    assert b.users == ["fan", "dup"] and b.no_ratings == ["nobody"] and b.none_liked == ["grump"]
    assert lp.tolist() == [0, 2, 4] and lc.tolist() == [5, 3, 9, 9] and lw.tolist() == [1, 1, 1, 1]
    assert ep.tolist() == [0, 3, 4] and ec.tolist() == [3, 5, 7, 9]           # every rated catalog movie, sorted, unique
    w = RatingsBatcher(cat, weight_fn=lambda r: max(0.0, r - 3.0))
    w.add_user("fan", [("00000005", 5.0), ("00000007", 2.0), ("00000003", 3.5)])
    assert w.build()[2].tolist() == [2.0, 0.5]


def test_retriever_and_search_engine_cpu(golden):
    m = build_catalog_f64(golden["catalogs"]["content1536_bf16"])
    ids = synth.row_ids(m.shape[0])
    cat = OracleCatalog(ids, m)
    q = synth.query_f32(1, 1536)
    nodes = ContentRetriever(cat).retrieve(q)
    want = ora.single_query(ora.catalog_frame(ids, m), q, k=10)
    assert [n.node_id for n in nodes] == [w[0] for w in want]
    np.testing.assert_allclose([n.score for n in nodes], [w[1] for w in want], rtol=1e-12)
    assert [n.score for n in nodes] == sorted((n.score for n in nodes), reverse=True)
    eng = ExactSearchEngine(ContentRetriever(cat, 5), embed=lambda text: q, respond=lambda msg, ns: f"{msg}:{len(ns)}")
    resp = eng.chat(message="space", chat_history=[])
    assert resp.response == "space:5" and [n.node_id for n in resp.source_nodes] == [w[0] for w in want[:5]]


# ---------------------------------------------------------------- GPU -----------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_save_load_roundtrip_gpu(tmp_path, dtype):
    import torch
    from robot_ebert_b200 import CatalogStore
    from robot_ebert_b200.store_io import load_catalog, save_catalog
    m = synth.catalog_rows_f32(3, 0, 5000, 96, scale_rows=True)
    a = CatalogStore.from_host(synth.row_ids(5000), m, dtype)
    path = str(tmp_path / "c.rbc")
    save_catalog(a, path)
    b = load_catalog(path)
    assert (b.n, b.d, b.ld, b.dtype, b.ids) == (a.n, a.d, a.ld, a.dtype, a.ids)
    assert torch.equal(a.rows[:a.n].view(torch.int16 if dtype == "bf16" else torch.float32),
                       b.rows[:b.n].view(torch.int16 if dtype == "bf16" else torch.float32))
    assert torch.equal(a.inv_norm[:a.n], b.inv_norm[:b.n]) and torch.equal(a.norm64[:a.n], b.norm64[:b.n])
    q = synth.query_f32(1, 96)
    ra, rb = a.recommend(query=q, k=10), b.recommend(query=q, k=10)
    np.testing.assert_array_equal(ra[0], rb[0])
    np.testing.assert_array_equal(ra[1], rb[1])


@pytest.mark.gpu
def test_upsert_equals_rebuild_gpu():
    import torch
    from robot_ebert_b200 import CatalogStore
    from robot_ebert_b200.store_io import from_chroma_result, upsert
    m = synth.catalog_rows_f32(4, 0, 300, 32)
    ids = [str(i) for i in range(0, 600, 2)]                              # unsorted as strings: "10" < "2"
    a = from_chroma_result({"ids": ids, "embeddings": m.tolist()}, dtype="bf16")
    assert a.ids == sorted(ids)
    new_ids = ["4", "7", "1001", "10"]
    new_rows = synth.catalog_rows_f32(5, 0, 4, 32)
    b = upsert(a, new_ids, new_rows)
    full = dict(zip(ids, m))
    full.update(dict(zip(new_ids, new_rows)))
    want = CatalogStore.from_host(list(full), np.stack(list(full.values())), "bf16")
    assert b.ids == want.ids and b.n == 302
    assert torch.equal(b.rows[:b.n].view(torch.int16), want.rows[:want.n].view(torch.int16))
    assert torch.equal(b.norm64[:b.n], want.norm64[:want.n])
    assert a.n == 300                                                     # the old store is untouched


@pytest.mark.gpu
def test_retriever_gpu_matches_oracle(golden):
    from robot_ebert_b200 import CatalogStore
    spec = golden["catalogs"]["content1536"]
    m = build_catalog_f32(spec)
    ids = synth.row_ids(m.shape[0])
    store = CatalogStore.from_host(ids, m, spec["dtype"])
    q = synth.query_f32(1, 1536)
    nodes = ContentRetriever(store).retrieve(q, exclude_ids=[ids[3]])
    want = ora.single_query(ora.catalog_frame(ids, build_catalog_f64(spec)), q, exclude_ids=[ids[3]], k=10)
    assert [n.node_id for n in nodes] == [w[0] for w in want]
    np.testing.assert_allclose([n.score for n in nodes], [w[1] for w in want], rtol=1e-9)


@pytest.mark.gpu
def test_recommend_all_matches_reference_goldens_gpu(golden):
    """Offline all-users job over the production-shape catalog == the reference's per-user get_user_recs outputs."""
    from robot_ebert_b200 import CatalogStore
    spec = golden["catalogs"]["collab32"]
    m = build_catalog_f32(spec)
    store = CatalogStore.from_host(synth.row_ids(m.shape[0]), m, spec["dtype"])
    cases = [c for c in golden["user_recs"] if c["catalog"] == "collab32" and c["k"] == 10]
    b = RatingsBatcher(store)
    for c in cases:
        b.add_user(c["user_id"], [tuple(r) for r in c["ratings"]])
    out = recommend_all(store, b, k=10)
    for c in cases:
        assert [i for i, _ in out[c["user_id"]]] == [e[0] for e in c["expect"]]
        np.testing.assert_allclose([s for _, s in out[c["user_id"]]], [e[1] for e in c["expect"]], rtol=1e-9)


@pytest.mark.gpu
def test_recommend_all_batched_path_gpu():
    from robot_ebert_b200 import CatalogStore
    n, d = 50_000, 256
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
    store.ids = synth.row_ids(n)
    m = store.rows[:n, :d].double().cpu().numpy()
    b = RatingsBatcher(store)
    users = synth.user_ratings(2, n, 40)
    for u, (rated, rts) in enumerate(users):
        b.add_user(f"u{u}", [(store.ids[r], float(x)) for r, x in zip(rated, rts)])
    out = recommend_all(store, b, k=10, batch_size=16)
    for u, (rated, rts) in enumerate(users):
        if f"u{u}" not in b.users:
            continue
        want_rows, want_scores = ora.recommend_rows(m, rated[rts >= 3.5], rated, 10)
        assert [int(i) for i, _ in out[f"u{u}"]] == want_rows.tolist()
        np.testing.assert_allclose([s for _, s in out[f"u{u}"]], want_scores, rtol=1e-9)
