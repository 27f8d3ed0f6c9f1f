"""The oracle against the unmodified reference's own outputs (tests/golden/), CPU only."""
import numpy as np
import pandas as pd
import pytest

from oracle import reference_scoring as ora
from robot_ebert_b200 import synth
from tests.helpers import build_catalog_f64


def _frame(golden, name):
    m = build_catalog_f64(golden["catalogs"][name])
    return ora.catalog_frame(synth.row_ids(m.shape[0]), m), m


def test_golden_has_cases(golden):
    assert len(golden["user_recs"]) >= 25 and len(golden["search"]) >= 8


def test_user_recs_matches_reference(golden):
    """oracle.user_recs_ranked == reference get_user_recs (lib.py:32-63), ids and float64 scores exact."""
    frames = {}
    for case in golden["user_recs"]:
        emb = frames.setdefault(case["catalog"], _frame(golden, case["catalog"]))[0]
        ratings = pd.DataFrame(case["ratings"], columns=["tmdb_id", "rating"])
        if not case["ratings"]:
            assert case["expect"] == []          # lib.py:39-40 returns [] before any arithmetic
            continue
        if case.get("raises"):
            with pytest.raises(ValueError):
                ora.user_recs_ranked(emb, ratings, case["k"])
            continue
        got = ora.user_recs_ranked(emb, ratings, case["k"])
        assert [g[0] for g in got] == [e[0] for e in case["expect"]], case["user_id"]
        assert [g[1] for g in got] == [e[1] for e in case["expect"]], case["user_id"]


def test_row_form_matches_reference(golden):
    """The array (row-index) oracle used at large N reproduces the reference's ids; scores to 1e-14."""
    cache = {}
    for case in golden["user_recs"]:
        if case.get("raises") or not case["ratings"]:
            continue
        m = cache.setdefault(case["catalog"], _frame(golden, case["catalog"]))[1]
        n = m.shape[0]
        rated = [(int(i), r) for i, r in case["ratings"] if int(i) < n]
        liked = np.array([i for i, r in rated if r >= ora.LIKED_MOVIE_SCORE], dtype=np.int64)
        excl = np.array([i for i, _ in rated], dtype=np.int64)
        rows, scores = ora.recommend_rows(m, liked, excl, case["k"])
        assert [synth.row_ids(n)[r] for r in rows] == [e[0] for e in case["expect"]], case["user_id"]
        np.testing.assert_allclose(scores, [e[1] for e in case["expect"]], rtol=0, atol=1e-14)


def test_rerank_matches_reference(golden):
    """oracle.rerank == reference run_search blend (lib.py:85-121)."""
    for case in golden["search"]:
        emb = _frame(golden, case["catalog"])[0]
        nodes = sorted(case["nodes"], key=lambda x: x[0])          # lib.py:75
        ids = [i for i, _ in nodes]
        if case["user_id"]:
            rated = pd.DataFrame(case["ratings"], columns=["tmdb_id", "rating"])
            liked = rated[rated["rating"] >= ora.LIKED_MOVIE_SCORE]["tmdb_id"].to_list()
            got = ora.rerank(emb, ids, [s for _, s in nodes], liked)
        else:
            got = ora.rerank(emb, ids, [s for _, s in nodes], None, [case["popularity"][i] for i in ids])
        assert [g[0] for g in got] == [e[0] for e in case["expect"]]
        assert [g[1] for g in got] == [e[1] for e in case["expect"]]


def test_single_query_is_L1_case(golden):
    """single_query(q) equals user_recs with one liked row equal to q (both are lib.py:51-55 with L=1)."""
    emb, m = _frame(golden, "content1536_bf16")
    ids = synth.row_ids(m.shape[0])
    got = ora.single_query(emb, m[5], exclude_ids=[ids[5]], k=10)
    ratings = pd.DataFrame([(ids[5], 5.0)], columns=["tmdb_id", "rating"])
    ref = ora.user_recs_ranked(emb, ratings, 10)
    assert got == ref
    rows, scores = ora.query_rows(m, m[5], np.array([5]), 10)
    assert [ids[r] for r in rows] == [g[0] for g in got]
    np.testing.assert_allclose(scores, [g[1] for g in got], atol=1e-14, rtol=0)


def test_stable_tie_break_is_row_ascending():
    """Duplicate rows tie exactly; the contract order is (score desc, id asc)."""
    m = synth.catalog_rows_f32(5, 0, 64, 32).astype(np.float64)
    m[40] = m[3]
    m[20] = m[3]
    emb = ora.catalog_frame(synth.row_ids(64), m)
    got = ora.single_query(emb, m[3], exclude_ids=[], k=4)
    assert [g[0] for g in got[:3]] == ["00000003", "00000020", "00000040"]
    rows, _ = ora.query_rows(m, m[3], None, 4)
    assert rows[:3].tolist() == [3, 20, 40]


def test_keep_mask_predicate():
    m = synth.catalog_rows_f32(6, 0, 200, 32).astype(np.float64)
    g, y = synth.movie_metadata(3, 0, 200)
    keep = ((g & 0b1011) != 0) & (y >= 1960) & (y <= 2000)
    emb = ora.catalog_frame(synth.row_ids(200), m)
    q = synth.query_f32(1, 32)
    got = ora.single_query(emb, q, exclude_ids=["00000007"], k=10, keep_mask=keep)
    rows, _ = ora.query_rows(m, q, np.array([7]), 10, keep_mask=keep)
    assert [int(i) for i, _ in got] == rows.tolist()
    assert all(keep[r] for r in rows)
