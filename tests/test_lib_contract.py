"""Drop-in contract: robot_ebert_b200.lib.get_user_recs / run_search against the unmodified reference's outputs,
and the FastAPI JSON shape of the two hot-path routes (api/users.py:150-155, api/search.py:12-17)."""
import inspect

import numpy as np
import pytest

from robot_ebert_b200 import lib, synth
from robot_ebert_b200.models import ChatMessage, Recommendation, SearchRequest, SearchResponse
from tests.helpers import FakeChatEngine, FakeSql, OracleCatalog, build_catalog_f32, build_catalog_f64, fake_movie


def _make_catalog(golden, name, gpu):
    spec = golden["catalogs"][name]
    if gpu:
        from robot_ebert_b200 import CatalogStore
        m = build_catalog_f32(spec)
        return CatalogStore.from_host(synth.row_ids(m.shape[0]), m, spec["dtype"])
    m = build_catalog_f64(spec)
    return OracleCatalog(synth.row_ids(m.shape[0]), m)


def _check_user_recs(golden, gpu):
    cats = {}
    for case in golden["user_recs"]:
        cat = cats.setdefault(case["catalog"], _make_catalog(golden, case["catalog"], gpu))
        sql = FakeSql()
        sql.ratings[case["user_id"]] = [tuple(r) for r in case["ratings"]]
        sql.movies = {i: fake_movie(i) for i in cat.ids}
        lib.configure(catalog=cat, sql=sql, strict_reference_errors=True)
        if case.get("raises"):
            with pytest.raises(ValueError):
                lib.get_user_recs(case["user_id"], case["k"])
            continue
        got = lib.get_user_recs(case["user_id"], case["k"])
        assert all(isinstance(r, Recommendation) for r in got)
        assert [r.movie.tmdb_id for r in got] == [e[0] for e in case["expect"]], case["user_id"]
        np.testing.assert_allclose([r.score for r in got], [e[1] for e in case["expect"]], rtol=1e-9, atol=1e-15)


def _check_search(golden, gpu):
    for case in golden["search"]:
        cat = _make_catalog(golden, case["catalog"], gpu)
        sql, chat = FakeSql(), FakeChatEngine()
        chat.nodes = [tuple(n) for n in case["nodes"]]
        sql.movies = {i: fake_movie(i, case["popularity"].get(i, 1.0)) for i in cat.ids}
        if case["user_id"]:
            sql.ratings[case["user_id"]] = [tuple(r) for r in case["ratings"]]
        lib.configure(catalog=cat, sql=sql, chat_engine=chat)
        resp = lib.run_search([ChatMessage(role="user", content="q")], user_id=case["user_id"])
        assert isinstance(resp, SearchResponse) and resp.message == case["message"]
        assert [r.movie.tmdb_id for r in resp.recommendations] == [e[0] for e in case["expect"]]
        np.testing.assert_allclose([r.score for r in resp.recommendations], [e[1] for e in case["expect"]], rtol=1e-12)


def test_signatures_match_reference():
    """lib.py:32 and lib.py:66."""
    s = inspect.signature(lib.get_user_recs)
    assert list(s.parameters) == ["user_id", "k"] and s.parameters["k"].default == 10
    s = inspect.signature(lib.run_search)
    assert list(s.parameters) == ["chat_messages", "user_id", "k"]
    assert s.parameters["user_id"].default is None and s.parameters["k"].default == 10
    assert (lib.LIKED_MOVIE_SCORE, lib.QUERY_SCORE_WEIGHT, lib.SIMILARITY_TOP_K) == (3.5, 0.90, 10)   # constants.py:19-21


def test_host_logic_user_recs_cpu(golden):
    _check_user_recs(golden, gpu=False)


def test_host_logic_search_cpu(golden):
    _check_search(golden, gpu=False)


def test_empty_liked_fix_forward_flag(golden):
    cat = _make_catalog(golden, "tiny_explicit", gpu=False)
    sql, chat = FakeSql(), FakeChatEngine()
    chat.nodes = [(cat.ids[i], 0.8 - 0.01 * i) for i in range(5)]
    sql.movies = {i: fake_movie(i) for i in cat.ids}
    sql.ratings["u"] = [(cat.ids[0], 1.0)]
    lib.configure(catalog=cat, sql=sql, chat_engine=chat, strict_reference_errors=True)
    with pytest.raises(ValueError):                      # the reference's behaviour (lib.py:101-106 falls through)
        lib.run_search([ChatMessage(content="q")], user_id="u")
    lib.configure(strict_reference_errors=False)
    resp = lib.run_search([ChatMessage(content="q")], user_id="u")
    assert [r.score for r in resp.recommendations] == sorted((0.8 - 0.01 * i for i in range(5)), reverse=True)
    lib.configure(strict_reference_errors=True)


def _routes_app():
    """The two hot-path routes exactly as the reference wires them."""
    from typing import List
    from fastapi import FastAPI
    app = FastAPI()

    @app.get("/users/{user_id}/recommendations/")
    def get_user_recommendations(user_id: str, k: int = 10) -> List[Recommendation]:      # api/users.py:150-155
        return lib.get_user_recs(user_id=user_id, k=k)

    @app.post("/search/")
    def search(search_request: SearchRequest) -> SearchResponse:                           # api/search.py:12-17
        return lib.run_search(chat_messages=search_request.chat_messages, user_id=search_request.user_id)
    return app


def _check_routes(golden, gpu):
    from fastapi.testclient import TestClient
    case = golden["user_recs"][0]
    cat = _make_catalog(golden, case["catalog"], gpu)
    sql, chat = FakeSql(), FakeChatEngine()
    sql.ratings[case["user_id"]] = [tuple(r) for r in case["ratings"]]
    sql.movies = {i: fake_movie(i) for i in cat.ids}
    chat.nodes = [(cat.ids[i], 0.9 - 0.01 * i) for i in range(10)]
    lib.configure(catalog=cat, sql=sql, chat_engine=chat)
    client = TestClient(_routes_app())
    r = client.get(f"/users/{case['user_id']}/recommendations/", params={"k": case["k"]})
    assert r.status_code == 200
    body = r.json()
    assert [b["movie"]["tmdb_id"] for b in body] == [e[0] for e in case["expect"]]
    assert set(body[0]) == {"movie", "score"} and len(body[0]["movie"]) == 16 and isinstance(body[0]["score"], float)
    assert [b["score"] for b in body] == sorted((b["score"] for b in body), reverse=True)
    assert client.get("/users/nobody/recommendations/").json() == []                      # lib.py:39-40
    r = client.post("/search/", json={"chat_messages": [{"role": "user", "content": "space movies"}],
                                       "user_id": case["user_id"]})
    assert r.status_code == 200
    body = r.json()
    assert set(body) == {"message", "recommendations"} and len(body["recommendations"]) == 10


def test_fastapi_json_contract_cpu(golden):
    _check_routes(golden, gpu=False)


@pytest.mark.gpu
def test_user_recs_gpu(golden):
    _check_user_recs(golden, gpu=True)


@pytest.mark.gpu
def test_search_gpu(golden):
    _check_search(golden, gpu=True)


@pytest.mark.gpu
def test_fastapi_json_contract_gpu(golden):
    _check_routes(golden, gpu=True)
