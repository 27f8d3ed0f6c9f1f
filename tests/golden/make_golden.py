#!/usr/bin/env python
"""Generate tests/golden/reference_lib_golden.json by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  It imports the real
/root/reference/src/backend/app/lib.py and /root/reference/src/shared/models.py; the third-party
modules those files import but that are not installed here (sqlalchemy, llama_index) and the two
sibling modules that need network services at import (backend.app.constants: OpenAI + Chroma +
CloudSQL; backend.app.database: sqlalchemy tables) are replaced by in-memory stubs that carry the
same names.  The arithmetic that runs — pandas isin/difference/sort_values, sklearn
cosine_similarity, the 0.9/0.1 blend — is the reference's own code, line for line.

Usage:  python tests/golden/make_golden.py            (rewrites the JSON next to this file)
"""
from __future__ import annotations

import collections
import contextlib
import importlib.util
import json
import os
import sys
import types
from datetime import date

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)

from robot_ebert_b200 import synth  # noqa: E402


# ------------------------------------------------------------------ stubs ------------------
def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


class _Cond:
    def __init__(self, op, col, val):
        self.op, self.col, self.val = op, col, val


class _Col:
    def __init__(self, table, name):
        self.table, self.name = table, name

    def __eq__(self, other):  # noqa: D105 - mimics sqlalchemy column operators
        return _Cond("eq", self.name, other)

    def in_(self, vals):
        return _Cond("in", self.name, list(vals))

    __hash__ = None


class _Table:
    def __init__(self, name, cols):
        self.name = name
        self.c = types.SimpleNamespace(**{c: _Col(name, c) for c in cols})


class _Select:
    def __init__(self, table):
        self.table, self.cond, self.order = table, None, None

    def where(self, cond):
        self.cond = cond
        return self

    def order_by(self, col):
        self.order = col.name
        return self


RatingRow = collections.namedtuple("RatingRow", ["user_id", "tmdb_id", "rating", "updated_at"])
MOVIE_FIELDS = ["tmdb_id", "tmdb_homepage", "title", "language", "release_date", "runtime", "director", "actors",
                "genres", "keywords", "overview", "budget", "revenue", "popularity", "vote_average", "vote_count"]
MovieRow = collections.namedtuple("MovieRow", MOVIE_FIELDS)


class FakeEngine:
    """Stands in for the SQLAlchemy engine (constants.py:26): serves ratings and movies from dicts."""

    def __init__(self):
        self.ratings = {}   # user_id -> list[RatingRow]
        self.movies = {}    # tmdb_id -> MovieRow

    @contextlib.contextmanager
    def begin(self):
        yield self

    def execute(self, stmt):
        if stmt.table.name == "ratings":
            rows = list(self.ratings.get(stmt.cond.val, []))
        else:
            rows = [self.movies[i] for i in set(stmt.cond.val) if i in self.movies]
            rows.sort(key=lambda r: r.tmdb_id)          # ORDER BY tmdb_id (lib.py:27)
        return types.SimpleNamespace(all=lambda: rows)


class FakeChatEngine:
    """Stands in for the llama-index chat engine (constants.py:47-53)."""

    def __init__(self):
        self.nodes = []

    def chat(self, message, chat_history):
        nodes = [types.SimpleNamespace(node_id=i, score=s) for i, s in self.nodes]
        return types.SimpleNamespace(source_nodes=nodes, response="stub reply")


def load_reference():
    from pydantic import BaseModel

    class ChatMessage(BaseModel):
        role: str = "user"
        content: str = ""

    class MessageRole:
        USER = "user"

    _mod("sqlalchemy", select=lambda t: _Select(t))
    _mod("llama_index")
    _mod("llama_index.llms", ChatMessage=ChatMessage, MessageRole=MessageRole)
    engine, chat = FakeEngine(), FakeChatEngine()
    backend = _mod("backend")
    app = _mod("backend.app")
    backend.app = app
    app.database = _mod("backend.app.database",
                        ratings=_Table("ratings", ["user_id", "tmdb_id", "rating", "updated_at"]),
                        movies=_Table("movies", MOVIE_FIELDS))
    consts = _mod("backend.app.constants", engine=engine, openai_client=None, users_collab_collection=None,
                  movies_collab_collection=None, movies_content_chat_engine=chat,
                  movies_collab_embeddings=pd.DataFrame(),
                  LIKED_MOVIE_SCORE=3.5, QUERY_SCORE_WEIGHT=0.90)   # constants.py:19-20
    app.constants = consts
    sys.path.insert(0, os.path.join(REF, "src"))
    spec = importlib.util.spec_from_file_location("backend.app.lib", os.path.join(REF, "src/backend/app/lib.py"))
    lib = importlib.util.module_from_spec(spec)
    sys.modules["backend.app.lib"] = lib
    spec.loader.exec_module(lib)
    return lib, engine, chat, ChatMessage


# ------------------------------------------------------------------ inputs -----------------
def fake_movie(tmdb_id: str) -> MovieRow:
    h = int(synth.splitmix64(np.array([int(tmdb_id) + 7], dtype=np.uint64))[0])
    return MovieRow(tmdb_id=tmdb_id, tmdb_homepage=f"https://www.themoviedb.org/movie/{int(tmdb_id)}", title=f"movie {tmdb_id}",
                    language="en", release_date=date(1920 + h % 104, 1, 1), runtime=60 + h % 120, director="d",
                    actors=["a"], genres=["g"], keywords=["k"], overview="o", budget=h % 1000, revenue=h % 5000,
                    popularity=float((h >> 20) % 10000) / 37.0, vote_average=float(h % 100) / 10.0, vote_count=h % 999)


def build_catalog(spec):
    if "matrix" in spec:
        m = np.asarray(spec["matrix"], dtype=np.float32)
    else:
        m = synth.catalog_rows_f32(spec["seed"], 0, spec["n"], spec["d"], spec.get("scale_rows", False))
        for dst, src in spec.get("dup_rows", []):
            m[dst] = m[src]
        for r in spec.get("zero_rows", []):
            m[r] = 0.0
    return synth.quantise(m, spec["dtype"])


CATALOGS = {
    "collab32": dict(seed=11, n=2269, d=32, dtype="fp32"),                       # production shape (ipynb:232,:1241)
    "collab32_scaled": dict(seed=12, n=600, d=32, dtype="fp32", scale_rows=True, zero_rows=[17]),
    "content1536": dict(seed=13, n=384, d=1536, dtype="fp32", scale_rows=True),
    "content1536_bf16": dict(seed=13, n=384, d=1536, dtype="bf16", scale_rows=True),
    "odd_d": dict(seed=14, n=257, d=50, dtype="bf16"),
    "tiny_explicit": dict(dtype="fp32", matrix=[[1, 0, 0, 0], [0, 1, 0, 0], [1, 1, 0, 0], [0, 0, 0, 0], [2, 0, 0, 0],
                                                [-1, 0, 0, 0], [0.5, 0.5, 0.5, 0.5], [3, 4, 0, 0], [0, 0, 1, 0], [1, 2, 3, 4],
                                                [4, 3, 2, 1], [1, 1, 1, 1]]),
}


def main():
    lib, engine, chat, ChatMessage = load_reference()
    out = {"generator": "tests/golden/make_golden.py", "reference": "src/backend/app/lib.py:32-63,66-121",
           "catalogs": CATALOGS, "user_recs": [], "search": []}

    def set_catalog(name):
        m = build_catalog(CATALOGS[name])
        ids = synth.row_ids(m.shape[0])
        # constants.py:55-56: DataFrame(data=<list of lists of Python floats>, index=ids)
        frame = pd.DataFrame(data=m.tolist(), index=ids)
        lib.movies_collab_embeddings = frame
        engine.movies = {i: fake_movie(i) for i in ids}
        return ids

    # ---- get_user_recs ---------------------------------------------------------------------
    cases = [("collab32", 21, 4, 10), ("collab32", 22, 3, 25), ("collab32_scaled", 23, 3, 10), ("content1536", 24, 3, 10),
             ("content1536_bf16", 24, 3, 10), ("odd_d", 25, 3, 7), ("content1536_bf16", 26, 2, 100)]
    for cat, useed, nusers, k in cases:
        ids = set_catalog(cat)
        n = len(ids)
        for u, (rows, rts) in enumerate(synth.user_ratings(useed, n, nusers, mean_rated=min(133.0, n / 4))):
            uid = f"user-{cat}-{useed}-{u}"
            rated_ids = [ids[r] for r in rows]
            rated = [(i, float(x)) for i, x in zip(rated_ids, rts)]
            if u == 0:
                rated.append(("99999999", 5.0))        # rated movie absent from the catalog (lib.py:44)
            engine.ratings[uid] = [RatingRow(uid, i, x, None) for i, x in rated]
            rec = {"catalog": cat, "user_id": uid, "k": k, "ratings": rated}
            try:
                res = lib.get_user_recs(uid, k)
                rec["expect"] = [[r.movie.tmdb_id, r.score] for r in res]
            except ValueError as e:                      # no liked movies
                rec["raises"] = "ValueError"
                rec["message"] = str(e)[:80]
            out["user_recs"].append(rec)

    # edge cases on the explicit catalog
    ids = set_catalog("tiny_explicit")
    edge = {
        "no-ratings": [],
        "none-liked": [(ids[0], 1.0), (ids[1], 3.0)],
        "one-liked": [(ids[0], 4.0)],
        "fewer-than-k": [(i, 4.0) for i in ids[:9]],
        "all-rated": [(i, 5.0) for i in ids],
        "zero-row-liked": [(ids[3], 5.0), (ids[9], 4.5)],
    }
    for name, rated in edge.items():
        uid = f"user-edge-{name}"
        engine.ratings[uid] = [RatingRow(uid, i, x, None) for i, x in rated]
        rec = {"catalog": "tiny_explicit", "user_id": uid, "k": 5, "ratings": rated}
        try:
            res = lib.get_user_recs(uid, 5)
            rec["expect"] = [[r.movie.tmdb_id, r.score] for r in res]
        except ValueError as e:
            rec["raises"] = "ValueError"
            rec["message"] = str(e)[:80]
        out["user_recs"].append(rec)

    # ---- run_search re-rank ----------------------------------------------------------------
    for cat, useed in [("collab32", 31), ("content1536_bf16", 32)]:
        ids = set_catalog(cat)
        n = len(ids)
        rng = np.random.default_rng(useed)
        for u, (rows, rts) in enumerate(synth.user_ratings(useed, n, 2, mean_rated=60)):
            match_rows = rng.choice(n, size=10, replace=False)
            chat.nodes = [(ids[r], float(s)) for r, s in zip(match_rows, rng.uniform(0.7, 0.9, size=10))]
            for with_user in (True, False):
                uid = f"user-search-{cat}-{u}" if with_user else None
                rated = [(ids[r], float(x)) for r, x in zip(rows, rts)]
                if with_user:
                    engine.ratings[uid] = [RatingRow(uid, i, x, None) for i, x in rated]
                with contextlib.redirect_stdout(open(os.devnull, "w")):
                    resp = lib.run_search([ChatMessage(role="user", content="q")], user_id=uid)
                out["search"].append({"catalog": cat, "user_id": uid, "ratings": rated if with_user else None,
                                      "nodes": chat.nodes,
                                      "popularity": {i: engine.movies[i].popularity for i, _ in chat.nodes},
                                      "expect": [[r.movie.tmdb_id, r.score] for r in resp.recommendations],
                                      "message": resp.message})

    path = os.path.join(HERE, "reference_lib_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["user_recs"]), "user_recs,", len(out["search"]), "search")


if __name__ == "__main__":
    main()
