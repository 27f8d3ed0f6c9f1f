"""Shared test helpers: rebuild golden catalogs from their specs (same code path as make_golden.py)."""
import numpy as np

from robot_ebert_b200 import synth


def build_catalog_f32(spec) -> np.ndarray:
    """The fp32 values handed to the catalog store (before dtype quantisation)."""
    if "matrix" in spec:
        return np.asarray(spec["matrix"], dtype=np.float32)
    m = synth.catalog_rows_f32(spec["seed"], 0, spec["n"], spec["d"], spec.get("scale_rows", False))
    for dst, src in spec.get("dup_rows", []):
        m[dst] = m[src]
    for r in spec.get("zero_rows", []):
        m[r] = 0.0
    return m


def build_catalog_f64(spec) -> np.ndarray:
    """The values the catalog of spec['dtype'] stores, upcast to float64 (the oracle's input)."""
    return synth.quantise(build_catalog_f32(spec), spec["dtype"])


# ---------------------------------------------------------------------------------------------
# Test doubles for the host-side logic (CPU tests): same interface as CatalogStore, arithmetic by the oracle.
# ---------------------------------------------------------------------------------------------
class OracleCatalog:
    """CatalogStore look-alike whose arithmetic is oracle/reference_scoring.py — tests only."""

    def __init__(self, ids, matrix_f64, row_base=0):
        self.ids, self.m, self.row_base = list(ids), np.asarray(matrix_f64, dtype=np.float64), row_base
        self.n, self.d = self.m.shape
        self._row = {i: r + row_base for r, i in enumerate(self.ids)}

    def row_of(self, tmdb_id):
        return self._row.get(tmdb_id)

    def id_of(self, row):
        return self.ids[row - self.row_base]

    def recommend(self, *, query=None, liked_rows=None, weights=None, exclude_rows=None, k=10, row_filter=None):
        from oracle import reference_scoring as ora
        if liked_rows is not None:
            if len(liked_rows) == 0:
                raise ValueError("Found array with 0 sample(s)")
            return ora.recommend_rows(self.m, np.asarray(liked_rows), exclude_rows, k)
        return ora.query_rows(self.m, np.asarray(query, dtype=np.float64), exclude_rows, k)

    def build_profiles(self, row_ptr, col, w=None):
        from oracle.reference_scoring import _normalize_rows
        out = []
        for u in range(len(row_ptr) - 1):
            rows = np.asarray(col[row_ptr[u]:row_ptr[u + 1]], dtype=np.int64)
            out.append(_normalize_rows(self.m[rows]).mean(axis=0))
        p = np.stack(out)
        return p.astype(np.float32), p, None

    def score_subset(self, p64, sub_rows):
        from oracle.reference_scoring import _normalize_rows
        return np.asarray(p64) @ _normalize_rows(self.m[np.asarray(sub_rows, dtype=np.int64)]).T


class FakeSql:
    """In-memory stand-in for the two SQL reads (the reference's tests swap the engine the same way)."""

    def __init__(self):
        self.ratings, self.movies = {}, {}

    def user_ratings(self, user_id):
        return list(self.ratings.get(user_id, []))

    def get_movies(self, tmdb_ids):
        return sorted((self.movies[i] for i in set(tmdb_ids) if i in self.movies), key=lambda m: m.tmdb_id)


class FakeChatEngine:
    def __init__(self):
        self.nodes = []

    def chat(self, message, chat_history):
        import types
        return types.SimpleNamespace(source_nodes=[types.SimpleNamespace(node_id=i, score=s) for i, s in self.nodes],
                                     response="stub reply")


def fake_movie(tmdb_id: str, popularity: float = 1.0):
    from datetime import date
    from robot_ebert_b200.models import Movie
    return Movie(tmdb_id=tmdb_id, tmdb_homepage=f"https://www.themoviedb.org/movie/{int(tmdb_id)}", title=f"movie {tmdb_id}",
                 language="en", release_date=date(2000, 1, 1), runtime=90, director="d", actors=["a"], genres=["g"],
                 keywords=["k"], overview="o", budget=1, revenue=2, popularity=popularity, vote_average=5.0, vote_count=10)
