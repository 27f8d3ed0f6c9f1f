"""ORACLE — test infrastructure, not product code.

A CPU restatement of robot-ebert's recommendation scoring arithmetic.  Citations are into
/root/reference/.  The reference module itself (src/backend/app/lib.py) cannot be imported without
sqlalchemy / llama_index / chromadb / API keys, but its arithmetic is five pandas + scikit-learn
expressions, and both libraries are installed, so the expressions below are the reference's own,
executed on DataFrames the caller supplies instead of on SQL rows and a Chroma collection.

PARITY PIN.  The reference's tests hold no golden vector for this path, so the pin
is the reference ITSELF run in the build container: tests/golden/make_golden.py imports the real
/root/reference/src/backend/app/lib.py under stubbed third-party modules, runs its unmodified
`get_user_recs` / `run_search` on seeded inputs, and commits the outputs under tests/golden/.
tests/test_oracle.py checks every function here against those fixtures.

Precision policy: float64 throughout, on exactly the values the GPU catalog stores
(fp32 values, or bf16-rounded values, upcast to float64).

Ordering contract.  `sort_values` at lib.py:55 uses pandas' default quicksort, which is not stable,
so on exact score ties the reference's own result is implementation-defined.  The deterministic
order it does guarantee is the final `sorted(..., reverse=True)` over a tmdb_id-sorted list
(lib.py:55,63): (score desc, tmdb_id string asc).  `kind="stable"` below realises that order on
the id-sorted `unrated` index, and is what the CUDA path is held to bit-exactly.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd
from sklearn.metrics.pairwise import cosine_similarity

# src/backend/app/constants.py:19-21
LIKED_MOVIE_SCORE = 3.5
QUERY_SCORE_WEIGHT = 0.90
SIMILARITY_TOP_K = 10


def catalog_frame(ids: Sequence[str], matrix: np.ndarray) -> pd.DataFrame:
    """constants.py:55-56 — DataFrame(data=embeddings, index=ids); Python floats => float64."""
    return pd.DataFrame(data=np.asarray(matrix, dtype=np.float64), index=list(ids))


def user_recs(emb: pd.DataFrame, user_ratings: pd.DataFrame, k: int = 10, kind: str = "quicksort") -> pd.Series:
    """lib.py:42-55, verbatim apart from names.  Returns `recommended_movies` as at lib.py:55:
    a Series of <= k scores indexed by tmdb_id, sorted by tmdb_id (sort_index)."""
    # lib.py:43-44
    user_ratings = pd.DataFrame(user_ratings)
    user_ratings = user_ratings[user_ratings["tmdb_id"].isin(emb.index)]
    # lib.py:47-48
    liked_movies = user_ratings[user_ratings["rating"] >= LIKED_MOVIE_SCORE]["tmdb_id"]
    unrated_movies = emb.index.difference(user_ratings["tmdb_id"])
    # lib.py:51-52
    pairwise_similarities = cosine_similarity(emb.loc[liked_movies], emb)
    movie_scores = pd.Series(pairwise_similarities.mean(axis=0), index=emb.index)
    # lib.py:55
    return movie_scores.loc[unrated_movies].sort_values(ascending=False, kind=kind)[:k].sort_index()


def user_recs_ranked(emb: pd.DataFrame, user_ratings: pd.DataFrame, k: int = 10) -> List[Tuple[str, float]]:
    """lib.py:55-63 with the SQL join removed: (tmdb_id, score) in the order the route returns them.
    lib.py:63's `sorted(reverse=True)` is stable over the id-sorted list, i.e. (score desc, id asc)."""
    rec = user_recs(emb, user_ratings, k, kind="stable")
    pairs = list(zip(rec.index.values.tolist(), rec.values.tolist()))
    return sorted(pairs, key=lambda x: x[1], reverse=True)


def single_query(emb: pd.DataFrame, query: np.ndarray, exclude_ids: Sequence[str] = (), k: int = 10,
                 keep_mask: Optional[np.ndarray] = None) -> List[Tuple[str, float]]:
    """The L=1 form of lib.py:51-55: cosine_similarity(q[None, :], E), drop excluded labels, top-k.

    `keep_mask` (bool[N], optional) is the C5 genre/year predicate, applied exactly where the
    reference applies `unrated` (lib.py:55).  The reference has no such filter; see SURVEY.md §8d.
    """
    q = np.asarray(query, dtype=np.float64)[None, :]
    scores = pd.Series(cosine_similarity(q, emb).mean(axis=0), index=emb.index)
    allowed = emb.index.difference(pd.Index(list(exclude_ids)))
    if keep_mask is not None:
        allowed = allowed.intersection(emb.index[np.asarray(keep_mask, dtype=bool)])
    rec = scores.loc[allowed].sort_values(ascending=False, kind="stable")[:k].sort_index()
    pairs = list(zip(rec.index.values.tolist(), rec.values.tolist()))
    return sorted(pairs, key=lambda x: x[1], reverse=True)


def profile_scores(emb: pd.DataFrame, liked_ids: Sequence[str]) -> pd.Series:
    """lib.py:51-52 only: mean cosine of every catalog movie to the liked movies."""
    sims = cosine_similarity(emb.loc[list(liked_ids)], emb)
    return pd.Series(sims.mean(axis=0), index=emb.index)


def rerank(emb: pd.DataFrame, query_match_ids: Sequence[str], query_match_scores: Sequence[float],
           liked_ids: Optional[Sequence[str]], popularity: Optional[Sequence[float]] = None) -> List[Tuple[str, float]]:
    """lib.py:85-86,105-106,113-114,117,120-121 — the search re-rank blend.

    `query_match_ids` must already be sorted by id (lib.py:75).  With `liked_ids` the user score is
    the mean cosine to the liked movies (lib.py:105-106); with None it is the min-max scaled
    popularity of the matches (lib.py:113-114).  Returns (id, combined) sorted score desc.
    """
    query_movie_scores = pd.Series(data=list(query_match_scores), index=list(query_match_ids))
    if liked_ids is not None:
        sims = cosine_similarity(emb.loc[list(liked_ids)], emb.loc[list(query_match_ids)])
        user_movie_scores = pd.Series(sims.mean(axis=0), index=emb.loc[list(query_match_ids)].index)
    else:
        user_movie_scores = pd.Series(data=list(popularity), index=list(query_match_ids))
        user_movie_scores = (user_movie_scores - user_movie_scores.min()) / (user_movie_scores.max() - user_movie_scores.min())
    combined = (QUERY_SCORE_WEIGHT * query_movie_scores + (1 - QUERY_SCORE_WEIGHT) * user_movie_scores).sort_index()
    pairs = list(zip(combined.index.values.tolist(), combined.values.tolist()))
    return sorted(pairs, key=lambda x: x[1], reverse=True)


# ---------------------------------------------------------------------------------------------
# Array forms: the same arithmetic on row indices instead of string labels, for sizes where a
# string-indexed DataFrame is impractical (N >= 1M).  Checked against the DataFrame forms above in
# tests/test_oracle.py.  sklearn's cosine_similarity = normalize(X) @ normalize(Y).T with
# row_norms = sqrt(einsum('ij,ij->i')) and zero norms replaced by 1 (_handle_zeros_in_scale).
# ---------------------------------------------------------------------------------------------

def _normalize_rows(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    norms = np.sqrt(np.einsum("ij,ij->i", x, x))
    norms[norms == 0.0] = 1.0
    return x / norms[:, None]


def scores_rows(matrix: np.ndarray, lhs: np.ndarray, chunk: int = 65536) -> np.ndarray:
    """mean_i cos(lhs_i, matrix_j) for every row j, float64, catalog streamed in chunks."""
    lhs_n = _normalize_rows(np.atleast_2d(lhs))
    n = matrix.shape[0]
    out = np.empty(n, dtype=np.float64)
    for s in range(0, n, chunk):
        blk = _normalize_rows(matrix[s:s + chunk])
        out[s:s + chunk] = (lhs_n @ blk.T).mean(axis=0)
    return out


def topk_rows(scores: np.ndarray, k: int, exclude_rows: Optional[np.ndarray] = None,
              keep_mask: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(rows int64[k'], scores f64[k']) under (score desc, row asc); k' = min(k, #allowed)."""
    n = scores.shape[0]
    allowed = np.ones(n, dtype=bool) if keep_mask is None else np.asarray(keep_mask, dtype=bool).copy()
    if exclude_rows is not None and len(exclude_rows):
        allowed[np.asarray(exclude_rows, dtype=np.int64)] = False
    rows = np.nonzero(allowed)[0]
    s = scores[rows]
    order = np.lexsort((rows, -s))[:k]
    return rows[order].astype(np.int64), s[order]


def recommend_rows(matrix: np.ndarray, liked_rows: np.ndarray, exclude_rows: Optional[np.ndarray], k: int,
                   keep_mask: Optional[np.ndarray] = None, weights: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Row-index form of user_recs_ranked.  `weights` generalises lib.py:52's unweighted mean to
    sum_i w_i cos_i / sum_i w_i (the reference is w = 1[rating >= 3.5])."""
    lhs = np.asarray(matrix[np.asarray(liked_rows, dtype=np.int64)], dtype=np.float64)
    if weights is None:
        return topk_rows(scores_rows(matrix, lhs), k, exclude_rows, keep_mask)
    w = np.asarray(weights, dtype=np.float64)
    sims = _normalize_rows(lhs) @ _normalize_rows(matrix).T
    return topk_rows((w[:, None] * sims).sum(axis=0) / w.sum(), k, exclude_rows, keep_mask)


def query_rows(matrix: np.ndarray, query: np.ndarray, exclude_rows: Optional[np.ndarray], k: int,
               keep_mask: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Row-index form of single_query."""
    return topk_rows(scores_rows(matrix, np.asarray(query, dtype=np.float64)[None, :]), k, exclude_rows, keep_mask)
