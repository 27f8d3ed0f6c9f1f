"""Drop-in for the scoring code of the reference's src/backend/app/lib.py.

Same function names, arguments, defaults, return types and error behaviour as the reference:

    get_user_recs(user_id: str, k: int = 10) -> List[Recommendation]                     (lib.py:32)
    run_search(chat_messages, user_id: Optional[str] = None, k: int = 10) -> SearchResponse   (lib.py:66)

What changes is only what happens between the SQL read and the SQL join: the pandas / scikit-learn block
(lib.py:43-55 and :105-106) becomes one call into the HBM-resident CatalogStore.  Everything the reference keeps in
module globals (constants.py:26-56: SQL engine, chat engine, the embeddings DataFrame) is installed here with
`configure(...)`; the SQL and chat sides stay on the host exactly as in the reference.
"""
from __future__ import annotations

from typing import List, Optional, Protocol, Sequence, Tuple

import numpy as np

from .models import ChatMessage, Movie, Recommendation, SearchResponse

# src/backend/app/constants.py:19-21 — carried over verbatim
LIKED_MOVIE_SCORE = 3.5
QUERY_SCORE_WEIGHT = 0.90
SIMILARITY_TOP_K = 10


class RatingsAndMovies(Protocol):
    """The two SQL reads on the path (lib.py:36-38 and lib.py:23-29)."""

    def user_ratings(self, user_id: str) -> Sequence[Tuple[str, float]]:
        """(tmdb_id, rating) rows of `SELECT * FROM ratings WHERE user_id = :user_id`."""

    def get_movies(self, tmdb_ids: List[str]) -> List[Movie]:
        """`SELECT * FROM movies WHERE tmdb_id IN (...) ORDER BY tmdb_id` as Movie objects."""


class SqlAlchemyStore:
    """The reference's own SQL layer (database.py tables + an Engine), for deployments that have SQLAlchemy."""

    def __init__(self, engine, database):
        self.engine, self.database = engine, database

    def user_ratings(self, user_id):
        from sqlalchemy import select
        with self.engine.begin() as cnx:
            rows = cnx.execute(select(self.database.ratings).where(self.database.ratings.c.user_id == user_id)).all()
        return [(r.tmdb_id, r.rating) for r in rows]

    def get_movies(self, tmdb_ids):
        from sqlalchemy import select
        with self.engine.begin() as cnx:
            st = select(self.database.movies).where(self.database.movies.c.tmdb_id.in_(tmdb_ids)).order_by(self.database.movies.c.tmdb_id)
            return [Movie(**{k: v for k, v in row._asdict().items() if k in Movie.model_fields}) for row in cnx.execute(st).all()]


class _State:
    catalog = None          # CatalogStore (or ShardedCatalog) over the movies-collab embeddings (constants.py:55-56)
    sql: Optional[RatingsAndMovies] = None
    chat_engine = None      # object with .chat(message=, chat_history=) -> (.source_nodes[*].node_id/.score, .response)
    strict_reference_errors = True


def configure(catalog=None, sql: Optional[RatingsAndMovies] = None, chat_engine=None,
              strict_reference_errors: Optional[bool] = None) -> None:
    """Install the process-global collaborators (the reference builds them at import, constants.py:26-56)."""
    if catalog is not None:
        _State.catalog = catalog
    if sql is not None:
        _State.sql = sql
    if chat_engine is not None:
        _State.chat_engine = chat_engine
    if strict_reference_errors is not None:
        _State.strict_reference_errors = strict_reference_errors


def get_movies(tmdb_ids: List[str]) -> List[Movie]:
    """get a list of Movie objects sorted by ID (lib.py:23-29)"""
    return _State.sql.get_movies(tmdb_ids)


def _rated_rows(ratings: Sequence[Tuple[str, float]]):
    """lib.py:43-47: the user's ratings restricted to catalog movies -> (rated rows, liked rows)."""
    cat = _State.catalog
    rated, liked = [], []
    for tmdb_id, rating in ratings:
        row = cat.row_of(tmdb_id)
        if row is None:                      # lib.py:44 — drop ratings of movies without an embedding
            continue
        rated.append(row)
        if rating >= LIKED_MOVIE_SCORE:      # lib.py:47
            liked.append(row)
    return np.asarray(rated, dtype=np.int64), np.asarray(liked, dtype=np.int64)


def get_user_recs(user_id: str, k: int = 10) -> List[Recommendation]:
    """get a list of movie recommendations based on a user's collaborative filtering embedding (lib.py:32-63)"""
    cat = _State.catalog
    ratings = _State.sql.user_ratings(user_id)                               # lib.py:36-38 — the one ratings read
    if not ratings:
        return []                                                            # lib.py:39-40
    rated, liked = _rated_rows(ratings)
    # lib.py:48-55 on the GPU: mean cosine to the liked movies, rated movies masked, top-k.
    # With no liked movie this raises ValueError exactly like sklearn does in the reference.
    rows, scores = cat.recommend(liked_rows=liked, exclude_rows=rated, k=k)
    ids = [cat.id_of(int(r)) for r in rows]
    score_of = dict(zip(ids, scores.tolist()))
    movies = get_movies(tmdb_ids=ids)                                        # lib.py:58, sorted by tmdb_id
    recommendations = [Recommendation(movie=m, score=score_of[m.tmdb_id]) for m in movies]
    return sorted(recommendations, key=lambda x: x.score, reverse=True)     # lib.py:63 (stable: ties stay id-ascending)


def run_search(chat_messages: List[ChatMessage], user_id: Optional[str] = None, k: int = 10) -> SearchResponse:
    """get a list of movie recommendations based on a user's search query embedding (lib.py:66-125)"""
    cat = _State.catalog
    message = chat_messages[-1].content
    chat_history = chat_messages[:-1]
    chat_response = _State.chat_engine.chat(message=message, chat_history=chat_history)          # lib.py:74
    source_nodes = sorted(chat_response.source_nodes, key=lambda x: x.node_id)                   # lib.py:75
    query_match_movies = [m.node_id for m in source_nodes]
    query_scores = np.asarray([m.score for m in source_nodes], dtype=np.float64)                 # lib.py:85-86
    query_movies = get_movies(tmdb_ids=query_match_movies)                                       # lib.py:89

    if user_id:
        _, liked = _rated_rows(_State.sql.user_ratings(user_id))                                 # lib.py:94-98
        if len(liked) == 0 and not _State.strict_reference_errors:
            user_scores = query_scores                                                           # the intent of lib.py:101-102
        else:
            if len(liked) == 0:
                raise ValueError("Found array with 0 sample(s): user has no liked movies in the catalog")
            rows = np.asarray([cat.row_of(i) for i in query_match_movies], dtype=np.int64)
            _, p64, _ = cat.build_profiles(np.array([0, len(liked)], dtype=np.int64), liked)
            user_scores = cat.score_subset(p64, rows)[0]                                         # lib.py:105-106
    else:
        pop = np.asarray([m.popularity for m in query_movies], dtype=np.float64)                 # lib.py:113-114
        user_scores = (pop - pop.min()) / (pop.max() - pop.min())
    combined = QUERY_SCORE_WEIGHT * query_scores + (1 - QUERY_SCORE_WEIGHT) * user_scores        # lib.py:117
    recommendations = [Recommendation(movie=m, score=float(s)) for m, s in zip(query_movies, combined)]   # lib.py:120
    recommendations = sorted(recommendations, key=lambda x: x.score, reverse=True)               # lib.py:121
    return SearchResponse(message=chat_response.response, recommendations=recommendations)
