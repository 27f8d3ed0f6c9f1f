"""Ratings -> CSR batcher for offline "recommend for every user" jobs.

Turns many users' (tmdb_id, rating) rows — the `ratings` table of src/backend/app/database.py:82-90 — into the
ragged CSR the batched tensor-core path consumes, applying exactly the per-user logic of lib.py:43-48:
ratings of movies without an embedding are dropped (lib.py:44), liked = rating >= 3.5 (lib.py:47, weight 1 by
default), excluded = every rated movie (lib.py:48).  Users the reference could not serve are reported, not guessed:
no ratings -> [] (lib.py:39-40); ratings but none liked -> the reference raises (sklearn).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

LIKED_MOVIE_SCORE = 3.5   # constants.py:19


class RatingsBatcher:
    def __init__(self, catalog, weight_fn: Optional[Callable[[float], float]] = None):
        """weight_fn maps a rating to a non-negative weight (0 = not liked); default 1[rating >= 3.5]."""
        self.catalog = catalog
        self.weight_fn = weight_fn or (lambda r: 1.0 if r >= LIKED_MOVIE_SCORE else 0.0)
        self.users: List[str] = []
        self.no_ratings: List[str] = []
        self.none_liked: List[str] = []
        self._liked: List[np.ndarray] = []
        self._w: List[np.ndarray] = []
        self._excl: List[np.ndarray] = []

    def add_user(self, user_id: str, ratings: Iterable[Tuple[str, float]]) -> bool:
        ratings = list(ratings)
        if not ratings:
            self.no_ratings.append(user_id)
            return False
        rows, w = [], []
        for tmdb_id, rating in ratings:
            r = self.catalog.row_of(tmdb_id)
            if r is None:
                continue
            rows.append(r)
            w.append(self.weight_fn(float(rating)))
        rows, w = np.asarray(rows, dtype=np.int64), np.asarray(w, dtype=np.float32)
        liked = w > 0
        if not liked.any():
            self.none_liked.append(user_id)
            return False
        self.users.append(user_id)
        self._liked.append(rows[liked])
        self._w.append(w[liked])
        self._excl.append(np.unique(rows))
        return True

    def __len__(self):
        return len(self.users)

    def build(self):
        """(liked_ptr int64[b+1], liked_col int32, liked_w float32, excl_ptr int64[b+1], excl_col int32)."""
        lp = np.zeros(len(self.users) + 1, dtype=np.int64)
        ep = np.zeros(len(self.users) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in self._liked], out=lp[1:])
        np.cumsum([len(x) for x in self._excl], out=ep[1:])
        cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt))
        return lp, cat(self._liked, np.int32), cat(self._w, np.float32), ep, cat(self._excl, np.int32)


def recommend_all(catalog, batcher: RatingsBatcher, k: int = 10, batch_size: int = 4096) -> Dict[str, List[Tuple[str, float]]]:
    """get_user_recs for every batched user: {user_id: [(tmdb_id, score), ...]} ordered like lib.py:63."""
    lp, lc, lw, ep, ec = batcher.build()
    out: Dict[str, List[Tuple[str, float]]] = {u: [] for u in batcher.no_ratings}
    uniform = bool(np.all(lw == 1.0))
    for s in range(0, len(batcher), batch_size):
        e = min(len(batcher), s + batch_size)
        sub_lp, sub_ep = lp[s:e + 1] - lp[s], ep[s:e + 1] - ep[s]
        rows, scores, counts = catalog.recommend_batch(
            liked_ptr=sub_lp, liked_col=lc[lp[s]:lp[e]], liked_w=None if uniform else lw[lp[s]:lp[e]],
            excl_ptr=sub_ep, excl_col=ec[ep[s]:ep[e]], k=k)
        for i, u in enumerate(batcher.users[s:e]):
            out[u] = [(catalog.id_of(int(r)), float(x)) for r, x in zip(rows[i, :counts[i]], scores[i, :counts[i]])]
    return out
