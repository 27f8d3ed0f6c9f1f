"""Exact content retrieval in place of the Chroma / hnswlib top-10.

The reference searches its 1536-d `movies-content` collection only through llama-index -> Chroma -> hnswlib, an
approximate index (src/backend/app/constants.py:40-53, lib.py:74-75).  The single-query kernel IS that operation,
exactly: cosine against every row, top-k.  `ContentRetriever.retrieve` returns what lib.py:75,85-86 consume —
objects with `.node_id` (tmdb_id) and `.score` — where score = cosine similarity (Chroma's cosine distance is
1 - cos, create-embeddings.ipynb:1302-1314; the reference's distance->score mapping inside llama-index is unpinned, so
the mapping is stated here rather than copied).  `ExactSearchEngine` gives `run_search` a `.chat()` with the same
shape as the llama-index chat engine, with the embedding model and the reply writer injected (they are hosted APIs in
the reference and stay outside this repo).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

SIMILARITY_TOP_K = 10      # constants.py:21


@dataclass
class NodeWithScore:
    node_id: str
    score: float


@dataclass
class ChatResponse:
    response: str
    source_nodes: List[NodeWithScore]


class ContentRetriever:
    """`movies_content_retriever` (constants.py:43-46) without llama-index, Chroma or hnswlib."""

    def __init__(self, catalog, similarity_top_k: int = SIMILARITY_TOP_K):
        self.catalog, self.similarity_top_k = catalog, similarity_top_k

    def retrieve(self, query_embedding: Sequence[float], exclude_ids: Sequence[str] = ()) -> List[NodeWithScore]:
        rows_ex = [r for r in (self.catalog.row_of(i) for i in exclude_ids) if r is not None]
        rows, scores = self.catalog.recommend(query=np.asarray(query_embedding, dtype=np.float32),
                                              exclude_rows=np.asarray(rows_ex, dtype=np.int64) if rows_ex else None,
                                              k=self.similarity_top_k)
        return [NodeWithScore(node_id=self.catalog.id_of(int(r)), score=float(s)) for r, s in zip(rows, scores)]


class ExactSearchEngine:
    """Drop-in for `movies_content_chat_engine` as used at lib.py:74: `.chat(message=, chat_history=)` ->
    object with `.source_nodes` and `.response`."""

    def __init__(self, retriever: ContentRetriever, embed: Callable[[str], Sequence[float]],
                 respond: Optional[Callable[[str, List[NodeWithScore]], str]] = None):
        self.retriever, self.embed, self.respond = retriever, embed, respond

    def chat(self, message: str, chat_history=None) -> ChatResponse:
        nodes = self.retriever.retrieve(self.embed(message))
        text = self.respond(message, nodes) if self.respond else ""
        return ChatResponse(response=text, source_nodes=nodes)
