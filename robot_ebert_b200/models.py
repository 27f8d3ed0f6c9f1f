"""Response contract of the two hot-path routes, unchanged from the reference.

Mirrors src/shared/models.py:33-49 (Movie), :73-75 (Recommendation), :77-80 (SearchRequest), :82-84 (SearchResponse).
`ChatMessage` comes from llama_index in the reference (models.py:5); that package is optional here, so a structurally
identical pydantic model (role, content) stands in when it is absent.
"""
from __future__ import annotations

from datetime import date
from typing import List, Optional

from pydantic import BaseModel

try:  # the reference's own type, when llama-index is installed
    from llama_index.llms import ChatMessage, MessageRole  # type: ignore  # noqa: F401
except Exception:  # pragma: no cover - llama-index is not in this image
    class MessageRole:  # type: ignore
        USER = "user"
        ASSISTANT = "assistant"
        SYSTEM = "system"

    class ChatMessage(BaseModel):  # type: ignore
        role: str = MessageRole.USER
        content: Optional[str] = ""


class Movie(BaseModel):
    tmdb_id: str
    tmdb_homepage: str
    title: str
    language: str
    release_date: date
    runtime: int
    director: str
    actors: Optional[List[str]]
    genres: Optional[List[str]]
    keywords: Optional[List[str]]
    overview: str
    budget: int
    revenue: int
    popularity: float
    vote_average: float
    vote_count: int


class Recommendation(BaseModel):
    movie: Movie
    score: float


class SearchRequest(BaseModel):
    chat_messages: List[ChatMessage]
    user_id: Optional[str] = None
    k: Optional[int] = 10


class SearchResponse(BaseModel):
    message: str
    recommendations: List[Recommendation]
