"""Counter-based synthetic inputs, bit-identical on CPU (numpy, here) and GPU (csrc/synth.cu).

SURVEY.md §8(d) asks for a catalog whose value at (row r, col c) is a pure function of
(seed, r, c) so any shard or chunk can be regenerated anywhere.  Box-Muller would need
logf/cosf, which differ between libm and CUDA, so the value is an Irwin-Hall(4) variate built
from integer arithmetic only:

    h = splitmix64(seed * K + r * D + c)
    s = sum of the four 16-bit fields of h                (0 .. 262140)
    v = float32(s - 131070) * float32(1 / 37837.2273)     (mean 0, variance 1)
    v *= 2 ** ((splitmix64(seed ^ ROWSALT + r) % 9) - 4)  (optional exact per-row scale)

Every step is exact integer work or a single IEEE fp32 multiply, hence bit-identical.  The bf16
variant is round-to-nearest-even of the fp32 value, also done with integer ops.

Shapes follow the reference's data: ids are tmdb_id *strings* (src/backend/app/constants.py:56),
ratings sit on a 0.5 grid with ~63.7 % >= 3.5 and ~133 ratings per user
(notebooks/create-embeddings.ipynb:961-975).
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED_MUL = np.uint64(0xD6E8FEB86659FD93)
ROW_SALT = np.uint64(0xA0761D6478BD642F)
IH4_SCALE = np.float32(1.0 / 37837.2273)  # 1/sqrt(4 * (65536**2 - 1) / 12)

# empirical MovieLens-small histogram on the 0.5 grid (create-embeddings.ipynb:963-975), normalised
RATING_GRID = np.arange(0.5, 5.01, 0.5)
RATING_PROB = np.array([0.0137, 0.0279, 0.0178, 0.0749, 0.0551, 0.1988, 0.1303, 0.2660, 0.0848, 0.1307])
RATING_PROB = RATING_PROB / RATING_PROB.sum()


def splitmix64(z: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def row_scale_exponent(seed: int, rows: np.ndarray) -> np.ndarray:
    """Per-row power-of-two exponent in [-4, 4]."""
    with np.errstate(over="ignore"):
        h = splitmix64((np.uint64(seed) ^ ROW_SALT) + rows.astype(np.uint64))
    return (h % np.uint64(9)).astype(np.int32) - 4


def catalog_rows_f32(seed: int, row0: int, nrows: int, d: int, scale_rows: bool = False) -> np.ndarray:
    """Rows [row0, row0+nrows) of the synthetic fp32 catalog with D = d columns."""
    rows = np.arange(row0, row0 + nrows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = np.uint64(seed) * SEED_MUL
        idx = rows[:, None] * np.uint64(d) + np.arange(d, dtype=np.uint64)[None, :]
        h = splitmix64(base + idx)
    m = np.uint64(0xFFFF)
    s = (h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))
    v = (s.astype(np.int64) - 131070).astype(np.float32) * IH4_SCALE
    if scale_rows:
        e = row_scale_exponent(seed, rows)
        v = v * np.exp2(e.astype(np.float32))[:, None]
    return v


def query_f32(seed: int, d: int) -> np.ndarray:
    """One synthetic query vector (row 0 of the stream with this seed)."""
    return catalog_rows_f32(seed, 0, 1, d)[0]


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns (finite inputs)."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = (b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)
    return r.astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def quantise(x: np.ndarray, dtype: str) -> np.ndarray:
    """The values a catalog of `dtype` actually stores, as float64 (what the oracle must score)."""
    if dtype == "fp32":
        return np.asarray(x, dtype=np.float32).astype(np.float64)
    if dtype == "bf16":
        return bf16_bits_to_f32(f32_to_bf16_bits(x)).astype(np.float64)
    raise ValueError(f"dtype must be 'fp32' or 'bf16', got {dtype!r}")


def row_ids(n: int, width: int = 8) -> list:
    """tmdb_id strings whose lexicographic order equals row order."""
    return [str(r).zfill(width) for r in range(n)]


def user_ratings(seed: int, n_catalog: int, n_users: int, mean_rated: float = 133.0, max_rated: int | None = None):
    """Ragged per-user ratings: list of (row_ids int64[], ratings float64[]) per user.

    Count ~ clipped log-normal with the reference's mean; rows uniform without replacement; ratings
    from the MovieLens 0.5-grid histogram.
    """
    rng = np.random.Generator(np.random.Philox(seed))
    out = []
    hi = min(n_catalog, max_rated if max_rated is not None else n_catalog)
    sigma = 0.9
    mu = np.log(mean_rated) - 0.5 * sigma * sigma
    for _ in range(n_users):
        cnt = int(np.clip(np.round(rng.lognormal(mu, sigma)), 1, hi))
        rows = np.sort(rng.choice(n_catalog, size=cnt, replace=False)).astype(np.int64)
        rts = rng.choice(RATING_GRID, size=cnt, p=RATING_PROB)
        out.append((rows, rts))
    return out


def movie_metadata(seed: int, row0: int, nrows: int):
    """C5 side columns: genre_bits uint32 (1-4 of 19 bits) and year uint16 in [1920, 2023]."""
    rows = np.arange(row0, row0 + nrows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(seed) * SEED_MUL + rows)
        g = np.zeros(nrows, dtype=np.uint32)
        nb = (h & np.uint64(3)).astype(np.int64) + 1
        hh = h
        for j in range(4):
            hh = splitmix64(hh)
            bit = (hh % np.uint64(19)).astype(np.uint32)
            g |= np.where(j < nb, np.uint32(1) << bit, np.uint32(0)).astype(np.uint32)
        year = (1920 + (splitmix64(hh) % np.uint64(104))).astype(np.uint16)
    return g, year
