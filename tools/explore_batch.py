#!/usr/bin/env python
"""Randomised exploration of recommend_batch against the oracle (development aid; the fixed cases live in tests/)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import reference_scoring as ora
from robot_ebert_b200 import CatalogStore, RowFilter, synth
seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bad = 0
for t in range(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    rng = np.random.default_rng(seed0 * 1000 + t)
    d = int(rng.choice([64, 128, 192, 256, 512, 1536]))
    k = int(rng.choice([1, 10, 33, 50, 100, 150, 240]))
    n = int(rng.integers(70_000 if k <= 100 else 140_000, 260_000))
    b = int(rng.choice([1, 7, 128, 129, 255, 256, 257, 300, 600]))
    use_pred = bool(rng.random() < 0.4)
    dup = bool(rng.random() < 0.3)
    m = synth.catalog_rows_f32(int(rng.integers(0, 1000)), 0, n, d, scale_rows=True)
    if dup:
        src = rng.integers(0, n, size=200); dst = rng.integers(0, n, size=200); m[dst] = m[src]
    store = CatalogStore.from_host(None, m, "bf16")
    stored = store.rows[:n, :d].to(torch.float64).cpu().numpy()
    unit = stored / np.maximum(np.linalg.norm(stored, axis=1, keepdims=True), 1e-300)
    q = rng.standard_normal((b, d)).astype(np.float32)
    if dup:
        q[0] = m[src[0]]
    ptr = [0]; cols = []
    for u in range(b):
        c = np.sort(rng.choice(n, size=int(rng.integers(0, 400)), replace=False)); cols.append(c); ptr.append(ptr[-1] + len(c))
    rf = keep = None
    if use_pred:
        g, y = synth.movie_metadata(3, 0, n); store.set_metadata(g, y)
        rf = RowFilter(genre_any=0b111, year_lo=1950, year_hi=2010); keep = ((g & 0b111) != 0) & (y >= 1950) & (y <= 2010)
    rows, scores, counts, info = store.recommend_batch(queries=q, excl_ptr=np.array(ptr), excl_col=np.concatenate(cols) if ptr[-1] else np.zeros(0, np.int32), k=k, row_filter=rf, return_info=True)
    qn = q.astype(np.float64); qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    nbad = 0
    for u in range(b):
        wr, ws = ora.topk_rows(unit @ qn[u], k + 4, cols[u], keep)
        ok = counts[u] == min(k, len(wr)) and np.allclose(scores[u, :counts[u]], ws[:k], rtol=1e-9, atol=1e-13)
        if ok and not np.array_equal(rows[u, :counts[u]], wr[:k]):
            diff = np.nonzero(rows[u, :counts[u]] != wr[:counts[u]])[0]
            ok = all(((np.abs(ws - ws[p]) > 0) & (np.abs(ws - ws[p]) < 1e-12)).any() for p in diff)
        nbad += not ok
    bad += nbad
    print(f"t={t} n={n} d={d} b={b} k={k} pred={use_pred} dup={dup} reruns={(info['status'] != 0).sum()} mismatches={nbad}", flush=True)
    del store
print("TOTAL MISMATCHES", bad)
