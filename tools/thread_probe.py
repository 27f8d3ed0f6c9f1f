#!/usr/bin/env python
"""Concurrent requests from host threads against one catalog (FastAPI threadpool pattern): correctness + throughput."""
import json, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth
out = []
for name, n, d, dtype in [("production collab 2269 x 32 fp32", 2269, 32, "fp32"), ("100k x 1536 bf16", 100_000, 1536, "bf16"), ("1M x 1536 bf16", 1_000_000, 1536, "bf16")]:
    store = CatalogStore.synthetic(0, n, d, dtype, scale_rows=True)
    users = synth.user_ratings(2, n, 16, mean_rated=min(133, n // 8))
    reqs = [(r[x >= 3.5] if (x >= 3.5).any() else r[:1], r) for r, x in users]
    base = [store.recommend(liked_rows=l, exclude_rows=r, k=10) for l, r in reqs]
    for nthreads in (1, 2, 4, 8, 16):
        per = max(50, 2000 // nthreads) if n < 500_000 else 200
        errors = []
        def worker(t):
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for i in range(per):
                    u = (t + i) % len(reqs)
                    rows, _ = store.recommend(liked_rows=reqs[u][0], exclude_rows=reqs[u][1], k=10)
                    if not np.array_equal(rows, base[u][0]): errors.append((t, i))
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]; dt = time.perf_counter() - t0
        out.append({"catalog": name, "threads": nthreads, "requests_per_s": round(nthreads * per / dt, 1), "wrong_results": len(errors)})
    del store
print(json.dumps(out, indent=1))
