#!/usr/bin/env python
"""Fixed-cost probe: time rebert_gemv_topk / finalize / full recommend() over a range of catalog sizes."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth, _native as nat
lib = nat.load()
dev = torch.device("cuda:0")
out = []
for n, d, dtype in [(2269, 32, "fp32"), (2368, 1536, "bf16"), (37888, 1536, "bf16"), (303104, 1536, "bf16"), (1_000_000, 1536, "bf16"), (1_250_000, 1536, "bf16")]:
    store = CatalogStore.synthetic(0, n, d, dtype, device=dev)
    q = synth.query_f32(1, d)
    excl = np.random.default_rng(1).choice(n, size=133, replace=False)
    K = 10; kc = lib.rebert_candidates_for_k(K)
    ptr, ne = store.stage_inputs(q, None, None, excl, K, kc)
    s = store._scratch()
    f = nat.Filter(); f.exclude_rows, f.n_exclude = ptr, ne
    st = torch.cuda.current_stream().cuda_stream
    ob = s.d_out.data_ptr()
    def gemv(): nat.check(lib.rebert_gemv_topk(C.byref(store._c), s.qn32.data_ptr(), C.byref(f), kc, s.ws.data_ptr(), s.ws.numel(), s.cand.data_ptr(), st))
    def fin(): nat.check(lib.rebert_finalize_topk(C.byref(store._c), s.qn64.data_ptr(), s.cand.data_ptr(), kc, K, ob, ob + 8*K, ob + 16*K, ob + 16*K + 8, st))
    res = {"n": n, "d": d, "dtype": dtype}
    for name, fn in [("gemv_us", gemv), ("finalize_us", fin), ("both_us", lambda: (gemv(), fin()))]:
        for _ in range(20): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200): fn()
        b.record(); torch.cuda.synchronize()
        res[name] = round(a.elapsed_time(b) / 200 * 1e3, 2)
    for _ in range(20): store.recommend(query=q, exclude_rows=excl, k=K)
    t0 = time.perf_counter()
    for _ in range(200): store.recommend(query=q, exclude_rows=excl, k=K)
    res["recommend_e2e_us"] = round((time.perf_counter() - t0) / 200 * 1e6, 1)
    esz = 2 if dtype == "bf16" else 4
    res["ideal_us_at_7TBs"] = round(n * store.ld * esz / 7.0e12 * 1e6, 2)
    out.append(res)
    del store
print(json.dumps(out, indent=1))
