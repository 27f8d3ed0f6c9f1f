#!/usr/bin/env python
"""Single-GPU sweep over the BASELINE.json single-query configs: kernel time (CUDA events), roofline fraction, e2e."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth, _native as nat
lib = nat.load()
dev = torch.device("cuda:0")
peak = 6546.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = []
for name, n, d, dtype, k in [("C2 1M x 1536 fp32 top-10", 1_000_000, 1536, "fp32", 10), ("C2 1M x 1536 bf16 top-10", 1_000_000, 1536, "bf16", 10),
                             ("1M x 1536 bf16 top-100", 1_000_000, 1536, "bf16", 100), ("10M x 1536 bf16 top-50 (C5 k)", 10_000_000, 1536, "bf16", 50),
                             ("production collab 2269 x 32 fp32 top-10", 2269, 32, "fp32", 10), ("production content 2264 x 1536 fp32 top-10", 2264, 1536, "fp32", 10),
                             ("C1 10k x 1536 fp32 top-10", 10_000, 1536, "fp32", 10)]:
    store = CatalogStore.synthetic(0, n, d, dtype, device=dev)
    q = synth.query_f32(1, d)
    excl = np.random.default_rng(1).choice(n, size=133, replace=False)
    kc = lib.rebert_candidates_for_k(k)
    ptr, ne = store.stage_inputs(q, None, None, excl, k, kc)
    s = store._scratch()
    f = nat.Filter(); f.exclude_rows, f.n_exclude = ptr, ne
    st = torch.cuda.current_stream().cuda_stream
    def gemv(): nat.check(lib.rebert_gemv_topk(C.byref(store._c), s.qn32.data_ptr(), C.byref(f), kc, s.ws.data_ptr(), s.ws.numel(), s.cand.data_ptr(), st))
    iters = 30 if n >= 5_000_000 else 300
    for _ in range(10): gemv()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): gemv()
    b.record(); torch.cuda.synchronize()
    kms = a.elapsed_time(b) / iters
    for _ in range(10): store.recommend(query=q, exclude_rows=excl, k=k)
    t0 = time.perf_counter()
    for _ in range(iters): rows, scores, info = store.recommend(query=q, exclude_rows=excl, k=k, return_info=True)
    e2e = (time.perf_counter() - t0) / iters
    bytes_ = n * store.ld * (2 if dtype == "bf16" else 4)
    out.append({"config": name, "kernel_ms": round(kms, 4), "kernel_GBs": round(bytes_ / kms / 1e6, 1), "frac_of_measured_hbm_peak": round(bytes_ / kms / 1e6 / peak, 3),
                "kernel_qps": round(1e3 / kms, 1), "e2e_ms": round(e2e * 1e3, 4), "e2e_qps": round(1 / e2e, 1), "kc": info["kc"], "proven_exact": bool(info["proven_exact"])})
    del store
print(json.dumps(out, indent=1))
