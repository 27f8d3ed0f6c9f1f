#!/usr/bin/env python
"""Where one request's wall time goes, end to end with host buffers (CatalogStore.recommend -> rebert_recommend_host):
Python wrapper, packing into the pinned block, enqueueing the kernels, waiting for the completion token (= launch latency +
kernels + the PCIe write of the result), unpacking.  Medians over many requests.
    python tools/latency_breakdown.py [rows dim dtype] ..."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth

shapes = [(2269, 32, "fp32"), (2264, 1536, "fp32"), (100_000, 1536, "bf16"), (1_250_000, 1536, "bf16")]
if len(sys.argv) > 3:
    shapes = [(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])]
for n, d, dt in shapes:
    store = CatalogStore.synthetic(0, n, d, dt)
    rated, rts = synth.user_ratings(3, n, 1, mean_rated=34)[0]
    liked = rated[rts >= 3.5] if (rts >= 3.5).any() else rated[:1]
    q = synth.query_f32(1, d)
    out = {"rows": n, "dim": d, "dtype": dt, "n_liked": int(len(liked)), "n_rated": int(len(rated))}
    for name, kw in (("query", dict(query=q, exclude_rows=rated)), ("profile", dict(liked_rows=liked, exclude_rows=rated))):
        for _ in range(200):
            store.recommend(k=10, **kw)
        reps = 2000 if n < 500_000 else 300
        tot, parts = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            _, _, info = store.recommend(k=10, return_info=True, **kw)
            tot.append((time.perf_counter() - t0) * 1e6)
            parts.append(info["host_us"])
        med = lambda v: round(float(np.median(v)), 2)
        c_call = [p["pack"] + p["enqueue"] + p["wait"] + p["unpack"] for p in parts]
        out[name] = {"wall_us": med(tot), "python_wrapper_us": med(np.array(tot) - np.array(c_call)), "pack_us": med([p["pack"] for p in parts]),
                     "enqueue_us": med([p["enqueue"] for p in parts]), "wait_us": med([p["wait"] for p in parts]),
                     "unpack_us": med([p["unpack"] for p in parts]), "p99_wall_us": round(float(np.percentile(tot, 99)), 2)}
    if "prefilter" in sys.argv[4:]:
        # the int8-shadow route of the same query request, beside the back-to-back device-resident time of its two launches
        store.enable_prefilter()
        kw = dict(query=q, exclude_rows=rated, prefilter=True)
        for _ in range(50):
            store.recommend(k=10, **kw)
        tot, parts = [], []
        for _ in range(200):
            t0 = time.perf_counter()
            _, _, info = store.recommend(k=10, return_info=True, **kw)
            tot.append((time.perf_counter() - t0) * 1e6)
            parts.append(info["host_us"])
        assert info["prefilter"]
        ptr, ne = store.stage_inputs(q, None, None, rated, 10, 256)
        for _ in range(10):
            store.enqueue_fused(10, 256, ptr, ne, None, prefilter=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            store.enqueue_fused(10, 256, ptr, ne, None, prefilter=True)
        b.record(); torch.cuda.synchronize()
        one = []
        for _ in range(20):                                  # one request at a time, device-resident: no overlap with a successor
            a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a1.record(); store.enqueue_fused(10, 256, ptr, ne, None, prefilter=True); b1.record(); torch.cuda.synchronize()
            one.append(a1.elapsed_time(b1) * 1e3)
        out["prefilter_query"] = {"wall_us": med(tot), "wait_us": med([p["wait"] for p in parts]), "enqueue_us": med([p["enqueue"] for p in parts]),
                                  "attempts": info["attempts"], "device_back_to_back_us": round(a.elapsed_time(b) / 50 * 1e3, 2),
                                  "device_one_at_a_time_us": med(one)}
    print(json.dumps(out), flush=True)
