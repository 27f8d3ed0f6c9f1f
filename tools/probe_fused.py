#!/usr/bin/env python
"""A/B probe of the request path (round 2): streaming kernel + cluster exact pass vs the round-1 form (streaming kernel with
its own last-CTA merge + one-CTA exact pass), knobs on/off, end-to-end latency of query / liked-rows requests, tiny catalogs.  Prints JSON lines; writes gpurun_out/fused.json.

    python tools/probe_fused.py [max_rows]
"""
import ctypes as C, json, os, sys, time
os.environ["REBERT_GEMV_TUNE"] = "1"          # make the library re-read its knobs at every launch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth, _native as nat

lib = nat.load()
dev = torch.device("cuda:0")
max_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
out = []


def emit(d):
    out.append(d)
    print(json.dumps(d), flush=True)


def ev_time(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def wall_time(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters * 1e6


# ---------------------------------------------------------------- device-resident: fused vs two kernels, knobs
for n, dtype, K in [(1_250_000, "bf16", 10), (1_250_000, "bf16", 100), (1_000_000, "fp32", 10), (10_000_000, "bf16", 10)]:
    if n > max_rows:
        continue
    store = CatalogStore.synthetic(0, n, 1536, dtype, device=dev)
    q = synth.query_f32(1, 1536)
    excl = np.random.default_rng(1).choice(n, size=133, replace=False)
    kc = lib.rebert_candidates_for_k(K)
    ptr, ne = store.stage_inputs(q, None, None, excl, K, kc)
    torch.cuda.synchronize()
    s = store._scratch()
    iters = 30 if n >= 5_000_000 else 200
    res = {"rows": n, "dtype": dtype, "k": K, "kc": kc, "ideal_us_at_7.3TBs": round(n * 1536 * (2 if dtype == "bf16" else 4) / 7.3e12 * 1e6, 1)}
    two = lambda: store.enqueue_topk(K, kc, ptr, ne)
    fused = lambda: store.enqueue_fused(K, kc, ptr, ne)
    two()
    ref = s.d_out.cpu().numpy().copy()
    same = True
    rounds = {}
    for rnd in range(3):
        for name, env, fn in [("two_kernel", {}, two), ("fused", {}, fused),
                              ("fused_no_early_tma", {"REBERT_GEMV_EARLY_TMA": "0"}, fused),
                              ("fused_no_prune", {"REBERT_GEMV_MERGE_PRUNE": "0"}, fused),
                              ("fused_no_pdl_overlap", {"REBERT_GEMV_EARLY_TMA": "0", "REBERT_GEMV_MERGE_PRUNE": "0"}, fused)]:
            os.environ.update(env)
            rounds.setdefault(name, []).append(ev_time(fn, iters))
            same &= bool(np.array_equal(ref, s.d_out.cpu().numpy()))
            for k_ in env:
                os.environ.pop(k_)
    for name, v in rounds.items():
        res[f"us_{name}"] = [round(min(v), 2), round(float(np.median(v)), 2)]
    res["identical_results"] = same
    # end to end through the public API
    (rated, rts), = synth.user_ratings(2, n, 1)
    liked = rated[rts >= 3.5]
    res["e2e_query_us"] = round(wall_time(lambda: store.recommend(query=q, exclude_rows=excl, k=K), iters), 1)
    res["e2e_profile_us"] = round(wall_time(lambda: store.recommend(liked_rows=liked, exclude_rows=rated, k=K), iters), 1)
    res["n_liked"], res["n_rated"] = int(len(liked)), int(len(rated))
    emit(res)
    del store
    torch.cuda.empty_cache()

# ---------------------------------------------------------------- tiny catalogs (the reference's production shapes): pure latency
for n, d, dtype in [(2269, 32, "fp32"), (2264, 1536, "fp32"), (2264, 1536, "bf16"), (10_000, 1536, "fp32")]:
    store = CatalogStore.synthetic(0, n, d, dtype, device=dev)
    q = synth.query_f32(1, d)
    (rated, rts), = synth.user_ratings(2, n, 1)
    liked = rated[rts >= 3.5]
    res = {"rows": n, "dim": d, "dtype": dtype, "n_liked": int(len(liked)), "n_rated": int(len(rated))}
    res["e2e_query_us"] = round(wall_time(lambda: store.recommend(query=q, exclude_rows=rated, k=10), 500, 20), 1)
    res["e2e_profile_us"] = round(wall_time(lambda: store.recommend(liked_rows=liked, exclude_rows=rated, k=10), 500, 20), 1)
    emit(res)

os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "fused.json"), "w") as fh:
    json.dump(out, fh, indent=1)
