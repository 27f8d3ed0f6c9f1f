#!/usr/bin/env python
"""A/B probe of the GEMV knobs (REBERT_GEMV_CTA_HINT, REBERT_GEMV_DYN_PCT).

For every (rows, k) case: times rebert_gemv_topk by CUDA events under each knob combination, interleaved in rounds so
that clock drift hits all combinations alike, and checks that the candidate keys are identical under all of them
(the knob may only change speed, never the result).  Prints one JSON document."""
import ctypes as C, json, os, sys
os.environ["REBERT_GEMV_TUNE"] = "1"          # make the library re-read its knobs at every launch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth, _native as nat

lib = nat.load()
dev = torch.device("cuda:0")
cases = [(1_250_000, "bf16", 10), (1_000_000, "bf16", 10), (1_000_000, "bf16", 100), (1_000_000, "fp32", 10),
         (1_250_000, "bf16", 50), (10_000_000, "bf16", 10), (10_000_000, "bf16", 100)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] <= int(sys.argv[1])]
combos = ["hint0_dyn12", "hint1_dyn0", "hint1_dyn6", "hint1_dyn12", "hint1_dyn25", "hint1_dyn100"]


def apply(combo):
    h, d = combo.split("_")
    os.environ["REBERT_GEMV_CTA_HINT"] = h[4:]
    os.environ["REBERT_GEMV_DYN_PCT"] = d[3:]

out = []
stores = {}
for n, dtype, K in cases:
    if (n, dtype) not in stores:
        stores.clear()
        torch.cuda.empty_cache()
        stores[(n, dtype)] = CatalogStore.synthetic(0, n, 1536, dtype, device=dev)
    store = stores[(n, dtype)]
    q = synth.query_f32(1, 1536)
    excl = np.random.default_rng(1).choice(n, size=133, replace=False)
    kc = lib.rebert_candidates_for_k(K)
    ptr, ne = store.stage_inputs(q, None, None, excl, K, kc)
    s = store._scratch()
    f = nat.Filter(); f.exclude_rows, f.n_exclude = ptr, ne
    st = torch.cuda.current_stream().cuda_stream

    def gemv():
        nat.check(lib.rebert_gemv_topk(C.byref(store._c), s.qn32.data_ptr(), C.byref(f), kc, s.ws.data_ptr(), s.ws.numel(),
                                       s.cand.data_ptr(), st))

    iters = 25 if n >= 5_000_000 else 150
    times = {c: [] for c in combos}
    keys = {}
    for rnd in range(4):
        for hint in (combos if rnd % 2 == 0 else combos[::-1]):     # alternate the order: no combination always runs first
            apply(hint)
            for _ in range(5):
                gemv()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                gemv()
            b.record()
            torch.cuda.synchronize()
            times[hint].append(a.elapsed_time(b) / iters * 1e3)
            keys[hint] = s.cand.cpu().numpy().copy()
    same = all(np.array_equal(keys[combos[0]], keys[c]) for c in combos)
    esz = 2 if dtype == "bf16" else 4
    res = {"rows": n, "dtype": dtype, "k": K, "kc": kc, "identical_candidates": bool(same),
           "ideal_us_at_7TBs": round(n * 1536 * esz / 7.0e12 * 1e6, 1)}
    for c in combos:
        res[f"us_{c}"] = [round(min(times[c]), 2), round(float(np.median(times[c])), 2)]      # [min, median] over the rounds
    out.append(res)
    print(json.dumps(res), flush=True)
os.environ.pop("REBERT_GEMV_CTA_HINT", None)
os.environ.pop("REBERT_GEMV_DYN_PCT", None)
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "knobs.json"), "w") as fh:
    json.dump(out, fh, indent=1)
