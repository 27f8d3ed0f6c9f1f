#!/usr/bin/env python
"""Tuning sweep of the GEMV streaming knobs (stage bytes / stages / L2 policy) on one catalog size."""
import ctypes as C, json, os, sys
os.environ["REBERT_GEMV_TUNE"] = "1"          # make the library re-read its knobs at every launch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth, _native as nat
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
lib = nat.load(); dev = torch.device("cuda:0")
store = CatalogStore.synthetic(0, n, 1536, "bf16", device=dev)
q = synth.query_f32(1, 1536); excl = np.random.default_rng(1).choice(n, size=133, replace=False)
K = 10; kc = lib.rebert_candidates_for_k(K)
ptr, ne = store.stage_inputs(q, None, None, excl, K, kc)
s = store._scratch(); f = nat.Filter(); f.exclude_rows, f.n_exclude = ptr, ne
st = torch.cuda.current_stream().cuda_stream
def gemv(): nat.check(lib.rebert_gemv_topk(C.byref(store._c), s.qn32.data_ptr(), C.byref(f), kc, s.ws.data_ptr(), s.ws.numel(), s.cand.data_ptr(), st))
iters = 40 if n >= 5_000_000 else 300
res = []
for sb in (24576, 36864, 49152, 61440, 73728):
    for stg in (2, 3, 4, 6, 8):
        for pol in (0, 1):
            os.environ["REBERT_GEMV_STAGE_BYTES"] = str(sb); os.environ["REBERT_GEMV_STAGES"] = str(stg); os.environ["REBERT_GEMV_L2_POLICY"] = str(pol)
            try:
                for _ in range(5): gemv()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters): gemv()
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / iters
                res.append((round(n * 3072 / ms / 1e6, 1), sb, stg, pol, round(ms, 4)))
            except Exception as e:
                res.append((0, sb, stg, pol, repr(e)[:60]))
res.sort(reverse=True)
for r in res[:12]: print(r)
print("worst", res[-3:])
