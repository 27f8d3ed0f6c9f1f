#!/usr/bin/env python
"""BASELINE config 3: batched users x catalog on the tcgen05 path.  Prints one JSON line (not the driver's bench)."""
import argparse, ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth, _native as nat

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--dim", type=int, default=1536)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--int8", action="store_true", help="int8 operands over the prefilter shadow (tcgen05 kind::i8)")
args = ap.parse_args()
dev = torch.device("cuda:0")
store = CatalogStore.synthetic(0, args.rows, args.dim, "bf16", device=dev)
q = synth.catalog_rows_f32(11, 0, args.batch, args.dim)
qn32, qn64, qbf = store.prepare_queries(q)
lib = nat.load()
users = synth.user_ratings(2, args.rows, args.batch)
ep = np.zeros(args.batch + 1, dtype=np.int64); ec = []
for u, (rated, _) in enumerate(users):
    ec.append(rated); ep[u + 1] = ep[u] + len(rated)
ept = torch.from_numpy(ep).to(dev); ect = torch.from_numpy(np.concatenate(ec).astype(np.int32)).to(dev)
shadow = None
if args.int8:
    store.enable_prefilter()
    shadow = store.quantize_queries(qn32)
plan = store.gemm_plan(args.batch, args.k, shadow=args.int8)
ws = torch.empty(lib.rebert_gemm_workspace_bytes(C.byref(store._c), C.byref(plan)), dtype=torch.uint8, device=dev)
out_rows = torch.empty((args.batch, args.k), dtype=torch.int64, device=dev)
out_scores = torch.empty((args.batch, args.k), dtype=torch.float64, device=dev)
out_count = torch.empty(args.batch, dtype=torch.int32, device=dev)
out_status = torch.empty(args.batch, dtype=torch.int32, device=dev)
def step():
    store.enqueue_batch(plan, qbf, qn64, ept, ect, ws, out_rows, out_scores, out_count, out_status, shadow=shadow)
for _ in range(args.warmup): step()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(args.steps): step()
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / args.steps
flops = 2.0 * args.batch * args.rows * args.dim
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1590.0}
status = out_status.cpu().numpy()
print(json.dumps({"workload": f"{args.batch} users x {args.rows} x {args.dim} bf16 top-{args.k}", "int8_operands": args.int8, "ms_per_batch": ms,
                  "queries_per_s": args.batch / (ms * 1e-3), "tflops_whole_pipeline": flops / (ms * 1e-3) / 1e12,
                  "frac_of_measured_bf16_peak": flops / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
                  "plan": {f: getattr(plan, f) for f, _ in plan._fields_}, "fallback_queries": int((status != 0).sum()),
                  "status_hist": {int(k): int(v) for k, v in zip(*np.unique(status, return_counts=True))}}))
