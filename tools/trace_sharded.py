#!/usr/bin/env python
"""Phase stamps of the row-sharded request path (torchrun): per rank, the streaming kernel's %globaltimer stamps and the
exact-pass cluster kernel's (entry, stream complete, winners selected, exact scores in CTA 0, ranked, end = after the
NVLink exchange + merge), for a few back-to-back device-resident requests.
    torchrun --nproc-per-node N tools/trace_sharded.py [rows_total] [k]"""
import json, os, sys
os.environ["REBERT_GEMV_TUNE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from robot_ebert_b200 import synth, _native as nat
from robot_ebert_b200.sharding import ShardedCatalog
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sc = ShardedCatalog.synthetic(0, n, 1536, "bf16", device=dev)
store = sc.backend.store
lib = nat.load()
q = synth.query_f32(1, 1536); excl = np.random.default_rng(1).choice(n, size=133, replace=False)
kc = lib.rebert_candidates_for_k(K)
sc.backend._excl = store.stage_inputs(q, None, None, excl, K, kc)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
L = 6
bufs = [torch.zeros((sms + 1) * 8, dtype=torch.int64, device=dev) for _ in range(L)]
fbufs = [torch.zeros(16, dtype=torch.int64, device=dev) for _ in range(L)]
for _ in range(20): sc.enqueue(K, kc)
dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(L):
    os.environ["REBERT_GEMV_TRACE"] = hex(bufs[i].data_ptr()); os.environ["REBERT_FIN_TRACE"] = hex(fbufs[i].data_ptr())
    sc.enqueue(K, kc)
b.record(); torch.cuda.synchronize()
os.environ.pop("REBERT_GEMV_TRACE"); os.environ.pop("REBERT_FIN_TRACE")
rows = [{"rank": rank, "rows_per_shard": store.n, "event_us_per_request": round(a.elapsed_time(b) / L * 1e3, 2)}]
prev_end = None
for i in range(L):
    t = bufs[i].cpu().numpy().reshape(sms + 1, 8); ft = fbufs[i].cpu().numpy()
    t0 = t[:sms, 0].min()
    r = {"launch": i, "entry_first_last": [0.0, round(float(t[:sms, 0].max() - t0) / 1e3, 2)],
         "first_tile_med": round(float(np.median(t[:sms, 2]) - t0) / 1e3, 2), "stream_end_med_max": [round(float(np.median(t[:sms, 4]) - t0) / 1e3, 2), round(float(t[:sms, 4].max() - t0) / 1e3, 2)],
         "published_max": round(float(t[:sms, 5].max() - t0) / 1e3, 2)}
    for j, nm in enumerate(["c_entry", "c_stream_complete", "c_winners", "c_exact", "c_ranked", "c_end_after_exchange"]):
        r[nm] = round(float(ft[j] - t0) / 1e3, 2)
    if prev_end is not None:
        r["gap_from_prev_end_us"] = round(float(t0 - prev_end) / 1e3, 2)
    prev_end = ft[5]
    rows.append(r)
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/trace_sharded_rank{rank}.jsonl", "w") as fh:
    for r in rows: fh.write(json.dumps(r) + "\n")
dist.barrier(); dist.destroy_process_group()
