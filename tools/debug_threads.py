#!/usr/bin/env python
"""Debug aid: N ranks x T threads of sharded requests on their own channels, one log line per request.
   torchrun --nproc-per-node 2 tools/debug_threads.py [threads] [per_thread]"""
import json, os, sys, threading, time, traceback
os.environ.setdefault("REBERT_EXCHANGE_TIMEOUT_MS", "3000")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from robot_ebert_b200 import synth
from robot_ebert_b200.sharding import ShardedCatalog
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_multigpu import _requests

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
P = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, d = 120_001, 256
sc = ShardedCatalog.synthetic(0, n, d, "bf16", scale_rows=True, device=torch.device("cuda", lr))
reqs = _requests(n, d, T * P)
log, lock = [], threading.Lock()

def serve(t):
    torch.cuda.set_device(lr)
    with torch.cuda.stream(torch.cuda.Stream()):
        for j in range(P):
            i = t * P + j
            kind = "query" if "query" in reqs[i] else ("wprofile" if "weights" in reqs[i] else "profile")
            seq0 = sc.backend._seq[t]
            try:
                r, s, info = sc.recommend(channel=t, return_info=True, **reqs[i])
                rec = {"rank": rank, "t": t, "i": i, "kind": kind, "seq_before": seq0, "seq_after": sc.backend._seq[t],
                       "attempts": info.get("attempts"), "kc": info["kc"], "margin": info["margin"], "top": int(r[0])}
            except BaseException as e:  # noqa: BLE001
                rec = {"rank": rank, "t": t, "i": i, "kind": kind, "seq_before": seq0, "seq_after": sc.backend._seq[t], "error": repr(e)[:160]}
                with lock:
                    log.append(rec)
                return
            with lock:
                log.append(rec)

th = [threading.Thread(target=serve, args=(t,)) for t in range(T)]
for x in th: x.start()
for x in th: x.join()
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/debug_threads_rank{rank}.jsonl", "w") as fh:
    for rec in log:
        fh.write(json.dumps(rec) + "\n")
time.sleep(1)
os._exit(0)
