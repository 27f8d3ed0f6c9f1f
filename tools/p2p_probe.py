import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from robot_ebert_b200 import synth
from robot_ebert_b200.sharding import ShardedCatalog
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
sc = ShardedCatalog.synthetic(0, 200_000, 256, "bf16", device=torch.device("cuda", rank))
q = synth.query_f32(1, 256)
r = sc.recommend(query=q, k=10)
s = sc.backend.store._scratch()
K = 10
def timed(fn, n=200):
    for _ in range(10): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
t_x = timed(lambda: sc.backend.exchange_merge(s.d_out, K))
buf = sc._gather_buf(K, s.d_out)
def nccl():
    dist.all_gather_into_tensor(buf.view(-1), s.d_out); sc.backend.merge(buf, K)
t_n = timed(nccl)
print(rank, f"exchange_merge {t_x:.1f} us   nccl all_gather+merge {t_n:.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
