#!/usr/bin/env python
"""Where does a GEMV launch spend its time?  Per-CTA %globaltimer stamps (REBERT_GEMV_TRACE = device buffer address):
0 kernel entry, 1 prologue done, 2 first tile landed, 3 warp 0 finished its last tile, 4 all warps finished,
5 CTA keys published; extra row: end of the last CTA's merge (rebert_gemv_topk form).  With `fused` the request path is
traced instead (rebert_recommend_device) and the cluster kernel behind the streaming kernel adds its own stamps
(REBERT_FIN_TRACE): 0 entry, 1 streaming kernel complete, 2 winners selected, 3 exact scores in CTA 0, 4 ranked, 5 end.
Prints the spread of every stamp relative to the earliest kernel entry, for a few launches back to back, beside the
CUDA-event time of the same launches.   trace_gemv.py <rows> <k> [i8] [fused] [dim=D] [fp32]"""
import ctypes as C, json, os, sys
os.environ["REBERT_GEMV_TUNE"] = "1"          # make the library re-read its knobs at every launch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth, _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
I8 = "i8" in sys.argv[3:]                                  # trace the int8 prefilter shadow's fast pass (kc = 256)
FUSED = "fused" in sys.argv[3:]                            # the one-launch request path (exact pass in the kernel's tail)
lib = nat.load(); dev = torch.device("cuda:0")
D = next((int(a[4:]) for a in sys.argv[3:] if a.startswith("dim=")), 1536)
DT = "fp32" if "fp32" in sys.argv[3:] else "bf16"
store = CatalogStore.synthetic(0, n, D, DT, device=dev)
q = synth.query_f32(1, D); excl = np.random.default_rng(1).choice(n, size=min(133, n // 4), replace=False)
kc = 256 if I8 else lib.rebert_candidates_for_k(K)
if I8:
    store.enable_prefilter()
cat = store._c8 if I8 else store._c
ptr, ne = store.stage_inputs(q, None, None, excl, K, kc)
s = store._scratch(); f = nat.Filter(); f.exclude_rows, f.n_exclude = ptr, ne
st = torch.cuda.current_stream().cuda_stream
def gemv():
    if FUSED:
        store.enqueue_fused(K, kc, ptr, ne, None, prefilter=I8)
    else:
        nat.check(lib.rebert_gemv_topk(C.byref(cat), s.qn32.data_ptr(), C.byref(f), kc, s.ws.data_ptr(), s.ws.numel(), s.cand.data_ptr(), st))
sms = torch.cuda.get_device_properties(dev).multi_processor_count
L = 6
bufs = [torch.zeros((sms + 1) * 8, dtype=torch.int64, device=dev) for _ in range(L)]
fbufs = [torch.zeros(16, dtype=torch.int64, device=dev) for _ in range(L)]
for _ in range(20): gemv()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(L):
    os.environ["REBERT_GEMV_TRACE"] = hex(bufs[i].data_ptr())
    os.environ["REBERT_FIN_TRACE"] = hex(fbufs[i].data_ptr())
    gemv()
b.record(); torch.cuda.synchronize()
os.environ.pop("REBERT_GEMV_TRACE")
os.environ.pop("REBERT_FIN_TRACE")
ob = s.d_out.data_ptr()
def fin(): nat.check(lib.rebert_finalize_topk(C.byref(store._c), s.qn64.data_ptr(), s.cand.data_ptr(), kc, K, ob, ob + 8*K, ob + 16*K, ob + 16*K + 8, st))
fin(); torch.cuda.synchronize()
fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fa.record()
for _ in range(20): fin()
fb.record(); torch.cuda.synchronize()
print(json.dumps({"finalize_us_at_kc": kc, "us": round(fa.elapsed_time(fb) / 20 * 1e3, 2)}))
print(json.dumps({"rows": n, "k": K, "i8": I8, "fused": FUSED, "event_us_per_launch": round(a.elapsed_time(b) / L * 1e3, 2)}))
names = ["entry", "prologue_done", "first_tile", "warp0_done", "all_warps_done", "keys_published"]
fnames = ["entry", "stream_complete", "winners_selected", "exact_in_cta0", "ranked", "end"]
prev_end = None
for i in range(L):
    t = bufs[i].cpu().numpy().reshape(sms + 1, 8)
    live = np.nonzero(t[:sms, 0])[0]                      # a small catalog launches fewer CTAs than SMs
    t0 = t[live, 0].min()
    row = {"launch": i, "ctas": int(len(live))}
    if prev_end is not None:
        row["gap_from_prev_end_us"] = round((t0 - prev_end) / 1e3, 2)
    for j, nm in enumerate(names):
        v = (t[live, j] - t0) / 1e3
        row[nm] = [round(float(v.min()), 2), round(float(np.median(v)), 2), round(float(v.max()), 2)]
    slow = live[np.argmax(t[live, 5])]
    row["slowest_cta"] = {"block": int(slow), "stamps": [round(float(x - t0) / 1e3, 2) for x in t[slow, :6]]}
    if FUSED:
        ft = fbufs[i].cpu().numpy()
        row["cluster_kernel"] = {nm: round(float(ft[j] - t0) / 1e3, 2) for j, nm in enumerate(fnames)}
        row["cluster_detail"] = {"heads_tails_landed": round(float(ft[8] - t0) / 1e3, 2),
                                 "threshold_known": round(float(ft[9] - t0) / 1e3, 2), "survivors_gathered": round(float(ft[10] - t0) / 1e3, 2),
                                 "cta0_exact_done": round(float(ft[7] - t0) / 1e3, 2)}
        end = ft[5]
    else:
        row["final_merge_end"] = round(float(t[sms, 0] - t0) / 1e3, 2)
        end = t[sms, 0]
    prev_end = end
    print(json.dumps(row))
