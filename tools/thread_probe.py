import os, sys, threading, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, synth
n, d = 60_000, 256
store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True)
users = synth.user_ratings(2, n, 8)
base = [store.recommend(liked_rows=r[x >= 3.5], exclude_rows=r, k=10) for r, x in users]
errors = []
def worker(u):
    try:
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for it in range(25):
                r, x = users[u]
                rows, scores = store.recommend(liked_rows=r[x >= 3.5], exclude_rows=r, k=10)
                if not np.array_equal(rows, base[u][0]):
                    errors.append((u, it, "rows differ", rows.tolist(), base[u][0].tolist()))
                    return
    except Exception as e:
        errors.append((u, traceback.format_exc()))
ts = [threading.Thread(target=worker, args=(u,)) for u in range(8)]
[t.start() for t in ts]; [t.join() for t in ts]
for e in errors[:4]: print(e)
print("errors:", len(errors))
