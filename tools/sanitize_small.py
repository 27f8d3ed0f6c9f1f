"""Small end-to-end run for compute-sanitizer (memcheck): every kernel family once, tiny sizes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, RowFilter, synth
for dtype, n, d in [("fp32", 3000, 32), ("bf16", 5000, 1536), ("bf16", 4000, 50), ("fp32", 2000, 2200)]:
    st = CatalogStore.synthetic(0, n, d, dtype, scale_rows=True)
    q = synth.query_f32(1, d)
    ex = np.random.default_rng(0).choice(n, 100, replace=False)
    print(dtype, n, d, st.recommend(query=q, exclude_rows=ex, k=10)[0][:3], st.recommend(liked_rows=ex[:20], exclude_rows=ex, k=100)[0][:3])
st = CatalogStore.synthetic(0, 20000, 128, "bf16")
g, y = synth.movie_metadata(3, 0, 20000); st.set_metadata(g, y)
print(st.recommend(query=synth.query_f32(1, 128), k=50, row_filter=RowFilter(genre_any=3, year_lo=1950, year_hi=2000))[0][:3])
qs = synth.catalog_rows_f32(11, 0, 200, 128)
r = st.recommend_batch(queries=qs, k=10)
print("batch", r[0][:2, :3])
_, p64, _ = st.build_profiles(np.array([0, 3]), np.array([1, 2, 3]))
print(st.score_subset(p64, np.array([5, 6, 7])))
torch.cuda.synchronize()
print("sanitize run ok")
