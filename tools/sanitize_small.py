"""Small end-to-end run of every kernel family on the ASSERTION build (REBERT_DEBUG=1 -> librebert_b200_debug.so, device-side
bounds checks at every computed index; stands in for compute-sanitizer, which the GPU pool does not offer).  Also usable
under compute-sanitizer memcheck where that exists.  A failed assertion traps the kernel and the run dies with a CUDA error."""
import os, sys
os.environ.setdefault("REBERT_DEBUG", "1")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_ebert_b200 import CatalogStore, RowFilter, synth, _native as nat
assert nat.LIB_PATH.endswith("_debug.so") or os.environ["REBERT_DEBUG"] != "1", nat.LIB_PATH
for dtype, n, d in [("fp32", 3000, 32), ("bf16", 5000, 1536), ("bf16", 4000, 50), ("fp32", 2000, 2200), ("fp32", 3, 32), ("bf16", 70001, 256)]:
    st = CatalogStore.synthetic(0, n, d, dtype, scale_rows=True)
    q = synth.query_f32(1, d)
    ex = np.random.default_rng(0).choice(n, min(100, n // 2), replace=False)
    for k in (1, 10, 100, 240):
        r1 = st.recommend(query=q, exclude_rows=ex, k=k)
        r2 = st.recommend(liked_rows=ex[:20], exclude_rows=ex, k=k)
    if n >= 3000:
        st.enable_prefilter()
        r3 = st.recommend(query=q, exclude_rows=ex, k=10, prefilter=True)
        assert np.array_equal(r3[0], st.recommend(query=q, exclude_rows=ex, k=10, prefilter=False)[0])
    print(dtype, n, d, r1[0][:3], r2[0][:3])
st = CatalogStore.synthetic(0, 20000, 128, "bf16")
g, y = synth.movie_metadata(3, 0, 20000); st.set_metadata(g, y)
print(st.recommend(query=synth.query_f32(1, 128), k=50, row_filter=RowFilter(genre_any=3, year_lo=1950, year_hi=2000))[0][:3])
qs = synth.catalog_rows_f32(11, 0, 200, 128)
r = st.recommend_batch(queries=qs, k=10)
print("batch", r[0][:2, :3])
big = CatalogStore.synthetic(0, 150_000, 256, "bf16", scale_rows=True)
big.enable_prefilter()
qb = synth.catalog_rows_f32(11, 0, 300, 256)
a = big.recommend_batch(queries=qb, k=50, prefilter=True)
b = big.recommend_batch(queries=qb, k=50, prefilter=False)
assert np.array_equal(a[0], b[0])
print("batch int8 == bf16", a[0][:1, :3])
_, p64, _ = st.build_profiles(np.array([0, 3]), np.array([1, 2, 3]))
print(st.score_subset(p64, np.array([5, 6, 7])))
torch.cuda.synchronize()
print("sanitize run ok (assertion build)" if os.environ["REBERT_DEBUG"] == "1" else "sanitize run ok")
