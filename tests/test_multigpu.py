"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2`): row-sharded result == single-GPU result, ids bit-exact."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_guarded(name, rank, world, port, *args):
    """Picklable worker entry (spawn): runs the named worker; an exception becomes a queue item, so the parent fails at
    once instead of waiting for a timeout."""
    import traceback
    out_q = args[-1]
    try:
        globals()[name](rank, world, port, *args)
    except BaseException:  # noqa: BLE001
        out_q.put((rank, "ERROR", traceback.format_exc()))
        raise


def _collect(out_q, procs, world, timeout=300):
    results = []
    for _ in range(world):
        item = out_q.get(timeout=timeout)
        if len(item) >= 2 and item[1] == "ERROR":
            for p in procs:
                p.kill()
            pytest.fail(f"worker {item[0]} raised:\n{item[2]}")
        results.append(item)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return results


def _worker(rank, world, port, n, d, dtype, out_q):
    import torch.distributed as dist
    from robot_ebert_b200 import synth
    from robot_ebert_b200.sharding import ShardedCatalog
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sc = ShardedCatalog.synthetic(0, n, d, dtype, scale_rows=True, device=torch.device("cuda", rank))
        q = synth.query_f32(1, d)
        excl = np.random.default_rng(3).choice(n, size=133, replace=False)
        r1 = sc.recommend(query=q, exclude_rows=excl, k=10)
        (rated, rts), = synth.user_ratings(2, n, 1)
        r2 = sc.recommend(liked_rows=rated[rts >= 3.5], exclude_rows=rated, k=50)
        r3 = sc.recommend(query=q, exclude_rows=excl, k=300)                 # k > 240: sharded bisection + sweep route
        r4 = sc.recommend(liked_rows=rated[rts >= 3.5], exclude_rows=rated, k=1000)
        out_q.put((rank, r1[0].tolist(), r1[1].tolist(), r2[0].tolist(), r2[1].tolist(), r3[0].tolist(), r3[1].tolist(),
                   r4[0].tolist(), r4[1].tolist()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_sharded_equals_single_gpu(dtype):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from robot_ebert_b200 import CatalogStore, synth
    n, d = 200_003, 1536
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_run_guarded, args=("_worker", r, world, port, n, d, dtype, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = _collect(out_q, procs, world)
    store = CatalogStore.synthetic(0, n, d, dtype, scale_rows=True, device="cuda:0")
    q = synth.query_f32(1, d)
    excl = np.random.default_rng(3).choice(n, size=133, replace=False)
    w1 = store.recommend(query=q, exclude_rows=excl, k=10)
    (rated, rts), = synth.user_ratings(2, n, 1)
    w2 = store.recommend(liked_rows=rated[rts >= 3.5], exclude_rows=rated, k=50)
    w3 = store.recommend(query=q, exclude_rows=excl, k=300)
    w4 = store.recommend(liked_rows=rated[rts >= 3.5], exclude_rows=rated, k=1000)
    assert len(w3[0]) == 300 and len(w4[0]) == 1000
    for rank, r1r, r1s, r2r, r2s, r3r, r3s, r4r, r4s in results:
        assert r1r == w1[0].tolist() and r2r == w2[0].tolist(), rank
        np.testing.assert_allclose(r1s, w1[1], rtol=1e-12)
        np.testing.assert_allclose(r2s, w2[1], rtol=1e-12)
        assert r3r == w3[0].tolist() and r4r == w4[0].tolist(), rank
        np.testing.assert_allclose(r3s, w3[1], rtol=1e-12)
        np.testing.assert_allclose(r4s, w4[1], rtol=1e-12)


def _batch_worker(rank, world, port, n, d, b, k, out_q):
    import torch.distributed as dist
    from robot_ebert_b200 import synth
    from robot_ebert_b200.sharding import ShardedCatalog
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sc = ShardedCatalog.synthetic(0, n, d, "bf16", scale_rows=True, device=torch.device("cuda", rank))
        lp, lc, ep, ec = _csr_users(n, b)
        from robot_ebert_b200 import RowFilter
        g, y = synth.movie_metadata(3, sc.backend.store.row_base, sc.backend.store.n)      # this shard's slice of the side columns
        sc.backend.store.set_metadata(g, y)
        r = sc.recommend_batch(liked_ptr=lp, liked_col=lc, excl_ptr=ep, excl_col=ec, k=k,
                               row_filter=RowFilter(genre_any=0b1011, year_lo=1960, year_hi=2000))
        out_q.put((rank, r[0].tolist(), r[1].tolist(), r[2].tolist()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _csr_users(n, b):
    from robot_ebert_b200 import synth
    lp, lc, ep, ec = [0], [], [0], []
    for rated, rts in synth.user_ratings(2, n, b):
        liked = rated[rts >= 3.5]
        if len(liked) == 0:
            liked = rated[:1]
        lc.append(liked); lp.append(lp[-1] + len(liked)); ec.append(rated); ep.append(ep[-1] + len(rated))
    return np.array(lp), np.concatenate(lc), np.array(ep), np.concatenate(ec)


def test_sharded_batch_equals_single_gpu_batch():
    """BASELINE config 5 shape at test size: CSR profiles + exclusions, sharded tcgen05 pass == one-GPU result."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from robot_ebert_b200 import CatalogStore
    n, d, b, k = 140_000, 512, 200, 50
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_run_guarded, args=("_batch_worker", r, world, port, n, d, b, k, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = _collect(out_q, procs, world)
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True, device="cuda:0")
    lp, lc, ep, ec = _csr_users(n, b)
    from robot_ebert_b200 import RowFilter, synth
    g, y = synth.movie_metadata(3, 0, n)
    store.set_metadata(g, y)
    want = store.recommend_batch(liked_ptr=lp, liked_col=lc, excl_ptr=ep, excl_col=ec, k=k,
                                 row_filter=RowFilter(genre_any=0b1011, year_lo=1960, year_hi=2000))
    keep = ((g & 0b1011) != 0) & (y >= 1960) & (y <= 2000)
    assert keep[want[0][want[0] >= 0]].all()
    for rank, rows, scores, counts in results:
        assert rows == want[0].tolist() and counts == want[2].tolist(), rank
        np.testing.assert_allclose(scores, want[1], rtol=1e-12)


def _tie_worker(rank, world, port, out_q):
    import torch.distributed as dist
    from robot_ebert_b200 import CatalogStore
    from robot_ebert_b200.sharding import CudaShardBackend, ShardedCatalog, ShardPlan
    from robot_ebert_b200 import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n = 3000
        m = synth.catalog_rows_f32(0, 0, n, 1, scale_rows=True)                    # d = 1: every cosine is exactly +-1
        row0, cnt = ShardPlan(n, world).range(rank)
        store = CatalogStore.from_host(None, m[row0:row0 + cnt], "fp32", device=torch.device("cuda", rank), row_base=row0)
        backend = CudaShardBackend(store)
        backend.setup_p2p()
        sc = ShardedCatalog(backend, n)
        rows, scores, info = sc.recommend(query=np.ones(1, dtype=np.float32), exclude_rows=np.arange(0, n, 7), k=10, return_info=True)
        out_q.put((rank, rows.tolist(), scores.tolist(), bool(info["proven_exact"])))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_mass_ties_use_the_exact_sweep():
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import reference_scoring as ora
    from robot_ebert_b200 import synth
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_run_guarded, args=("_tie_worker", r, world, port, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = _collect(out_q, procs, world)
    m = synth.catalog_rows_f32(0, 0, 3000, 1, scale_rows=True).astype(np.float64)
    want_rows, want_scores = ora.query_rows(m, np.ones(1), np.arange(0, 3000, 7), 10)
    for rank, rows, scores, proven in results:
        assert proven and rows == want_rows.tolist(), rank
        np.testing.assert_allclose(scores, want_scores, rtol=1e-12)


def test_sharded_batch_small_shards_fall_back_to_single_query_path():
    """Shards below the tensor-core path's minimum size: recommend_batch must still answer (per-query sharded path)."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from robot_ebert_b200 import CatalogStore, RowFilter, synth
    n, d, b, k = 9_000 * world, 128, 6, 10
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_run_guarded, args=("_batch_worker", r, world, port, n, d, b, k, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = _collect(out_q, procs, world)
    store = CatalogStore.synthetic(0, n, d, "bf16", scale_rows=True, device="cuda:0")
    lp, lc, ep, ec = _csr_users(n, b)
    g, y = synth.movie_metadata(3, 0, n)
    store.set_metadata(g, y)
    want = store.recommend_batch(liked_ptr=lp, liked_col=lc, excl_ptr=ep, excl_col=ec, k=k,
                                 row_filter=RowFilter(genre_any=0b1011, year_lo=1960, year_hi=2000))
    for rank, rows, scores, counts in results:
        assert rows == want[0].tolist() and counts == want[2].tolist(), rank
        np.testing.assert_allclose(scores, want[1], rtol=1e-12)
