"""Sharding layer host logic on CPU: ShardPlan arithmetic and a world_size-2 gloo run of ShardedCatalog with an
oracle-backed shard backend (tests only) — results must equal the unsharded oracle on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from robot_ebert_b200 import synth
from robot_ebert_b200.catalog import unpack_result
from robot_ebert_b200.sharding import ShardedCatalog, ShardPlan


def test_shard_plan_partitions_rows():
    for n, w in [(10, 3), (10_000_000, 8), (7, 8), (0, 2), (1_000_003, 4)]:
        plan = ShardPlan(n, w)
        ranges = [plan.range(g) for g in range(w)]
        assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n
        for g in range(w - 1):
            assert ranges[g][0] + ranges[g][1] == ranges[g + 1][0]
        assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1
        for row in ([0, n - 1, n // 2, n // 3] if n else []):
            g = plan.owner(row)
            assert ranges[g][0] <= row < ranges[g][0] + ranges[g][1]
    with pytest.raises(ValueError):
        ShardPlan(10, 0)


class OracleShardBackend:
    """Test double with the CudaShardBackend interface; arithmetic by the oracle on CPU tensors."""

    def __init__(self, m_local, row0):
        from oracle.reference_scoring import _normalize_rows
        self.unit = _normalize_rows(m_local)
        self.row0, self.n, self.d = row0, m_local.shape[0], m_local.shape[1]
        self.device = torch.device("cpu")

    def stage(self, query, liked_rows, weights, exclude_rows, k, kc):
        self.excl = None if exclude_rows is None else np.asarray(exclude_rows, dtype=np.int64)
        if liked_rows is None:
            q = np.asarray(query, dtype=np.float64)
            nrm = np.linalg.norm(q)
            self.vec = q / (nrm if nrm else 1.0)
            return None
        liked = np.asarray(liked_rows, dtype=np.int64)
        w = np.ones(len(liked)) if weights is None else np.asarray(weights, dtype=np.float64)
        mine = (liked >= self.row0) & (liked < self.row0 + self.n)
        part = (self.unit[liked[mine] - self.row0] * w[mine, None]).sum(axis=0)
        return torch.from_numpy(np.concatenate([part, [w.sum()]]))

    def set_profile(self, summed):
        s = summed.numpy()
        self.vec = s[:self.d] / (s[self.d] / self._world)

    def _allowed_scores(self):
        scores = self.unit @ self.vec
        ok = np.ones(self.n, dtype=bool)
        if self.excl is not None:
            ok[self.excl[(self.excl >= self.row0) & (self.excl < self.row0 + self.n)] - self.row0] = False
        return scores, ok

    def count_local(self, threshold, row_filter):
        scores, ok = self._allowed_scores()
        return int(np.count_nonzero(ok & (scores >= threshold)))

    def sweep_local(self, threshold, row_filter):
        scores, ok = self._allowed_scores()
        rows = np.nonzero(ok & (scores >= threshold))[0]
        return rows.astype(np.int64) + self.row0, scores[rows]

    def local_topk(self, k, kc, row_filter):
        from oracle.reference_scoring import topk_rows
        scores = self.unit @ self.vec
        excl = None
        if self.excl is not None:
            excl = self.excl[(self.excl >= self.row0) & (self.excl < self.row0 + self.n)] - self.row0
        rows, sc = topk_rows(scores, k, excl)
        words = np.zeros(2 * k + 2, dtype=np.int64)
        words[:len(rows)] = rows + self.row0
        words[len(rows):k] = -1
        words[k:k + len(rows)] = sc.view(np.int64)
        words[2 * k:2 * k + 1].view(np.int32)[0] = len(rows)
        words[2 * k + 1:].view(np.float64)[0] = np.inf
        return torch.from_numpy(words)

    def merge(self, gathered, k):
        g = gathered.numpy()
        rows, scores = [], []
        for l in range(g.shape[0]):
            r, s, _ = unpack_result(g[l], k)
            rows.append(r)
            scores.append(s)
        rows, scores = np.concatenate(rows), np.concatenate(scores)
        order = np.lexsort((rows, -scores))[:k]
        words = np.zeros(2 * k + 2, dtype=np.int64)
        words[:len(order)] = rows[order]
        words[k:k + len(order)] = scores[order].view(np.int64)
        words[2 * k:2 * k + 1].view(np.int32)[0] = len(order)
        words[2 * k + 1:].view(np.float64)[0] = np.inf
        return torch.from_numpy(words)

    def fetch(self, packed, k):
        return unpack_result(packed.numpy(), k)


def _worker(rank, world, port, n, d, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(n, world)
        row0, cnt = plan.range(rank)
        m_local = synth.catalog_rows_f32(0, row0, cnt, d, scale_rows=True).astype(np.float64)
        sc = ShardedCatalog(OracleShardBackend(m_local, row0), n)
        q = synth.query_f32(1, d)
        excl = np.random.default_rng(3).choice(n, size=40, replace=False)
        r1 = sc.recommend(query=q, exclude_rows=excl, k=10)
        (rated, rts), = synth.user_ratings(2, n, 1, mean_rated=60)
        liked = rated[rts >= 3.5]
        r2 = sc.recommend(liked_rows=liked, exclude_rows=rated, k=25)
        # k beyond the register-list kernel: sharded threshold bisection + sweep (query form, then profile form with
        # fewer allowed rows than k)
        r3 = sc.recommend(query=q, exclude_rows=excl, k=300)
        r4 = sc.recommend(liked_rows=liked, exclude_rows=np.setdiff1d(np.arange(n), np.arange(0, n, 11)), k=400)
        out_q.put((rank, r1[0].tolist(), r1[1].tolist(), r2[0].tolist(), r2[1].tolist(), r3[0].tolist(), r3[1].tolist(),
                   r4[0].tolist(), r4[1].tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_unsharded_oracle():
    from oracle import reference_scoring as ora
    n, d, world = 3001, 48, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out_q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = synth.catalog_rows_f32(0, 0, n, d, scale_rows=True).astype(np.float64)
    q = synth.query_f32(1, d).astype(np.float64)
    excl = np.random.default_rng(3).choice(n, size=40, replace=False)
    want1 = ora.query_rows(m, q, excl, 10)
    (rated, rts), = synth.user_ratings(2, n, 1, mean_rated=60)
    want2 = ora.recommend_rows(m, rated[rts >= 3.5], rated, 25)
    want3 = ora.query_rows(m, q, excl, 300)
    allowed4 = np.arange(0, n, 11)
    want4 = ora.recommend_rows(m, rated[rts >= 3.5], np.setdiff1d(np.arange(n), allowed4), 400)
    assert len(want3[0]) == 300 and len(want4[0]) == len(allowed4) < 400
    for rank, r1_rows, r1_scores, r2_rows, r2_scores, r3_rows, r3_scores, r4_rows, r4_scores in results:
        assert r3_rows == want3[0].tolist(), rank
        np.testing.assert_allclose(r3_scores, want3[1], rtol=1e-12)
        assert r4_rows == want4[0].tolist(), rank
        np.testing.assert_allclose(r4_scores, want4[1], rtol=1e-12)
        assert r1_rows == want1[0].tolist(), rank
        np.testing.assert_allclose(r1_scores, want1[1], rtol=1e-12)
        assert r2_rows == want2[0].tolist(), rank
        np.testing.assert_allclose(r2_scores, want2[1], rtol=1e-12)
