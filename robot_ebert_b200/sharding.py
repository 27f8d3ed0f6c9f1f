"""Row-sharding layer: one process per GPU, catalog split row-wise, local top-k + all-gather + merge.

SURVEY.md §8(e): score(j) depends only on catalog row j and the replicated query, so rank g owns the contiguous
rows [g*N/G, (g+1)*N/G) and the only exchange step is an all-gather of each rank's k best (row, score) pairs
(k*16+16 bytes per rank), followed by the same (score desc, row asc) merge on every rank — so every rank returns
the result a single GPU would.  Building a user profile from liked rows scattered over the shards adds one
all-reduce(sum) of the fp64 partial profile [ld+1].

The collective plumbing is torch.distributed (NCCL on GPUs).  The arithmetic is behind a small backend object so
the host logic can be exercised with gloo on CPU in tests; the product backend (`CudaShardBackend`) is CUDA-only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _native as nat
from .catalog import CatalogStore, RowFilter, _on_device, large_k_search, sorted_csr, sorted_unique_i32, unpack_result


class ShardPlan:
    """Contiguous, balanced row ranges; rank g owns [start(g), start(g+1))."""

    def __init__(self, n_total: int, world: int):
        if n_total < 0 or world <= 0:
            raise ValueError("bad shard plan")
        self.n_total, self.world = int(n_total), int(world)

    def start(self, g: int) -> int:
        return (self.n_total * g) // self.world

    def range(self, g: int) -> Tuple[int, int]:
        """(row0, nrows) of rank g."""
        return self.start(g), self.start(g + 1) - self.start(g)

    def owner(self, row: int) -> int:
        if not 0 <= row < self.n_total:
            raise ValueError("row out of range")
        g = min(self.world - 1, (row * self.world) // max(self.n_total, 1))
        while row < self.start(g):
            g -= 1
        while row >= self.start(g + 1):
            g += 1
        return g


class CudaShardBackend:
    """Product backend: the local shard is a CatalogStore; every step is a kernel enqueued on the current stream."""

    K_MAX = 240          # largest k of the single-query path

    def __init__(self, store: CatalogStore):
        self.store = store
        self.device = store.device
        self._merged = {}
        self._host = {}
        self.exchange = "nccl"       # becomes "p2p" once setup_p2p() has mapped the peers' buffers
        self._seq = 0

    def setup_p2p(self, group=None) -> bool:
        """Map one small symmetric buffer per rank (torch symmetric memory = CUDA VMM + fabric handles) so that the
        exchange step becomes ONE kernel of P2P stores + flags + merge.  Falls back to NCCL all-gather if the
        platform cannot provide peer mappings.  Collective: every rank must call it."""
        lib = nat.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ok = torch.zeros(1, dtype=torch.int32, device=self.device)
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = lib.rebert_exchange_buffer_bytes(world, self.K_MAX)
            if nbytes == 0:
                raise RuntimeError("world size not supported by the exchange kernel")
            buf = symm.empty(nbytes // 8, dtype=torch.int64, device=self.device)
            buf.zero_()
            hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            self._symm_buf, self._symm_hdl = buf, hdl
            self._peer_ptrs = (C.c_uint64 * world)(*[int(p) for p in hdl.buffer_ptrs])
            self._p2p_out = {}
            self._p2p_world, self._p2p_rank = world, rank
            ok.fill_(1)
        except Exception as e:  # noqa: BLE001 - any failure of the plumbing means: keep NCCL
            self._p2p_error = repr(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all or nothing, and a barrier after the zero-fill
        torch.cuda.synchronize(self.device)
        self.exchange = "p2p" if int(ok.item()) == 1 else "nccl"
        return self.exchange == "p2p"

    def exchange_merge(self, local: torch.Tensor, k: int) -> torch.Tensor:
        """P2P exchange + merge of the packed local result (int64 [2k+2]) -> packed merged result (+1 word: error flag)."""
        lib = nat.load()
        out = self._p2p_out.get(k)
        if out is None:
            out = self._p2p_out[k] = torch.zeros(2 * k + 3, dtype=torch.int64, device=self.device)   # last word: peer-timeout flag
        self._seq += 1
        nat.check(lib.rebert_exchange_merge(self._peer_ptrs, self._p2p_world, self._p2p_rank, k, self.K_MAX,
                                            self._seq & 0xFFFFFFFF or 1, local.data_ptr(), out.data_ptr(),
                                            out.data_ptr() + 8 * (2 * k + 2), torch.cuda.current_stream().cuda_stream))
        return out

    def recommend_query_host(self, query, exclude_rows, k: int, kc: int, row_filter: Optional[RowFilter]):
        """One C call for a query request on the P2P path (rebert_recommend_host_sharded): zero-copy request, local fast +
        exact pass, fused exchange + merge, result written into pinned memory, one stream sync.  Collective in effect:
        every rank must call it with the same arguments.  Returns (rows, scores, margin)."""
        lib = nat.load()
        st = self.store
        s = st._scratch()
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.shape != (st.d,):
            raise ValueError(f"query must have shape ({st.d},)")
        ex, ne = None, 0
        if exclude_rows is not None and len(exclude_rows):
            ex = sorted_unique_i32(exclude_rows)
            ne = int(ex.shape[0])
        s.ensure_host_sharded(ne, k)
        f = st._filter_struct(row_filter)
        cnt, margin = C.c_int32(0), C.c_double(0.0)
        self._seq += 1
        with _on_device(self.device):
            rc = lib.rebert_recommend_host_sharded(
                C.byref(st._c), q.ctypes.data, None if ex is None else ex.ctypes.data, ne, None if f is None else C.byref(f), k, kc,
                s.sne_cap, s.shpin.data_ptr(), s.shpin.numel(), s.shdev.data_ptr(), s.shdev.numel(), self._peer_ptrs, self._p2p_world,
                self._p2p_rank, self.K_MAX, self._seq & 0xFFFFFFFF or 1, s.sh_rows.ctypes.data, s.sh_scores.ctypes.data,
                C.byref(cnt), C.byref(margin), torch.cuda.current_stream().cuda_stream)
        nat.check(rc)
        st.last_h2d_bytes = 4 * st.d + 4 * ne            # read by the staging kernel straight from the pinned block
        n = cnt.value
        return s.sh_rows[:n].copy(), s.sh_scores[:n].copy(), margin.value

    def stage(self, query, liked_rows, weights, exclude_rows, k, kc):
        st = self.store
        self._excl = st.stage_inputs(query, liked_rows, weights, exclude_rows, k, kc, profile_partial_only=True)
        if liked_rows is None:
            return None
        s = st._scratch()
        return torch.cat([s.sum64, s.wsum])                      # fp64 [ld + 1] partial profile

    def set_profile(self, summed: torch.Tensor):
        st = self.store
        s = st._scratch()
        s.sum64.copy_(summed[:st.ld])
        # wsum is summed over ranks too, but every rank added the full weight sum: divide it back
        s.wsum.copy_(summed[st.ld:] / self._world)
        st.finalize_profile()

    def sweep_local(self, threshold: float, row_filter: Optional[RowFilter]):
        ptr, ne = self._excl
        return self.store.sweep_above(threshold, ptr, ne, row_filter)

    def count_local(self, threshold: float, row_filter: Optional[RowFilter]) -> int:
        ptr, ne = self._excl
        return self.store.count_above(threshold, ptr, ne, row_filter)

    def local_topk(self, k: int, kc: int, row_filter: Optional[RowFilter]) -> torch.Tensor:
        ptr, ne = self._excl
        return self.store.enqueue_topk(k, kc, ptr, ne, row_filter)

    def merge(self, gathered: torch.Tensor, k: int) -> torch.Tensor:
        """gathered int64 [G, 2k+2] packed per-rank results -> packed merged result int64 [2k+2] (margin = min)."""
        lib = nat.load()
        g = gathered.shape[0]
        out = self._merged.get(k)
        if out is None:
            out = self._merged[k] = torch.empty(2 * k + 2, dtype=torch.int64, device=self.device)
        base, w = gathered.data_ptr(), 2 * k + 2
        ob = out.data_ptr()
        nat.check(lib.rebert_merge_topk(base, base + 8 * k, base + 16 * k, w, w, 2 * w, g, 1, k, ob, ob + 8 * k, ob + 16 * k,
                                        torch.cuda.current_stream().cuda_stream))
        out[2 * k + 1:2 * k + 2].view(torch.float64).copy_(gathered[:, 2 * k + 1].view(torch.float64).min().reshape(1))
        return out

    def fetch(self, packed: torch.Tensor, k: int):
        """ONE pinned D2H copy of the packed result (and, on the P2P path, its trailing error word), then a stream sync."""
        n = packed.numel()
        host = self._host.get(n)
        if host is None:
            host = self._host[n] = torch.empty(n, dtype=torch.int64).pin_memory()
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        words = host.numpy()
        if n == 2 * k + 3 and int(words[2 * k + 2:].view(np.int32)[0]) != 0:
            raise RuntimeError(f"peer {int(words[2 * k + 2:].view(np.int32)[0]) - 1} did not deliver its result to the exchange kernel")
        return unpack_result(words, k)


class ShardedCatalog:
    """The sharding layer.  Every rank calls recommend() with the same arguments and gets the same answer."""

    def __init__(self, backend, n_total: int, group=None):
        self.backend = backend
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.plan = ShardPlan(n_total, self.world)
        backend._world = self.world
        self._gather = {}

    @classmethod
    def synthetic(cls, seed: int, n_total: int, d: int, dtype: str = "bf16", scale_rows: bool = False, device=None,
                  group=None) -> "ShardedCatalog":
        plan = ShardPlan(n_total, dist.get_world_size(group))
        row0, n = plan.range(dist.get_rank(group))
        store = CatalogStore.synthetic(seed, n, d, dtype, scale_rows=scale_rows, device=device, row0=row0)
        backend = CudaShardBackend(store)
        backend.setup_p2p(group)
        return cls(backend, n_total, group)

    def _gather_buf(self, k: int, like: torch.Tensor) -> torch.Tensor:
        buf = self._gather.get(k)
        if buf is None:
            buf = self._gather[k] = torch.empty((self.world, 2 * k + 2), dtype=torch.int64, device=like.device)
        return buf

    def enqueue(self, k: int, kc: int, row_filter=None) -> torch.Tensor:
        """Device-resident step for an already staged query: local top-k -> all-gather -> merge (packed result)."""
        local = self.backend.local_topk(k, kc, row_filter)
        if getattr(self.backend, "exchange", "nccl") == "p2p":
            return self.backend.exchange_merge(local, k)             # one kernel: P2P stores + flags + merge
        buf = self._gather_buf(k, local)
        dist.all_gather_into_tensor(buf.view(-1), local, group=self.group)
        return self.backend.merge(buf, k)

    def recommend(self, *, query=None, liked_rows=None, weights=None, exclude_rows=None, k: int = 10, row_filter=None,
                  return_info: bool = False):
        if (query is None) == (liked_rows is None):
            raise ValueError("pass exactly one of query / liked_rows")
        lib = nat.load()
        kc = lib.rebert_candidates_for_k(k)
        eps = getattr(getattr(self.backend, "store", None), "fast_eps", 0.0)
        if kc == 0:
            # k beyond the register-list kernel (k > 240): sharded threshold bisection + sweep, still exact (slower route)
            rows, scores = self._recommend_large_k(query, liked_rows, weights, exclude_rows, k, row_filter, eps)
            if return_info:
                return rows, scores, {"kc": 0, "margin": float("inf"), "proven_exact": True, "exact_sweep": True}
            return rows, scores
        fast_host = query is not None and getattr(self.backend, "exchange", "nccl") == "p2p"
        while True:
            if fast_host:
                # query request on the P2P path: the whole step is ONE C call per rank (no torch ops, no NCCL)
                rows, scores, margin = self.backend.recommend_query_host(query, exclude_rows, k, kc, row_filter)
            else:
                partial = self.backend.stage(query, liked_rows, weights, exclude_rows, k, kc)
                if partial is not None:
                    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.group)
                    self.backend.set_profile(partial)
                packed = self.enqueue(k, kc, row_filter)
                rows, scores, margin = self.backend.fetch(packed, k)
            if margin > eps or kc >= 256:
                break
            kc = min(256, kc * 4)
        proven, swept = margin > eps, False
        if not proven and len(rows) == k and hasattr(self.backend, "sweep_local"):
            # mass ties (see CatalogStore._exact_sweep): every rank sweeps its shard against the global k-th score
            if fast_host:
                self.backend.stage(query, None, None, exclude_rows, k, kc)     # the sweep reads the query from torch scratch
            res = self._exact_sweep(float(scores[k - 1]) - 2.0 * eps, k, row_filter)
            if res is not None:
                rows, scores = res
                proven = swept = True
        if return_info:
            return rows, scores, {"kc": kc, "margin": margin, "proven_exact": proven, "exact_sweep": swept}
        return rows, scores

    def _recommend_large_k(self, query, liked_rows, weights, exclude_rows, k, row_filter, eps):
        """Sharded form of CatalogStore._recommend_large_k.  The bisection runs on the all-reduced count, so every rank
        takes the same branches and the collectives stay aligned; the survivors are all-gathered and ordered alike."""
        if liked_rows is not None and len(liked_rows) == 0:
            raise ValueError("Found array with 0 sample(s): user has no liked movies in the catalog")
        if k > CatalogStore.SWEEP_CAP // 2:
            raise ValueError(f"k={k} is too large (limit {CatalogStore.SWEEP_CAP // 2})")
        partial = self.backend.stage(query, liked_rows, weights, exclude_rows, 1, 32)
        if partial is not None:
            dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.group)
            self.backend.set_profile(partial)
        dev = self.backend.device

        def count(thr):
            c = torch.tensor([self.backend.count_local(thr, row_filter)], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.group)
            return int(c.item())

        return large_k_search(count, lambda thr: self._sweep_gather(thr, row_filter), k, eps)

    def _exact_sweep(self, threshold: float, k: int, row_filter):
        """Sharded form of the exact fallback: local sweeps, one all-gather of the (padded) survivors, same order on every rank."""
        res = self._sweep_gather(threshold, row_filter)
        if res is None or len(res[0]) < k:
            return None
        rr, sc = res
        order = np.lexsort((rr, -sc))[:k]
        return rr[order], sc[order]

    def _sweep_gather(self, threshold: float, row_filter):
        """(global rows, exact fp64 scores) of every allowed row of ANY shard whose fast score >= threshold, identical on
        every rank; None if some shard overflowed its sweep buffer."""
        local = self.backend.sweep_local(threshold, row_filter)
        dev = self.backend.device
        n_loc = torch.tensor([-1 if local is None else len(local[0])], dtype=torch.int64, device=dev)
        counts = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, n_loc, group=self.group)
        counts = counts.cpu().numpy()
        if (counts < 0).any():
            return None
        cap = int(counts.max())
        if cap == 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float64)
        mine = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
        if len(local[0]):
            mine[:len(local[0])] = torch.from_numpy(np.ascontiguousarray(local[0], dtype=np.int64)).to(dev)
            mine[cap:cap + len(local[0])] = torch.from_numpy(np.ascontiguousarray(local[1], dtype=np.float64).view(np.int64)).to(dev)
        allr = torch.empty((self.world, 2 * cap), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allr.view(-1), mine, group=self.group)
        allr = allr.cpu().numpy()
        rr = np.concatenate([allr[g, :counts[g]] for g in range(self.world)])
        sc = np.concatenate([allr[g, cap:cap + counts[g]].view(np.float64) for g in range(self.world)])
        return rr, sc

    # ------------------------------------------------------------------ batched (tensor-core) path ----------
    def batch_context(self, qbf: torch.Tensor, qn64: torch.Tensor, k: int, excl_ptr=None, excl_col=None, row_filter=None):
        """Allocate everything a batched step needs (device tensors only) for prepared queries [b, ld]."""
        import ctypes as C
        store: CatalogStore = self.backend.store
        lib = nat.load()
        dev = store.device
        b = qbf.shape[0]
        plan = store.gemm_plan(b, k)
        ep = ec = None
        if excl_ptr is not None:
            excl_ptr, excl_col = sorted_csr(excl_ptr, excl_col)
            ep = torch.from_numpy(excl_ptr).to(dev)
            ec = torch.from_numpy(excl_col).to(dev)
        hb = (b + 1) // 2
        words = 2 * b * k + 2 * hb                                       # rows | scores | counts | status
        ctx = {"b": b, "k": k, "plan": plan, "row_filter": row_filter, "qbf": qbf, "qn64": qn64, "ep": ep, "ec": ec, "hb": hb, "words": words,
               "ws": torch.empty(lib.rebert_gemm_workspace_bytes(C.byref(store._c), C.byref(plan)), dtype=torch.uint8, device=dev),
               "local": torch.empty(words, dtype=torch.int64, device=dev),
               "gathered": torch.empty((self.world, words), dtype=torch.int64, device=dev),
               "m_rows": torch.empty((b, k), dtype=torch.int64, device=dev),
               "m_scores": torch.empty((b, k), dtype=torch.float64, device=dev),
               "m_count": torch.empty(b, dtype=torch.int32, device=dev)}
        return ctx

    def batch_step(self, ctx) -> None:
        """Device-resident batched step: per-rank tcgen05 pass -> ONE packed all-gather -> merge.  No host sync."""
        store: CatalogStore = self.backend.store
        lib = nat.load()
        b, k, hb, words, local = ctx["b"], ctx["k"], ctx["hb"], ctx["words"], ctx["local"]
        o_rows, o_scores = local[:b * k], local[b * k:2 * b * k].view(torch.float64)
        o_count = local[2 * b * k:2 * b * k + hb].view(torch.int32)
        o_status = local[2 * b * k + hb:].view(torch.int32)
        store.enqueue_batch(ctx["plan"], ctx["qbf"], ctx["qn64"], ctx["ep"], ctx["ec"], ctx["ws"], o_rows, o_scores, o_count, o_status,
                            ctx.get("row_filter"))
        dist.all_gather_into_tensor(ctx["gathered"].view(-1), local, group=self.group)
        base = ctx["gathered"].data_ptr()
        nat.check(lib.rebert_merge_topk(base, base + 8 * b * k, base + 16 * b * k, words, words, 2 * words, self.world, b, k,
                                        ctx["m_rows"].data_ptr(), ctx["m_scores"].data_ptr(), ctx["m_count"].data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))

    def recommend_batch(self, *, queries=None, liked_ptr=None, liked_col=None, liked_w=None, excl_ptr=None, excl_col=None,
                        k: int = 10, row_filter=None, return_info: bool = False):
        """Sharded form of CatalogStore.recommend_batch: every rank runs the tcgen05 pass over its rows, the per-rank
        [b, k] results are all-gathered in ONE packed buffer and merged on every rank.  CUDA backend only."""
        store: CatalogStore = self.backend.store
        dev = store.device
        if (queries is None) == (liked_ptr is None):
            raise ValueError("pass exactly one of queries / liked CSR")
        with torch.cuda.device(dev):
            if queries is not None:
                qn32, qn64, qbf = store.prepare_queries(queries)
            else:
                lp = np.asarray(liked_ptr, dtype=np.int64)
                if np.any(np.diff(lp) == 0):
                    raise ValueError("Found array with 0 sample(s): a user has no liked movies in the catalog")
                qn32, qn64, qbf = store.build_profiles(
                    lp, liked_col, liked_w, reduce_fn=lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group))
            try:
                ctx = self.batch_context(qbf, qn64, k, excl_ptr, excl_col, row_filter)
            except nat.NativeError as e:
                if e.code != nat.ERR_UNSUPPORTED:
                    raise
                ctx = None           # shard too small for the tensor-core scheme (same verdict on every rank: equal shard sizes)
            if ctx is None:
                return self._recommend_batch_loop(queries, liked_ptr, liked_col, liked_w, excl_ptr, excl_col, k, row_filter,
                                                  return_info)
            self.batch_step(ctx)
            b, hb = ctx["b"], ctx["hb"]
            status = ctx["gathered"][:, 2 * b * k + hb:].contiguous().view(torch.int32)[:, :b].max(dim=0).values   # any rank unsure
            rows, scores = ctx["m_rows"].cpu().numpy(), ctx["m_scores"].cpu().numpy()
            counts, status = ctx["m_count"].cpu().numpy(), status.cpu().numpy()
        ecp = None if excl_ptr is None else np.asarray(excl_ptr, dtype=np.int64)
        eca = None if excl_col is None else np.asarray(excl_col)
        for u in np.nonzero(status)[0]:                                  # identical on every rank -> collectives stay aligned
            ex = None if ecp is None else eca[ecp[u]:ecp[u + 1]]
            if queries is not None:
                r, sc = self.recommend(query=np.asarray(queries[u]), exclude_rows=ex, k=k, row_filter=row_filter)
            else:
                lpn = np.asarray(liked_ptr, dtype=np.int64)
                lw = None if liked_w is None else np.asarray(liked_w)[lpn[u]:lpn[u + 1]]
                r, sc = self.recommend(liked_rows=np.asarray(liked_col)[lpn[u]:lpn[u + 1]], weights=lw, exclude_rows=ex, k=k,
                                       row_filter=row_filter)
            rows[u, :], scores[u, :] = -1, -np.inf
            rows[u, :len(r)], scores[u, :len(r)], counts[u] = r, sc, len(r)
        if return_info:
            return rows, scores, counts, {"status": status}
        return rows, scores, counts

    def _recommend_batch_loop(self, queries, liked_ptr, liked_col, liked_w, excl_ptr, excl_col, k, row_filter, return_info):
        """Small shards: every query of the batch goes through the sharded single-query path (identical on all ranks)."""
        b = len(queries) if queries is not None else len(liked_ptr) - 1
        rows = np.full((b, k), -1, dtype=np.int64)
        scores = np.full((b, k), -np.inf, dtype=np.float64)
        counts = np.zeros(b, dtype=np.int32)
        ecp = None if excl_ptr is None else np.asarray(excl_ptr, dtype=np.int64)
        eca = None if excl_col is None else np.asarray(excl_col)
        lpn = None if liked_ptr is None else np.asarray(liked_ptr, dtype=np.int64)
        for u in range(b):
            ex = None if ecp is None else eca[ecp[u]:ecp[u + 1]]
            if queries is not None:
                r, sc = self.recommend(query=np.asarray(queries[u]), exclude_rows=ex, k=k, row_filter=row_filter)
            else:
                lw = None if liked_w is None else np.asarray(liked_w)[lpn[u]:lpn[u + 1]]
                r, sc = self.recommend(liked_rows=np.asarray(liked_col)[lpn[u]:lpn[u + 1]], weights=lw, exclude_rows=ex, k=k,
                                       row_filter=row_filter)
            rows[u, :len(r)], scores[u, :len(r)], counts[u] = r, sc, len(r)
        if return_info:
            return rows, scores, counts, {"status": np.zeros(b, dtype=np.int32), "plan": None}
        return rows, scores, counts
