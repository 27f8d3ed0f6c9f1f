"""Catalog wire / on-disk format, loader and upsert.

The reference rebuilds its catalog on every process start from Chroma -> Python lists -> DataFrame
(src/backend/app/constants.py:55-56), and fills Chroma with `collection.upsert(ids=..., embeddings=...)`
(notebooks/create-embeddings.ipynb:1250).  Here:

  from_chroma_result   the notebook / collection.get() shape {"ids": [...], "embeddings": [[...]]} -> CatalogStore
  save_catalog         raw dump of the HBM layout: header, id table, rows [n, ld] in the stored dtype, inv_norm, norm64
  load_catalog         start-up = stream the dump into HBM (no arithmetic, no re-normalisation)
  upsert               Chroma's upsert semantics on an immutable store: returns a NEW store (ids stay string-sorted)

File layout (little endian):  8-byte magic "RBCAT01\\0" | uint64 header_len | header JSON | pad to 4096 |
rows bytes | pad to 4096 | inv_norm fp32[n] | pad to 4096 | norm64 fp64[n].  The id table lives in the JSON header.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

MAGIC = b"RBCAT01\0"
_ALIGN = 4096


def _pad(n: int) -> int:
    return (-n) % _ALIGN


def write_catalog_file(path: str, ids: Optional[Sequence[str]], rows: np.ndarray, inv_norm: np.ndarray, norm64: np.ndarray,
                       d: int, dtype: str, row_base: int = 0) -> None:
    """rows: uint16 [n, ld] (bf16 bit patterns) or float32 [n, ld]."""
    n, ld = rows.shape
    want = np.uint16 if dtype == "bf16" else np.float32
    if rows.dtype != want:
        raise ValueError(f"rows must be {want} for dtype {dtype}")
    header = json.dumps({"n": int(n), "d": int(d), "ld": int(ld), "dtype": dtype, "row_base": int(row_base),
                         "ids": None if ids is None else list(ids)}).encode()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(np.uint64(len(header)).tobytes())
        f.write(header)
        f.write(b"\0" * _pad(16 + len(header)))
        b = np.ascontiguousarray(rows).tobytes()
        f.write(b)
        f.write(b"\0" * _pad(len(b)))
        b = np.ascontiguousarray(inv_norm[:n], dtype=np.float32).tobytes()
        f.write(b)
        f.write(b"\0" * _pad(len(b)))
        f.write(np.ascontiguousarray(norm64[:n], dtype=np.float64).tobytes())


def read_catalog_file(path: str) -> Tuple[Dict, np.ndarray, np.ndarray, np.ndarray]:
    """(header, rows memmap [n, ld], inv_norm memmap, norm64 memmap) — nothing is copied until sliced."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a rebert catalog file")
        hlen = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
        header = json.loads(f.read(hlen).decode())
    n, ld = header["n"], header["ld"]
    esize, npdt = (2, np.uint16) if header["dtype"] == "bf16" else (4, np.float32)
    off = 16 + hlen + _pad(16 + hlen)
    rows = np.memmap(path, dtype=npdt, mode="r", offset=off, shape=(n, ld))
    off += n * ld * esize + _pad(n * ld * esize)
    inv = np.memmap(path, dtype=np.float32, mode="r", offset=off, shape=(n,))
    off += n * 4 + _pad(n * 4)
    nrm = np.memmap(path, dtype=np.float64, mode="r", offset=off, shape=(n,))
    return header, rows, inv, nrm


def from_chroma_result(result: Dict, dtype: str = "fp32", device=None):
    """`collection.get(include=["embeddings"])` (constants.py:55) or the notebook's {"id", "values"} frame -> CatalogStore."""
    from .catalog import CatalogStore
    ids = result["ids"] if "ids" in result else result["id"]
    emb = result["embeddings"] if "embeddings" in result else result["values"]
    return CatalogStore.from_host(ids, np.asarray(emb, dtype=np.float32), dtype=dtype, device=device)


def save_catalog(store, path: str) -> None:
    import torch
    rows = store.rows[:store.n]
    rows_np = rows.view(torch.int16).cpu().numpy().view(np.uint16) if store.dtype == "bf16" else rows.cpu().numpy()
    write_catalog_file(path, store.ids, rows_np, store.inv_norm[:store.n].cpu().numpy(), store.norm64[:store.n].cpu().numpy(),
                       store.d, store.dtype, store.row_base)


def load_catalog(path: str, device=None, chunk_rows: int = 1 << 18):
    """Stream a dump into HBM.  The stored values, inv_norm and norm64 are taken as-is: no kernel runs."""
    import torch
    from .catalog import CatalogStore
    dev = CatalogStore._require_cuda(device)
    header, rows, inv, nrm = read_catalog_file(path)
    n, d, ld, dtype = header["n"], header["d"], header["ld"], header["dtype"]
    if CatalogStore.layout(n, d, dtype)[0] != ld:
        raise ValueError(f"{path}: row stride {ld} does not match this build's layout for d={d} {dtype}")
    _, d_rows, d_inv, d_nrm = CatalogStore._alloc(n, d, dtype, dev)
    for s in range(0, n, chunk_rows):
        e = min(n, s + chunk_rows)
        host = np.array(rows[s:e])                       # private, writable copy of this chunk of the memmap
        blk = torch.from_numpy(host.view(np.int16) if dtype == "bf16" else host)
        dst = d_rows[s:e].view(torch.int16) if dtype == "bf16" else d_rows[s:e]
        dst.copy_(blk, non_blocking=False)
    d_inv[:n].copy_(torch.from_numpy(np.array(inv)))
    d_nrm[:n].copy_(torch.from_numpy(np.array(nrm)))
    torch.cuda.synchronize(dev)
    return CatalogStore(d_rows, d_inv, d_nrm, n, d, ld, dtype, header.get("row_base", 0), header.get("ids"))


def upsert(store, ids: Sequence[str], embeddings: np.ndarray):
    """Chroma-style upsert (create-embeddings.ipynb:1250): rows with an existing id are replaced, new ids are added.
    The store is immutable (requests may be in flight), so this returns a NEW CatalogStore, ids string-sorted."""
    import torch
    from . import _native as nat
    from .catalog import CatalogStore
    import ctypes as C
    if store.ids is None:
        raise ValueError("upsert needs a catalog with an id table")
    emb = np.asarray(embeddings, dtype=np.float32)
    ids = [str(i) for i in ids]
    if emb.ndim != 2 or emb.shape[0] != len(ids) or emb.shape[1] != store.d or len(set(ids)) != len(ids):
        raise ValueError("upsert: embeddings must be [len(ids), d] with unique ids")
    new_of = {i: j for j, i in enumerate(ids)}
    all_ids = sorted(set(store.ids) | set(ids))
    old_of = {i: r for r, i in enumerate(store.ids)}
    n = len(all_ids)
    src_old = np.array([old_of.get(i, -1) if i not in new_of else -1 for i in all_ids], dtype=np.int64)
    src_new = np.array([new_of.get(i, -1) for i in all_ids], dtype=np.int64)
    dev = store.device
    lib = nat.load()
    ld, rows, inv_norm, norm64 = CatalogStore._alloc(n, store.d, store.dtype, dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        keep = np.nonzero(src_old >= 0)[0]
        if len(keep):   # unchanged rows keep their stored bits (device-to-device gather; data movement only)
            rows[torch.from_numpy(keep).to(dev)] = store.rows[torch.from_numpy(src_old[keep]).to(dev)]
        put = np.nonzero(src_new >= 0)[0]
        if len(put):    # new / replaced rows go through the same conversion as catalog load
            tmp_ld, tmp, _, _ = CatalogStore._alloc(len(put), store.d, store.dtype, dev)
            src = torch.from_numpy(np.ascontiguousarray(emb[src_new[put]])).to(dev)
            nat.check(lib.rebert_catalog_store_rows(src.data_ptr(), len(put), store.d, nat.DTYPES[store.dtype], tmp.data_ptr(), tmp_ld, st))
            rows[torch.from_numpy(put).to(dev)] = tmp[:len(put)]
        nat.check(lib.rebert_catalog_norms(rows.data_ptr(), n, ld, nat.DTYPES[store.dtype], inv_norm.data_ptr(), norm64.data_ptr(), st))
        torch.cuda.current_stream().synchronize()
    return CatalogStore(rows, inv_norm, norm64, n, store.d, ld, store.dtype, store.row_base, all_ids)
