"""Multi-GPU host logic: one process per GPU, W replicated, probes sharded, results gathered on the host.

The path shards by independent units: probe relaxation needs NO data-path collective;
each rank relaxes a contiguous index range and the per-rank unique-attractor tables are merged on the
host in ascending probe index so that first-seen StateIndex and Hits equal the single-stream answer
(main.go:226 iterates results in index order; datacollector/UniqueRelaxedStateHandler.go:41-66).
Learning shards the TARGETS and all-reduces dW (NCCL, inside the C library: hn_comm_attach).

torch.distributed is used only as plumbing (gather of small arrays); works on gloo (CPU tests) and nccl.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) range of rank: sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def merge_unique_tables(tables: List[dict], numpy_only: bool = False) -> dict:
    """Merge per-rank unique tables.

    tables[r] = {"offset": global index of the rank's first probe,
                 "first":  int32 [U_r]  local index of first occurrence (first-seen order),
                 "hits":   int32 [U_r],
                 "hash":   uint64 [U_r][2] 128-bit hash of the canonical final state of that first occurrence,
                 "states": optional uint64 [U_r][words] canonical-comparable packed states (for verification)}
    Returns {"first": global StateIndex per unique attractor in first-seen order, "hits", "rank_of", "local_id"}.
    Equality is decided on the 128-bit hash; where `states` are supplied, hash-equal entries from different
    ranks are verified on the full state (a mismatch raises).  Without `states` the merge runs in the C library
    (hn_merge_unique, O(n)); `numpy_only=True` forces the numpy restatement below (tests compare the two)."""
    if not numpy_only and all(t.get("states") is None for t in tables):   # the library's O(n) host merge
        return _merge_native(tables)
    firsts, hits, hashes, ranks, lids, states = [], [], [], [], [], []
    for r, t in enumerate(tables):
        u = len(t["first"])
        firsts.append(np.asarray(t["first"], dtype=np.int64) + int(t["offset"]))
        hits.append(np.asarray(t["hits"], dtype=np.int64))
        hashes.append(np.asarray(t["hash"], dtype=np.uint64).reshape(u, 2))
        ranks.append(np.full(u, r, dtype=np.int32))
        lids.append(np.arange(u, dtype=np.int32))
        if t.get("states") is not None:
            states.append(np.asarray(t["states"], dtype=np.uint64).reshape(u, -1))
    first = np.concatenate(firsts) if firsts else np.zeros(0, np.int64)
    hit = np.concatenate(hits) if hits else np.zeros(0, np.int64)
    hsh = np.concatenate(hashes) if hashes else np.zeros((0, 2), np.uint64)
    rk = np.concatenate(ranks) if ranks else np.zeros(0, np.int32)
    lid = np.concatenate(lids) if lids else np.zeros(0, np.int32)
    st = np.concatenate(states) if len(states) == len(tables) and states else None
    if first.size == 0:
        return {"first": first.astype(np.int32), "hits": hit.astype(np.int32), "rank_of": rk, "local_id": lid}
    # Every rank lists its attractors in ascending first index and ranks are concatenated in ascending offset, so
    # `first` is already ascending: a STABLE sort on the low hash word keeps the earliest member first in a group.
    h0, h1 = np.ascontiguousarray(hsh[:, 0]), np.ascontiguousarray(hsh[:, 1])
    order = np.argsort(h0, kind="stable")
    h0s, h1s = h0[order], h1[order]
    same0 = h0s[1:] == h0s[:-1]
    ascending = bool(np.all(first[1:] >= first[:-1]))
    if not ascending or bool(np.any(same0 & (h1s[1:] != h1s[:-1]))):
        order = np.lexsort((first, h1, h0))                      # general case: full sort, earliest first in a group
        h0s, h1s = h0[order], h1[order]
        same0 = h0s[1:] == h0s[:-1]
    new_group = np.ones(order.size, dtype=bool)
    new_group[1:] = ~(same0 & (h1s[1:] == h1s[:-1]))
    starts = np.flatnonzero(new_group)
    gid = np.cumsum(new_group) - 1
    n_groups = starts.size
    g_first = np.minimum.reduceat(first[order], starts)
    g_hits = np.add.reduceat(hit[order], starts)
    rep = order[new_group]                                        # earliest member of each group
    if st is not None:
        same = (st[order] == st[rep][gid]).all(axis=1)
        if not same.all():
            raise RuntimeError("128-bit state-hash collision between distinct attractors")
    out_order = np.argsort(g_first, kind="stable")                # first-seen order == ascending first index
    return {"first": g_first[out_order].astype(np.int32), "hits": g_hits[out_order].astype(np.int32),
            "rank_of": rk[rep][out_order], "local_id": lid[rep][out_order]}


def _merge_native(tables: List[dict]) -> dict:
    """merge_unique_tables through hn_merge_unique (hash grouping in C, O(n))."""
    import ctypes as C
    from . import capi
    counts = np.array([len(t["first"]) for t in tables], dtype=np.int64)
    starts = np.concatenate(([0], np.cumsum(counts)))
    n = int(starts[-1])
    if n == 0:
        z = np.zeros(0, np.int32)
        return {"first": z, "hits": z.copy(), "rank_of": z.copy(), "local_id": z.copy()}
    first = np.empty(n, np.int64)
    hits = np.empty(n, np.int32)
    hsh = np.empty((n, 2), np.uint64)
    for r, t in enumerate(tables):
        a, b = int(starts[r]), int(starts[r + 1])
        np.add(np.asarray(t["first"]), int(t["offset"]), out=first[a:b], casting="unsafe")
        hits[a:b] = t["hits"]
        hsh[a:b] = np.asarray(t["hash"], dtype=np.uint64).reshape(b - a, 2)
    f_out, h_out, r_out = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.int64)
    ng = C.c_int64()
    capi.check(capi.load().hn_merge_unique(n, capi.ptr(first), capi.ptr(hits), capi.ptr(hsh), capi.ptr(f_out),
                                           capi.ptr(h_out), capi.ptr(r_out), C.byref(ng)))
    rep = r_out[:ng.value]
    rank_of = (np.searchsorted(starts[1:], rep, side="right")).astype(np.int32)
    return {"first": f_out[:ng.value].astype(np.int32), "hits": h_out[:ng.value].astype(np.int32),
            "rank_of": rank_of, "local_id": (rep - starts[rank_of]).astype(np.int32)}


def canonical_states(packed: np.ndarray, n: int, domain: int) -> np.ndarray:
    """Sign-canonical form used by the dedup key: bipolar states with unit 0 set are inverted
    (s and -s share one energy profile bit for bit, which is the reference's key)."""
    p = np.ascontiguousarray(packed, dtype=np.uint64).copy()
    if domain != 0 or p.size == 0:
        return p
    inv = (p[:, 0] & np.uint64(1)).astype(bool)
    p[inv] = ~p[inv]
    if n % 64:
        p[:, -1] &= np.uint64((1 << (n % 64)) - 1)
    return p


def gather_unique_tables(local: dict, group=None) -> Optional[dict]:
    """All ranks call; rank 0 returns the merged table, the others None.  `local` as in merge_unique_tables.
    Tables travel as ONE padded int64 tensor per rank (first | hits << 32, hash lo, hash hi) through dist.gather —
    on the device for the nccl backend, on the host for gloo; tables that carry `states` use the object gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if local.get("states") is not None:
        gathered: List[Optional[dict]] = [None] * world
        dist.gather_object(local, gathered if rank == 0 else None, dst=0, group=group)
        return merge_unique_tables(gathered) if rank == 0 else None  # type: ignore[arg-type]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    u = len(local["first"])
    head = torch.tensor([u, int(local["offset"])], dtype=torch.int64, device=dev)
    heads = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(heads, head, group=group)
    heads = [h.cpu().numpy() for h in heads]
    umax = max(1, max(int(h[0]) for h in heads))
    pay = np.zeros((umax, 3), dtype=np.int64)
    pay[:u, 0] = np.asarray(local["first"], dtype=np.int64) | (np.asarray(local["hits"], dtype=np.int64) << 32)
    pay[:u, 1:] = np.ascontiguousarray(local["hash"], dtype=np.uint64).reshape(u, 2).view(np.int64)
    mine = torch.from_numpy(pay).to(dev)
    bufs = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, bufs, dst=0, group=group)
    if rank != 0:
        return None
    tables = []
    for r in range(world):
        ur = int(heads[r][0])
        t = bufs[r][:ur].cpu().numpy()
        tables.append({"offset": int(heads[r][1]), "first": (t[:, 0] & 0xFFFFFFFF).astype(np.int32),
                       "hits": (t[:, 0] >> 32).astype(np.int32), "hash": np.ascontiguousarray(t[:, 1:]).view(np.uint64)})
    return merge_unique_tables(tables)


def _merge_outputs(net, n):
    """Host arrays hn_merge_unique_device writes into: kept on the network and reused while large enough (fresh arrays
    of a million rows are page-faulted in on every call).  The returned table's arrays are views of them."""
    cap, bufs = getattr(net, "_merge_out", (0, None))
    if cap < n:
        cap, bufs = n, tuple(np.zeros(n, np.int64) for _ in range(3))
        net._merge_out = (cap, bufs)
    return bufs


def gather_unique_tables_device(net, index_offset: int, group=None) -> Optional[dict]:
    """Cross-GPU merge of unique tables WITHOUT the host in the data path (nccl backend): every rank exports its table
    — global first index, hits, key words, packed final states, zero-field marks — into device tensors
    (hn_export_unique_device), the tensors are gathered on rank 0's GPU, and rank 0 runs the library's two-level
    unique table over the concatenation (hn_merge_unique_device: equal states compared as states, different states
    merged only after the float-profile comparison).  All ranks call; rank 0 returns {"first", "hits", "rank_of",
    "local_id"} (numpy), the others None."""
    import ctypes as C
    import os
    import time
    import torch
    import torch.distributed as dist
    from . import capi
    lib = capi.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    trace = os.environ.get("HN_MERGE_TRACE") and rank == 0
    marks = []
    def mark(name):
        if trace:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))
    mark("start")
    dev = torch.device("cuda", torch.cuda.current_device())
    nu, pk = C.c_int32(), C.c_int32()
    capi.check(lib.hn_export_unique_device(net._h, 0, None, None, None, None, None, C.byref(nu), C.byref(pk)))
    u, w = nu.value, capi.words(net.dimension)
    head = torch.tensor([u, pk.value], dtype=torch.int64, device=dev)
    heads = [torch.zeros_like(head) for _ in range(world)]
    dist.all_gather(heads, head, group=group)
    heads = torch.stack(heads).cpu().numpy()
    mark("heads")
    counts = heads[:, 0].astype(np.int64)
    umax = max(1, int(counts.max()))
    # one int64 payload per rank: [first | hits | zero-field | 4 key words | packed state]
    pay = torch.zeros((umax, 7 + w), dtype=torch.int64, device=dev)
    first = torch.zeros(umax, dtype=torch.int64, device=dev)
    hits = torch.zeros(umax, dtype=torch.int32, device=dev)
    keys = torch.zeros((umax, 4), dtype=torch.int64, device=dev)
    states = torch.zeros((umax, w), dtype=torch.int64, device=dev)
    zf = torch.zeros(umax, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    capi.check(lib.hn_export_unique_device(net._h, int(index_offset), first.data_ptr(), hits.data_ptr(), keys.data_ptr(),
                                           states.data_ptr(), zf.data_ptr(), C.byref(nu), C.byref(pk)))
    pay[:, 0], pay[:, 1], pay[:, 2] = first, hits.to(torch.int64), zf.to(torch.int64)
    pay[:, 3:7], pay[:, 7:] = keys, states
    mark("export")
    bufs = [torch.empty_like(pay) for _ in range(world)] if rank == 0 else None
    dist.gather(pay, bufs, dst=0, group=group)
    mark("gather")
    if rank != 0:
        return None
    allp = torch.cat([bufs[r][:int(counts[r])] for r in range(world)]).contiguous()
    n = allp.shape[0]
    F = allp[:, 0].contiguous()
    H = allp[:, 1].to(torch.int32).contiguous()
    Z = allp[:, 2].to(torch.uint8).contiguous()
    K = allp[:, 3:7].contiguous()
    S = allp[:, 7:].contiguous()
    torch.cuda.synchronize()
    mark("concat")
    f_out, h_out, r_out = _merge_outputs(net, max(n, 1))
    ng = C.c_int64()
    capi.check(lib.hn_merge_unique_device(net._h, n, F.data_ptr(), H.data_ptr(), K.data_ptr(), S.data_ptr(), Z.data_ptr(),
                                          int(heads[0, 1]), capi.ptr(f_out), capi.ptr(h_out), capi.ptr(r_out), C.byref(ng)))
    g = ng.value
    mark("merge")
    if trace:
        print("[merge trace] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f} ms" for a, b in zip(marks, marks[1:])), flush=True)
    return MergedUniqueTable(f_out[:g], h_out[:g], r_out[:g], np.concatenate(([0], np.cumsum(counts))))


class MergedUniqueTable(dict):
    """Result of gather_unique_tables_device: "first" (global index of the first probe of every attractor, in first-seen
    order) and "hits" as written by hn_merge_unique_device.  "rank_of" / "local_id" — which rank's table holds the
    attractor's representative state, and at which row — are derived on first use: with one attractor per probe
    (config 3: a million rows at 8 GPUs) deriving them costs more host time than the merge itself."""

    def __init__(self, first, hits, rep, starts):
        super().__init__(first=first, hits=hits)
        self._rep, self._starts = rep, starts

    def __missing__(self, key):
        if key not in ("rank_of", "local_id"):
            raise KeyError(key)
        rank_of = np.searchsorted(self._starts[1:], self._rep, side="right").astype(np.int32)
        self["rank_of"] = rank_of
        self["local_id"] = (self._rep - self._starts[rank_of]).astype(np.int32)
        return self[key]
