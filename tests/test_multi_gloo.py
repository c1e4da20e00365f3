"""CPU, world_size 2, gloo: the N>1 host logic — contiguous probe sharding, per-rank relaxation (the oracle
stands in for the GPU here), gather + merge of the unique-attractor tables — must reproduce the unsharded
single-stream answer (first-seen StateIndex and Hits).  Also: dW of target shards sums to the full dW."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from hopfield_network_go_b200 import multigpu
from oracle import orc


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _problem():
    n, p, b = 48, 3, 90       # few targets, small N -> many probes share attractors
    rng = orc.Rng(99)
    t = rng.generate_states(0, n, p)
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    net = orc.Net(n); net.hebbian(t); net.normalize(); net.set_targets(t)
    return net, probes, pool, n


def _local_table(net, probes, pool, n, begin, end):
    res = net.relax_batch(probes[begin:end].copy(), pool, threads=1)
    aid, first, hits = orc.unique_table(res["energies"])
    packed = multigpu.canonical_states(orc.pack(res["final"][first]), n, 0)
    # stand-in for the device hash: any deterministic 128-bit function of the canonical state
    hsh = np.stack([(packed * np.arange(1, packed.shape[1] + 1, dtype=np.uint64)).sum(axis=1),
                    (packed ^ np.uint64(0x9E3779B97F4A7C15)).sum(axis=1)], axis=1)
    return {"offset": begin, "first": first, "hits": hits, "hash": hsh, "states": packed}


def _worker(rank, world, port, q, with_states=True):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net, probes, pool, n = _problem()
    begin, end = multigpu.shard_range(probes.shape[0], rank, world)
    local = _local_table(net, probes, pool, n, begin, end)
    if not with_states:            # the tensor gather + hn_merge_unique path (what bench.py uses)
        local.pop("states")
    merged = multigpu.gather_unique_tables(local)
    if rank == 0:
        q.put((merged["first"].tolist(), merged["hits"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n, w in ((10, 3), (7, 8), (0, 2), (131072, 8), (5, 1)):
        r = [multigpu.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in r]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("with_states", [True, False])
def test_two_rank_gloo_merge_equals_single_stream(built, with_states):
    net, probes, pool, n = _problem()
    res = net.relax_batch(probes.copy(), pool, threads=2)
    _, first, hits = orc.unique_table(res["energies"])
    assert len(first) < probes.shape[0]          # the case really has repeated attractors
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, with_states)) for r in range(2)]
    for p in procs: p.start()
    got_first, got_hits = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert got_first == first.tolist() and got_hits == hits.tolist()


def test_merge_detects_hash_collisions():
    t0 = {"offset": 0, "first": [0], "hits": [1], "hash": np.array([[1, 2]], np.uint64), "states": np.array([[5]], np.uint64)}
    t1 = {"offset": 4, "first": [0], "hits": [2], "hash": np.array([[1, 2]], np.uint64), "states": np.array([[6]], np.uint64)}
    with pytest.raises(RuntimeError):
        multigpu.merge_unique_tables([t0, t1])
    t1["states"] = np.array([[5]], np.uint64)
    m = multigpu.merge_unique_tables([t0, t1])
    assert m["first"].tolist() == [0] and m["hits"].tolist() == [3]


def test_target_shards_sum_to_full_dw(built):
    """Delta/Hebbian dW is integer valued: per-shard partial sums add up to the full update in
    any order, so the NCCL all-reduce keeps every replica of W bitwise identical."""
    n, p = 40, 9
    t = orc.Rng(5).generate_states(0, n, p)
    full = orc.Net(n, force_symmetric=False, force_zero_diag=False); full.hebbian(t)
    acc = np.zeros((n, n))
    for r in range(4):
        b, e = multigpu.shard_range(p, r, 4)
        part = orc.Net(n, force_symmetric=False, force_zero_diag=False)
        if e > b:
            part.hebbian(t[b:e])
        acc += part.w
    assert np.array_equal(acc, full.w)


def test_native_merge_equals_numpy_merge():
    """hn_merge_unique (O(n) hash grouping in the library) == the numpy restatement, including tables whose offsets
    are not ascending and attractors shared by several ranks."""
    import numpy as np
    from hopfield_network_go_b200 import multigpu
    rs = np.random.default_rng(11)
    pool = rs.integers(0, 2**63, size=(700, 2), dtype=np.uint64)
    for trial in range(6):
        tabs = []
        for r in range(5):
            u = int(rs.integers(0, 400))
            idx = rs.choice(700, size=u, replace=False)
            first = np.sort(rs.choice(10000, size=u, replace=False)).astype(np.int32)
            off = r * 10000 if trial < 4 else (4 - r) * 10000
            tabs.append({"offset": off, "first": first, "hits": rs.integers(1, 50, size=u).astype(np.int32), "hash": pool[idx]})
        a = multigpu.merge_unique_tables(tabs)
        b = multigpu.merge_unique_tables(tabs, numpy_only=True)
        assert set(a) == set(b)
        for k in a:
            assert np.array_equal(a[k], b[k]), (trial, k)
        assert int(a["hits"].sum()) == sum(int(t["hits"].sum()) for t in tabs)
        assert np.all(np.diff(a["first"]) > 0)
