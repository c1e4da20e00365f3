"""Worker for tests/test_gpu_multi.py — launched with torchrun, one rank per GPU.
Targets are sharded across ranks; every rank runs Delta epochs on its shard with the library's NCCL
all-reduce of dW attached; every replica of W must equal the oracle's full-batch W bit for bit.
Probes are sharded too and the merged unique table must equal the single-stream answer."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopfield_network_go_b200 import capi, multigpu  # noqa: E402
from hopfield_network_go_b200 import hopfieldnetwork as hn  # noqa: E402
from oracle import orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n, p, b = 192, 24, 120
    rng = orc.Rng(2024)
    t = rng.generate_states(0, n, p)
    net = hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(1).SetNetworkLearningRule(
        hn.DeltaLearningRule).SetNetworkLearningMethod(hn.FullSetMethod).SetSeed(1).SetDevice(local).Build()
    uid = (C.c_uint8 * 128)()
    if rank == 0:
        capi.check(capi.load().hn_comm_unique_id(uid))
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)
    uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
    capi.check(capi.load().hn_comm_attach(net._h, uid, rank, world))
    onet = orc.Net(n)
    lo, hi = multigpu.shard_range(p, rank, world)
    for epoch in range(3):
        noised = t.copy()
        for s in range(p):
            rng.apply_noise(1, 0.1, noised[s])
            noised[s] = np.where(noised[s] <= 0, -1.0, 1.0)
        perms = rng.fresh_perms(n, p)
        onet.delta(t, noised, perms)                       # oracle: all targets
        net.delta_epoch(t[lo:hi], noised[lo:hi], perms[lo:hi])   # GPU: this rank's shard + all-reduce
        assert np.array_equal(net.GetMatrix(), onet.w), f"rank {rank}: W differs after epoch {epoch}"
    capi.check(capi.load().hn_comm_detach(net._h))
    # probes sharded, unique tables merged on the host
    onet.normalize()
    onet.set_targets(t)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(t))
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    plo, phi = multigpu.shard_range(b, rank, world)
    res = net.relax_batch(probes[plo:phi], pool, dedup=True)
    merged = multigpu.gather_unique_tables({"offset": plo, "first": res.UniqueFirst, "hits": res.UniqueHits,
                                            "hash": res.StateHash[res.UniqueFirst],
                                            "states": multigpu.canonical_states(res.FinalStates[res.UniqueFirst], n, 0)})
    if rank == 0:
        ores = onet.relax_batch(probes.copy(), pool, threads=4)
        _, first, hits = orc.unique_table(ores["energies"])
        assert merged["first"].tolist() == first.tolist() and merged["hits"].tolist() == hits.tolist()
    dist.barrier()
    net.close()
    # the same checks bench.py runs before its timed region at n_gpus > 1 (sharded explicit epochs, sharded LearnStates
    # with the all-reduce(min) early exit, unique tables merged on rank 0's GPU)
    import bench
    verdict = bench.multi_gpu_parity(hn, dist, rank, world, local)
    assert verdict == "ok", verdict
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} unique={len(first)} parity={verdict}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
