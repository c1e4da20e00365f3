#!/usr/bin/env python
"""bench.py — probe states relaxed / second at N=4096 (BASELINE.json metric), one rank per GPU.

A step = one pass of the hot path (initial fields GEMM -> relaxation kernel -> unique-attractor
table) over one batch of synthetic probes.  Workload at every N: BASELINE config 3 per GPU
(N=4096, Hebbian, 400 targets, 2^17 probes per GPU, shared 100-row schedule, dedup on) — weak scaling.

  value : probes/s with the probes, schedule and W already resident in HBM (CUDA-event time)
  e2e   : probes/s through hn_relax_batch with HOST buffers (H2D of probes+schedule, D2H of results
          inside the timed region)
  --impl reference : the CPU restatement of the reference (oracle/, all host cores) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "probe_states_relaxed_per_sec_N4096"
UNIT = "probes/s"


def synth(n, p, b, seed):
    """i.i.d. fair +-1 states (what states/StateGenerator.go:44-52 yields), packed."""
    rs = np.random.Generator(np.random.PCG64(seed))
    w = (n + 63) // 64
    def bits(count):
        a = rs.integers(0, 2**63, size=(count, w), dtype=np.int64).astype(np.uint64) << np.uint64(1)
        a |= rs.integers(0, 2, size=(count, w), dtype=np.int64).astype(np.uint64)
        if n % 64:
            a[:, -1] &= np.uint64((1 << (n % 64)) - 1)
        return a
    targets = bits(p)
    probes = bits(b)
    pool = np.empty((100, n), dtype=np.uint16)
    lst = np.arange(n, dtype=np.uint16)
    for r in range(100):          # one persistent list shuffled once per sweep (HopfieldNetwork.go:374,393)
        rs.shuffle(lst)
        pool[r] = lst
    return targets, probes, pool


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0])); self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def run_reference(args, cfg):
    """CPU restatement of the reference (oracle, port) with every host core, bounded sample."""
    from oracle import orc
    n, p = cfg["N"], cfg["P"]
    cores = os.cpu_count() or 1
    targets, probes, pool = synth(n, p, max(4 * cores, 64), 0x5EED + 3)
    t_f64 = orc.unpack(targets, n)
    net = orc.Net(n)
    net.hebbian(t_f64)
    net.normalize()
    net.set_targets(t_f64)
    sample = cores  # one probe per core per step: O(1 s)/probe/core at N=4096
    times = []
    for it in range(args.warmup + args.steps):
        pr = orc.unpack(probes[(it * sample) % probes.shape[0]:][:sample], n)
        if pr.shape[0] < sample:
            pr = orc.unpack(probes[:sample], n)
        t0 = time.perf_counter()
        net.relax_batch(pr, pool, threads=cores, want_energies=True, faithful_cost=True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    tot = sum(times)
    v = sample * len(times) / tot
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], **{k: cfg[k] for k in ("N", "P")}, "probes_per_gpu": cfg["B"],
                       "probes_per_step": sample, "note": "bounded sample of the same workload per step (CPU)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} probes per step x {len(times)} steps of the same workload, oracle "
                                       "C restatement of the reference algorithm as written (N dots/sweep + 2-3 gemv)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--probes", type=int, default=1 << 17, help="probes per GPU per step")
    ap.add_argument("--dim", type=int, default=4096)
    ap.add_argument("--targets", type=int, default=400)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-delta", action="store_true", help="skip the Delta-rule epoch timing")
    ap.add_argument("--no-f32", action="store_true", help="skip the extra measurements of the other column streams")
    ap.add_argument("--columns", default="auto", choices=["auto", "f64", "f32", "i16"],
                    help="column stream of the relaxation kernel (auto = library default: exact int16 stream when available)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = {"name": f"config3: N={args.dim} Hebbian P={args.targets}, {args.probes} probes/GPU, shared 100-row schedule, dedup",
           "N": args.dim, "P": args.targets, "B": args.probes}

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, cfg)
        return

    import torch
    import torch.distributed as dist
    from hopfield_network_go_b200 import hopfieldnetwork as hn

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n, p, b = cfg["N"], cfg["P"], cfg["B"]
    _, probes, _ = synth(n, p, b, 0x5EED + 3 + 1000 * (rank + 1))
    targets, _, pool = synth(n, p, 1, 0x5EED + 3)  # same W / schedule on every rank
    net = hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(1).SetNetworkLearningRule(
        hn.HebbianLearningRule).SetNetworkLearningMethod(hn.FullSetMethod).SetSeed(1).SetDevice(local).SetColumnStream(
        {"f64": 0, "f32": 1, "i16": 2, "auto": 3}[args.columns]).Build()
    net.LearnStates(targets)
    stream = "i16" if net.integer_stream_active() else ("f32" if args.columns == "f32" else "f64")
    # pinned host staging for the e2e leg
    probes_pinned = torch.from_numpy(probes).pin_memory()
    probes_np = probes_pinned.numpy()
    net.upload_perm_pool(pool)
    net.upload_probes(probes_np)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        net.relax_resident(dedup=True)
    barrier()
    sampler.start()
    dev_ms, relax_ms, fields_ms, post_ms, launches = [], [], [], [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev_ms.append(net.relax_resident(dedup=True))
        t = net.last_timing()
        relax_ms.append(t.relax_ms)
        fields_ms.append(t.fields_ms)
        post_ms.append(t.post_ms)
        launches += t.launches
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    # counters for the roofline (same every step: deterministic)
    import ctypes as C
    from hopfield_network_go_b200 import capi
    out = capi.HnRelaxOut()
    capi.check(capi.load().hn_download_results(net._h, C.byref(out)))
    flips, sweeps = out.total_flips, out.total_sweeps
    alg_bytes = 8 * n * flips + 8 * n * sweeps + b * (n // 4 + 4 * p + 16)
    step_ms = float(np.sum(dev_ms))
    # e2e: host buffers through hn_relax_batch
    e2e_t = []
    n_unique_global = None
    for it in range(1 + min(args.steps, 2)):
        barrier()
        t1 = time.perf_counter()
        res = net.relax_batch(probes_np, pool, dedup=True, reuse_pinned_buffers=True)   # page-locked result buffers of the network
        if world > 1:   # host-side merge of the per-GPU unique-attractor tables (first-seen order, hits)
            from hopfield_network_go_b200 import multigpu
            merged = multigpu.gather_unique_tables({"offset": rank * b, "first": res.UniqueFirst, "hits": res.UniqueHits,
                                                    "hash": res.StateHash[res.UniqueFirst]})
            if rank == 0:
                n_unique_global = len(merged["first"])
        barrier()
        if it > 0:
            e2e_t.append(time.perf_counter() - t1)
    h2d = probes_np.nbytes + pool.nbytes
    d2h = res.FinalStates.nbytes + res.NumSteps.nbytes + res.Stable.nbytes + res.DistancesToTargets.nbytes + \
        res.Flips.nbytes + res.MarginHits.nbytes + res.AttractorId.nbytes + 2 * 4 * len(res.UniqueFirst) + res.StateHash.nbytes
    # the same step with the float32 / exact-int16 column streams (bit-identical results, 1/2 and 1/4 of the column
    # bytes): reported beside the headline, never as the headline
    alt_ms = {"f64": 0.0, "f32": 0.0, "i16": 0.0}
    alt_relax_ms = dict(alt_ms)
    if not args.no_f32:
        for name, mode in (("f64", 0), ("f32", 1), ("i16", 2)):
            if name == stream:
                continue
            alt = hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(1).SetNetworkLearningRule(
                hn.HebbianLearningRule).SetNetworkLearningMethod(hn.FullSetMethod).SetSeed(1).SetDevice(local).SetColumnStream(mode).Build()
            alt.LearnStates(targets)          # in-library learning: the integer structure is captured at normalisation
            alt.upload_perm_pool(pool)
            alt.upload_probes(probes_np)
            alt.relax_resident(dedup=True)
            alt_ms[name] = alt.relax_resident(dedup=True)
            alt_relax_ms[name] = alt.last_timing().relax_ms
            alt.close()
    tmax = torch.tensor([step_ms, float(np.sum(relax_ms)), float(np.mean(e2e_t)), alt_ms["f64"], alt_ms["f32"], alt_ms["i16"],
                         alt_relax_ms["f64"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    step_ms_max, relax_ms_max, e2e_max, f64_ms_max, f32_ms_max, i16_ms_max, f64_relax_ms_max = [float(x) for x in tmax.cpu()]
    delta_ms = {}
    if not args.no_delta:   # collective when world > 1: every rank takes part
        if world == 1:
            delta_ms["config2_N1024_P100_inversion0.1"] = measure_delta_epoch(hn, local, 1024, 100, 5)
        delta_ms[f"config5_N16384_P{128 * world}_sharded_128_per_gpu"] = measure_delta_epoch(
            hn, local, 16384, 128, 2, dist=dist if world > 1 else None, rank=rank, world=world)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        achieved = alg_bytes * args.steps / (relax_ms_max * 1e-3) / 1e9
        col_bytes = {"f64": 8, "f32": 4, "i16": 2}[stream]
        field_bytes = 4 if stream == "i16" else 8
        streamed = col_bytes * n * flips + field_bytes * n * sweeps     # what this stream really moves L2->SM per launch
        kernel = {"f64": "relax_kernel<32,double,128,4>", "f32": "relax_kernel<32,float,128,4>", "i16": "relax_kernel<32,short,128,9>"}[stream]
        notes = {
            "f64": "bounded by the L2->SM fabric: every flip moves its 8N-byte column L2->SM (~10 TB/s, ~34 B/clk/SM)",
            "f32": "float32 shadow of W: 4N bytes per flip, float64 fields, exact re-evaluation inside the grown margin",
            "i16": "exact integer stream: int32 fields and int16 columns (2N bytes per flip) give the reference's float64 "
                   "decisions bit for bit; exact ties are re-evaluated in float64 in the reference's summation order; the int16 "
                   "image of W (32 MiB) lives in L2, so DRAM is idle and the kernel is bound by instruction issue (75 % of issue "
                   "slots, ncu) and the dependent chain of a flip, not by bytes"}
        line = {"metric": METRIC, "value": world * b * args.steps / (step_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64" if stream != "i16" else "f64 (decisions taken on an exact int32 image of the float64 fields)",
                "data": "synthetic",
                "config": {"workload": cfg["name"], "N": n, "P": p, "probes_per_gpu": b, "column_stream": f"{args.columns} -> {stream}",
                           "l2": "inputs larger than L2 (W 128 MiB + 2-4 GiB of initial fields streamed per step)"},
                "e2e": {"value": world * b / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                             "traffic": None, "kernel": kernel,
                             "note": "per GPU; achieved = ALGORITHMIC bytes of the reference algorithm (8N per flip + 8N per sweep, "
                                     "SURVEY 8d) over the relaxation kernel time; >1 because W stays in L2 (see traffic) and "
                                     "this stream moves fewer bytes than the algorithm names (streamed_bytes_per_launch); "
                                     + notes[stream] + "; profiles/r01_relax_i16_final_v2_ncu_summary.txt, profiles/r01_relax_f64_ncu_summary.txt",
                             "peak_source": "measured" if peaks else "fallback",
                             "algorithmic_bytes_per_launch": int(alg_bytes), "streamed_bytes_per_launch": int(streamed),
                             "streamed_gbs": streamed * args.steps / (relax_ms_max * 1e-3) / 1e9,
                             "flips_per_probe": flips / b, "sweeps_per_probe": sweeps / b},
                "clocks": sampler.summary(), "wall_s": wall,
                "stage_ms_per_step": {"fields_gemm": float(np.mean(fields_ms)), "relax": float(np.mean(relax_ms)),
                                      "dedup": float(np.mean(post_ms))},
                "alt_streams_note": "same step, same inputs, bit-identical results (tests/test_gpu_f32_columns.py, "
                                    "tests/test_gpu_i16_columns.py, tests/test_gpu_auto_columns.py)",
                "margin_hits_per_probe": out.total_margin_hits / b,
                "unique_attractors": int(n_unique_global if n_unique_global is not None else out.n_unique)}
        for name, ms in (("f64", f64_ms_max), ("f32", f32_ms_max), ("i16", i16_ms_max)):
            if ms > 0:
                line[f"value_{name}_column_stream"] = world * b / (ms * 1e-3)
        traffic = {}
        try:   # DRAM bytes of the relaxation kernels from the committed ncu captures of this command (per launch)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_relax_traffic.json")))
        except Exception:
            pass
        tr = traffic.get(stream, {})
        if tr.get("probes_per_launch") == b and tr.get("N") == n:
            line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
            line["roofline"]["traffic_source"] = tr["source"]
        if stream != "f64" and f64_relax_ms_max > 0:   # the float64 stream's own roofline beside the headline
            ach64 = alg_bytes / (f64_relax_ms_max * 1e-3) / 1e9
            t64 = traffic.get("f64", {})
            line["roofline_f64_stream"] = {"bound": "hbm", "achieved": ach64, "peak": hbm, "unit": "GB/s", "frac": ach64 / hbm,
                                           "traffic": t64.get("dram_bytes_per_launch") if t64.get("probes_per_launch") == b and t64.get("N") == n else None,
                                           "kernel": "relax_kernel<32,double,128,4>", "note": notes["f64"]}
        if not args.no_delta:
            line["delta_epoch_ms"] = dict(delta_ms, note="wall clock (max over ranks) through hn_delta_epoch + hn_energy_profiles "
                                          "with host buffers (H2D of packed targets, noised copies and permutations inside the timed "
                                          "region); with n_gpus > 1 the config-5 targets are sharded 128 per GPU and dW (2 GiB) is "
                                          "all-reduced over NCCL every epoch")
        if not args.no_cpu:
            from oracle import orc
            cores = os.cpu_count() or 1
            t_f64 = orc.unpack(targets, n)
            onet = orc.Net(n)
            onet.hebbian(t_f64)
            onet.normalize()
            onet.set_targets(t_f64)
            sample = 2 * cores
            pr = orc.unpack(probes[:sample], n)
            t2 = time.perf_counter()
            onet.relax_batch(pr, pool, threads=cores, want_energies=True, faithful_cost=True)
            dt = time.perf_counter() - t2
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} probes of the same batch, oracle C restatement of the "
                                              "reference algorithm as written, one probe per worker thread"}
        print(json.dumps(line))
    net.close()
    if world > 1:
        dist.destroy_process_group()


def measure_delta_epoch(hn, local, n, p, epochs, noise_scale=0.1, dist=None, rank=0, world=1):
    """Second BASELINE metric: Delta-rule epoch time = one pass of `delta` (LearningRule.go:90-114) over all
    targets + the per-epoch target diagnostics of fullSetLearningMethod (LearningMethod.go:45-58).
    Noise / update orders are drawn host-side before the timed region.  With world > 1 the p*world targets are
    sharded (p per GPU) and the library all-reduces dW over NCCL inside hn_delta_epoch."""
    import ctypes as C
    from hopfield_network_go_b200 import capi
    rs = np.random.Generator(np.random.PCG64(0x5EED + 2))
    w = (n + 63) // 64
    ptot = p * world
    targets = rs.integers(0, 2**63, size=(ptot, w), dtype=np.int64).astype(np.uint64) << np.uint64(1) | \
        rs.integers(0, 2, size=(ptot, w), dtype=np.int64).astype(np.uint64)
    net = hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(1).SetNetworkLearningRule(
        hn.DeltaLearningRule).SetNetworkLearningMethod(hn.FullSetMethod).SetSeed(2).SetDevice(local).Build()
    if world > 1:
        uid = (C.c_uint8 * 128)()
        if rank == 0:
            capi.check(capi.load().hn_comm_unique_id(uid))
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0)
        uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
        capi.check(capi.load().hn_comm_attach(net._h, uid, rank, world))
    lo, hi = rank * p, (rank + 1) * p
    k = int(n * noise_scale)
    times = []
    for ep in range(epochs + 2):
        noised = targets[lo:hi].copy()
        for s in range(p):      # MaximalInversion: int(N*ratio) units among 0..N-2 (NoiseApplication.go:64-75)
            idx = rs.permutation(n - 1)[:k]
            np.bitwise_xor.at(noised[s], idx // 64, np.uint64(1) << (idx % 64).astype(np.uint64))
        perms = np.stack([rs.permutation(n).astype(np.uint16) for _ in range(p)])
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        net.delta_epoch(targets[lo:hi], noised, perms)
        net.energy_profiles(targets[lo:hi], want_energies=False)
        dt = time.perf_counter() - t0
        if ep >= 2:
            times.append(dt * 1e3)
    ms = float(np.mean(times))
    if world > 1:
        import torch
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        capi.check(capi.load().hn_comm_detach(net._h))
    net.close()
    return ms


if __name__ == "__main__":
    main()
