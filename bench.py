#!/usr/bin/env python
"""bench.py — probe states relaxed / second at N=4096 and Delta-rule epoch ms (BASELINE.json metric), one rank per GPU.

A step = one pass of the hot path (initial fields contraction -> relaxation -> unique-attractor table) over one batch
of synthetic probes.  Workload at every GPU count: BASELINE config 3 per GPU (N=4096, Hebbian, 400 targets, 2^17 probes
per GPU, shared 100-row schedule, dedup on) — weak scaling.

  value : probes/s with the probes, schedule and W already resident in HBM (CUDA-event time inside the library)
  e2e   : probes/s through hn_relax_batch with HOST buffers (H2D of probes + schedule, D2H of the results, and — with
          several GPUs — the merge of the per-GPU unique tables) inside the timed region, over all --steps
  roofline / roofline_l2 / issue_frac : the HBM line SURVEY 8(d) defines, and the same kernel against what binds it
  delta_epoch : the second half of the metric, with its own roofline and cpu_baseline
  multi_gpu_parity (n_gpus > 1) : sharded learning and sharded probing checked against the oracle before the timed region
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


def load_json(*path):
    try:
        return json.load(open(os.path.join(ROOT, *path)))
    except Exception:
        return {}


def run_reference(args, cfg):
    """CPU restatement of the reference (oracle, port) with every host core, bounded sample: every worker thread
    relaxes `per_core` probes of the same workload per step, so a step costs the mean probe, not the slowest one."""
    from oracle import orc
    n, p = cfg["N"], cfg["P"]
    cores = os.cpu_count() or 1
    per_core = 2
    sample = per_core * cores
    targets, probes, pool = synth(n, p, max(4 * sample, 64), 0x5EED + 3)
    t_f64 = orc.unpack(targets, n)
    net = orc.Net(n)
    net.hebbian(t_f64)
    net.normalize()
    net.set_targets(t_f64)
    times = []
    for it in range(args.warmup + args.steps):
        lo = (it * sample) % (probes.shape[0] - sample + 1)
        pr = orc.unpack(probes[lo:lo + sample], n)
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
                             "sample": f"{sample} probes per step ({per_core} per worker thread) x {len(times)} steps of the "
                                       "same workload, oracle C restatement of the reference algorithm as written "
                                       "(N dots/sweep + 2-3 gemv)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def attach_comm(capi, net, dist, rank, world):
    import ctypes as C
    uid = (C.c_uint8 * 128)()
    if rank == 0:
        capi.check(capi.load().hn_comm_unique_id(uid))
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)
    uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
    capi.check(capi.load().hn_comm_attach(net._h, uid, rank, world))


def multi_gpu_parity(hn, dist, rank, world, local):
    """Sharded learning and sharded probing on a small network, checked against the oracle (the checker, outside the
    timed region): (1) Delta epochs with the targets sharded and the library's NCCL all-reduce of dW -> every replica's W
    equals the oracle's full-batch W bit for bit; (2) the whole sharded LearnStates (all-reduce(min) early exit) -> same
    LearnStateData rows as the oracle, replicas bit-identical; (3) probes sharded, per-GPU unique tables merged on rank
    0's GPU -> the oracle's single-stream table.  Returns "ok" or the first mismatch (rank 0)."""
    import torch
    from hopfield_network_go_b200 import capi, multigpu
    from oracle import orc
    bad = []
    n, p, b = 192, 24, 160
    rng = orc.Rng(2024)
    t = rng.generate_states(0, n, p)
    mk = lambda **kw: hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(kw.get("epochs", 1)).SetNetworkLearningRule(
        hn.DeltaLearningRule).SetNetworkLearningMethod(hn.FullSetMethod).SetSeed(kw.get("seed", 1)).SetDevice(local).SetLearningNoiseMethod(
        hn.MaximalInversion).SetLearningNoiseRatio(0.1).Build()
    # (1) explicit sharded Delta epochs
    net = mk()
    attach_comm(capi, net, dist, rank, world)
    onet = orc.Net(n)
    lo, hi = multigpu.shard_range(p, rank, world)
    for epoch in range(3):
        noised = t.copy()
        for s in range(p):
            rng.apply_noise(1, 0.1, noised[s])
            noised[s] = np.where(noised[s] <= 0, -1.0, 1.0)
        perms = rng.fresh_perms(n, p)
        onet.delta(t, noised, perms)
        if hi > lo:
            net.delta_epoch(t[lo:hi], noised[lo:hi], perms[lo:hi])
        else:
            bad.append("empty shard in the explicit-epoch check")
        if not np.array_equal(net.GetMatrix(), onet.w):
            bad.append(f"rank {rank}: W != oracle W after sharded Delta epoch {epoch}")
            break
    capi.check(capi.load().hn_comm_detach(net._h))
    # (2) sharded LearnStates: same seed on every rank and in the oracle
    net2 = mk(epochs=12, seed=77)
    attach_comm(capi, net2, dist, rank, world)
    o2 = orc.Net(n)
    nrows, rows, _ = o2.learn_states(t.copy(), 2, 0, 12, 1, 0.1, orc.Rng(77), cap=12 * p)
    lsd = net2.LearnStates(t.copy(), shard=(lo, hi))
    mine = [(d.Epoch, d.TargetStateIndex, d.Stable) for d in lsd]
    want = [r for r in rows if lo <= r[1] < hi]
    if mine != want:
        bad.append(f"rank {rank}: sharded LearnStates rows differ from the oracle's")
    w2 = net2.GetMatrix()
    if not np.allclose(w2, o2.w, rtol=1e-12, atol=0):
        bad.append(f"rank {rank}: sharded LearnStates W differs from the oracle's beyond 1e-12")
    wsum = torch.tensor(np.frombuffer(w2.tobytes(), dtype=np.int64).sum(dtype=np.int64).reshape(1), device="cuda")
    wall = [torch.zeros_like(wsum) for _ in range(world)]
    dist.all_gather(wall, wsum)
    if len({int(x.item()) for x in wall}) != 1:
        bad.append("replicas of W differ after sharded LearnStates")
    capi.check(capi.load().hn_comm_detach(net2._h))
    net2.close()
    # (3) sharded probes, unique tables merged on rank 0's GPU
    onet.normalize()
    onet.set_targets(t)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(t))
    probes = np.concatenate([rng.generate_states(0, n, b - 2 * p), t, -t])   # repeats: the targets and their inverses
    pool = rng.perm_pool(n, 100)
    plo, phi = multigpu.shard_range(probes.shape[0], rank, world)
    res = net.relax_batch(probes[plo:phi], pool, dedup=True)
    merged = multigpu.gather_unique_tables_device(net, plo)
    ores = onet.relax_batch(probes.copy(), pool, threads=4)
    if not np.array_equal(res.FinalStates, orc.pack(ores["final"][plo:phi])) or not np.array_equal(res.NumSteps, ores["num_steps"][plo:phi]):
        bad.append(f"rank {rank}: relaxation of the probe shard differs from the oracle")
    if rank == 0:
        _, first, hits = orc.unique_table(ores["energies"])
        if merged["first"].tolist() != first.tolist() or merged["hits"].tolist() != hits.tolist():
            bad.append("merged unique table differs from the oracle's single-stream table")
    net.close()
    flag = torch.tensor([len(bad)], device="cuda", dtype=torch.int64)
    dist.all_reduce(flag)
    msgs = [None] * world
    dist.all_gather_object(msgs, bad)
    if int(flag.item()) == 0:
        return "ok"
    return "; ".join(m for ms in msgs for m in ms)[:500]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--probes", type=int, default=1 << 17, help="probes per GPU per step")
    ap.add_argument("--dim", type=int, default=4096)
    ap.add_argument("--targets", type=int, default=400)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
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

    import ctypes as C
    import torch
    import torch.distributed as dist
    from hopfield_network_go_b200 import capi, multigpu
    from hopfield_network_go_b200 import hopfieldnetwork as hn

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    parity = None
    if world > 1:
        try:
            parity = multi_gpu_parity(hn, dist, rank, world, local)
        except Exception as e:   # a crash here must not hide the timing; it is reported as the parity result
            parity = f"exception on rank {rank}: {type(e).__name__}: {e}"[:300]
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
    dev_ms, stage = [], {}
    launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev_ms.append(net.relax_resident(dedup=True))
        t = net.last_timing()
        if os.environ.get("HN_MERGE_TRACE") and rank == 0:
            print(f"[resident trace] device {dev_ms[-1]:.2f}: dense {t.dense_ms:.2f} fields {t.fields_ms:.2f} relax {t.relax_ms:.2f} post {t.post_ms:.2f}", flush=True)
        for k in ("fields_ms", "relax_ms", "post_ms", "dense_ms", "dense_kernel_ms", "dense_gemm_ms", "dense_scan_ms"):
            stage.setdefault(k, []).append(getattr(t, k, 0.0))
        dense_sweeps, dense_flips = t.dense_sweeps, t.dense_flips
        launches += t.launches
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    # counters for the roofline (same every step: deterministic)
    out = capi.HnRelaxOut()
    capi.check(capi.load().hn_download_results(net._h, C.byref(out)))
    flips, sweeps = out.total_flips, out.total_sweeps
    # the share of the sparse kernel (the dominant one): what the dense lockstep sweeps did not do.  Every probe of this
    # workload is still unstable after the dense sweeps, so each of them counts dense_sweeps sweeps there.
    sp_flips, sp_sweeps = flips - dense_flips, sweeps - b * dense_sweeps
    alg_bytes = 8 * n * sp_flips + 8 * n * sp_sweeps + b * (n // 4 + 4 * p + 16)
    alg_bytes_all = 8 * n * flips + 8 * n * sweeps + b * (n // 4 + 4 * p + 16)
    step_ms = float(np.sum(dev_ms))
    # e2e: host buffers through hn_relax_batch (+ the cross-GPU merge of the unique tables), every step timed
    e2e_t = []
    n_unique_global = None
    for it in range(1 + args.steps):
        barrier()
        t1 = time.perf_counter()
        res = net.relax_batch(probes_np, pool, dedup=True, reuse_pinned_buffers=True)   # page-locked result buffers of the network
        if os.environ.get("HN_MERGE_TRACE"):
            tt = net.last_timing()
            print(f"[e2e trace] rank {rank} relax_batch {1e3 * (time.perf_counter() - t1):.2f} ms (device {res.device_ms:.2f}: dense {tt.dense_ms:.2f} "
                  f"fields {tt.fields_ms:.2f} relax {tt.relax_ms:.2f} post {tt.post_ms:.2f})", flush=True)
        if world > 1:   # per-GPU unique tables merged on rank 0's GPU (first-seen order, hits)
            merged = multigpu.gather_unique_tables_device(net, rank * b)
            if rank == 0:
                n_unique_global = len(merged["first"])
        if os.environ.get("HN_MERGE_TRACE"):
            print(f"[e2e trace] rank {rank} before the closing barrier {1e3 * (time.perf_counter() - t1):.2f} ms", flush=True)
        barrier()
        if it > 0:
            e2e_t.append(time.perf_counter() - t1)
    h2d = probes_np.nbytes + pool.nbytes
    d2h = res.FinalStates.nbytes + res.NumSteps.nbytes + res.Stable.nbytes + res.DistancesToTargets.nbytes + \
        res.Flips.nbytes + res.MarginHits.nbytes + res.AttractorId.nbytes + 2 * 4 * len(res.UniqueFirst) + res.StateHash.nbytes
    verified, refuted = net.unique_table_stats()
    # the same step with the other column streams (bit-identical results): reported beside the headline, never as it
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
    relax_total = float(np.sum(stage["relax_ms"]))
    tmax = torch.tensor([step_ms, relax_total, float(np.mean(e2e_t)), alt_ms["f64"], alt_ms["f32"], alt_ms["i16"],
                         alt_relax_ms["f64"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    step_ms_max, relax_ms_max, e2e_max, f64_ms_max, f32_ms_max, i16_ms_max, f64_relax_ms_max = [float(x) for x in tmax.cpu()]
    delta = {}
    if not args.no_delta:   # collective when world > 1: every rank takes part
        if world == 1:
            # config 2 runs 100 epochs: all of them are timed (the sweep's flips shrink as the network learns its targets)
            delta["config2_N1024_P100_inversion0.1"] = measure_delta_epoch(hn, local, 1024, 100, 100, cpu=not args.no_cpu)
            delta["config5_N16384_P1024_one_gpu"] = measure_delta_epoch(hn, local, 16384, 1024, 2, cpu=False)
        delta[f"config5_N16384_P{128 * world}_sharded_128_per_gpu"] = measure_delta_epoch(
            hn, local, 16384, 128, 2, dist=dist if world > 1 else None, rank=rank, world=world, cpu=(world == 1 and not args.no_cpu))
    if rank == 0:
        peaks = load_json("MEASURED_PEAKS.json")
        own = load_json("profiles", "r02_peaks.json")
        hbm = peaks.get("hbm_gbs", 6650.0)
        achieved = alg_bytes * args.steps / (relax_ms_max * 1e-3) / 1e9
        col_bytes = {"f64": 8, "f32": 4, "i16": 2}[stream]
        field_bytes = 4 if stream == "i16" else 8
        streamed = col_bytes * n * sp_flips + field_bytes * n * sp_sweeps     # what the sparse kernel moves L2->SM per launch
        kernel = {"f64": "relax_kernel<32,double,128,4>", "f32": "relax_kernel<32,float,128,4>", "i16": "relax_kernel<32,short,128,9>"}[stream]
        # measured L2->SM read bandwidth with this kernel's access shape (tools/peaks_microbench.cu -> profiles/r02_peaks.json)
        def l2_peak(mib):
            c = [e["gbs"] for e in own.get("l2_read", []) if e["matrix_mib"] == mib and not e["lockstep"] and e["ctas_per_sm"] == 9]
            return c[0] if c else None
        l2pk = l2_peak(32 if stream == "i16" else 128)
        issue = load_json("profiles", "r02_relax_issue.json").get(stream, {})
        line = {"metric": METRIC, "value": world * b * args.steps / (step_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64" if stream != "i16" else "f64 (decisions taken on an exact int32 image of the float64 fields)",
                "data": "synthetic",
                "config": {"workload": cfg["name"], "N": n, "P": p, "probes_per_gpu": b, "column_stream": f"{args.columns} -> {stream}",
                           "l2": "inputs larger than L2 (W 128 MiB + 2 GiB of initial fields streamed per step)"},
                "e2e": {"value": world * b / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "timed_steps": len(e2e_t),
                        "note": "hn_relax_batch with page-locked host buffers; with n_gpus > 1 the per-GPU unique tables are "
                                "gathered over NCCL and merged on rank 0's GPU (hn_merge_unique_device) inside the timed region"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                             "traffic": None, "kernel": kernel,
                             "note": "per GPU, the dominant kernel (sparse relaxation, after the dense lockstep sweeps).  SURVEY 8(d)'s "
                                     "line: ALGORITHMIC bytes of the reference algorithm for the flips and sweeps this kernel performs (one "
                                     "float64 column per flip + one pass over the fields per sweep) over its time, against the measured HBM "
                                     "copy peak.  HBM does not bind this path (traffic << algorithmic bytes: the weight image is L2 "
                                     "resident and narrower than float64), so the fraction exceeds 1; roofline_l2 and issue_frac are the "
                                     "fractions of the resources that do bind, roofline_dense describes the tensor-core phase",
                             "peak_source": "measured" if peaks else "fallback",
                             "algorithmic_bytes_per_launch": int(alg_bytes), "algorithmic_bytes_whole_relaxation": int(alg_bytes_all),
                             "whole_relaxation_gbs": alg_bytes_all * args.steps / ((relax_ms_max + float(np.sum(stage["dense_ms"]))) * 1e-3) / 1e9,
                             "flips_per_probe": flips / b, "sweeps_per_probe": sweeps / b,
                             "flips_per_probe_in_this_kernel": sp_flips / b, "sweeps_per_probe_in_this_kernel": sp_sweeps / b},
                "roofline_l2": {"bound": "l2_to_sm", "achieved": streamed * args.steps / (relax_ms_max * 1e-3) / 1e9, "peak": l2pk,
                                "unit": "GB/s", "streamed_bytes_per_launch": int(streamed),
                                "peak_source": "profiles/r02_peaks.json: tools/peaks_microbench.cu, independent pseudo-random rows "
                                               "of the same width from a matrix of the same size, one 128-bit load per thread, 9 CTAs/SM",
                                "note": "bytes the column stream really moves L2->SM (2N per flip + 4N per sweep on the integer "
                                        "stream) over the relaxation time; with the dense first sweeps the tensor-core phase moves "
                                        "its weight tiles by TMA and is not counted here"},
                "issue_frac": issue.get("issue_slots_busy_frac"),
                "issue_frac_source": issue.get("source"),
                "roofline_dense": dense_roofline(n, b, dense_sweeps, dense_flips, stage, own, load_json("profiles", "r02_relax_issue.json")),
                "clocks": sampler.summary(), "wall_s": wall,
                "stage_ms_per_step": {"fields_gemm": float(np.mean(stage["fields_ms"])), "relax": float(np.mean(stage["relax_ms"])),
                                      "dense_sweeps": float(np.mean(stage["dense_ms"])),
                                      "dense_sweep_kernel": float(np.mean(stage["dense_kernel_ms"])),
                                      "dense_fields_gemm": float(np.mean(stage["dense_gemm_ms"])),
                                      "dense_stability_scan": float(np.mean(stage["dense_scan_ms"])),
                                      "dedup": float(np.mean(stage["post_ms"]))},
                "dense_sweeps_per_step": int(dense_sweeps),
                "margin_hits_per_probe": out.total_margin_hits / b,
                "unique_attractors": int(n_unique_global if n_unique_global is not None else out.n_unique),
                "unique_table": {"different_state_merges_verified": verified, "refuted": refuted}}
        if line["roofline_l2"]["peak"]:
            line["roofline_l2"]["frac"] = line["roofline_l2"]["achieved"] / line["roofline_l2"]["peak"]
        if parity is not None:
            line["multi_gpu_parity"] = parity
        for name, ms in (("f64", f64_ms_max), ("f32", f32_ms_max), ("i16", i16_ms_max)):
            if ms > 0:
                line[f"value_{name}_column_stream"] = world * b / (ms * 1e-3)
        traffic = load_json("profiles", "r02_relax_traffic.json") or load_json("profiles", "r01_relax_traffic.json")
        tr = traffic.get(stream, {})
        if tr.get("probes_per_launch") == b and tr.get("N") == n:
            line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
            line["roofline"]["traffic_source"] = tr["source"]
        if stream != "f64" and f64_relax_ms_max > 0:   # the float64 stream's own rooflines beside the headline
            ach64 = alg_bytes / (f64_relax_ms_max * 1e-3) / 1e9
            t64 = traffic.get("f64", {})
            pk64 = l2_peak(128)
            line["roofline_f64_stream"] = {"bound": "hbm", "achieved": ach64, "peak": hbm, "unit": "GB/s", "frac": ach64 / hbm,
                                           "traffic": t64.get("dram_bytes_per_launch") if t64.get("probes_per_launch") == b and t64.get("N") == n else None,
                                           "kernel": "relax_kernel<32,double,128,4>",
                                           "l2_to_sm_frac": (ach64 / pk64) if pk64 else None, "l2_to_sm_peak": pk64,
                                           "note": "float64 columns: the streamed bytes ARE the algorithmic bytes (8N per flip); W (128 MiB) "
                                                   "exceeds L2, so the stream is served partly from DRAM"}
        if not args.no_delta:
            line["delta_epoch"] = delta
            line["delta_epoch_ms"] = {k: v["ms"] for k, v in delta.items()}
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
                                    "sample": f"first {sample} probes of the same batch (2 per worker thread), oracle C restatement "
                                              "of the reference algorithm as written"}
        print(json.dumps(line))
    net.close()
    if world > 1:
        dist.destroy_process_group()


def dense_roofline(n, b, sweeps, dense_flips, stage, own, ncu):
    """The tensor-core phase: `sweeps` lockstep sweeps per step, each an N x N x B int8 contraction computed block by block
    (32 positions x 32 probes, K cut in four quarters to fill the 128-row MMA: one quarter of the issued MACs are fields),
    plus the exact fields contraction (N x N x B) after every sweep."""
    if not sweeps:
        return None
    km, gm = float(np.mean(stage["dense_kernel_ms"])), float(np.mean(stage["dense_gemm_ms"]))
    useful = 2.0 * n * n * b * sweeps
    pk = {e["n"]: e["tops"] for e in own.get("tcgen05_i8", [])}
    tiles = (b + 31) // 32
    np_, rp = (n + 511) // 512 * 512, (n + 31) // 32 * 32
    stream = float(tiles) * rp * np_ * sweeps
    tma = max([e["gbs"] for e in own.get("tma_stream", [])], default=None)
    d = ncu.get("dense_sweep", {})
    return {"bound": "tensor", "unit": "TOP/s", "kernel": "dense_sweep_kernel",
            "achieved_useful": useful / (km * 1e-3) / 1e12, "achieved_issued": 4 * useful / (km * 1e-3) / 1e12,
            "peak": pk.get(128) or pk.get(256), "peak_source": "profiles/r02_peaks.json: tcgen05.mma.kind::i8 M=128 N=128 K=32 issue rate on resident tiles",
            "frac_issued": (4 * useful / (km * 1e-3) / 1e12) / (pk.get(128) or pk.get(256)) if (pk.get(128) or pk.get(256)) else None,
            "tensor_pipe_active_frac": d.get("tensor_pipe_active_frac"), "tensor_pipe_source": d.get("source"),
            "weight_stream": {"bytes_per_step": stream, "gbs": stream / (km * 1e-3) / 1e9, "tma_stream_peak_gbs": tma,
                              "note": "every tile of 32 probes streams the whole byte image (N x N) through TMA once per sweep; L2 resident"},
            "flips_per_probe": dense_flips / b, "sweeps": int(sweeps),
            "fields_gemm": {"kernel": "igemm_tc05_kernel<true>", "top_s": useful / (gm * 1e-3) / 1e12 if gm > 0 else None, "peak": pk.get(256),
                            "frac": (useful / (gm * 1e-3) / 1e12) / pk[256] if gm > 0 and 256 in pk else None,
                            "note": "B x N x N per dense sweep with the stability scan fused into the epilogue; the int32 fields "
                                    "(2 GiB at config 3) are stored only after the last dense sweep.  ncu: neither L2 nor the tensor "
                                    "pipe saturated — the bits -> int8 expansion of the probe tile bounds a K step (DESIGN.md K1)"}}


def measure_delta_epoch(hn, local, n, p, epochs, noise_scale=0.1, dist=None, rank=0, world=1, cpu=False):
    """Second BASELINE metric: Delta-rule epoch time = one pass of `delta` (LearningRule.go:90-114) over all targets +
    the per-epoch target diagnostics of fullSetLearningMethod (LearningMethod.go:45-58: EnergyProfile rows, Stable).
    Wall clock through hn_delta_epoch + hn_energy_profiles with HOST buffers, the P x N float64 energy rows included;
    noise / update orders are drawn host-side before the timed region.  With world > 1 the p*world targets are sharded
    (p per GPU) and the library all-reduces dW over NCCL inside hn_delta_epoch.  Returns the time, the
    device-time breakdown, a roofline per stage and (cpu=True) the oracle's time for the same epoch."""
    from hopfield_network_go_b200 import capi
    rs = np.random.Generator(np.random.PCG64(0x5EED + 2))
    w = (n + 63) // 64
    ptot = p * world
    targets = rs.integers(0, 2**63, size=(ptot, w), dtype=np.int64).astype(np.uint64) << np.uint64(1) | \
        rs.integers(0, 2, size=(ptot, w), dtype=np.int64).astype(np.uint64)
    net = hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(1).SetNetworkLearningRule(
        hn.DeltaLearningRule).SetNetworkLearningMethod(hn.FullSetMethod).SetSeed(2).SetDevice(local).Build()
    if world > 1:
        attach_comm(capi, net, dist, rank, world)
    lo, hi = rank * p, (rank + 1) * p
    k = int(n * noise_scale)
    times, stages = [], []
    last_inputs = None
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
        net.energy_profiles(targets[lo:hi], want_energies=True)
        dt = time.perf_counter() - t0
        if ep >= 2:
            times.append(dt * 1e3)
            lt = net.last_learn_timing()
            stages.append({f: getattr(lt, f) for f in ("sweep_fields_ms", "sweep_ms", "dw_gemm_ms", "allreduce_ms", "update_ms",
                                                       "energy_gemm_ms", "energy_epilogue_ms", "sweep_flips")})
        last_inputs = (noised, perms)
    ms = float(np.mean(times))
    if world > 1:
        import torch
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        capi.check(capi.load().hn_comm_detach(net._h))
    st = {f: float(np.mean([s[f] for s in stages])) for f in stages[0]}
    own = load_json("profiles", "r02_peaks.json")
    peaks = load_json("MEASURED_PEAKS.json")
    fp64_peak = own.get("dmma_sync_tflops")
    hbm = peaks.get("hbm_gbs", 6650.0)
    gemm_flop = 2.0 * n * n * p
    tf = lambda t_ms: gemm_flop / (t_ms * 1e-3) / 1e12 if t_ms > 0 else None
    sweep_bytes = 8.0 * n * st["sweep_flips"] + p * (8.0 * n + n / 4)       # one float64 column per flip + one pass over the fields
    res = {"ms": ms, "epochs_timed": len(times), "ms_first_epoch": times[0], "ms_last_epoch": times[-1], "targets_per_gpu": p, "n_gpus": world, "device_stage_ms": {k2: v for k2, v in st.items() if k2 != "sweep_flips"},
           "sweep_flips_per_target": st["sweep_flips"] / p,
           "roofline": {"fp64_tensor": {"bound": "tensor", "unit": "TFLOP/s (equivalent: 2 N^2 P per contraction)", "peak": fp64_peak,
                                        "peak_source": "profiles/r02_peaks.json: mma.sync.m8n8k4.f64 issue rate (tcgen05 has no f64 kind)",
                                        "note": "the reference's contractions are float64; with exact-arithmetic weights (every Delta / Hebbian "
                                                "epoch from a zero matrix) the library computes all three as exact int8 tensor-core "
                                                "contractions (tcgen05), so the equivalent rate exceeds the FP64 pipe's peak",
                                        "sweep_fields_gemm": tf(st["sweep_fields_ms"]), "dw_gemm": tf(st["dw_gemm_ms"]),
                                        "energies_gemm": tf(st["energy_gemm_ms"]), "flop_per_gemm": gemm_flop},
                        "sweep": {"bound": "hbm", "unit": "GB/s", "peak": hbm,
                                  "achieved": sweep_bytes / (st["sweep_ms"] * 1e-3) / 1e9 if st["sweep_ms"] > 0 else None,
                                  "algorithmic_bytes": sweep_bytes},
                        "update": {"bound": "hbm", "unit": "GB/s", "peak": hbm,
                                   "achieved": 5 * 8.0 * n * n / (st["update_ms"] * 1e-3) / 1e9 if st["update_ms"] > 0 else None,
                                   "note": "W += lr dW (2 reads + 1 write) then zero diagonal + symmetrise (1 read + 1 write): 40 N^2 bytes"}},
           "note": "ms = wall clock (max over ranks) with host buffers: H2D of packed targets, noised copies and permutations, "
                   "D2H of the P x N float64 EnergyProfile rows; device_stage_ms = CUDA events inside the library"}
    for g in ("sweep_fields_gemm", "dw_gemm", "energies_gemm"):
        v = res["roofline"]["fp64_tensor"][g]
        if v and fp64_peak:
            res["roofline"]["fp64_tensor"][g + "_frac"] = v / fp64_peak
    if cpu and rank == 0:
        # the oracle's delta epoch + per-target diagnostics, single-threaded like LearningRule.go:90-114; on a bounded
        # number of targets where a full epoch would take minutes, scaled linearly (each target costs the same)
        from oracle import orc
        ps = p if n <= 2048 else 2
        onet = orc.Net(n)
        onet.set_weights(net.GetMatrix())          # the same (un-normalised) weights the timed epochs ran on
        t_f = orc.unpack(targets[lo:lo + ps], n)
        nz = orc.unpack(last_inputs[0][:ps], n)
        t2 = time.perf_counter()
        onet.delta(t_f, nz, last_inputs[1][:ps])
        for s in range(ps):
            onet.all_unit_energies(t_f[s])
            onet.state_is_stable(t_f[s])
        dt = (time.perf_counter() - t2) * 1e3 * (p / ps)
        res["cpu_baseline"] = {"value": dt, "unit": "ms", "cores": 1, "kind": "port",
                               "sample": f"{ps} of the {p} targets of one epoch (scaled by {p / ps:g}), oracle C restatement of delta + "
                                         "AllUnitEnergies + StateIsStable per target, single-threaded as the reference's learning loop is"}
    net.close()
    return res


if __name__ == "__main__":
    main()
