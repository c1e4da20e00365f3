#!/usr/bin/env python
"""hopfield_main.py — the reference's experiment driver (main.go:124-287) on top of the B200 package.

Same flags (Go style, single dash), same phases (learn -> save -> analyse targets -> probe -> dump), same six
parquet outputs with the row structs of hopfieldnetwork/datacollector/*.go (README.md:18-137 of the reference):
relaxationResult.pq, targetStateProbe.pq, uniqueStates.pq, learnStateData.pq, networkSummary.pq and, with
-allowIntensiveDataCollection, relaxationHistory.pq.  SURVEY 8f #1.

Differences, all deliberate: additive flags -seed / -device; runs are reproducible; `-threads` is recorded in the
summary but compute runs on the GPU; REPEATED double columns are written as parquet LIST columns (pyarrow);
matrix.bin / targetStates.bin and -targetStatesFile / -probeStatesFile use gonum's binary encoding as restated in
hopfield_network_go_b200/gonumio.py (unverified against a Go build)."""
from __future__ import annotations

import argparse
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DOMAIN_NAMES = ["BipolarDomain", "BinaryDomain"]                        # domain/domainenum_string.go
RULE_NAMES = ["HebbianLearningRule", "BipolarMappedHebbianLearningRule", "DeltaLearningRule",
              "BipolarMappedDeltaLearningRule", "ThermalDeltaLearningRule", "BipolarMappedThermalDeltaLearningRule"]
NOISE_NAMES = ["None", "MaximalInversion", "RandomSubMaximalInversion", "GaussianApplication"]


def go_bool(v):
    return str(v).lower() in ("1", "t", "true", "yes")


def parse(argv=None):
    ap = argparse.ArgumentParser(prefix_chars="-", description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    a = ap.add_argument                                                   # main.go:31-63
    a("-forceSymmetric", type=go_bool, default=True)
    a("-randomMatrixInit", type=go_bool, nargs="?", const=True, default=False)
    a("-domain", type=int, default=0)
    a("-dimension", type=int, default=100)
    a("-unitsUpdated", type=int, default=1)
    a("-learningMethod", type=int, default=0)
    a("-learningRule", type=int, default=0)
    a("-epochs", type=int, default=100)
    a("-numTargetStates", type=int, default=1)
    a("-numProbeStates", type=int, default=1000)
    a("-learningRate", type=float, default=1.0)
    a("-learningNoiseMethod", type=int, default=0)
    a("-learningNoiseScale", type=float, default=0.0)
    a("-threads", type=int, default=1)
    a("-dataDir", default="data/hopfieldData")
    a("-allowIntensiveDataCollection", type=go_bool, nargs="?", const=True, default=False)
    a("-verbose", type=go_bool, nargs="?", const=True, default=False)
    a("-targetStatesFile", default="")        # main.go:46
    a("-probeStatesFile", default="")         # main.go:48
    a("-seed", type=int, default=0x5EED)      # additive
    a("-device", type=int, default=0)         # additive
    return ap.parse_args(argv)


def main(argv=None):
    import pyarrow as pa
    import pyarrow.parquet as pq
    from hopfield_network_go_b200 import capi, gonumio
    from hopfield_network_go_b200 import hopfieldnetwork as hn

    args = parse(argv)
    n, domain = args.dimension, args.domain
    if os.path.isdir(args.dataDir):                                       # main.go:105-108
        shutil.rmtree(args.dataDir)
    os.makedirs(args.dataDir)
    log = (lambda *m: print(*m)) if args.verbose else (lambda *m: None)

    network = (hn.NewHopfieldNetworkBuilder().SetNetworkDomain(domain).SetNetworkDimension(n)
               .SetRandMatrixInit(args.randomMatrixInit).SetForceSymmetric(args.forceSymmetric)
               .SetNetworkLearningMethod(args.learningMethod).SetNetworkLearningRule(args.learningRule)
               .SetEpochs(args.epochs).SetMaximumRelaxationIterations(100).SetMaximumRelaxationUnstableUnits(0)
               .SetLearningRate(args.learningRate).SetLearningNoiseMethod(args.learningNoiseMethod)
               .SetLearningNoiseRatio(args.learningNoiseScale).SetUnitsUpdatedPerStep(args.unitsUpdated)
               .SetAllowIntensiveDataCollection(args.allowIntensiveDataCollection)
               .SetSeed(args.seed).SetDevice(args.device).Build())            # main.go:131-148
    generator = hn.Rand(args.seed + 1)                                    # states.StateGenerator, main.go:150-155
    low = 0.0 if domain == 1 else -1.0

    def f64(packed):
        return capi.unpack_states(packed, n, domain)

    def lists(rows):
        return pa.array([r.tolist() for r in rows], type=pa.list_(pa.float64()))

    # LEARNING PHASE (main.go:157-188)
    if args.targetStatesFile:                                             # main.go:166-177 (the file overrides generation)
        targets = capi.pack_states(np.ascontiguousarray(gonumio.LoadVectorCollection(args.targetStatesFile)))
        args.numTargetStates = len(targets)
    else:
        targets = generator.CreateStateCollection(n, args.numTargetStates)
    learn = network.LearnStates(targets, want_energies=True)
    pq.write_table(pa.table({
        "Epoch": pa.array([d.Epoch for d in learn], pa.int32()),
        "TargetStateIndex": pa.array([d.TargetStateIndex for d in learn], pa.int32()),
        "EnergyProfile": lists([d.EnergyProfile for d in learn]),
        "Stable": pa.array([d.Stable for d in learn], pa.bool_())}), os.path.join(args.dataDir, "learnStateData.pq"))
    gonumio.SaveMatrix(network.GetMatrix(), os.path.join(args.dataDir, "matrix.bin"))                 # main.go:187
    gonumio.SaveVectorCollection(f64(targets), os.path.join(args.dataDir, "targetStates.bin"))        # main.go:188

    # target analysis (main.go:190-204)
    e, _, stab, _ = network.energy_profiles(targets)
    pq.write_table(pa.table({
        "TargetStateIndex": pa.array(range(len(targets)), pa.int32()),
        "IsStable": pa.array(stab.astype(bool)),
        "State": lists(f64(targets)), "EnergyProfile": lists(e)}), os.path.join(args.dataDir, "targetStateProbe.pq"))
    log(f"learned {len(targets)} targets, {int(stab.sum())} stable")

    # PROBING PHASE (main.go:206-221)
    if args.probeStatesFile:                                              # main.go:211-219
        probes = capi.pack_states(np.ascontiguousarray(gonumio.LoadVectorCollection(args.probeStatesFile)))
        args.numProbeStates = len(probes)
    else:
        probes = generator.CreateStateCollection(n, args.numProbeStates)
    res = network.ConcurrentRelaxStates(probes, args.threads, dedup=True, want_energies=True,
                                        history=args.allowIntensiveDataCollection)

    # DATA PROCESSING (main.go:223-257)
    finals = f64(res.FinalStates)
    dist = res.DistancesToTargets.astype(np.float64)
    idx = np.arange(len(probes), dtype=np.int32)
    pq.write_table(pa.table({
        "StateIndex": pa.array(idx, pa.int32()), "Stable": pa.array(res.Stable.astype(bool)),
        "NumSteps": pa.array(res.NumSteps, pa.int32()), "FinalState": lists(finals),
        "DistancesToTargets": lists(dist), "EnergyProfile": lists(res.EnergyProfiles)}),
        os.path.join(args.dataDir, "relaxationResult.pq"))
    u = res.UniqueFirst                                                    # UniqueRelaxedStateHandler.go:41-73
    pq.write_table(pa.table({
        "StateIndex": pa.array(u, pa.int32()), "Stable": pa.array(res.Stable[u].astype(bool)),
        "FinalState": lists(finals[u]), "DistancesToTargets": lists(dist[u]),
        "EnergyProfile": lists(res.EnergyProfiles[u]), "Hits": pa.array(res.UniqueHits, pa.int32())}),
        os.path.join(args.dataDir, "uniqueStates.pq"))
    if args.allowIntensiveDataCollection:                                  # main.go:242-256
        si, st, states = [], [], []
        for i in range(len(probes)):
            k = int(res.NumSteps[i])
            si += [i] * k
            st += list(range(k))
            states.append(res.History[i, :k])
        hist_packed = np.concatenate(states) if states else np.zeros((0, capi.words(n)), np.uint64)
        he = network.energy_profiles(hist_packed)[0] if len(hist_packed) else np.zeros((0, n))
        pq.write_table(pa.table({"StateIndex": pa.array(si, pa.int32()), "StepIndex": pa.array(st, pa.int32()),
                                 "State": lists(f64(hist_packed)), "EnergyProfile": lists(he)}),
                       os.path.join(args.dataDir, "relaxationHistory.pq"))

    # summary (main.go:259-279; ForceZeroBias is never populated by the reference)
    pq.write_table(pa.table({
        "NetworkDomain": [DOMAIN_NAMES[domain]], "NetworkDimension": pa.array([n], pa.int32()),
        "LearningRule": [RULE_NAMES[args.learningRule]], "Epochs": pa.array([args.epochs], pa.int32()),
        "MaximumRelaxationIterations": pa.array([100], pa.int32()), "LearningRate": [args.learningRate],
        "LearningNoiseMethod": [NOISE_NAMES[args.learningNoiseMethod]], "LearningNoiseScale": [args.learningNoiseScale],
        "UnitsUpdated": pa.array([args.unitsUpdated], pa.int32()), "ForceSymmetricWeightMatrix": [bool(args.forceSymmetric)],
        "ForceZeroBias": [False], "Threads": pa.array([args.threads], pa.int32()),
        "TargetStates": pa.array([args.numTargetStates], pa.int32()), "ProbeStates": pa.array([args.numProbeStates], pa.int32())}),
        os.path.join(args.dataDir, "networkSummary.pq"))
    log(f"relaxed {len(probes)} probes: {int(res.Stable.sum())} stable, {len(u)} unique attractors; data in {args.dataDir}")
    network.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
