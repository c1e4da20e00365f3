"""-m gpu: the experiment driver (tools/hopfield_main.py, mirror of main.go) writes the reference's six parquet
outputs with the datacollector row schemas; contents agree with the oracle run on the same seeded inputs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flag_parsing_matches_main_go_defaults():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import hopfield_main
    a = hopfield_main.parse([])
    assert (a.dimension, a.numTargetStates, a.numProbeStates, a.epochs, a.threads, a.learningRule, a.domain) == (100, 1, 1000, 100, 1, 0, 0)
    assert a.forceSymmetric is True and a.dataDir == "data/hopfieldData"     # main.go:33-47,58-59
    b = hopfield_main.parse(["-dimension", "64", "-learningRule", "2", "-forceSymmetric=false", "-allowIntensiveDataCollection"])
    assert b.dimension == 64 and b.learningRule == 2 and b.forceSymmetric is False and b.allowIntensiveDataCollection is True


@pytest.mark.gpu
def test_cli_end_to_end(built, tmp_path):
    import pyarrow.parquet as pq
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import hopfield_main
    from hopfield_network_go_b200 import hopfieldnetwork as hn
    from oracle import orc
    d = str(tmp_path / "data")
    n, p, b, seed = 100, 10, 200, 77
    assert hopfield_main.main(["-dimension", str(n), "-numTargetStates", str(p), "-numProbeStates", str(b), "-epochs", "100",
                               "-dataDir", d, "-seed", str(seed), "-allowIntensiveDataCollection"]) == 0
    files = sorted(os.listdir(d))
    for f in ("learnStateData.pq", "networkSummary.pq", "relaxationHistory.pq", "relaxationResult.pq", "targetStateProbe.pq", "uniqueStates.pq"):
        assert f in files
    rr = pq.read_table(os.path.join(d, "relaxationResult.pq"))
    assert rr.column_names == ["StateIndex", "Stable", "NumSteps", "FinalState", "DistancesToTargets", "EnergyProfile"]
    us = pq.read_table(os.path.join(d, "uniqueStates.pq"))
    assert us.column_names == ["StateIndex", "Stable", "FinalState", "DistancesToTargets", "EnergyProfile", "Hits"]
    sm = pq.read_table(os.path.join(d, "networkSummary.pq")).to_pylist()[0]
    assert sm["NetworkDomain"] == "BipolarDomain" and sm["LearningRule"] == "HebbianLearningRule" and sm["ProbeStates"] == b
    # oracle on the same inputs (same generator stream, same schedule stream)
    gen = orc.Rng(seed + 1)
    t = gen.generate_states(0, n, p)
    onet = orc.Net(n)
    onet.learn_states(t.copy(), 0, 0, 100, 0, 0.0, orc.Rng(seed))
    probes = gen.generate_states(0, n, b)
    sched = orc.Rng(seed)            # Hebbian learning draws nothing: the schedule is the first draw of the network RNG
    pool = sched.perm_pool(n, 100)
    from hopfield_network_go_b200 import gonumio
    assert "matrix.bin" in files and "targetStates.bin" in files                       # main.go:187-188
    assert np.array_equal(gonumio.LoadVectorCollection(os.path.join(d, "targetStates.bin")), t)
    onet.set_weights(gonumio.LoadMatrix(os.path.join(d, "matrix.bin")))   # identical bits for the trajectories
    ores = onet.relax_batch(probes.copy(), pool, threads=4)
    assert np.array_equal(np.array(rr["FinalState"].to_pylist()), ores["final"])
    assert rr["NumSteps"].to_pylist() == ores["num_steps"].tolist()
    assert rr["Stable"].to_pylist() == [bool(x) for x in ores["stable"]]
    assert np.array_equal(np.array(rr["DistancesToTargets"].to_pylist()), ores["distances"])
    _, first, hits = orc.unique_table(ores["energies"])
    assert us["StateIndex"].to_pylist() == first.tolist() and us["Hits"].to_pylist() == hits.tolist()
    hist = pq.read_table(os.path.join(d, "relaxationHistory.pq"))
    assert hist.num_rows == int(ores["num_steps"].sum())


@pytest.mark.gpu
def test_cli_loads_state_files(built, tmp_path):
    """-targetStatesFile / -probeStatesFile (main.go:46-48,166-177,211-219) override random generation."""
    import pyarrow.parquet as pq
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import hopfield_main
    from hopfield_network_go_b200 import gonumio
    from oracle import orc
    n, p, b = 64, 5, 40
    gen = orc.Rng(5)
    t, probes = gen.generate_states(0, n, p), gen.generate_states(0, n, b)
    gonumio.SaveVectorCollection(t, str(tmp_path / "t.bin"))
    gonumio.SaveVectorCollection(probes, str(tmp_path / "p.bin"))
    d = str(tmp_path / "out")
    assert hopfield_main.main(["-dimension", str(n), "-numTargetStates", "99", "-numProbeStates", "7", "-dataDir", d, "-seed", "3",
                               "-targetStatesFile", str(tmp_path / "t.bin"), "-probeStatesFile", str(tmp_path / "p.bin")]) == 0
    sm = pq.read_table(os.path.join(d, "networkSummary.pq")).to_pylist()[0]
    assert sm["TargetStates"] == p and sm["ProbeStates"] == b                     # counts come from the files
    assert np.array_equal(gonumio.LoadVectorCollection(os.path.join(d, "targetStates.bin")), t)
    onet = orc.Net(n)
    onet.set_weights(gonumio.LoadMatrix(os.path.join(d, "matrix.bin")))
    onet.set_targets(t)
    ores = onet.relax_batch(probes.copy(), orc.Rng(3).perm_pool(n, 100), threads=4)
    rr = pq.read_table(os.path.join(d, "relaxationResult.pq"))
    assert np.array_equal(np.array(rr["FinalState"].to_pylist()), ores["final"])
    assert rr["NumSteps"].to_pylist() == ores["num_steps"].tolist()
