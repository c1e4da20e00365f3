"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them.  GPU (-m gpu): the CUDA path reproduces them through the C ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import orc

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    return dict(np.load(path))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(built, path):
    g = _load(path)
    n, domain = int(g["n"]), int(g["domain"])
    t = orc.unpack(g["targets"], n, domain)
    net = orc.Net(n, domain=domain)
    nrows, rows, _ = net.learn_states(t.copy(), int(g["rule"]), 0, int(g["epochs"]), int(g["noise"]),
                                      float(g["noise_scale"]), orc.Rng(int(g["learn_seed"])), cap=int(g["epochs"]) * len(t))
    assert np.array_equal(net.w, g["weights"])
    assert np.array_equal(np.array(rows, dtype=np.int32).reshape(-1, 3), g["learn_rows"])
    res = net.relax_batch(orc.unpack(g["probes"], n, domain), g["pool"], serial_stream=bool(g["serial"]))
    assert np.array_equal(orc.pack(res["final"]), g["final"])
    for k in ("num_steps", "stable", "flips", "margin_hits", "energies"):
        assert np.array_equal(res[k], g[k]), k
    assert np.array_equal(res["distances"].astype(np.int32), g["distances"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_reproduces_golden(built, path):
    from hopfield_network_go_b200 import hopfieldnetwork as hn
    g = _load(path)
    n, domain = int(g["n"]), int(g["domain"])
    net = (hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetNetworkDomain(domain).SetEpochs(int(g["epochs"]))
           .SetNetworkLearningRule(int(g["rule"])).SetNetworkLearningMethod(hn.FullSetMethod)
           .SetLearningNoiseMethod(int(g["noise"])).SetLearningNoiseRatio(float(g["noise_scale"]))
           .SetSeed(int(g["learn_seed"])).Build())
    lsd = net.LearnStates(g["targets"])
    assert np.array_equal(np.array([(d.Epoch, d.TargetStateIndex, int(d.Stable)) for d in lsd], dtype=np.int32).reshape(-1, 3),
                          g["learn_rows"])
    assert np.array_equal(net.GetMatrix(), g["weights"])                       # quarter-integer weights: the Frobenius norm is reproduced bit for bit
    net.SetMatrix(g["weights"])                                                # identical bits for the trajectories
    res = net.relax_batch(g["probes"], g["pool"], serial_stream=bool(g["serial"]), dedup=True, want_energies=True)
    assert np.array_equal(res.FinalStates, g["final"])
    assert np.array_equal(res.NumSteps, g["num_steps"])
    assert np.array_equal(res.Stable, g["stable"])
    assert np.array_equal(res.DistancesToTargets, g["distances"])
    assert np.array_equal(res.Flips, g["flips"])
    assert np.allclose(res.EnergyProfiles, g["energies"], rtol=1e-9, atol=1e-300)
    assert np.array_equal(res.UniqueFirst, g["unique_first"]) and np.array_equal(res.UniqueHits, g["unique_hits"])
    assert np.array_equal(res.AttractorId, g["attractor_id"])
    net.close()
