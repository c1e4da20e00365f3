"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): learning (Hebbian + Delta), relaxation in the
three column-stream modes with dedup + history, energy profiles.  Checks results against the oracle as well."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopfield_network_go_b200 import hopfieldnetwork as hn
from oracle import orc

n, p, b = 200, 12, 24
rng = orc.Rng(5)
t = rng.generate_states(0, n, p)
probes = rng.generate_states(0, n, b)
pool = rng.perm_pool(n, 100)
for mode in (0, 1, 2):
    net = (hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(5).SetNetworkLearningRule(hn.DeltaLearningRule)
           .SetNetworkLearningMethod(hn.FullSetMethod).SetLearningNoiseMethod(hn.MaximalInversion).SetLearningNoiseRatio(0.1)
           .SetSeed(3).SetColumnStream(mode).Build())
    net.LearnStates(t.copy())
    onet = orc.Net(n, w=net.GetMatrix()); onet.set_targets(t)
    ores = onet.relax_batch(probes.copy(), pool)
    res = net.relax_batch(probes, pool, dedup=True, history=True, want_energies=True)
    assert np.array_equal(res.FinalStates, orc.pack(ores["final"])) and np.array_equal(res.NumSteps, ores["num_steps"])
    net.energy_profiles(probes)
    net.hebbian_epoch(t)
    net.close()
print("SANITIZER_CASE_OK")
