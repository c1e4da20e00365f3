"""Small config-3 case for profiling the dense lockstep sweeps: N=4096, Hebbian P=400, one tile per SM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import synth
from hopfield_network_go_b200 import hopfieldnetwork as hn

b = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 32
columns = {"auto": 3, "f64": 0, "f32": 1, "i16": 2}[sys.argv[2] if len(sys.argv) > 2 else "auto"]   # hn_config.column_stream
targets, probes, pool = synth(4096, 400, b, 0x5EED + 3)
net = hn.NewHopfieldNetworkBuilder().SetNetworkDimension(4096).SetEpochs(1).SetNetworkLearningRule(hn.HebbianLearningRule).SetNetworkLearningMethod(
    hn.FullSetMethod).SetSeed(1).SetColumnStream(columns).Build()
net.LearnStates(targets)
net.upload_perm_pool(pool)
net.upload_probes(probes)
for _ in range(2):
    ms = net.relax_resident(dedup=True)
t = net.last_timing()
print(f"B={b} total {ms:.2f} ms dense {t.dense_ms:.2f} ms ({t.dense_sweeps} sweeps) relax {t.relax_ms:.2f} ms")
net.close()
