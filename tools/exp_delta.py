"""Per-epoch wall time of hn_delta_epoch at the config-5 shape (N=16384, P=128), to see what varies between epochs."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hopfield_network_go_b200 import hopfieldnetwork as hn, capi
import ctypes as C

n, p = int(os.environ.get("N", 16384)), int(os.environ.get("P", 128))
rs = np.random.Generator(np.random.PCG64(0x5EED + 2))
w = (n + 63) // 64
targets = rs.integers(0, 2**63, size=(p, w), dtype=np.int64).astype(np.uint64) << np.uint64(1) | rs.integers(0, 2, size=(p, w), dtype=np.int64).astype(np.uint64)
net = hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(1).SetNetworkLearningRule(hn.DeltaLearningRule).SetNetworkLearningMethod(hn.FullSetMethod).SetSeed(2).Build()
k = int(n * 0.1)
for ep in range(8):
    noised = targets.copy()
    for s in range(p):
        idx = rs.permutation(n - 1)[:k]
        np.bitwise_xor.at(noised[s], idx // 64, np.uint64(1) << (idx % 64).astype(np.uint64))
    perms = np.stack([rs.permutation(n).astype(np.uint16) for _ in range(p)])
    t0 = time.perf_counter()
    rel = net.delta_epoch(targets, noised, perms)
    t1 = time.perf_counter()
    e, uns, stab, _ = net.energy_profiles(targets, want_energies=False)
    t2 = time.perf_counter()
    tb, tf = C.c_double(), C.c_double()
    capi.check(capi.load().hn_get_margin(net._h, C.byref(tb), C.byref(tf)))
    diff = int(np.unpackbits((rel ^ targets).view(np.uint8)).sum())
    print(f"epoch {ep}: delta {1e3*(t1-t0):7.2f} ms  energies {1e3*(t2-t1):6.2f} ms  tau_base {tb.value:.3e}  relaxed!=target bits {diff}  stable {int(stab.sum())}/{p}")
net.close()
