"""-m gpu: seeded random sweep over small configurations — dimension, target count, domain, learning rule / method,
noise method, epochs, symmetry / diagonal flags, column stream — each one through the whole public path
(LearnStates -> relax with dedup + energies) against the oracle.  Small sizes make ties, empty shards of the
iterative method, all-stable early exits and degenerate matrices frequent."""
import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from hopfield_network_go_b200 import hopfieldnetwork as hn
from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net

pytestmark = pytest.mark.gpu

NON_THERMAL = [hn.HebbianLearningRule, hn.BipolarMappedHebbianLearningRule, hn.DeltaLearningRule,
               hn.BipolarMappedDeltaLearningRule]


def _case(seed):
    rs = np.random.default_rng(seed)
    n = int(rs.choice([2, 3, 5, 8, 16, 17, 31, 33, 48, 64, 65, 70, 96, 130]))
    p = int(rs.integers(1, max(2, n // 3) + 1))
    return dict(n=n, p=p, domain=int(rs.integers(0, 2)), rule=int(rs.choice(NON_THERMAL)), method=int(rs.integers(0, 2)),
                epochs=int(rs.integers(1, 5)), noise=int(rs.integers(0, 4)), scale=float(rs.choice([0.0, 0.1, 0.3, 0.9])),
                sym=bool(rs.integers(0, 2)), zdiag=bool(rs.integers(0, 2)), stream=int(rs.choice([0, 1, 3])),
                max_unstable=int(rs.choice([0, 0, 1, 3])), max_iter=int(rs.choice([100, 100, 3])), b=int(rs.integers(1, 60)))


@pytest.mark.parametrize("seed", range(160))
def test_random_small_configuration(built, seed):
    c = _case(seed)
    n, p, domain = c["n"], c["p"], c["domain"]
    rng = orc.Rng(9000 + seed)
    t = rng.generate_states(domain, n, p)
    onet = orc.Net(n, domain=domain, force_symmetric=c["sym"], force_zero_diag=c["zdiag"], max_unstable=c["max_unstable"],
                   max_iter=c["max_iter"])
    nrows, rows, en = onet.learn_states(t.copy(), c["rule"], c["method"], c["epochs"], c["noise"], c["scale"],
                                        orc.Rng(seed), cap=4096, want_energies=True)
    net = gpu_net(n, domain, rule=c["rule"], method=c["method"], epochs=c["epochs"], seed=seed,
                  SetLearningNoiseMethod=c["noise"], SetLearningNoiseRatio=c["scale"], SetForceSymmetric=c["sym"],
                  SetForceZeroDiagonal=c["zdiag"], SetColumnStream=c["stream"],
                  SetMaximumRelaxationUnstableUnits=c["max_unstable"], SetMaximumRelaxationIterations=c["max_iter"])
    lsd = net.LearnStates(t.copy(), want_energies=True, cap=4096)
    assert len(lsd) == nrows, c
    assert [(d.Epoch, d.TargetStateIndex, d.Stable) for d in lsd] == rows, c
    w_o, w_g = onet.w, net.GetMatrix()
    if not np.all(np.isfinite(w_o)):                       # all-zero matrix normalised by a zero norm: NaN on both sides
        assert np.array_equal(np.isnan(w_o), np.isnan(w_g)), c
        net.close()
        return
    assert np.allclose(w_g, w_o, rtol=1e-9, atol=0), c
    for d, e in zip(lsd, en):
        assert np.allclose(d.EnergyProfile, e, rtol=1e-9, atol=1e-300), c
    # relaxation on the GPU's own weights (bit-identical inputs for both sides)
    onet.set_weights(w_g)
    probes = rng.generate_states(domain, n, c["b"])
    pool = rng.perm_pool(n, c["max_iter"])
    ores = onet.relax_batch(probes.copy(), pool, threads=2, want_energies=True)
    res = net.relax_batch(probes, pool, dedup=True, want_energies=True)
    assert_relax_equal(res, ores, n)
    assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-300), c
    _, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits), c
    net.close()


def _case2(seed):
    rs = np.random.default_rng(1000 + seed)
    n = int(rs.choice([200, 257, 300, 511, 513, 700, 1024, 1025, 1500, 2049]))
    return dict(n=n, p=int(rs.integers(2, max(3, n // 12))), domain=int(rs.integers(0, 2)),
                rule=int(rs.choice([hn.HebbianLearningRule, hn.DeltaLearningRule])), epochs=int(rs.integers(1, 4)),
                stream=int(rs.integers(0, 4)), mode=str(rs.choice(["plain", "offsets", "serial", "history"])),
                max_unstable=int(rs.choice([0, 0, 2])), b=int(rs.integers(2, 20 if n > 1024 else 48)))


@pytest.mark.parametrize("seed", range(40))
def test_random_medium_configuration(built, seed):
    """Every thread geometry up to N = 2049 (one, two and four warps per probe), every column stream, with offsets /
    the serial stream / history capture, learned in the library, against the oracle."""
    import ctypes as C
    c = _case2(seed)
    n, p, domain, b = c["n"], c["p"], c["domain"], c["b"]
    rng = orc.Rng(7000 + seed)
    t = rng.generate_states(domain, n, p)
    net = gpu_net(n, domain, rule=c["rule"], epochs=c["epochs"], seed=seed, SetColumnStream=c["stream"],
                  SetLearningNoiseMethod=hn.MaximalInversion, SetLearningNoiseRatio=0.1,
                  SetMaximumRelaxationUnstableUnits=c["max_unstable"])
    net.LearnStates(t.copy())
    onet = orc.Net(n, domain=domain, max_unstable=c["max_unstable"], w=net.GetMatrix())
    onet.set_targets(t)
    probes = rng.generate_states(domain, n, b)
    kw_o, kw_g = {}, {}
    pool = rng.perm_pool(n, 100)
    if c["mode"] == "offsets":
        pool = rng.perm_pool(n, 140)
        off = (np.arange(b) * 7 % 40).astype(np.int32)
        kw_o["offsets"] = off; kw_g["offsets"] = off
    elif c["mode"] == "serial":
        pool = rng.perm_pool(n, 100 * b)
        kw_o["serial_stream"] = True; kw_g["serial_stream"] = True
    elif c["mode"] == "history":
        kw_g["history"] = True
    ores = onet.relax_batch(probes.copy(), pool, threads=1 if c["mode"] == "serial" else 4, want_energies=True, **kw_o)
    res = net.relax_batch(probes, pool, dedup=True, want_energies=True, **kw_g)
    assert_relax_equal(res, ores, n)
    assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-300), c
    _, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits), c
    if c["mode"] == "history":
        for i in range(min(b, 6)):
            st = probes[i].copy()
            hist = np.zeros((onet.c.max_iter + 1, n))
            info = orc.OrcRelaxInfo()
            orc.lib().orc_relax_state(onet.ref(), st.ctypes.data_as(C.c_void_p), pool.ctypes.data_as(C.c_void_p), pool.shape[0], 0,
                                      C.byref(info), None, None, hist.ctypes.data_as(C.c_void_p), 0)
            k = info.num_steps
            assert res.NumSteps[i] == k and np.array_equal(res.History[i, :k], orc.pack(hist[:k])), c
    net.close()
