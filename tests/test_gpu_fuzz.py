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
