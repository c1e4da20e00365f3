"""-m gpu: HN_COLUMNS_F32 — the relaxation kernel streams a float32-rounded shadow of W (4N bytes per flip) while
the fields stay float64; the rounding is covered by the decision margin and every decision inside the margin is
re-evaluated exactly on the float64 matrix.  Results must be IDENTICAL to the oracle (and to HN_COLUMNS_F64)."""
import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net, make_problem

pytestmark = pytest.mark.gpu
F32 = capi.HN_COLUMNS_F32


@pytest.mark.parametrize("n,p,b,seed,domain", [(100, 10, 200, 3, 0), (512, 50, 64, 27, 0), (1024, 100, 96, 6, 0),
                                               (192, 16, 96, 12, 1), (257, 26, 64, 5, 0)])
def test_f32_columns_match_oracle(built, n, p, b, seed, domain):
    onet, k2, pool, probes, targets = make_problem(n, p, b, seed, domain=domain)
    ores = onet.relax_batch(probes.copy(), pool, threads=8)
    net = gpu_net(n, domain, SetColumnStream=F32)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True, want_energies=True)
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])          # dyadic weights: only exact ties are inside
    assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-300)
    _, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    net.close()


@pytest.mark.parametrize("mode", [capi.HN_COLUMNS_F64, F32])
def test_non_dyadic_asymmetric_weights_all_probes(built, mode):
    """Gaussian-initialised, non-symmetric, non-dyadic W: fields are continuous, so some decisions genuinely fall
    inside the (wider, in F32 mode) margin; the exact re-evaluation must make EVERY probe match the oracle."""
    n, p, b = 192, 16, 160
    rng = orc.Rng(131)
    t = rng.generate_states(0, n, p)
    w0 = np.array([rng.norm_float64() * 0.01 for _ in range(n * n)]).reshape(n, n)
    onet = orc.Net(n, force_symmetric=False, w=w0)
    onet.hebbian(t)
    onet.normalize()
    onet.set_targets(t)
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8)
    net = gpu_net(n, SetForceSymmetric=False, SetColumnStream=mode)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(t))
    res = net.relax_batch(probes, pool)
    assert_relax_equal(res, ores, n)
    net.close()


def test_f32_columns_config3_sample_and_delta_sweep(built):
    n, p, b = 4096, 400, 32
    onet, k2, pool, probes, targets = make_problem(n, p, b, 16)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=False, fast_k2=k2)
    net = gpu_net(n, SetColumnStream=F32)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool)
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    net.close()
    # the Delta sweep (UpdateState on un-normalised, large-magnitude dyadic W) through the same kernel
    n2, p2 = 256, 20
    rng = orc.Rng(9)
    t = rng.generate_states(0, n2, p2)
    o2 = orc.Net(n2)
    g2 = gpu_net(n2, SetColumnStream=F32)
    for _ in range(3):
        noised = t.copy()
        for s in range(p2):
            rng.apply_noise(1, 0.1, noised[s])
        perms = rng.fresh_perms(n2, p2)
        rel_o = o2.delta(t, noised, perms)
        rel_g = g2.delta_epoch(t, noised, perms)
        assert np.array_equal(rel_g, orc.pack(rel_o)) and np.array_equal(g2.GetMatrix(), o2.w)
    g2.close()
