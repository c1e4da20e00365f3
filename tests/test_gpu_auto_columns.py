"""-m gpu: HN_COLUMNS_AUTO (the default configuration) — the exact integer stream whenever the integer structure of
the weights is known, the float64 stream otherwise; results IDENTICAL to the oracle either way."""
import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from hopfield_network_go_b200 import hopfieldnetwork as hn
from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net, make_problem

pytestmark = pytest.mark.gpu
AUTO = capi.HN_COLUMNS_AUTO


def _relax_and_compare(net, onet, rng, domain, n, b, dedup=False):
    probes = rng.generate_states(domain, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=True)
    res = net.relax_batch(probes, pool, dedup=dedup, want_energies=True)
    assert_relax_equal(res, ores, n)
    assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-300)
    if dedup:
        _, first, hits = orc.unique_table(ores["energies"])
        assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)


def test_builder_default_is_auto(built):
    assert hn.NewHopfieldNetworkBuilder().columnStream == AUTO


@pytest.mark.parametrize("rule,domain,n", [(hn.HebbianLearningRule, 0, 320), (hn.DeltaLearningRule, 0, 256),
                                           (hn.DeltaLearningRule, 1, 192), (hn.HebbianLearningRule, 0, 1000)])
def test_auto_uses_integer_stream_after_learning(built, rule, domain, n):
    p, b = max(4, n // 12), 96
    rng = orc.Rng(77 + n)
    t = rng.generate_states(domain, n, p)
    net = gpu_net(n, domain, rule=rule, epochs=10, seed=4, SetColumnStream=AUTO,
                  SetLearningNoiseMethod=hn.MaximalInversion, SetLearningNoiseRatio=0.1)
    net.LearnStates(t.copy())
    assert net.integer_stream_active()
    onet = orc.Net(n, domain=domain, w=net.GetMatrix())
    onet.set_targets(t)
    _relax_and_compare(net, onet, rng, domain, n, b, dedup=True)
    net.close()


def test_auto_for_supplied_weights_and_fallback_for_thermal_weights(built):
    n, p, b = 256, 20, 64
    onet, k2, pool, probes, targets = make_problem(n, p, b, 41)
    net = gpu_net(n, SetColumnStream=AUTO)
    net.SetMatrix(onet.w)                                  # caller-supplied weights: W = fl(c*K) is recovered from the bits
    assert net.integer_stream_active()
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool)
    assert_relax_equal(res, onet.relax_batch(probes.copy(), pool, threads=8), n)
    net.SetIntegerStructure(k2 * 0.5)                      # declared: the integer stream takes over, same answers
    assert net.integer_stream_active()
    assert_relax_equal(net.relax_batch(probes, pool), onet.relax_batch(probes.copy(), pool, threads=8), n)
    net.close()

    rng = orc.Rng(5)
    t = rng.generate_states(0, n, p)
    th = gpu_net(n, rule=hn.ThermalDeltaLearningRule, epochs=3, seed=2, SetColumnStream=AUTO,
                 SetLearningNoiseMethod=hn.MaximalInversion, SetLearningNoiseRatio=0.1)
    th.LearnStates(t.copy())
    assert not th.integer_stream_active()                  # thermal dW is not integer valued
    o2 = orc.Net(n, w=th.GetMatrix())
    o2.set_targets(t)
    _relax_and_compare(th, o2, rng, 0, n, b)
    th.close()


def test_auto_falls_back_for_random_initial_matrix(built):
    n, p, b = 200, 12, 64
    rng = orc.Rng(19)
    t = rng.generate_states(0, n, p)
    net = gpu_net(n, seed=6, SetColumnStream=AUTO, SetRandMatrixInit=True)
    net.LearnStates(t.copy())
    assert not net.integer_stream_active()
    onet = orc.Net(n, w=net.GetMatrix())
    onet.set_targets(t)
    _relax_and_compare(net, onet, rng, 0, n, b)
    net.close()
