"""-m gpu: edge cases of the hot path through the C ABI — empty and single-probe batches, tiny and word-ragged
dimensions, no learned states, one-sweep budgets, probes that are already attractors — every one against the oracle,
in the float64 stream and (where weights are learned in the library) in the default AUTO stream."""
import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from hopfield_network_go_b200 import hopfieldnetwork as hn
from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net, make_problem

pytestmark = pytest.mark.gpu


def test_empty_batch(built):
    n = 100
    onet, _, pool, probes, targets = make_problem(n, 5, 4, 1)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(np.zeros((0, capi.words(n)), dtype=np.uint64), pool, dedup=True)
    assert res.FinalStates.shape == (0, capi.words(n)) and res.NumSteps.shape == (0,)
    assert len(res.UniqueFirst) == 0 and res.total_flips == 0
    res1 = net.relax_batch(probes[:1], pool, dedup=True)            # a single probe
    assert_relax_equal(res1, onet.relax_batch(probes[:1].copy(), pool), n)
    assert res1.UniqueFirst.tolist() == [0] and res1.UniqueHits.tolist() == [1]
    net.close()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129])
@pytest.mark.parametrize("stream", [capi.HN_COLUMNS_F64, capi.HN_COLUMNS_AUTO])
def test_tiny_and_word_ragged_dimensions(built, n, stream):
    p, b = max(1, n // 8), 40
    rng = orc.Rng(500 + n)
    t = rng.generate_states(0, n, p)
    net = gpu_net(n, SetColumnStream=stream, SetForceZeroDiagonal=(n > 1))
    net.LearnStates(t.copy())                                        # Hebbian in the library, normalised
    onet = orc.Net(n, w=net.GetMatrix(), force_zero_diag=(n > 1))
    onet.set_targets(t)
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=4, want_energies=True)
    res = net.relax_batch(probes, pool, dedup=True, want_energies=True)
    assert_relax_equal(res, ores, n)
    assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-300)
    _, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    net.close()


def test_no_learned_states_means_no_distances(built):
    n, b = 96, 16
    onet, _, pool, probes, _ = make_problem(n, 6, b, 9)
    net = gpu_net(n)
    net.SetMatrix(onet.w)                                            # weights but no target states registered
    res = net.relax_batch(probes, pool)
    assert res.DistancesToTargets is None
    onet.set_targets(np.zeros((0, n)))
    ores = onet.relax_batch(probes.copy(), pool)
    assert np.array_equal(res.FinalStates, orc.pack(ores["final"])) and np.array_equal(res.NumSteps, ores["num_steps"])
    net.close()


@pytest.mark.parametrize("max_iter", [1, 2])
def test_one_sweep_budget(built, max_iter):
    """maximumRelaxationIterations = 1: one sweep, NumSteps = 2 whether or not the state came out stable
    (HopfieldNetwork.go:382-441; Stable=false => NumSteps = maxIter + 1)."""
    n, p, b = 200, 30, 64
    onet, _, pool, probes, targets = make_problem(n, p, b, 21)
    onet.c.max_iter = max_iter
    net = gpu_net(n, SetMaximumRelaxationIterations=max_iter)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    ores = onet.relax_batch(probes.copy(), pool[:max_iter], threads=4)
    res = net.relax_batch(probes, pool[:max_iter])
    assert_relax_equal(res, ores, n)
    assert (res.NumSteps <= max_iter + 1).all() and ((res.Stable == 1) | (res.NumSteps == max_iter + 1)).all()
    net.close()


def test_probes_that_are_attractors_already(built):
    """A stored pattern of a lightly loaded network and its inverse are fixed points: one sweep without a flip."""
    n, p = 512, 8
    onet, _, pool, _, targets = make_problem(n, p, 1, 31)
    net = gpu_net(n, SetColumnStream=capi.HN_COLUMNS_AUTO)
    net.LearnStates(targets.copy())
    probes = np.concatenate([targets, -targets])
    res = net.relax_batch(probes, pool, dedup=True)
    onet2 = orc.Net(n, w=net.GetMatrix())
    onet2.set_targets(targets)
    ores = onet2.relax_batch(probes.copy(), pool, threads=4)
    assert_relax_equal(res, ores, n)
    assert (res.Flips == 0).all() and (res.NumSteps == 2).all() and (res.Stable == 1).all()
    assert np.array_equal(res.FinalStates, orc.pack(probes))
    assert res.UniqueHits.tolist() == [2] * p                        # s and -s are one attractor (same energy profile)
    assert (res.DistancesToTargets.min(axis=1) == 0).all()
    net.close()


@pytest.mark.parametrize("domain", [0, 1])
def test_relaxation_on_the_all_zero_matrix(built, domain):
    """A network that has learned nothing: every field is an exact zero, one sweep sends every unit to the low value
    and the state is stable (closed form in the library, zero_matrix_relax_kernel); serial stream and history too."""
    n, b = 70, 12
    rng = orc.Rng(3)
    t = rng.generate_states(domain, n, 3)
    onet = orc.Net(n, domain=domain)
    onet.set_targets(t)
    probes = rng.generate_states(domain, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, want_energies=True)
    net = gpu_net(n, domain)
    net.SetLearnedStates(orc.pack(t))
    for kw in ({}, {"serial_stream": True}):
        res = net.relax_batch(probes, pool, dedup=True, want_energies=True, history=True, **kw)
        assert_relax_equal(res, ores, n)
        assert np.array_equal(res.MarginHits, ores["margin_hits"])
        assert not res.FinalStates.any() and (res.NumSteps == 2).all() and res.Stable.all()
        assert np.allclose(res.EnergyProfiles, ores["energies"], atol=0)
        _, first, hits = orc.unique_table(ores["energies"])
        assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
        assert np.array_equal(res.History[:, 0], orc.pack(probes)) and not res.History[:, 1].any()
    with pytest.raises(capi.HopfieldError) as ei:
        net.relax_batch(probes, pool[:5], serial_stream=True)          # the serial stream needs one row per probe
    assert ei.value.code == capi.HN_E_POOL
    net.close()
