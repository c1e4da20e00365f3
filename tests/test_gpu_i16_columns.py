"""-m gpu: HN_COLUMNS_I16 — exact integer stream (int32 fields g = 2*Wraw*s, int16 columns, float64 re-evaluation at
exact ties).  Results must be IDENTICAL to the oracle; the mode must switch itself off whenever the integer structure
of the weights is unknown."""
import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from hopfield_network_go_b200 import hopfieldnetwork as hn
from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net, make_problem

pytestmark = pytest.mark.gpu
I16 = capi.HN_COLUMNS_I16


@pytest.mark.parametrize("n,p,b,seed,domain", [(100, 10, 200, 3, 0), (512, 50, 64, 27, 0), (1024, 100, 96, 6, 0),
                                               (192, 16, 96, 12, 1), (257, 26, 64, 5, 0), (4096, 400, 32, 16, 0)])
def test_i16_stream_matches_oracle(built, n, p, b, seed, domain):
    onet, k2, pool, probes, targets = make_problem(n, p, b, seed, domain=domain)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=n <= 1024, fast_k2=k2 if n > 1024 else None)
    net = gpu_net(n, domain, SetColumnStream=I16)
    net.SetMatrix(onet.w)
    assert net.integer_stream_active()                  # caller-supplied weights: W = fl(c*K) recovered from the matrix itself
    net.SetIntegerStructure(k2 * 0.5)                   # the explicit declaration still works
    assert net.integer_stream_active()
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True, want_energies=n <= 1024)
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])          # exactly the exact ties g == 0
    if n <= 1024:
        assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-300)
        _, first, hits = orc.unique_table(ores["energies"])
        assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    net.close()


@pytest.mark.parametrize("rule,domain", [(hn.HebbianLearningRule, 0), (hn.DeltaLearningRule, 0), (hn.DeltaLearningRule, 1)])
def test_i16_stream_after_in_library_learning(built, rule, domain):
    """LearnStates inside the library captures the un-normalised matrix at normalisation time; relaxation then runs
    on the integer stream and matches the oracle run on the very same weights."""
    n, p, b = 256, 20, 96
    rng = orc.Rng(301)
    t = rng.generate_states(domain, n, p)
    net = gpu_net(n, domain, rule=rule, epochs=12, seed=9, SetColumnStream=I16,
                  SetLearningNoiseMethod=hn.MaximalInversion, SetLearningNoiseRatio=0.1)
    net.LearnStates(t.copy())
    assert net.integer_stream_active()
    w = net.GetMatrix()
    onet = orc.Net(n, domain=domain, w=w)
    onet.set_targets(t)
    probes = rng.generate_states(domain, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8)
    res = net.relax_batch(probes, pool)
    assert_relax_equal(res, ores, n)
    # further learning on the normalised matrix destroys the known structure: the mode must switch itself off
    net.hebbian_epoch(t)
    assert not net.integer_stream_active()
    net.close()


def test_i16_rejects_wrong_structure(built):
    n = 64
    onet, k2, pool, probes, targets = make_problem(n, 6, 8, 41)
    net = gpu_net(n, SetColumnStream=I16)
    net.SetMatrix(onet.w)
    bad = k2 * 0.5
    bad[3, 5] += 1.0
    with pytest.raises(capi.HopfieldError):
        net.SetIntegerStructure(bad)                      # not proportional to W
    with pytest.raises(capi.HopfieldError):
        net.SetIntegerStructure(onet.w)                   # proportional, but 4*matrix is not integer valued
    assert net.integer_stream_active()                    # a refused declaration leaves the recovered structure in force
    net.close()


@pytest.mark.parametrize("n,p,b", [(2048, 100, 24), (1500, 40, 24), (5000, 60, 12), (8192, 24, 6), (16384, 16, 3)])
def test_i16_stream_thread_geometries(built, n, p, b):
    """Every thread geometry of the integer stream (32 units per thread; 64 / 128 / 256 / 512 threads per probe) and
    dimensions that are not a multiple of the tile, against the exact-integer oracle."""
    onet, k2, pool, probes, targets = make_problem(n, p, b, 91, pool_rows=100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=False, fast_k2=k2)
    net = gpu_net(n, SetColumnStream=I16)
    net.SetMatrix(onet.w)
    net.SetIntegerStructure(k2 * 0.5)
    assert net.integer_stream_active()
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True)
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    net.close()


def test_i16_stream_with_entries_beyond_a_byte(built):
    """Strongly correlated targets push |K| past 127: the two-plane (hi/lo byte) integer fields GEMM and int16 columns
    with large entries; small-P cases elsewhere take the single-plane GEMM."""
    n, p, b = 512, 300, 48
    rng = orc.Rng(17)
    base = rng.generate_states(0, n, 1)[0]
    t = np.tile(base, (p, 1))
    flip = rng.generate_states(0, n, p) > 0                 # each target = the base pattern with ~3 % of units inverted
    mask = np.zeros_like(flip)
    mask[:, : n // 16] = flip[:, : n // 16]
    t = np.ascontiguousarray(np.where(mask, -t, t))
    net = gpu_net(n, SetColumnStream=I16)
    net.LearnStates(t.copy())
    assert net.integer_stream_active()
    w = net.GetMatrix()
    assert np.abs(w).max() / np.abs(w[np.nonzero(w)]).min() > 127      # the integer image really needs more than a byte
    onet = orc.Net(n, w=w)
    onet.set_targets(t)
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8)
    res = net.relax_batch(probes, pool, dedup=True)
    assert_relax_equal(res, ores, n)
    net.close()
