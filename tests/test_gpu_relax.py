"""-m gpu: relaxation parity, CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.
Bar: final states, NumSteps, Stable, distances, unique table bit-exact; energies <= 1e-9 relative."""
import numpy as np
import pytest

from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net, make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,b,seed", [(100, 10, 200, 3), (256, 25, 128, 4), (257, 26, 64, 5), (1024, 100, 96, 6),
                                        (64, 6, 50, 7), (33, 3, 40, 8)])
def test_relax_matches_oracle(built, n, p, b, seed):
    onet, k2, pool, probes, targets = make_problem(n, p, b, seed)
    ores = onet.relax_batch(probes.copy(), pool, threads=8)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True, want_energies=True)
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-300)
    aid, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first)
    assert np.array_equal(res.UniqueHits, hits)
    assert np.array_equal(res.AttractorId, aid)
    net.close()


def test_relax_with_exact_ties(built):
    """N, P chosen so that exact ties of the true field occur (even P): decisions inside the margin are
    re-evaluated in the reference's summation order and must still match the oracle bit for bit."""
    found = False
    for seed in range(20, 40):
        onet, k2, pool, probes, targets = make_problem(512, 50, 64, seed)
        ores = onet.relax_batch(probes.copy(), pool, threads=8, fast_k2=k2)
        if ores["margin_hits"].sum() == 0:
            continue
        found = True
        net = gpu_net(512)
        net.SetMatrix(onet.w)
        net.SetLearnedStates(orc.pack(targets))
        res = net.relax_batch(probes, pool)
        assert_relax_equal(res, ores, 512)
        assert np.array_equal(res.MarginHits, ores["margin_hits"])
        net.close()
        break
    assert found, "no seed produced an exact tie; widen the search"


def test_serial_stream_config1(built):
    """BASELINE config 1 (N=100, Hebbian, 10 targets, 1000 probes) with threads=1 semantics:
    one persistent shuffled list and one RNG stream serve all probes in order."""
    n, p, b = 100, 10, 1000
    rng = orc.Rng(11)
    targets = rng.generate_states(0, n, p)
    onet = orc.Net(n)
    onet.learn_states(targets.copy(), 0, 0, 100, 0, 0.0, orc.Rng(1))
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 12000)
    ores = onet.relax_batch(probes.copy(), pool, serial_stream=True)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, serial_stream=True)
    assert_relax_equal(res, ores, n)
    net.close()


def test_offsets_and_pool_exhaustion(built):
    from hopfield_network_go_b200 import capi
    onet, k2, pool, probes, targets = make_problem(128, 12, 32, 9, pool_rows=140)
    offsets = (np.arange(32) % 40).astype(np.int32)
    ores = onet.relax_batch(probes.copy(), pool, offsets=offsets, threads=4)
    net = gpu_net(128)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, offsets=offsets)
    assert_relax_equal(res, ores, 128)
    # a pool that is too short for an unstable probe must fail loudly, not wrap around
    w = np.zeros((128, 128)); w[0, 1] = 1.0; w[1, 0] = -1.0   # frustrated pair: never stable
    net.SetMatrix(w)
    with pytest.raises(capi.HopfieldError) as ei:
        net.relax_batch(probes[:4], pool[:10])
    assert ei.value.code == capi.HN_E_POOL
    net.close()


def test_pinned_result_buffers_match_pageable(built):
    """hn_host_alloc staging: the same results arrive in the network's page-locked buffers, which are reused."""
    n, p, b = 320, 24, 128
    onet, _, pool, probes, targets = make_problem(n, p, b, 58)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    ref = net.relax_batch(probes, pool, dedup=True, want_energies=True)
    a = net.relax_batch(probes, pool, dedup=True, want_energies=True, reuse_pinned_buffers=True)
    for name in ("FinalStates", "NumSteps", "Stable", "DistancesToTargets", "Flips", "AttractorId", "UniqueFirst",
                 "UniqueHits", "StateHash", "EnergyProfiles"):
        assert np.array_equal(getattr(a, name), getattr(ref, name)), name
    addr = a.FinalStates.ctypes.data
    b2 = net.relax_batch(probes[:64], pool, dedup=True, reuse_pinned_buffers=True)      # smaller batch: same buffers
    assert b2.FinalStates.ctypes.data == addr
    assert np.array_equal(b2.FinalStates, ref.FinalStates[:64])
    net.close()
