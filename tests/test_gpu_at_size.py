"""-m gpu: the BASELINE configurations at their own sizes and the round-2 boundary additions.

 * config 3 (N=4096, Hebbian, P=400) with probes that REPEAT attractors (noisy copies of the targets, both signs):
   the unique table — first-seen index and hit count — against the oracle's table keyed on the float energy
   profile, on the exact integer stream and on the float64 stream;
 * config 2 (N=1024, Delta, 100 targets, inversion noise) relaxation of a few thousand probes;
 * config 4 (N=2048, Delta, Gaussian noise, asymmetric weights, relaxation history) as ONE run;
 * the multi-warp exact-arithmetic tie path (un-normalised weights, many exactly-zero fields);
 * UnitEnergy (the scalar loop, Binary's -1 per term), integer structure recovered from supplied weights.
"""
import ctypes as C

import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from hopfield_network_go_b200 import hopfieldnetwork as hn
from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net

pytestmark = pytest.mark.gpu


def _noisy_copies(rs, targets, copies, lo, hi, n, with_inverse=True):
    """`copies` noisy copies of every target (a fraction lo..hi of the units inverted); every fourth copy is of the
    INVERSE target, so that s / -s pairs meet in the table."""
    out = []
    for t in targets:
        for c in range(copies):
            x = t.copy() if (not with_inverse or c % 4 != 3) else -t
            k = int(n * rs.uniform(lo, hi))
            idx = rs.permutation(n)[:k]
            x[idx] = -x[idx]
            out.append(x)
    return np.ascontiguousarray(np.array(out))


@pytest.fixture(scope="module")
def config3_problem(built):
    """Oracle side of the config-3 case, computed once for both streams (Hebbian N=4096: ~15 s of CPU)."""
    n, p = 4096, 400
    rng = orc.Rng(0x5EED + 3)
    targets = rng.generate_states(0, n, p)
    pool = rng.perm_pool(n, 100)
    onet = orc.Net(n)
    onet.hebbian(targets)
    k2 = np.rint(onet.w * 2).astype(np.int32)
    onet.normalize()
    onet.set_targets(targets)
    rs = np.random.Generator(np.random.PCG64(99))
    probes = _noisy_copies(rs, targets[:48], 8, 0.05, 0.20, n)          # 384 probes, 48 (x2 signs) basins
    probes = np.concatenate([probes, rng.generate_states(0, n, 16)])   # + 16 random ones (own attractors)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=True, fast_k2=k2)
    aid, first, hits = orc.unique_table(ores["energies"])
    assert len(first) < len(probes) // 2, "the workload must repeat attractors"
    assert ores["margin_hits"].sum() > 0, "and it meets exact ties (even P)"
    return dict(n=n, targets=targets, pool=pool, w=onet.w, probes=probes, ores=ores, aid=aid, first=first, hits=hits)


@pytest.mark.parametrize("stream", [capi.HN_COLUMNS_AUTO, capi.HN_COLUMNS_F64])
def test_config3_unique_table_with_repeating_attractors(config3_problem, stream):
    c = config3_problem
    n = c["n"]
    net = gpu_net(n, SetColumnStream=stream)
    net.LearnStates(c["targets"].copy())             # learned in the library: AUTO runs the integer stream
    assert net.integer_stream_active() == (stream == capi.HN_COLUMNS_AUTO)
    assert np.allclose(net.GetMatrix(), c["w"], rtol=1e-9, atol=0)
    net.SetMatrix(c["w"])                            # the oracle's bits decide exact ties identically
    assert net.integer_stream_active() == (stream == capi.HN_COLUMNS_AUTO)
    res = net.relax_batch(c["probes"], c["pool"], dedup=True)
    assert_relax_equal(res, c["ores"], n)
    assert np.array_equal(res.MarginHits, c["ores"]["margin_hits"])
    assert np.array_equal(res.UniqueFirst, c["first"]) and np.array_equal(res.UniqueHits, c["hits"])
    assert np.array_equal(res.AttractorId, c["aid"])
    net.close()


def test_config2_relaxation_at_size(built):
    """N=1024, Delta rule, 100 targets, MaximalInversion 0.1: weights learned by the oracle (explicit epochs), 4096
    random probes relaxed on the default stream; everything against the oracle, unique table included."""
    n, p, b = 1024, 100, 4096
    rng = orc.Rng(0x5EED + 2)
    t = rng.generate_states(0, n, p)
    onet = orc.Net(n)
    net = gpu_net(n, rule=hn.DeltaLearningRule, SetColumnStream=capi.HN_COLUMNS_AUTO)
    for _ in range(6):
        noised = t.copy()
        for s in range(p):
            rng.apply_noise(1, 0.1, noised[s])
            noised[s] = np.where(noised[s] <= 0, -1.0, 1.0)
        perms = rng.fresh_perms(n, p)
        onet.delta(t, noised, perms)
        net.delta_epoch(t, noised, perms)
    assert np.array_equal(net.GetMatrix(), onet.w)
    k4 = np.rint(onet.w * 4).astype(np.int32)
    assert np.array_equal(k4 * 0.25, onet.w)
    onet.normalize()
    onet.set_targets(t)
    net.normalize()
    net.SetLearnedStates(orc.pack(t))
    assert net.integer_stream_active()
    net.SetMatrix(onet.w)          # the oracle's bits (Dlassq norm); the structure is recovered from the matrix itself
    assert net.integer_stream_active()
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=True, fast_k2=k4)
    res = net.relax_batch(probes, pool, dedup=True)
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    _, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    net.close()


def test_config4_at_size(built):
    """N=2048, Delta rule, Gaussian learning noise (sigma 0.5), forceSymmetric=false, relaxationHistory on, 256 probes."""
    n, p, b = 2048, 200, 256
    rng = orc.Rng(0x5EED + 4)
    t = rng.generate_states(0, n, p)
    onet = orc.Net(n, force_symmetric=False)
    net = gpu_net(n, rule=hn.DeltaLearningRule, SetForceSymmetric=False, SetColumnStream=capi.HN_COLUMNS_AUTO)
    for epoch in range(3):
        noised = t.copy()
        for s in range(p):
            rng.apply_noise(3, 0.5, noised[s])                      # GaussianApplication, then the activation
            noised[s] = np.where(noised[s] <= 0, -1.0, 1.0)
        perms = rng.fresh_perms(n, p)
        rel_o = onet.delta(t, noised, perms)
        rel_g = net.delta_epoch(t, noised, perms)
        assert np.array_equal(rel_g, orc.pack(rel_o)), f"relaxed copies differ at epoch {epoch}"
        assert np.array_equal(net.GetMatrix(), onet.w), f"weights differ at epoch {epoch}"
    assert not np.array_equal(onet.w, onet.w.T), "config 4 is the asymmetric case"
    k1 = np.rint(onet.w).astype(np.int32)
    assert np.array_equal(k1.astype(np.float64), onet.w)
    onet.normalize()
    onet.set_targets(t)
    net.normalize()
    assert np.allclose(net.GetMatrix(), onet.w, rtol=1e-9, atol=0)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(t))
    assert net.integer_stream_active()
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=True, fast_k2=k1)
    res = net.relax_batch(probes, pool, dedup=True, history=True, want_energies=True)
    assert_relax_equal(res, ores, n)
    assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-300)
    _, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    # history: every captured step of the first probes against the slow oracle's StateHistory
    for i in range(6):
        st = probes[i].copy()
        hist = np.zeros((onet.c.max_iter + 1, n))
        info = orc.OrcRelaxInfo()
        orc.lib().orc_relax_state(onet.ref(), st.ctypes.data_as(C.c_void_p), pool.ctypes.data_as(C.c_void_p),
                                  pool.shape[0], 0, C.byref(info), None, None, hist.ctypes.data_as(C.c_void_p), 0)
        k = info.num_steps
        assert res.NumSteps[i] == k
        assert np.array_equal(res.History[i, :k], orc.pack(hist[:k]))
    # last captured step == final state, first == the probe, for all of them
    ix = np.arange(b)
    assert np.array_equal(res.History[ix, res.NumSteps - 1], res.FinalStates)
    assert np.array_equal(res.History[:, 0], orc.pack(probes))
    # the same on the float64 stream
    f64 = gpu_net(n, rule=hn.DeltaLearningRule, SetForceSymmetric=False, SetColumnStream=capi.HN_COLUMNS_F64)
    f64.SetMatrix(onet.w)
    f64.SetLearnedStates(orc.pack(t))
    r2 = f64.relax_batch(probes, pool, dedup=True, history=True)
    assert_relax_equal(r2, ores, n)
    assert np.array_equal(r2.History, res.History)
    assert np.array_equal(r2.UniqueFirst, first) and np.array_equal(r2.UniqueHits, hits)
    f64.close()
    net.close()


@pytest.mark.parametrize("n,p", [(2048, 4), (4096, 2), (1280, 6)])
def test_exact_arithmetic_ties_multi_warp(built, n, p):
    """Un-normalised Hebbian weights with even P: every field is an even integer and exact zeros occur; probes
    with >= 2 warps (N > 1024) take the exact-arithmetic tie path over and over (the s_flag hand-off of ADVICE r1)."""
    rng = orc.Rng(600 + n)
    t = rng.generate_states(0, n, p)
    # a field is sum_mu +-m_mu - P s_k with every overlap m_mu even: it can vanish only if the number of targets at odd
    # Hamming distance from the state has the parity of P/2, a property of the target set; one unit flipped fixes it
    k = sum(int((t[0] != t[m]).sum()) % 2 for m in range(1, p))
    if k % 2 != (p // 2) % 2:
        t[1, 0] = -t[1, 0]
    onet = orc.Net(n)
    onet.hebbian(t)
    onet.set_targets(t)
    k1 = np.rint(onet.w).astype(np.int32)
    net = gpu_net(n, SetColumnStream=capi.HN_COLUMNS_F64)
    net.hebbian_epoch(t)                      # not normalised: quarter-integer weights, exact arithmetic mode
    assert np.array_equal(net.GetMatrix(), onet.w)
    net.SetLearnedStates(orc.pack(t))
    b = 512
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=True, fast_k2=k1)
    assert ores["margin_hits"].sum() > 10, "the case must meet exact ties"
    for _ in range(3):                        # a race would not show on every run
        res = net.relax_batch(probes, pool, dedup=True)
        assert_relax_equal(res, ores, n)
        assert np.array_equal(res.MarginHits, ores["margin_hits"])
        _, first, hits = orc.unique_table(ores["energies"])
        assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    net.close()


@pytest.mark.parametrize("domain", [0, 1])
def test_unit_energy_scalar_loop(built, domain):
    """hn_unit_energy == BipolarDomainManager.go:29-37 / BinaryDomainManager.go:43-52 (the -1 per term included)."""
    n, p = 160, 12
    rng = orc.Rng(900 + domain)
    t = rng.generate_states(domain, n, p)
    onet = orc.Net(n, domain=domain)
    onet.hebbian(t)
    onet.normalize()
    net = gpu_net(n, domain)
    net.SetMatrix(onet.w)
    states = rng.generate_states(domain, n, 5)
    for s in states:
        for i in (0, 1, 77, n - 1):
            assert net.UnitEnergy(s, i) == onet.unit_energy(s, i)
    if domain == 1:   # the quirk: the scalar loop is NOT AllUnitEnergies(state)[i] for Binary networks
        assert abs(net.UnitEnergy(states[0], 3) - net.AllUnitEnergies(states[0])[3]) > n / 2
    with pytest.raises(capi.HopfieldError):
        net.UnitEnergy(states[0], n)
    net.close()


def test_supplied_weights_recover_integer_structure(built):
    """hn_set_weights recovers W = fl(c*K) from the matrix alone (a matrix.bin resume runs on the integer stream);
    a random initial matrix or thermal weights do not pass the bit test."""
    n, p = 384, 30
    rng = orc.Rng(1234)
    t = rng.generate_states(0, n, p)
    for rule in ("hebbian", "delta"):
        onet = orc.Net(n)
        if rule == "hebbian":
            onet.hebbian(t)
        else:
            for _ in range(3):
                noised = t.copy()
                for s in range(p):
                    rng.apply_noise(1, 0.1, noised[s])
                    noised[s] = np.where(noised[s] <= 0, -1.0, 1.0)
                onet.delta(t, noised, rng.fresh_perms(n, p))
        onet.normalize()
        onet.set_targets(t)
        net = gpu_net(n, SetColumnStream=capi.HN_COLUMNS_AUTO)
        net.SetMatrix(onet.w)
        assert net.integer_stream_active(), rule
        net.SetLearnedStates(orc.pack(t))
        probes = rng.generate_states(0, n, 64)
        pool = rng.perm_pool(n, 100)
        ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=True)
        res = net.relax_batch(probes, pool, dedup=True)
        assert_relax_equal(res, ores, n)
        _, first, hits = orc.unique_table(ores["energies"])
        assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
        w = onet.w.copy()
        w[3, 5] = np.nextafter(w[3, 5], 1.0)      # one ulp off: no longer fl(c*K)
        w[5, 3] = w[3, 5]
        net.SetMatrix(w)
        assert not net.integer_stream_active()
        net.close()
    g = gpu_net(n, SetColumnStream=capi.HN_COLUMNS_AUTO)
    g.SetMatrix(np.random.default_rng(3).normal(0.0, 0.01, (n, n)))
    assert not g.integer_stream_active()
    g.close()


def test_learn_states_sharded_whole_range_is_learn_states(built):
    """Without a communicator hn_learn_states_sharded over [0, n) is hn_learn_states; a proper shard is refused."""
    n, p = 96, 10
    rng = orc.Rng(88)
    t = rng.generate_states(0, n, p)
    kw = dict(rule=hn.DeltaLearningRule, epochs=8, seed=5, SetLearningNoiseMethod=hn.MaximalInversion, SetLearningNoiseRatio=0.1)
    a, b = gpu_net(n, **kw), gpu_net(n, **kw)
    ra = a.LearnStates(t.copy(), want_energies=True)
    rb = b.LearnStates(t.copy(), want_energies=True, shard=(0, p))
    assert [(d.Epoch, d.TargetStateIndex, d.Stable) for d in ra] == [(d.Epoch, d.TargetStateIndex, d.Stable) for d in rb]
    assert np.array_equal(a.GetMatrix(), b.GetMatrix())
    with pytest.raises(capi.HopfieldError):
        b.LearnStates(t.copy(), shard=(0, p // 2))
    a.close()
    b.close()
