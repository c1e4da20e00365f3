"""-m gpu: learning-rule / energy / stability parity, CUDA path vs the CPU oracle.
Un-normalised weights are dyadic -> bit-exact; normalised weights <= 1e-9 relative."""
import numpy as np
import pytest

from hopfield_network_go_b200 import hopfieldnetwork as hn
from oracle import orc
from tests.helpers import gpu_net

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,domain", [(100, 10, 0), (257, 33, 0), (512, 64, 0), (96, 7, 1)])
def test_hebbian_epoch_bitwise(built, n, p, domain):
    rng = orc.Rng(21)
    t = rng.generate_states(domain, n, p)
    onet = orc.Net(n, domain=domain)
    onet.hebbian(t)
    onet.hebbian(t)  # re-applied per epoch (LearningMethod.go:41-58): W grows by X X^T each time
    net = gpu_net(n, domain)
    net.hebbian_epoch(t)
    net.hebbian_epoch(t)
    assert np.array_equal(net.GetMatrix(), onet.w)
    nrm_o = onet.normalize()
    nrm_g = net.normalize()
    # quarter-integer weights: the sum of squares is exact, as in the reference's Dlassq (LAPACK-3.10 variant, the default;
    # the classic recurrence is covered by test_both_dlassq_variants_are_bitwise), so the normalised W is bit-identical
    assert nrm_g == nrm_o
    assert np.array_equal(net.GetMatrix(), onet.w)
    net.close()


@pytest.mark.parametrize("stream", [0, 3])   # float64 column stream / AUTO: inside a learning run AUTO rebuilds the integer image of
                                             # the un-normalised weights every epoch (integer-path sweeps, fields and diagnostics)
@pytest.mark.parametrize("n,p,domain,sym", [(128, 12, 0, True), (200, 20, 0, False), (256, 30, 1, True), (1024, 100, 0, True)])
def test_delta_epochs_bitwise(built, n, p, domain, sym, stream):
    """Several Delta epochs with MaximalInversion noise; explicit noised copies + permutations (the
    RNG-derived inputs) are shared by oracle and GPU.  W, relaxed copies: bit-exact."""
    rng = orc.Rng(31)
    t = rng.generate_states(domain, n, p)
    onet = orc.Net(n, domain=domain, force_symmetric=sym)
    net = gpu_net(n, domain, SetForceSymmetric=sym, SetColumnStream=stream)
    for epoch in range(4):
        noised = t.copy()
        for s in range(p):
            rng.apply_noise(1, 0.1, noised[s])
            noised[s] = np.where(noised[s] <= 0, 0.0 if domain else -1.0, 1.0)
        perms = rng.fresh_perms(n, p)
        rel_o = onet.delta(t, noised, perms)
        rel_g = net.delta_epoch(t, noised, perms)
        assert np.array_equal(rel_g, orc.pack(rel_o)), f"relaxed copies differ at epoch {epoch}"
        assert np.array_equal(net.GetMatrix(), onet.w), f"weights differ at epoch {epoch}"
        e_g, unst_g, stab_g, se_g = net.energy_profiles(t)
        for s in range(p):
            e_o = onet.all_unit_energies(t[s])
            assert np.array_equal(e_g[s], e_o)  # dyadic weights: exact in any summation order
            st_o, cnt_o = onet.state_is_stable(t[s])
            assert bool(stab_g[s]) == st_o and unst_g[s] == cnt_o
    net.close()


def test_learn_states_full_loop_matches_oracle(built):
    """Whole LearnStates (activation, epochs with early exit, Frobenius normalisation) with the library's
    own restated RNG against the oracle's: same seed -> same noise/permutation draws -> same W."""
    n, p = 128, 14
    rng = orc.Rng(41)
    t = rng.generate_states(0, n, p)
    onet = orc.Net(n)
    nrows, rows, en = onet.learn_states(t.copy(), 2, 0, 30, 1, 0.1, orc.Rng(77), cap=30 * p, want_energies=True)
    net = gpu_net(n, rule=hn.DeltaLearningRule, epochs=30, seed=77, SetLearningNoiseMethod=hn.MaximalInversion,
                  SetLearningNoiseRatio=0.1)
    lsd = net.LearnStates(t.copy(), want_energies=True)
    assert len(lsd) == nrows
    assert [(d.Epoch, d.TargetStateIndex, d.Stable) for d in lsd] == rows
    for d, e in zip(lsd, en):
        assert np.array_equal(d.EnergyProfile, e)
    assert np.array_equal(net.GetMatrix(), onet.w)   # bitwise, through the normalisation (Delta from zero: quarter-integers)
    assert np.array_equal(net.GetLearnedStates(), orc.pack(t))
    net.close()


@pytest.mark.parametrize("n,p,rule", [(100, 10, 0), (257, 33, 0), (512, 40, 2), (1024, 100, 0)])
def test_both_dlassq_variants_are_bitwise(built, n, p, rule):
    """hn_set_norm_algorithm: the classic Dlassq recurrence and the LAPACK-3.10 rewrite (gonum carries one or the other
    depending on its release) are each reproduced bit for bit for quarter-integer weights."""
    rng = orc.Rng(71)
    t = rng.generate_states(0, n, p)
    for variant in (0, 1):
        onet = orc.Net(n)
        net = gpu_net(n, rule=rule, epochs=3, seed=9, SetNormAlgorithm=variant, SetLearningNoiseMethod=hn.MaximalInversion,
                      SetLearningNoiseRatio=0.1)
        if rule == 0:
            onet.hebbian(t)
            net.hebbian_epoch(t)
        else:
            r2 = orc.Rng(9)
            for _ in range(3):
                noised = t.copy()
                for s_ in range(p):
                    r2.apply_noise(1, 0.1, noised[s_])
                perms = r2.fresh_perms(n, p)
                onet.delta(t, noised, perms)
                net.delta_epoch(t, noised, perms)
        assert np.array_equal(net.GetMatrix(), onet.w)
        nrm_o = onet.normalize(variant)
        nrm_g = net.normalize()
        assert nrm_g == nrm_o, f"variant {variant}"
        assert np.array_equal(net.GetMatrix(), onet.w), f"variant {variant}"
        net.close()


@pytest.mark.parametrize("n,p,stream", [(96, 6, 0), (96, 6, 3), (2048, 8, 3), (1100, 4, 0)])
def test_learned_on_device_then_relaxed_with_ties(built, n, p, stream):
    """End to end without injecting the oracle's matrix: LearnStates on the device (Hebbian, even P: exact ties of the
    true field are common), then relaxation.  The re-evaluation of a tied field in the reference's summation order
    depends on the last bits of the normalised W, so this pins the device-side normalisation to the oracle's."""
    rng = orc.Rng(61)
    t = rng.generate_states(0, n, p)
    probes = rng.generate_states(0, n, 48)
    pool = rng.perm_pool(n, 100)
    # a field can vanish only if the number of targets at odd Hamming distance from target 0 has the parity of P/2
    # (all fields are = 2 mod 4 otherwise): flip one unit of target 1 when it does not
    k = int(np.sum(np.sum(t != t[0], axis=1) % 2))
    if k % 2 != (p // 2) % 2:
        t[1, 0] = -t[1, 0]
    onet = orc.Net(n)
    onet.hebbian(t)
    onet.normalize()
    onet.set_targets(t)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=False)
    net = gpu_net(n, SetColumnStream=stream)
    net.LearnStates(t.copy())
    assert np.array_equal(net.GetMatrix(), onet.w)
    res = net.relax_batch(probes, pool, dedup=True)
    from tests.helpers import assert_relax_equal
    assert_relax_equal(res, ores, n)
    assert int(np.sum(res.MarginHits)) > 0, "the case must contain decisions re-evaluated in the reference's order (exact ties)"
    net.close()


def test_iterative_batch_and_mapped_rules(built):
    n, p = 64, 15
    rng = orc.Rng(51)
    t = rng.generate_states(0, n, p)
    for rule in (hn.HebbianLearningRule, hn.BipolarMappedHebbianLearningRule, hn.BipolarMappedDeltaLearningRule):
        onet = orc.Net(n)
        nrows, rows, _ = onet.learn_states(t.copy(), rule, 1, 3, 1, 0.1, orc.Rng(5), cap=1000)
        net = gpu_net(n, rule=rule, method=hn.IterativeBatchMethod, epochs=3, seed=5,
                      SetLearningNoiseMethod=hn.MaximalInversion, SetLearningNoiseRatio=0.1)
        lsd = net.LearnStates(t.copy())
        assert [(d.Epoch, d.TargetStateIndex, d.Stable) for d in lsd] == rows, f"rule {rule}"
        assert np.allclose(net.GetMatrix(), onet.w, rtol=1e-9, atol=0), f"rule {rule}"
        net.close()


def test_mapped_rules_on_binary_network(built):
    """bipolarMappedDelta on a BINARY network: targets are re-activated to +-1 for the rule while the relaxed
    copies carry {0,1} values (LearningRule.go:117-126)."""
    n, p = 72, 9
    t = orc.Rng(52).generate_states(1, n, p)
    onet = orc.Net(n, domain=1)
    nrows, rows, _ = onet.learn_states(t.copy(), hn.BipolarMappedDeltaLearningRule, 0, 4, 1, 0.1, orc.Rng(6), cap=100)
    net = gpu_net(n, 1, rule=hn.BipolarMappedDeltaLearningRule, epochs=4, seed=6,
                  SetLearningNoiseMethod=hn.MaximalInversion, SetLearningNoiseRatio=0.1)
    lsd = net.LearnStates(t.copy())
    assert [(d.Epoch, d.TargetStateIndex, d.Stable) for d in lsd] == rows
    assert np.allclose(net.GetMatrix(), onet.w, rtol=1e-9, atol=0)
    net.close()


def test_energy_profiles_normalised_weights(built):
    n, p = 300, 31
    rng = orc.Rng(61)
    t = rng.generate_states(0, n, p)
    onet = orc.Net(n)
    onet.hebbian(t)
    onet.normalize()
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    probes = rng.generate_states(0, n, 40)
    e_g, unst_g, stab_g, se_g = net.energy_profiles(probes)
    for s in range(40):
        e_o = onet.all_unit_energies(probes[s])
        assert np.allclose(e_g[s], e_o, rtol=1e-9, atol=1e-18)
        st_o, cnt_o = onet.state_is_stable(probes[s])
        assert unst_g[s] == cnt_o and bool(stab_g[s]) == st_o
        assert abs(se_g[s] - onet.state_energy(probes[s])) <= 1e-9 * abs(onet.state_energy(probes[s]))
    net.close()


def test_update_state_and_gaussian_noise(built):
    """UpdateState (one sweep, fresh order) on states noised with GaussianApplication; the oracle and the
    library RNG restatements must draw the same normal variates."""
    n = 160
    rng_o, rng_g = orc.Rng(71), hn.Rand(71)
    t = orc.Rng(3).generate_states(0, n, 8)
    a, b = t.copy(), t.copy()
    for s in range(8):
        rng_o.apply_noise(3, 0.8, a[s])
        rng_g.apply_noise(hn.GaussianApplication, 0.8, b[s])
    assert np.array_equal(a, b)
    onet = orc.Net(n)
    onet.hebbian(t)
    onet.normalize()
    act = np.where(a <= 0, -1.0, 1.0)
    perms = rng_o.fresh_perms(n, 8)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    out = net.UpdateState(act, perms)
    ref = np.stack([onet.update_state(act[s], perms[s]) for s in range(8)])
    assert np.array_equal(out, orc.pack(ref))
    net.close()


@pytest.mark.parametrize("domain", [0, 1])
def test_thermal_delta_epochs(built, domain):
    """thermalDelta (LearningRule.go:129-158): per-target factor exp(-||W t||_2/(1+||W||_F)).  The relaxed copies
    are bit-exact; dW is no longer integer valued, so W matches within summation-order rounding."""
    n, p = 160, 14
    rng = orc.Rng(91)
    t = rng.generate_states(domain, n, p)
    onet = orc.Net(n, domain=domain)
    onet.hebbian(t)
    net = gpu_net(n, domain)
    net.hebbian_epoch(t)
    for epoch in range(3):
        net.SetMatrix(onet.w)                        # identical bits going into every epoch
        noised = t.copy()
        for s in range(p):
            rng.apply_noise(1, 0.15, noised[s])
            noised[s] = np.where(noised[s] <= 0, 0.0 if domain else -1.0, 1.0)
        perms = rng.fresh_perms(n, p)
        rel_o = onet.thermal_delta(t, noised, perms)
        rel_g = net.thermal_delta_epoch(t, noised, perms)
        assert np.array_equal(rel_g, orc.pack(rel_o)), f"relaxed copies differ at epoch {epoch}"
        w_g = net.GetMatrix()
        assert np.allclose(w_g, onet.w, rtol=1e-9, atol=1e-9 * np.abs(onet.w).max()), f"weights differ at epoch {epoch}"
        assert np.array_equal(np.diag(w_g), np.zeros(n)) and np.array_equal(w_g, w_g.T)
    net.close()


def test_learn_states_thermal_rule_full_loop(built):
    n, p = 96, 10
    t = orc.Rng(92).generate_states(0, n, p)
    onet = orc.Net(n)
    nrows, rows, _ = onet.learn_states(t.copy(), hn.ThermalDeltaLearningRule, 0, 6, 1, 0.1, orc.Rng(13), cap=100)
    net = gpu_net(n, rule=hn.ThermalDeltaLearningRule, epochs=6, seed=13, SetLearningNoiseMethod=hn.MaximalInversion,
                  SetLearningNoiseRatio=0.1)
    lsd = net.LearnStates(t.copy())
    assert [(d.Epoch, d.TargetStateIndex, d.Stable) for d in lsd] == rows
    assert np.allclose(net.GetMatrix(), onet.w, rtol=1e-8, atol=1e-9 * np.abs(onet.w).max())
    net.close()


@pytest.mark.parametrize("domain", [0, 1])
def test_update_state_on_zero_matrix(built, domain):
    """A fresh network has W = 0: every field is an exact zero, h <= 0 selects the low value for every unit
    (HopfieldNetwork.go:280-291 with BipolarDomainManager.go:10-22) — the closed-form path of sweep_once."""
    n, b = 300, 9
    rng = orc.Rng(8)
    s = rng.generate_states(domain, n, b)
    perms = rng.fresh_perms(n, b)
    onet = orc.Net(n, domain=domain)
    expect = np.stack([onet.update_state(s[i].copy(), perms[i]) for i in range(b)])
    assert (expect == (0.0 if domain else -1.0)).all()
    net = gpu_net(n, domain)
    out = net.UpdateState(s, perms)
    assert np.array_equal(out, orc.pack(expect)) and not out.any()
    net.close()
