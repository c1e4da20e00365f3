"""-m gpu: relaxation variants — binary domain, asymmetric W (resident W^T), random init (non-dyadic),
history capture, maxUnstable > 0, already-stable probes, properties at BASELINE size."""
import numpy as np
import pytest

from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net, make_problem

pytestmark = pytest.mark.gpu


def test_binary_domain(built):
    onet, k2, pool, probes, targets = make_problem(192, 16, 96, 12, domain=1)
    ores = onet.relax_batch(probes.copy(), pool, threads=8)
    net = gpu_net(192, 1)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True, want_energies=True)
    assert_relax_equal(res, ores, 192)
    assert np.allclose(res.EnergyProfiles, ores["energies"], rtol=1e-9, atol=1e-18)
    aid, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    net.close()


def test_asymmetric_random_init_weights(built):
    """forceSymmetric=false with Gaussian(0, 0.01) init + Hebbian: W is neither symmetric nor dyadic."""
    n, p, b = 160, 14, 80
    rng = orc.Rng(13)
    t = rng.generate_states(0, n, p)
    w0 = np.array([rng.norm_float64() * 0.01 for _ in range(n * n)]).reshape(n, n)
    onet = orc.Net(n, force_symmetric=False, w=w0)
    onet.hebbian(t)
    onet.normalize()
    onet.set_targets(t)
    assert not np.array_equal(onet.w, onet.w.T)
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8)
    net = gpu_net(n, SetForceSymmetric=False)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(t))
    res = net.relax_batch(probes, pool, want_energies=True)
    # non-dyadic weights: compare every probe with no decision inside the margin (none expected)
    ok = ores["margin_hits"] == 0
    assert ok.sum() >= b - 2
    assert np.array_equal(res.FinalStates[ok], orc.pack(ores["final"])[ok])
    assert np.array_equal(res.NumSteps[ok], ores["num_steps"][ok])
    assert np.array_equal(res.Stable[ok], ores["stable"][ok])
    assert np.array_equal(res.DistancesToTargets[ok], ores["distances"].astype(np.int32)[ok])
    assert np.allclose(res.EnergyProfiles[ok], ores["energies"][ok], rtol=1e-9, atol=1e-16)
    net.close()


def test_history_capture(built):
    n, p, b = 96, 9, 24
    onet, k2, pool, probes, targets = make_problem(n, p, b, 14)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    res = net.relax_batch(probes, pool, history=True)
    import ctypes as C
    for i in range(b):
        st = probes[i].copy()
        hist = np.zeros((onet.c.max_iter + 1, n))
        info = orc.OrcRelaxInfo()
        orc.lib().orc_relax_state(onet.ref(), st.ctypes.data_as(C.c_void_p), pool.ctypes.data_as(C.c_void_p), pool.shape[0], 0,
                                  C.byref(info), None, None, hist.ctypes.data_as(C.c_void_p), 0)
        k = info.num_steps
        assert res.NumSteps[i] == k
        assert np.array_equal(res.History[i, :k], orc.pack(hist[:k]))
    net.close()


def test_max_unstable_units_and_stable_probes(built):
    n, p = 128, 6
    onet, k2, pool, probes, targets = make_problem(n, p, 16, 15)
    # probes that are the (stable) targets themselves: still swept once -> NumSteps == 2 (no pre-check)
    ores = onet.relax_batch(targets.copy(), pool, threads=2)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(targets, pool)
    assert_relax_equal(res, ores, n)
    assert (res.NumSteps == 2).all() and res.Stable.all()
    assert (res.DistancesToTargets.diagonal() == 0).all()
    net.close()
    # tolerance of 5 unstable units
    onet2 = orc.Net(n, max_unstable=5, w=onet.w)
    onet2.set_targets(targets)
    o2 = onet2.relax_batch(probes.copy(), pool, threads=2)
    net2 = gpu_net(n, SetMaximumRelaxationUnstableUnits=5)
    net2.SetMatrix(onet.w)
    net2.SetLearnedStates(orc.pack(targets))
    r2 = net2.relax_batch(probes, pool)
    assert_relax_equal(r2, o2, n)
    net2.close()


def test_baseline_size_config3_sample(built):
    """N=4096, P=400 (BASELINE config 3) on a sample of probes against the exact-integer oracle, plus
    size-independent properties: s and -s relax to inverse states with identical step counts, distances
    are even integers in [0, N], Stable=false => NumSteps = maxIter+1."""
    n, p, b = 4096, 400, 48
    onet, k2, pool, probes, targets = make_problem(n, p, b, 16)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=False, fast_k2=k2)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True)
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    assert ((res.DistancesToTargets % 2) == 0).all() and (res.DistancesToTargets >= 0).all() and (res.DistancesToTargets <= n).all()
    assert ((res.Stable == 1) | (res.NumSteps == 101)).all()
    assert (res.NumSteps >= 2).all() and (res.NumSteps <= 101).all()
    net.close()


@pytest.mark.parametrize("n,p,b", [(8192, 24, 6), (16384, 16, 3), (5000, 60, 12)])
def test_large_dimensions(built, n, p, b):
    """UPT=32 / 64 register tilings (N up to 16384, BASELINE config 5 size) and a dimension that is not a
    multiple of the tile, against the exact-integer oracle."""
    onet, k2, pool, probes, targets = make_problem(n, p, b, 77, pool_rows=100)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=False, fast_k2=k2)
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True)
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    net.close()


def test_delta_epoch_config5_size(built):
    """One Delta epoch at N=16384 (2 GiB W): relaxed copies and W bit-exact against the oracle."""
    n, p = 16384, 6
    rng = orc.Rng(78)
    t = rng.generate_states(0, n, p)
    noised = t.copy()
    for s in range(p):
        rng.apply_noise(1, 0.05, noised[s])
    perms = rng.fresh_perms(n, p)
    onet = orc.Net(n)
    onet.hebbian(t)                      # a non-trivial W to sweep against
    net = gpu_net(n)
    net.hebbian_epoch(t)
    rel_o = onet.delta(t, noised, perms)
    rel_g = net.delta_epoch(t, noised, perms)
    assert np.array_equal(rel_g, orc.pack(rel_o))
    assert np.array_equal(net.GetMatrix(), onet.w)
    net.close()


@pytest.mark.parametrize("n,count,lo,hi", [(100, 37, -1.0, 1.0), (4096, 64, -1.0, 1.0), (257, 20, -0.25, 1.0), (64, 5, -3.0, 0.5)])
def test_device_side_probe_generation(built, n, count, lo, hi):
    """hn_generate_probes (StateGenerator.go:44-74 on the device, LCG jump-ahead) == the host generator, bit for bit,
    and the host stream continues where the device draws ended."""
    from hopfield_network_go_b200 import hopfieldnetwork as hn
    net = gpu_net(n)
    g_dev, g_host = hn.Rand(1234 + n), hn.Rand(1234 + n)
    g_dev.Uint64(); g_host.Uint64()                       # not at the start of the stream
    net.generate_probes(count, lo, hi, rng=g_dev)
    expect = g_host.CreateStateCollection(n, count, lo, hi)
    assert np.array_equal(net.download_probes(count), expect)
    assert g_dev.Uint64() == g_host.Uint64()
    o = orc.Rng(1234 + n)
    o.uint64()
    assert np.array_equal(expect, orc.pack(o.generate_states(0, n, count, lo, hi)))
    net.close()


def test_generated_probes_relax_like_uploaded_ones(built):
    n, p, b = 512, 40, 96
    onet, _, pool, _, targets = make_problem(n, p, b, 33)
    from hopfield_network_go_b200 import hopfieldnetwork as hn
    net = gpu_net(n)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    net.generate_probes(b, rng=hn.Rand(99))
    probes = net.download_probes(b)
    net.upload_perm_pool(pool)
    net.relax_resident(dedup=True)
    import ctypes as C
    from hopfield_network_go_b200 import capi
    finals = np.zeros_like(probes); steps = np.zeros(b, dtype=np.int32)
    out = capi.HnRelaxOut()
    out.final_states, out.num_steps = capi.ptr(finals), capi.ptr(steps)
    capi.check(capi.load().hn_download_results(net._h, C.byref(out)))
    ref = net.relax_batch(probes, pool, dedup=True)
    assert np.array_equal(finals, ref.FinalStates) and np.array_equal(steps, ref.NumSteps)
    ores = onet.relax_batch(orc.unpack(probes, n), pool, threads=8)
    assert_relax_equal(ref, ores, n)
    net.close()
