"""CPU: pin the oracle (the restated reference algorithm) with hand-derivable known answers, invariants,
the reference's quirks, and an independent numpy restatement.  The reference itself ships no
tests or golden vectors and cannot be built here -> PARITY UNPINNED against Go; these are the pins."""
import numpy as np
import pytest

from oracle import orc


def np_dot_unitary(x, y):
    """gonum f64.DotUnitary order, in numpy scalars."""
    a = [np.float64(0)] * 4
    n = len(x)
    i = 0
    while i + 4 <= n:
        for l in range(4):
            a[l] = a[l] + np.float64(x[i + l]) * np.float64(y[i + l])
        i += 4
    while i < n:
        a[0] = a[0] + np.float64(x[i]) * np.float64(y[i])
        i += 1
    return (a[0] + a[2]) + (a[1] + a[3])


def np_relax(w, state, pool, max_iter=100, max_unstable=0, offset=0):
    """Independent restatement of HopfieldNetwork.go:382-441 (bipolar) in plain Python/numpy."""
    n = len(state)
    s = state.copy()
    for step in range(1, max_iter + 1):
        for k in pool[offset + step - 1]:
            h = np_dot_unitary(w[k], s)
            s[k] = -1.0 if h <= 0.0 else 1.0
        e = np.array([-0.5 * (np_dot_unitary(w[i], s) * s[i]) for i in range(n)])
        if (e > 0).sum() <= max_unstable:
            return s, step + 1, True, e
    return s, max_iter + 1, False, e


def test_dot_unitary_order(built):
    rs = np.random.default_rng(0)
    for n in (1, 3, 4, 5, 7, 8, 101, 256):
        x, y = rs.standard_normal(n), rs.standard_normal(n)
        got = orc.lib().orc_dot_unitary(x.ctypes.data, y.ctypes.data, n)
        assert got == np_dot_unitary(x, y)


def test_hebbian_single_pattern_known_answer(built):
    """P=1: W = x x^T - I before normalisation, ||W||_F = sqrt(N^2-N); every probe with non-zero overlap
    lands on +-x after one sweep -> Stable, NumSteps = 2, distance 0."""
    n = 12
    x = np.array([1, -1, 1, 1, -1, -1, 1, -1, 1, 1, 1, -1.0])
    net = orc.Net(n)
    net.hebbian(x[None, :])
    assert np.array_equal(net.w, np.outer(x, x) - np.eye(n))
    nrm = net.normalize()
    assert nrm == pytest.approx(np.sqrt(n * n - n), rel=1e-15)
    assert np.linalg.norm(net.w) == pytest.approx(1.0, rel=1e-14)
    net.set_targets(x[None, :])
    rng = orc.Rng(3)
    probes = rng.generate_states(0, n, 40)
    pool = rng.perm_pool(n, 100)
    res = net.relax_batch(probes.copy(), pool)
    overlap = probes @ x
    nz = overlap != 0
    assert res["stable"][nz].all() and (res["num_steps"][nz] == 2).all()
    assert (res["distances"][nz, 0] == 0).all()
    assert np.array_equal(res["final"][nz], np.where(overlap[nz, None] > 0, x, -x))


def test_hand_computed_n4(built):
    """N=4, two patterns; every number below is computed by hand."""
    a, b = np.array([1.0, 1, -1, -1]), np.array([1.0, -1, 1, -1])
    net = orc.Net(4)
    net.hebbian(np.stack([a, b]))
    assert np.array_equal(net.w, np.array([[0, 0, 0, -2], [0, 0, -2, 0], [0, -2, 0, 0], [-2, 0, 0, 0.0]]))
    assert net.normalize() == 4.0          # sqrt(4 * 2^2)
    assert np.array_equal(net.w * 4, np.array([[0, 0, 0, -2], [0, 0, -2, 0], [0, -2, 0, 0], [-2, 0, 0, 0.0]]))
    assert np.array_equal(net.all_unit_energies(a), np.full(4, -0.25))   # e_i = -0.5*h_i*s_i, h = W s = 0.5*s
    assert net.state_energy(a) == -1.0
    assert net.state_is_stable(a) == (True, 0)
    s = np.array([1.0, 1, 1, 1])                   # h = (-.5,-.5,-.5,-.5): every unit unstable
    assert np.array_equal(net.all_unit_energies(s), np.full(4, 0.25))
    assert net.state_is_stable(s) == (False, 4)
    assert net.unit_energy(s, 0) == 0.25


def test_quirks(built):
    n = 10
    # last unit never inverted by MaximalInversion; count = int(N*ratio)  (NoiseApplication.go:64-75)
    for seed in range(30):
        s = np.ones(n)
        orc.Rng(seed).apply_noise(1, 0.35, s)
        assert (s == -1).sum() == 3 and s[-1] == 1.0
    # h == 0 -> -1 (bipolar) / 0 (binary); e == 0 counts as stable
    net = orc.Net(3)  # zero matrix: every field is 0
    pool = np.array([[0, 1, 2]] * 3, dtype=np.uint16)
    res = net.relax_batch(np.array([[1.0, 1, -1]]), pool)
    assert np.array_equal(res["final"][0], [-1, -1, -1]) and res["stable"][0] == 1 and res["num_steps"][0] == 2
    netb = orc.Net(3, domain=1)
    resb = netb.relax_batch(np.array([[1.0, 1, 0]]), pool)
    assert np.array_equal(resb["final"][0], [0, 0, 0])
    # Hebbian is re-applied every epoch until stable (LearningMethod.go:41-58)
    x = orc.Rng(1).generate_states(0, 8, 1)
    n1 = orc.Net(8); n1.hebbian(x)
    n2 = orc.Net(8); n2.hebbian(x); n2.hebbian(x)
    assert np.array_equal(n2.w, 2 * n1.w)
    # unstable after maxIter sweeps => NumSteps = maxIter + 1
    w = np.zeros((4, 4)); w[0, 1] = 1.0; w[1, 0] = -1.0
    nf = orc.Net(4, force_symmetric=False, w=w, max_iter=7)
    pool7 = np.tile(np.arange(4, dtype=np.uint16), (7, 1))
    r = nf.relax_batch(np.array([[1.0, 1, 1, 1]]), pool7)
    assert r["stable"][0] == 0 and r["num_steps"][0] == 8


def test_invariants_after_learning(built):
    n, p = 60, 6
    rng = orc.Rng(8)
    t = rng.generate_states(0, n, p)
    for rule in (0, 2):
        net = orc.Net(n)
        net.learn_states(t.copy(), rule, 0, 10, 1, 0.1, orc.Rng(4))
        assert np.linalg.norm(net.w) == pytest.approx(1.0, rel=1e-13)
        assert np.array_equal(np.diag(net.w), np.zeros(n))
        assert np.array_equal(net.w, net.w.T)
    # energy profile of s equals that of -s bit for bit (the dedup key collides by construction)
    s = rng.generate_states(0, n, 1)[0]
    assert np.array_equal(net.all_unit_energies(s), net.all_unit_energies(-s))
    # Manhattan with inversion = 2*min(dH, N-dH)
    net.set_targets(t)
    d = np.zeros(p)
    orc.lib().orc_distances_to_targets(net.ref(), s.ctypes.data, d.ctypes.data)
    dh = (t != s).sum(axis=1)
    assert np.array_equal(d, 2.0 * np.minimum(dh, n - dh))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_independent_numpy_restatement(built, seed):
    n, p, b = 24, 3, 6
    rng = orc.Rng(seed)
    t = rng.generate_states(0, n, p)
    net = orc.Net(n)
    net.hebbian(t)
    net.normalize()
    # numpy side learns independently: sum of outer products, zero diag, symmetrise, Frobenius scale
    w = sum(np.outer(x, x) for x in t)
    np.fill_diagonal(w, 0.0)
    w = 0.5 * (w + w.T)
    w = (1 / np.sqrt((w * w).sum())) * w
    assert np.allclose(net.w, w, rtol=1e-13, atol=0)
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 100)
    res = net.relax_batch(probes.copy(), pool)
    for i in range(b):
        s, steps, stable, e = np_relax(net.w, probes[i], pool)
        assert np.array_equal(res["final"][i], s)
        assert res["num_steps"][i] == steps and bool(res["stable"][i]) == stable
        assert np.array_equal(res["energies"][i], e)


def test_unique_table_first_seen_order_and_hits(built):
    e = np.array([[1.0, 2], [3, 4], [1, 2], [5, 6], [3, 4], [1, 2]])
    aid, first, hits = orc.unique_table(e)
    assert first.tolist() == [0, 1, 3] and hits.tolist() == [3, 2, 1] and aid.tolist() == [0, 1, 0, 2, 1, 0]
    z = np.array([[0.0, 1.0], [-0.0, 1.0]])       # fmt.Sprint prints "-0" vs "0": distinct keys
    assert len(orc.unique_table(z)[1]) == 2


def test_serial_stream_offsets(built):
    """threads=1: probe p starts at pool row sum_{q<p} sweeps(q)."""
    n, p, b = 40, 4, 12
    rng = orc.Rng(6)
    t = rng.generate_states(0, n, p)
    net = orc.Net(n); net.hebbian(t); net.normalize(); net.set_targets(t)
    probes = rng.generate_states(0, n, b)
    pool = rng.perm_pool(n, 400)
    serial = net.relax_batch(probes.copy(), pool, serial_stream=True)
    offs = np.concatenate([[0], np.cumsum(serial["num_steps"] - 1)[:-1]]).astype(np.int32)
    explicit = net.relax_batch(probes.copy(), pool, offsets=offs)
    for k in ("final", "num_steps", "stable", "distances"):
        assert np.array_equal(serial[k], explicit[k])


@pytest.mark.parametrize("n,p,domain", [(64, 6, 0), (96, 10, 0), (80, 8, 1), (128, 12, 0)])
def test_fast_oracle_equals_slow_oracle(built, n, p, domain):
    """oracle_fast.c (exact-integer fields, used for BASELINE-size parity) == the literal restatement."""
    rng = orc.Rng(100 + n)
    t = rng.generate_states(domain, n, p)
    net = orc.Net(n, domain=domain)
    net.hebbian(t)
    k2 = np.rint(net.w * 2).astype(np.int32)
    net.normalize()
    net.set_targets(t)
    probes = rng.generate_states(domain, n, 40)
    pool = rng.perm_pool(n, 100)
    slow = net.relax_batch(probes.copy(), pool, threads=4)
    fast = net.relax_batch(probes.copy(), pool, threads=4, fast_k2=k2)
    for k in ("final", "num_steps", "stable", "distances", "flips", "margin_hits", "energies"):
        assert np.array_equal(slow[k], fast[k]), k


def test_frobenius_variants_against_independent_restatements():
    """orc_frobenius_variant: 0 = classic Dlassq (sequential scale/ssq recurrence), 1 = the LAPACK-3.10 rewrite, which for
    weights of ordinary magnitude is a plain sequential sum of squares per row plus the running total.  Checked against
    Python restatements written from the algorithms' published descriptions, on a quarter-integer Hebbian matrix (where
    variant 1 must equal the exact integer sum) and on a generic float matrix."""
    import math
    import ctypes as C
    rng = orc.Rng(5)
    t = rng.generate_states(0, 60, 7)
    w_h = (t.T @ t - 7 * np.eye(60)) * 0.5
    w_f = np.random.default_rng(3).normal(size=(37, 37)) * 1e-3

    def classic(w):
        scale, ssq = 0.0, 1.0
        for x in w.ravel():
            a = abs(float(x))
            if a > 0:
                if scale < a:
                    r = scale / a
                    ssq = 1.0 + (ssq * r) * r
                    scale = a
                else:
                    r = a / scale
                    ssq += r * r
        return scale * math.sqrt(ssq)

    def lapack310(w):   # medium range only: amed accumulates x*x along the row, then the running total scl^2*sumsq
        scale, ssq = 0.0, 1.0
        for row in w:
            if ssq == 0.0:
                scale = 1.0
            if scale == 0.0:
                scale, ssq = 1.0, 0.0
            amed = 0.0
            for x in row:
                amed += float(x) * float(x)
            if ssq > 0.0:
                amed += scale * scale * ssq
            scale, ssq = 1.0, amed
        return scale * math.sqrt(ssq)

    for w in (w_h, w_f):
        w = np.ascontiguousarray(w, dtype=np.float64)
        n = w.shape[0]
        got0 = orc.lib().orc_frobenius_variant(w.ctypes.data_as(C.c_void_p), n, 0)
        got1 = orc.lib().orc_frobenius_variant(w.ctypes.data_as(C.c_void_p), n, 1)
        assert got0 == classic(w)
        assert got1 == lapack310(w)
    exact = math.sqrt(float(np.sum(np.rint(w_h * 4).astype(np.int64) ** 2)) / 16.0)
    n = w_h.shape[0]
    wc = np.ascontiguousarray(w_h, dtype=np.float64)
    assert orc.lib().orc_frobenius_variant(wc.ctypes.data_as(C.c_void_p), n, 1) == exact
