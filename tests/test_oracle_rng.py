"""CPU: pin the restated reference RNG (golang.org/x/exp/rand PCG source + Rand methods).
The Go module source is absent (go.mod:12), so the core generator is pinned against numpy's
independent implementation of the same published algorithm (PCG XSL-RR 128/64), and the product
library's own restatement (csrc/host_rng.cpp) is checked against the oracle's (oracle/xrand.c)."""
import numpy as np
import pytest

from oracle import orc

PCG_INC = (6364136223846793005 << 64) | 1442695040888963407


def numpy_pcg(seed):
    bg = np.random.PCG64()
    st = bg.state
    st["state"] = {"state": (seed << 64) | seed, "inc": PCG_INC}   # PCGSource.Seed: low = high = seed
    bg.state = st
    return bg


@pytest.mark.parametrize("seed", [0, 1, 12345, 2**63 + 17, 2**64 - 1])
def test_pcg_matches_numpy_pcg64(built, seed):
    r = orc.Rng(seed)
    mine = [r.uint64() for _ in range(64)]
    ref = [int(x) for x in numpy_pcg(seed).random_raw(64)]
    assert mine == ref


def test_uint64n_power_of_two_masks_and_rejection_is_unbiased(built):
    r, q = orc.Rng(9), numpy_pcg(9)
    for n in (1, 2, 64, 1 << 53):
        assert r.uint64n(n) == int(q.random_raw(1)[0]) & (n - 1)
    # non power of two: v % n of the same draw (rejection region is astronomically small for small n)
    for n in (3, 10, 100, 4096 - 1):
        assert r.uint64n(n) == int(q.random_raw(1)[0]) % n


def test_float64_is_53_bit_fraction(built):
    r, q = orc.Rng(5), numpy_pcg(5)
    for _ in range(100):
        f = r.float64()
        assert f == (int(q.random_raw(1)[0]) & ((1 << 53) - 1)) / float(1 << 53)
        assert 0.0 <= f < 1.0


def test_shuffle_is_fisher_yates_from_the_top(built):
    n = 50
    r, q = orc.Rng(77), numpy_pcg(77)
    lst = np.arange(n, dtype=np.uint16)
    r.shuffle_u16(lst)
    ref = list(range(n))
    for i in range(n - 1, 0, -1):            # Rand.Shuffle: j = Uint64n(i+1), swap(i, j)
        m = i + 1
        v = int(q.random_raw(1)[0])
        j = v & (m - 1) if (m & (m - 1)) == 0 else v % m
        ref[i], ref[j] = ref[j], ref[i]
    assert lst.tolist() == ref
    assert sorted(lst.tolist()) == list(range(n))


def test_norm_float64_statistics_and_ziggurat_constants(built):
    r = orc.Rng(2024)
    x = np.array([r.norm_float64() for _ in range(200000)])
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    assert abs((np.abs(x) > 1.959964).mean() - 0.05) < 0.003
    assert np.abs(x).max() > 3.442619855899  # the base-strip tail path is exercised


def test_library_rng_equals_oracle_rng(built):
    """Two independent restatements (product csrc/host_rng.cpp vs oracle/xrand.c) draw identical streams."""
    from hopfield_network_go_b200 import hopfieldnetwork as hn
    a, b = orc.Rng(31337), hn.Rand(31337)
    assert [a.uint64() for _ in range(32)] == [b.Uint64() for _ in range(32)]
    assert [a.uint64n(1000003) for _ in range(32)] == [b.Uint64n(1000003) for _ in range(32)]
    assert [a.float64() for _ in range(32)] == [b.Float64() for _ in range(32)]
    assert [a.norm_float64() for _ in range(2000)] == [b.NormFloat64() for _ in range(2000)]
    assert np.array_equal(a.perm_pool(257, 9), b.perm_pool(257, 9))
    for method, scale in ((1, 0.1), (2, 0.3), (3, 0.7)):
        s1 = np.where(np.arange(100) % 3 == 0, 1.0, -1.0)
        s2 = s1.copy()
        a.apply_noise(method, scale, s1)
        b.apply_noise(method, scale, s2)
        assert np.array_equal(s1, s2)
