"""-m gpu: the Blackwell-native integer fields contraction (tcgen05.mma.kind::i8, TMEM accumulators, TMA-fed weight
tiles; csrc/igemm_tc05.cuh) against plain integer arithmetic and against the mma.sync kernel it replaces."""
import os

import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from oracle import orc
from tests.helpers import gpu_net

pytestmark = pytest.mark.gpu


def _image(wraw, domain):
    """The int16 image the library builds from the un-normalised matrix: M = 2*scale*Wraw with the smallest scale in
    {0.5, 1, 2, 4} that makes scale*Wraw integer, and -2/delta added on the diagonal (kernels_misc.cuh: to_i16_kernel)."""
    for scale in (0.5, 1.0, 2.0, 4.0):
        k = wraw * scale
        if np.array_equal(k, np.rint(k)):
            break
    m = np.rint(2.0 * k).astype(np.int64)
    m[np.diag_indices_from(m)] += -2 if domain else -1
    return m


@pytest.mark.parametrize("n,p,b,domain", [(4096, 400, 300, 0), (1024, 100, 1000, 0), (100, 10, 37, 0), (257, 26, 129, 0),
                                          (1000, 51, 260, 0), (192, 16, 96, 1), (2048, 64, 1, 0), (6000, 40, 140, 0)])
def test_tc05_fields_equal_integer_arithmetic(built, n, p, b, domain):
    rng = orc.Rng(4000 + n)
    t = rng.generate_states(domain, n, p)
    onet = orc.Net(n, domain=domain)
    onet.hebbian(t)
    m = _image(onet.w, domain)
    assert np.abs(m).max() <= 127
    net = gpu_net(n, domain, SetColumnStream=capi.HN_COLUMNS_AUTO)
    net.LearnStates(t.copy())
    assert net.integer_stream_active()
    states = rng.generate_states(domain, n, b)
    got, kernel = net.debug_integer_fields(states)
    assert kernel == 1, "the tcgen05 kernel must be the one that ran"
    want = states.astype(np.int64) @ m.T
    assert np.array_equal(got.astype(np.int64), want)
    net.close()


def test_tc05_equals_mma_sync_kernel(built):
    n, p, b = 1536, 120, 777
    rng = orc.Rng(4242)
    t = rng.generate_states(0, n, p)
    states = rng.generate_states(0, n, b)
    net = gpu_net(n, SetColumnStream=capi.HN_COLUMNS_AUTO)
    net.LearnStates(t.copy())
    a, ka = net.debug_integer_fields(states)
    net.close()
    os.environ["HN_IGEMM"] = "mma_sync"
    try:
        old = gpu_net(n, SetColumnStream=capi.HN_COLUMNS_AUTO)
        old.LearnStates(t.copy())
        c, kc = old.debug_integer_fields(states)
        old.close()
    finally:
        del os.environ["HN_IGEMM"]
    assert (ka, kc) == (1, 0)
    assert np.array_equal(a, c)
