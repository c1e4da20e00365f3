"""-m gpu: the dense lockstep sweeps (csrc/dense.cuh: the first sweeps of a relaxation per POSITION on the tcgen05 path,
32 probes in lockstep) against the oracle and against the sparse-only path.  Everything must be identical: final
states, NumSteps, Stable, distances, flip and exact-tie counters, unique tables."""
import os

import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from oracle import orc
from tests.helpers import assert_relax_equal, gpu_net, make_problem

pytestmark = pytest.mark.gpu
AUTO = capi.HN_COLUMNS_AUTO


def _env(**kw):
    class Ctx:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kw}
            os.environ.update({k: str(v) for k, v in kw.items()})
        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return Ctx()


def _full_compare(a, b):
    for f in ("FinalStates", "NumSteps", "Stable", "DistancesToTargets", "Flips", "MarginHits", "AttractorId", "UniqueFirst", "UniqueHits"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.parametrize("n,p,b,seed,domain", [(4096, 400, 96, 16, 0), (2048, 100, 70, 3, 0), (2500, 120, 65, 5, 0),
                                               (3000, 61, 33, 7, 0), (2048, 64, 40, 9, 1), (4096, 41, 1, 11, 0)])
def test_dense_sweeps_match_oracle(built, n, p, b, seed, domain):
    onet, k2, pool, probes, targets = make_problem(n, p, b, seed, domain=domain)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=False, fast_k2=k2)
    net = gpu_net(n, domain, SetColumnStream=AUTO)
    net.SetMatrix(onet.w)
    assert net.integer_stream_active()
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True)
    assert net.last_timing().dense_sweeps >= 1, "the dense lockstep kernel must have run"
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    with _env(HN_DENSE=0):
        ref = gpu_net(n, domain, SetColumnStream=AUTO)
    ref.SetMatrix(onet.w)
    ref.SetLearnedStates(orc.pack(targets))
    r2 = ref.relax_batch(probes, pool, dedup=True)
    assert ref.last_timing().dense_sweeps == 0
    _full_compare(res, r2)
    net.close()
    ref.close()


@pytest.mark.parametrize("n,p,b,max_iter,max_unstable,rows", [(300, 20, 70, 100, 0, 6), (512, 40, 64, 2, 0, 6), (1000, 51, 40, 1, 0, 6),
                                                             (640, 30, 50, 100, 3, 2), (333, 8, 37, 5, 0, 9), (1024, 100, 96, 100, 0, 3)])
def test_dense_sweeps_small_dimensions_and_budgets(built, n, p, b, max_iter, max_unstable, rows):
    """Small, ragged dimensions (forced with HN_DENSE_MIN_N), sweep budgets shorter than the dense phase, a tolerance
    of unstable units, tie-rich small networks: dense + sparse == sparse only == oracle."""
    onet, k2, pool, probes, targets = make_problem(n, p, b, 100 + n)
    onet.c.max_iter, onet.c.max_unstable = max_iter, max_unstable
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=True)
    kw = dict(SetColumnStream=AUTO, SetMaximumRelaxationIterations=max_iter, SetMaximumRelaxationUnstableUnits=max_unstable)
    with _env(HN_DENSE_MIN_N=64, HN_DENSE_ROWS=rows, HN_DENSE_MIN_FLIPS=0):
        net = gpu_net(n, **kw)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True)
    assert net.last_timing().dense_sweeps >= 1
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    _, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    net.close()


def test_dense_sweeps_with_exact_ties(built):
    """Even P at N=2048: exact ties of the true field inside the dense sweeps are decided by the float64 row in gonum's
    order, exactly as the sparse kernel and the oracle decide them (noisy copies of targets: attractors repeat)."""
    n, p = 2048, 200
    rng = orc.Rng(7771)
    targets = rng.generate_states(0, n, p)
    pool = rng.perm_pool(n, 100)
    onet = orc.Net(n)
    onet.hebbian(targets)
    k2 = np.rint(onet.w * 2).astype(np.int32)
    onet.normalize()
    onet.set_targets(targets)
    rs = np.random.Generator(np.random.PCG64(5))
    probes = []
    for t in targets[:40]:
        for c in range(4):
            x = t.copy() if c != 3 else -t
            idx = rs.permutation(n)[: int(n * rs.uniform(0.05, 0.3))]
            x[idx] = -x[idx]
            probes.append(x)
    probes = np.ascontiguousarray(np.array(probes))
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=True, fast_k2=k2)
    assert ores["margin_hits"].sum() > 20
    net = gpu_net(n, SetColumnStream=AUTO)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    res = net.relax_batch(probes, pool, dedup=True)
    assert net.last_timing().dense_sweeps >= 1
    assert_relax_equal(res, ores, n)
    assert np.array_equal(res.MarginHits, ores["margin_hits"])
    _, first, hits = orc.unique_table(ores["energies"])
    assert np.array_equal(res.UniqueFirst, first) and np.array_equal(res.UniqueHits, hits)
    net.close()


@pytest.mark.parametrize("n,p,b,domain,dense", [(4096, 400, 200, 0, 1), (1000, 1030, 150, 0, 0), (257, 30, 77, 1, 0)])
def test_zero_copy_results_match_copied_results(built, n, p, b, domain, dense):
    """Page-locked result buffers: the relaxation kernel writes final states and distance rows straight into the caller's
    memory (relax.cuh epilogue, staged per probe).  Same arrays as the copy path and as the oracle, including a target
    count above one staging pass (1030 > 512)."""
    onet, k2, pool, probes, targets = make_problem(n, p, b, 21, domain=domain)
    ores = onet.relax_batch(probes.copy(), pool, threads=8, want_energies=False, fast_k2=k2)
    with _env(HN_DENSE_MIN_N=256 if dense else 1 << 20):
        net = gpu_net(n, domain, SetColumnStream=AUTO)
    net.SetMatrix(onet.w)
    net.SetLearnedStates(orc.pack(targets))
    copied = net.relax_batch(probes, pool, dedup=True)
    packed = orc.pack(probes)
    pinned_probes = net._result_buffer("probes_in", packed.shape, np.uint64, True)   # page-locked: read in place by the first dense sweep
    pinned_probes[...] = packed
    mapped = net.relax_batch(pinned_probes, pool, dedup=True, reuse_pinned_buffers=True)
    assert np.array_equal(net.download_probes(b), packed)   # ... and left in the resident probe buffer as well
    assert (net.last_timing().dense_sweeps >= 1) == bool(dense)
    assert_relax_equal(mapped, ores, n)
    _full_compare(mapped, copied)
    with _env(HN_ZERO_COPY=0):
        off = gpu_net(n, domain, SetColumnStream=AUTO)
    off.SetMatrix(onet.w)
    off.SetLearnedStates(orc.pack(targets))
    _full_compare(off.relax_batch(probes, pool, dedup=True, reuse_pinned_buffers=True), copied)
    net.close()
    off.close()
