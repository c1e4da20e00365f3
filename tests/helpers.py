"""Shared builders for the parity tests: identical seeded inputs for oracle and GPU."""
from __future__ import annotations

import numpy as np

from oracle import orc


def make_problem(n, p, b, seed, domain=0, rule="hebbian", pool_rows=100, normalize=True):
    """Targets / probes / pool from the restated reference RNG; W learned by the ORACLE.
    Returns (oracle net, integer matrix K2 with W = c*K2 (or None), pool, probes f64, targets f64)."""
    rng = orc.Rng(seed)
    targets = rng.generate_states(domain, n, p)
    probes = rng.generate_states(domain, n, b)
    pool = rng.perm_pool(n, pool_rows)
    net = orc.Net(n, domain=domain)
    k2 = None
    if rule == "hebbian":
        net.hebbian(targets)
        k2 = np.rint(net.w * 2).astype(np.int32)  # entries are multiples of 0.5 (exactly representable)
        assert np.array_equal(k2 * 0.5, net.w)
    if normalize:
        net.normalize()
    net.set_targets(targets)
    return net, k2, pool, probes, targets


def gpu_net(n, domain=0, **kw):
    from hopfield_network_go_b200 import hopfieldnetwork as hn
    b = hn.NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetNetworkDomain(domain).SetEpochs(kw.pop("epochs", 1))
    b.SetNetworkLearningRule(kw.pop("rule", hn.HebbianLearningRule)).SetNetworkLearningMethod(kw.pop("method", hn.FullSetMethod))
    b.SetSeed(kw.pop("seed", 1))
    kw.setdefault("SetColumnStream", 0)   # HN_COLUMNS_F64 unless a test asks for another stream (AUTO: test_gpu_auto_columns.py)
    for k, v in kw.items():
        getattr(b, k)(v)
    return b.Build()


def assert_relax_equal(res, ores, n, check_counters=True):
    assert np.array_equal(res.FinalStates, orc.pack(ores["final"])), "final states differ"
    assert np.array_equal(res.NumSteps, ores["num_steps"]), "NumSteps differ"
    assert np.array_equal(res.Stable, ores["stable"]), "Stable flags differ"
    if res.DistancesToTargets is not None:
        assert np.array_equal(res.DistancesToTargets, ores["distances"].astype(np.int32)), "distances differ"
    if check_counters:
        assert np.array_equal(res.Flips, ores["flips"]), "flip counters differ"
