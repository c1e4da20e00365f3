"""Regenerates tests/golden/*.npz from the CPU oracle (python tests/golden/make_golden.py).

The reference is Go and cannot be run here (no toolchain; gonum / x/exp sources absent), and it ships no
golden vectors of its own, so these fixtures freeze the ORACLE's outputs on fixed seeded
inputs: they pin the oracle against regressions and give the GPU tests a committed, oracle-independent target
on the GPU box.  Inputs are stored too, so a Go maintainer can replay them against the real reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def relax_case(name, n, p, b, seed, domain=0, rule=0, epochs=1, noise=0, noise_scale=0.0, serial=False):
    rng = orc.Rng(seed)
    targets = rng.generate_states(domain, n, p)
    net = orc.Net(n, domain=domain)
    nrows, rows, _ = net.learn_states(targets.copy(), rule, 0, epochs, noise, noise_scale, orc.Rng(seed + 1), cap=epochs * p)
    probes = rng.generate_states(domain, n, b)
    pool = rng.perm_pool(n, 100 if not serial else 40 * b)
    res = net.relax_batch(probes.copy(), pool, serial_stream=serial, threads=1)
    aid, first, hits = orc.unique_table(res["energies"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        n=n, domain=domain, rule=rule, epochs=epochs, noise=noise, noise_scale=noise_scale,
                        learn_seed=seed + 1, serial=int(serial),
                        targets=orc.pack(targets), probes=orc.pack(probes), pool=pool, weights=net.w,
                        learn_rows=np.array(rows, dtype=np.int32).reshape(-1, 3),
                        final=orc.pack(res["final"]), num_steps=res["num_steps"], stable=res["stable"],
                        distances=res["distances"].astype(np.int32), energies=res["energies"],
                        flips=res["flips"], margin_hits=res["margin_hits"],
                        attractor_id=aid, unique_first=first, unique_hits=hits)


if __name__ == "__main__":
    relax_case("hebbian_n100_p10", 100, 10, 64, 1001, epochs=100)                       # BASELINE config 1 shape
    relax_case("hebbian_n100_serial", 100, 10, 48, 1002, epochs=100, serial=True)       # threads=1 stream
    relax_case("delta_n128_inversion", 128, 12, 48, 1003, rule=2, epochs=25, noise=1, noise_scale=0.1)
    relax_case("delta_n96_gaussian", 96, 9, 40, 1004, rule=2, epochs=20, noise=3, noise_scale=0.5)
    relax_case("binary_n80_hebbian", 80, 6, 40, 1005, domain=1, epochs=3)
    print("golden fixtures written to", HERE)
