"""ctypes loader for the CPU oracle (oracle/build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg.  The product package never
imports this module.  PARITY UNPINNED (see oracle/hopfield_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("hopfield_oracle.c", "oracle_fast.c", "xrand.c", "hopfield_oracle.h", "xrand.h")
    ):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class OrcNet(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("domain", C.c_int32), ("force_symmetric", C.c_int32),
        ("force_zero_diag", C.c_int32), ("max_unstable", C.c_int32), ("max_iter", C.c_int32),
        ("lr", C.c_double), ("w", C.c_void_p), ("n_targets", C.c_int32), ("targets", C.c_void_p),
    ]


class OrcRelaxInfo(C.Structure):
    _fields_ = [("stable", C.c_int32), ("num_steps", C.c_int32), ("flips", C.c_int32),
                ("margin_hits", C.c_int32), ("min_abs_h", C.c_double)]


class OrcRng(C.Structure):
    _fields_ = [("low", C.c_uint64), ("high", C.c_uint64)]


class OrcLearnRow(C.Structure):
    _fields_ = [("epoch", C.c_int32), ("index", C.c_int32), ("stable", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_dot_unitary.restype = C.c_double
        L.orc_dot_unitary.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_state_energy.restype = C.c_double
        L.orc_unit_energy.restype = C.c_double
        L.orc_frobenius.restype = C.c_double
        L.orc_frobenius.argtypes = [C.c_void_p, C.c_int32]
        L.orc_normalize.restype = C.c_double
        L.orc_normalize_variant.restype = C.c_double
        L.orc_frobenius_variant.restype = C.c_double
        L.orc_frobenius_variant.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_rng_uint64.restype = C.c_uint64
        L.orc_rng_uint64n.restype = C.c_uint64
        L.orc_rng_uint64n.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_rng_float64.restype = C.c_double
        L.orc_rng_norm_float64.restype = C.c_double
        L.orc_rng_uniform.restype = C.c_double
        L.orc_rng_uniform.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_rng_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_relax_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_relax_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int]
        L.orc_relax_batch_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_learn_states.restype = C.c_int32
        L.orc_learn_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p]
        L.orc_unique_table.restype = C.c_int32
        L.orc_apply_noise.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int32]
        L.orc_generate_states.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                          C.c_double, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Rng:
    """rand.New(rand.NewSource(seed)) of golang.org/x/exp/rand (restated)."""

    def __init__(self, seed: int):
        self.s = OrcRng()
        lib().orc_rng_seed(C.byref(self.s), C.c_uint64(seed))

    def uint64(self) -> int:
        return lib().orc_rng_uint64(C.byref(self.s))

    def uint64n(self, n: int) -> int:
        return lib().orc_rng_uint64n(C.byref(self.s), n)

    def float64(self) -> float:
        return lib().orc_rng_float64(C.byref(self.s))

    def norm_float64(self) -> float:
        return lib().orc_rng_norm_float64(C.byref(self.s))

    def shuffle_u16(self, lst: np.ndarray) -> None:
        assert lst.dtype == np.uint16 and lst.flags.c_contiguous
        lib().orc_rng_shuffle_u16(C.byref(self.s), _p(lst), C.c_int32(lst.size))

    def perm_pool(self, n: int, rows: int) -> np.ndarray:
        """rows successive in-place shuffles of one persistent identity list
        (HopfieldNetwork.go:374,393)."""
        lst = np.arange(n, dtype=np.uint16)
        pool = np.empty((rows, n), dtype=np.uint16)
        for r in range(rows):
            self.shuffle_u16(lst)
            pool[r] = lst
        return pool

    def fresh_perms(self, n: int, rows: int) -> np.ndarray:
        """rows independent shuffles of a fresh identity list (HopfieldNetwork.go:281-282)."""
        out = np.empty((rows, n), dtype=np.uint16)
        for r in range(rows):
            lst = np.arange(n, dtype=np.uint16)
            self.shuffle_u16(lst)
            out[r] = lst
        return out

    def generate_states(self, domain: int, n: int, count: int, mn=-1.0, mx=1.0) -> np.ndarray:
        out = np.empty((count, n), dtype=np.float64)
        lib().orc_generate_states(C.byref(self.s), domain, n, count, mn, mx, _p(out))
        return out

    def apply_noise(self, method: int, scale: float, state: np.ndarray) -> None:
        assert state.dtype == np.float64 and state.flags.c_contiguous
        lib().orc_apply_noise(C.byref(self.s), method, scale, _p(state), state.size)


class Net:
    """orc_net wrapper holding numpy buffers alive."""

    def __init__(self, n, domain=0, force_symmetric=True, force_zero_diag=True, max_unstable=0,
                 max_iter=100, lr=1.0, w=None):
        self.n = n
        self.w = np.zeros((n, n), dtype=np.float64) if w is None else np.ascontiguousarray(w, dtype=np.float64).copy()
        self.targets = np.zeros((0, n), dtype=np.float64)
        self.c = OrcNet(n, domain, int(force_symmetric), int(force_zero_diag), max_unstable, max_iter,
                        lr, self.w.ctypes.data, 0, None)

    @property
    def domain(self):
        return self.c.domain

    def set_weights(self, w):
        self.w[...] = w

    def set_targets(self, t):
        self.targets = np.ascontiguousarray(t, dtype=np.float64).copy().reshape(-1, self.n)
        self.c.n_targets = self.targets.shape[0]
        self.c.targets = self.targets.ctypes.data if self.targets.size else None

    def ref(self):
        return C.byref(self.c)

    # ---- energies / stability
    def all_unit_energies(self, state):
        e = np.empty(self.n)
        lib().orc_all_unit_energies(self.ref(), _p(np.ascontiguousarray(state, dtype=np.float64)), _p(e))
        return e

    def state_is_stable(self, state):
        cnt = C.c_int32(0)
        st = lib().orc_state_is_stable(self.ref(), _p(np.ascontiguousarray(state, dtype=np.float64)), C.byref(cnt))
        return bool(st), cnt.value

    def state_energy(self, state):
        return lib().orc_state_energy(self.ref(), _p(np.ascontiguousarray(state, dtype=np.float64)))

    def unit_energy(self, state, i):
        return lib().orc_unit_energy(self.ref(), _p(np.ascontiguousarray(state, dtype=np.float64)), C.c_int32(i))

    def margin(self):
        a, b = C.c_double(), C.c_double()
        lib().orc_margin(self.ref(), C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- learning
    def hebbian(self, states):
        s = np.ascontiguousarray(states, dtype=np.float64)
        lib().orc_hebbian(self.ref(), _p(s), C.c_int32(s.shape[0]))

    def delta(self, states, noised, perms):
        s = np.ascontiguousarray(states, dtype=np.float64)
        nz = np.ascontiguousarray(noised, dtype=np.float64)
        pm = np.ascontiguousarray(perms, dtype=np.uint16)
        rel = np.empty_like(s)
        lib().orc_delta(self.ref(), _p(s), _p(nz), _p(pm), C.c_int32(s.shape[0]), _p(rel))
        return rel

    def thermal_delta(self, states, noised, perms):
        s = np.ascontiguousarray(states, dtype=np.float64)
        nz = np.ascontiguousarray(noised, dtype=np.float64)
        pm = np.ascontiguousarray(perms, dtype=np.uint16)
        rel = np.empty_like(s)
        lib().orc_thermal_delta(self.ref(), _p(s), _p(nz), _p(pm), C.c_int32(s.shape[0]), _p(rel))
        return rel

    def enforce_constraints(self):
        lib().orc_enforce_constraints(self.ref())

    def normalize(self, variant=1):
        """variant 1: the LAPACK-3.10 Dlassq (the default everywhere), 0: the classic recurrence (hopfield_oracle.c)"""
        return lib().orc_normalize_variant(self.ref(), C.c_int32(variant))

    def update_state(self, state, perm):
        s = np.ascontiguousarray(state, dtype=np.float64).copy()
        lib().orc_update_state(self.ref(), _p(s), _p(np.ascontiguousarray(perm, dtype=np.uint16)))
        return s

    def learn_states(self, states, rule, method, epochs, noise_method, noise_scale, rng: Rng, cap=0,
                     want_energies=False):
        """LearnStates; `states` (float64 [P][n]) are activated in place and appended to targets."""
        assert states.dtype == np.float64 and states.flags.c_contiguous
        p = states.shape[0]
        rows = (OrcLearnRow * max(cap, 1))()
        en = np.empty((max(cap, 1), self.n)) if want_energies else None
        nrows = lib().orc_learn_states(self.ref(), _p(states), p, rule, method, epochs, noise_method,
                                       noise_scale, C.byref(rng.s), cap, rows, _p(en))
        self.set_targets(np.concatenate([self.targets, states], axis=0))
        out_rows = [(rows[i].epoch, rows[i].index, bool(rows[i].stable)) for i in range(min(nrows, cap))]
        return nrows, out_rows, (en[:min(nrows, cap)] if want_energies else None)

    # ---- relaxation
    def relax_batch(self, states, pool, offsets=None, serial_stream=False, threads=1,
                    want_energies=True, faithful_cost=False, fast_k2=None):
        """Relax states (float64 [B][n]) in place. Returns dict of numpy arrays."""
        assert states.dtype == np.float64 and states.flags.c_contiguous
        b = states.shape[0]
        pool = np.ascontiguousarray(pool, dtype=np.uint16)
        info = (OrcRelaxInfo * max(b, 1))()
        p = self.c.n_targets
        dist = np.zeros((b, p), dtype=np.float64)
        en = np.zeros((b, self.n), dtype=np.float64) if want_energies else None
        off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int32)
        if fast_k2 is None:
            rc = lib().orc_relax_batch(self.ref(), _p(states), b, _p(pool), pool.shape[0], _p(off),
                                       int(serial_stream), threads, info, _p(dist), _p(en),
                                       int(faithful_cost))
        else:
            k2 = np.ascontiguousarray(fast_k2, dtype=np.int32)
            k2t = k2 if np.array_equal(k2, k2.T) else np.ascontiguousarray(k2.T)
            rc = lib().orc_relax_batch_fast(self.ref(), _p(k2), _p(k2t), _p(states), b, _p(pool),
                                            pool.shape[0], _p(off), int(serial_stream), threads, info,
                                            _p(dist), _p(en))
        if rc != 0:
            raise RuntimeError(f"oracle relax failed rc={rc} (pool exhausted?)")
        return {
            "stable": np.array([info[i].stable for i in range(b)], dtype=np.uint8),
            "num_steps": np.array([info[i].num_steps for i in range(b)], dtype=np.int32),
            "flips": np.array([info[i].flips for i in range(b)], dtype=np.int32),
            "margin_hits": np.array([info[i].margin_hits for i in range(b)], dtype=np.int32),
            "min_abs_h": np.array([info[i].min_abs_h for i in range(b)]),
            "distances": dist, "energies": en, "final": states,
        }


def unique_table(energies: np.ndarray):
    """handleUniqueRelaxedState over probes in index order -> (attractor_id, first, hits)."""
    e = np.ascontiguousarray(energies, dtype=np.float64)
    b, n = e.shape
    aid = np.empty(b, dtype=np.int32)
    first = np.empty(max(b, 1), dtype=np.int32)
    hits = np.empty(max(b, 1), dtype=np.int32)
    nu = lib().orc_unique_table(_p(e), b, n, _p(aid), _p(first), _p(hits))
    return aid, first[:nu].copy(), hits[:nu].copy()


def pack(states: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(states, dtype=np.float64)
    s2 = s.reshape(-1, s.shape[-1])
    out = np.zeros((s2.shape[0], (s2.shape[1] + 63) // 64), dtype=np.uint64)
    lib().orc_pack(s2.shape[1], s2.shape[0], _p(s2), _p(out))
    return out


def unpack(packed: np.ndarray, n: int, domain: int = 0) -> np.ndarray:
    pk = np.ascontiguousarray(packed, dtype=np.uint64).reshape(-1, (n + 63) // 64)
    out = np.empty((pk.shape[0], n), dtype=np.float64)
    lib().orc_unpack(n, domain, pk.shape[0], _p(pk), _p(out))
    return out
