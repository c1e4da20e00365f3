"""Host-side mirror of the reference's `hopfieldnetwork` package API over the C ABI.

Same names, argument meaning and error behaviour as the Go package
(hopfieldnetwork/HopfieldNetworkBuilder.go, HopfieldNetwork.go, LearningRule.go,
LearningMethod.go); every method forwards to libhopfield_b200.so through ctypes,
exactly as the cgo package in go/hopfieldnetwork does.  This module contains no
numerical code.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import capi
from .capi import check, ptr, words

# enums: values equal the reference's iota values
BipolarDomain, BinaryDomain = 0, 1                                     # domain/DomainEnum.go:6-9
(HebbianLearningRule, BipolarMappedHebbianLearningRule, DeltaLearningRule, BipolarMappedDeltaLearningRule,
 ThermalDeltaLearningRule, BipolarMappedThermalDeltaLearningRule) = range(6)   # LearningRule.go:32-39
FullSetMethod, IterativeBatchMethod = 0, 1                             # LearningMethod.go:22-25
NoiseNone, MaximalInversion, RandomSubMaximalInversion, GaussianApplication = range(4)  # NoiseApplication.go:24-36


class Rand:
    """rand.New(rand.NewSource(seed)) — host-side restatement living in the library."""

    def __init__(self, seed: int):
        self._h = C.c_void_p()
        check(capi.load().hn_rng_create(C.c_uint64(seed), C.byref(self._h)))

    def __del__(self):
        try:
            capi.load().hn_rng_destroy(self._h)
        except Exception:
            pass

    def Uint64(self) -> int:
        return capi.load().hn_rng_uint64(self._h)

    def Uint64n(self, n: int) -> int:
        return capi.load().hn_rng_uint64n(self._h, n)

    def Advance(self, draws: int) -> None:
        """Skip `draws` Uint64 values in O(log draws) (hn_rng_advance)."""
        check(capi.load().hn_rng_advance(self._h, draws))

    def Float64(self) -> float:
        return capi.load().hn_rng_float64(self._h)

    def NormFloat64(self) -> float:
        return capi.load().hn_rng_norm_float64(self._h)

    def ShuffleList(self, lst: np.ndarray) -> None:
        assert lst.dtype == np.uint16 and lst.flags.c_contiguous
        check(capi.load().hn_rng_shuffle_u16(self._h, ptr(lst), lst.size))

    def perm_pool(self, dimension: int, rows: int) -> np.ndarray:
        lst = np.zeros(dimension, dtype=np.uint16)
        pool = np.empty((rows, dimension), dtype=np.uint16)
        check(capi.load().hn_rng_perm_pool(self._h, dimension, rows, 1, ptr(lst), ptr(pool)))
        return pool

    def apply_noise(self, method: int, scale: float, state: np.ndarray) -> None:
        assert state.dtype == np.float64 and state.flags.c_contiguous
        check(capi.load().hn_rng_apply_noise(self._h, method, scale, ptr(state), state.size))

    def CreateStateCollection(self, dimension: int, count: int, rand_min=-1.0, rand_max=1.0) -> np.ndarray:
        """states.StateGenerator.CreateStateCollection (states/StateGenerator.go:65-74), packed."""
        out = np.zeros((count, words(dimension)), dtype=np.uint64)
        check(capi.load().hn_rng_generate_states(self._h, dimension, count, rand_min, rand_max, ptr(out)))
        return out


@dataclass
class LearnStateData:          # datacollector/LearnStateData.go:14-19
    Epoch: int
    TargetStateIndex: int
    EnergyProfile: Optional[np.ndarray]
    Stable: bool


@dataclass
class RelaxationBatchResult:
    """Compact, array-of-struct-free form of []*RelaxationResult (HopfieldNetwork.go:268-273)."""
    FinalStates: np.ndarray                 # packed uint64 [B][words]
    NumSteps: np.ndarray                    # int32 [B]  len(StateHistory)
    Stable: np.ndarray                      # uint8 [B]
    DistancesToTargets: Optional[np.ndarray]  # int32 [B][P]
    EnergyProfiles: Optional[np.ndarray]    # float64 [B][N]
    Flips: np.ndarray
    MarginHits: np.ndarray
    AttractorId: Optional[np.ndarray] = None
    UniqueFirst: Optional[np.ndarray] = None
    UniqueHits: Optional[np.ndarray] = None
    History: Optional[np.ndarray] = None
    StateHash: Optional[np.ndarray] = None
    total_flips: int = 0
    total_sweeps: int = 0
    total_margin_hits: int = 0
    device_ms: float = 0.0


class HopfieldNetwork:
    def __init__(self, cfg: capi.HnConfig, epochs: int, rule: int, method: int, noise_method: int,
                 noise_scale: float, units_updated: int, seed: int, intensive: bool):
        self._cfg = cfg
        self._h = C.c_void_p()
        check(capi.load().hn_create(C.byref(cfg), C.byref(self._h)))
        self.dimension = cfg.dimension
        self.domain = cfg.domain
        self.epochs, self.rule, self.method = epochs, rule, method
        self.noise_method, self.noise_scale = noise_method, noise_scale
        self.unitsUpdatedPerStep = units_updated
        self.allowIntensiveDataCollection = intensive
        self.randomGenerator = Rand(seed)

    def close(self):
        for addr, _ in getattr(self, "_pinned", {}).values():
            capi.load().hn_host_free(C.c_void_p(addr))
        self._pinned = {}
        if self._h:
            capi.load().hn_destroy(self._h)
            self._h = C.c_void_p()

    def _result_buffer(self, name, shape, dtype, pinned):
        """Result array: a fresh numpy array, or (pinned=True) a view of a page-locked buffer of this network that
        is REUSED by the next call with pinned=True (hn_host_alloc: results arrive at PCIe speed, no page faults)."""
        dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(shape)) * dtype.itemsize, 1)
        if not pinned:
            return np.zeros(shape, dtype=dtype)
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        addr, cap = self._pinned.get(name, (0, 0))
        if cap < nbytes:
            if addr:
                check(capi.load().hn_host_free(C.c_void_p(addr)))
            p = C.c_void_p()
            check(capi.load().hn_host_alloc(C.byref(p), nbytes))
            addr, cap = p.value, nbytes
            self._pinned[name] = (addr, cap)
        raw = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(addr))
        return raw[:int(np.prod(shape)) * dtype.itemsize].view(dtype).reshape(shape)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (HopfieldNetwork.go:93-116)
    def GetMatrix(self) -> np.ndarray:
        w = np.empty((self.dimension, self.dimension), dtype=np.float64)
        check(capi.load().hn_get_weights(self._h, ptr(w)))
        return w

    def SetMatrix(self, w: np.ndarray) -> None:  # additive extension (resume)
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.shape == (self.dimension, self.dimension)
        check(capi.load().hn_set_weights(self._h, ptr(w)))

    def SetIntegerStructure(self, w_unnormalized: np.ndarray) -> None:
        """Declare W = c * w_unnormalized (2*w_unnormalized integer) for the exact integer column stream."""
        w = np.ascontiguousarray(w_unnormalized, dtype=np.float64)
        check(capi.load().hn_set_integer_structure(self._h, ptr(w)))

    def integer_stream_active(self) -> bool:
        a = C.c_int32()
        check(capi.load().hn_integer_stream_active(self._h, C.byref(a)))
        return bool(a.value)

    def GetDimension(self) -> int:
        return self.dimension

    def GetLearnedStates(self) -> np.ndarray:
        n = C.c_int32()
        check(capi.load().hn_num_targets(self._h, C.byref(n)))
        out = np.zeros((n.value, words(self.dimension)), dtype=np.uint64)
        if n.value:
            check(capi.load().hn_get_targets(self._h, ptr(out)))
        return out

    def SetLearnedStates(self, packed: np.ndarray) -> None:
        p = np.ascontiguousarray(packed, dtype=np.uint64).reshape(-1, words(self.dimension))
        check(capi.load().hn_set_targets(self._h, ptr(p), p.shape[0]))

    def margin(self):
        a, b = C.c_double(), C.c_double()
        check(capi.load().hn_get_margin(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_integer_fields(self, states):
        """(fields int32 [n][N], kernel) of hn_debug_integer_fields: kernel 1 = tcgen05/TMEM/TMA, 0 = mma.sync."""
        p = self._packed(states)
        out = np.empty((p.shape[0], self.dimension), dtype=np.int32)
        k = C.c_int32()
        check(capi.load().hn_debug_integer_fields(self._h, ptr(p), p.shape[0], ptr(out), C.byref(k)))
        return out, k.value

    def unique_table_stats(self):
        """(verified, refuted) merges of DIFFERENT states in the unique table (hn_unique_table_stats)."""
        a, b = C.c_int64(), C.c_int64()
        check(capi.load().hn_unique_table_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def _packed(self, states) -> np.ndarray:
        s = np.asarray(states)
        if s.dtype == np.uint64:
            return np.ascontiguousarray(s).reshape(-1, words(self.dimension))
        return capi.pack_states(s)

    # ---- energies / stability (HopfieldNetwork.go:171-243)
    def energy_profiles(self, states, want_energies=True):
        p = self._packed(states)
        n = p.shape[0]
        e = np.empty((n, self.dimension)) if want_energies else None
        unst = np.empty(n, dtype=np.int32)
        stab = np.empty(n, dtype=np.uint8)
        se = np.empty(n)
        check(capi.load().hn_energy_profiles(self._h, ptr(p), n, ptr(e), ptr(unst), ptr(stab), ptr(se)))
        return e, unst, stab, se

    def AllUnitEnergies(self, state) -> np.ndarray:
        return self.energy_profiles(np.asarray(state).reshape(1, -1))[0][0]

    def StateIsStable(self, state) -> bool:
        return bool(self.energy_profiles(np.asarray(state).reshape(1, -1), False)[2][0])

    def AllStatesAreStable(self, states) -> bool:
        return bool(self.energy_profiles(states, False)[2].all())

    def StateEnergy(self, state) -> float:
        return float(self.energy_profiles(np.asarray(state).reshape(1, -1), False)[3][0])

    def UnitEnergy(self, state, unit_index: int) -> float:
        """HopfieldNetwork.go:185-187: the reference's scalar loop (Binary: -1 per term), hn_unit_energy."""
        p = self._packed(np.asarray(state).reshape(1, -1))
        out = np.empty(1)
        check(capi.load().hn_unit_energy(self._h, ptr(p), 1, int(unit_index), ptr(out)))
        return float(out[0])

    # ---- learning
    def hebbian_epoch(self, states) -> None:
        p = self._packed(states)
        check(capi.load().hn_hebbian_epoch(self._h, ptr(p), p.shape[0]))

    def delta_epoch(self, states, noised, perms) -> np.ndarray:
        p, q = self._packed(states), self._packed(noised)
        pm = np.ascontiguousarray(perms, dtype=np.uint16)
        rel = np.empty_like(p)
        check(capi.load().hn_delta_epoch(self._h, ptr(p), ptr(q), ptr(pm), p.shape[0], ptr(rel)))
        return rel

    def thermal_delta_epoch(self, states, noised, perms) -> np.ndarray:
        p, q = self._packed(states), self._packed(noised)
        pm = np.ascontiguousarray(perms, dtype=np.uint16)
        rel = np.empty_like(p)
        check(capi.load().hn_thermal_delta_epoch(self._h, ptr(p), ptr(q), ptr(pm), p.shape[0], ptr(rel)))
        return rel

    def rule_epoch(self, rule: int, states, noised=None, perms=None) -> Optional[np.ndarray]:
        """One application of any learning rule with explicit noised copies / update orders (hn_rule_epoch)."""
        p = self._packed(states)
        q = None if noised is None else self._packed(noised)
        pm = None if perms is None else np.ascontiguousarray(perms, dtype=np.uint16)
        rel = np.empty_like(p) if q is not None else None
        check(capi.load().hn_rule_epoch(self._h, int(rule), ptr(p), ptr(q), ptr(pm), p.shape[0], ptr(rel)))
        return rel

    def GetNetworkSummary(self) -> dict:
        """HopfieldNetworkSummary (HopfieldNetwork.go:124-149): the fields main.go:262-279 writes to networkSummary.pq."""
        return {"Matrix": self.GetMatrix(), "Dimension": self.dimension, "ForceSymmetric": bool(self._cfg.force_symmetric),
                "ForceZeroDiagonal": bool(self._cfg.force_zero_diagonal), "Epochs": self.epochs,
                "MaximumRelaxationUnstableUnits": self._cfg.max_relax_unstable_units,
                "MaximumRelaxationIterations": self._cfg.max_relax_iterations,
                "LearningNoiseScale": self.noise_scale, "UnitsUpdatedPerStep": self.unitsUpdatedPerStep}

    def enforce_constraints(self) -> None:
        check(capi.load().hn_enforce_constraints(self._h))

    def normalize(self) -> float:
        nrm = C.c_double()
        check(capi.load().hn_normalize_frobenius(self._h, C.byref(nrm)))
        return nrm.value

    def LearnStates(self, states, want_energies: bool = False, cap: Optional[int] = None,
                    shard: Optional[tuple] = None) -> List[LearnStateData]:
        """HopfieldNetwork.go:254-262. `states`: float64 [P][N] (activated here) or packed uint64.
        shard=(begin, end): this replica learns targets [begin, end) of `states` and the attached communicator
        all-reduces dW and the all-stable flag (hn_learn_states_sharded); rows returned are this shard's."""
        p = self._packed(states)
        n = p.shape[0]
        if cap is None:
            cap = self.epochs * n * (max(1, n // 5) if self.method == IterativeBatchMethod else 1)
        ep = np.zeros(max(cap, 1), dtype=np.int32)
        ix = np.zeros(max(cap, 1), dtype=np.int32)
        st = np.zeros(max(cap, 1), dtype=np.uint8)
        en = np.zeros((max(cap, 1), self.dimension)) if want_energies else None
        rows = C.c_int32()
        if shard is None:
            check(capi.load().hn_learn_states(self._h, ptr(p), n, self.rule, self.method, self.epochs,
                                              self.noise_method, self.noise_scale, self.randomGenerator._h, cap,
                                              ptr(ep), ptr(ix), ptr(st), ptr(en), C.byref(rows)))
        else:
            check(capi.load().hn_learn_states_sharded(self._h, ptr(p), n, int(shard[0]), int(shard[1]), self.rule,
                                                      self.method, self.epochs, self.noise_method, self.noise_scale,
                                                      self.randomGenerator._h, cap, ptr(ep), ptr(ix), ptr(st), ptr(en),
                                                      C.byref(rows)))
        k = min(rows.value, cap)
        return [LearnStateData(int(ep[i]), int(ix[i]), en[i] if want_energies else None, bool(st[i]))
                for i in range(k)]

    # ---- relaxation
    def UpdateState(self, states, perms) -> np.ndarray:
        """HopfieldNetwork.go:280-291 for a batch: one sweep per state in the given fresh order."""
        p = self._packed(states).copy()
        pm = np.ascontiguousarray(perms, dtype=np.uint16).reshape(p.shape[0], self.dimension)
        check(capi.load().hn_update_states(self._h, ptr(p), p.shape[0], ptr(pm)))
        return p

    def relax_batch(self, probes, perm_pool, offsets=None, serial_stream=False, dedup=False, history=False,
                    want_energies=False, want_distances=True, reuse_pinned_buffers=False) -> RelaxationBatchResult:
        """hn_relax_batch with host buffers.  reuse_pinned_buffers=True returns views of page-locked buffers owned by
        this network; they are overwritten by the next such call (copy what must outlive it)."""
        p = self._packed(probes)
        b = p.shape[0]
        pool = np.ascontiguousarray(perm_pool, dtype=np.uint16).reshape(-1, self.dimension)
        nt = C.c_int32()
        check(capi.load().hn_num_targets(self._h, C.byref(nt)))
        w = words(self.dimension)
        buf = lambda name, shape, dtype: self._result_buffer(name, shape, dtype, reuse_pinned_buffers)
        finals = buf("finals", (b, w), np.uint64)
        steps = buf("steps", (b,), np.int32)
        stable = buf("stable", (b,), np.uint8)
        flips = buf("flips", (b,), np.int32)
        margin = buf("margin", (b,), np.int32)
        dist = buf("dist", (b, nt.value), np.int32) if (want_distances and nt.value) else None
        en = buf("energies", (b, self.dimension), np.float64) if want_energies else None
        aid = buf("aid", (b,), np.int32) if dedup else None
        uf = buf("uf", (max(b, 1),), np.int32) if dedup else None
        uh = buf("uh", (max(b, 1),), np.int32) if dedup else None
        hist = buf("history", (b, self._cfg.max_relax_iterations + 1, w), np.uint64) if history else None
        out = capi.HnRelaxOut()
        out.final_states, out.num_steps, out.stable = ptr(finals), ptr(steps), ptr(stable)
        out.distances, out.energies, out.flips, out.margin_hits = ptr(dist), ptr(en), ptr(flips), ptr(margin)
        out.attractor_id, out.unique_first, out.unique_hits = ptr(aid), ptr(uf), ptr(uh)
        out.unique_cap = b if dedup else 0
        out.history = ptr(hist)
        sh = buf("hash", (b, 2), np.uint64) if dedup else None
        out.state_hash = ptr(sh)
        flags = (capi.HN_RELAX_SERIAL_STREAM if serial_stream else 0) | (capi.HN_RELAX_DEDUP if dedup else 0) | \
                (capi.HN_RELAX_HISTORY if history else 0)
        off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int32)
        check(capi.load().hn_relax_batch(self._h, ptr(p), b, ptr(pool), pool.shape[0], ptr(off), flags, C.byref(out)))
        nu = out.n_unique
        return RelaxationBatchResult(finals, steps, stable, dist, en, flips, margin, aid,
                                     uf[:nu].copy() if dedup else None, uh[:nu].copy() if dedup else None, hist, sh,
                                     out.total_flips, out.total_sweeps, out.total_margin_hits, out.device_ms)

    def ConcurrentRelaxStates(self, states, numThreads: int = 1, **kw) -> RelaxationBatchResult:
        """HopfieldNetwork.go:461-490.  numThreads is accepted and ignored for compute.
        The unit orders are drawn from the network RNG as one relaxation goroutine would draw
        them (one persistent list shuffled once per sweep) and shared by every probe."""
        pool = self.randomGenerator.perm_pool(self.dimension, self._cfg.max_relax_iterations)
        return self.relax_batch(states, pool, **kw)

    # ---- device-resident bulk path
    def upload_perm_pool(self, pool) -> None:
        pool = np.ascontiguousarray(pool, dtype=np.uint16).reshape(-1, self.dimension)
        check(capi.load().hn_upload_perm_pool(self._h, ptr(pool), pool.shape[0]))

    def upload_probes(self, probes) -> None:
        p = self._packed(probes)
        check(capi.load().hn_upload_probes(self._h, ptr(p), p.shape[0]))

    def generate_probes(self, count: int, rand_min: float = -1.0, rand_max: float = 1.0, rng: Optional["Rand"] = None) -> None:
        """states.StateGenerator on the device (StateGenerator.go:44-74): `count` random probes straight into the
        resident probe buffer, drawn from `rng` (default: the network's generator), which is advanced accordingly."""
        g = rng if rng is not None else self.randomGenerator
        check(capi.load().hn_generate_probes(self._h, g._h, count, rand_min, rand_max))

    def download_probes(self, count: int) -> np.ndarray:
        out = np.zeros((count, words(self.dimension)), dtype=np.uint64)
        check(capi.load().hn_download_probes(self._h, ptr(out), count))
        return out

    def relax_resident(self, dedup=False) -> float:
        ms = C.c_double()
        check(capi.load().hn_relax_resident(self._h, capi.HN_RELAX_DEDUP if dedup else 0, C.byref(ms)))
        return ms.value

    def last_learn_timing(self) -> capi.HnLearnTiming:
        t = capi.HnLearnTiming()
        check(capi.load().hn_last_learn_timing(self._h, C.byref(t)))
        return t

    def last_timing(self) -> capi.HnRelaxTiming:
        t = capi.HnRelaxTiming()
        check(capi.load().hn_last_relax_timing(self._h, C.byref(t)))
        return t


class HopfieldNetworkBuilder:
    """HopfieldNetworkBuilder.go:16-279 (same setters, same Build checks and messages)."""

    def __init__(self):  # NewHopfieldNetworkBuilder :40-55
        self.randMatrixInit = False
        self.dimension = 0
        self.domain = BipolarDomain
        self.forceSymmetric = True
        self.forceZeroDiagonal = True
        self.learningMethod = None
        self.learningRule = None
        self.epochs = 0
        self.maximumRelaxationUnstableUnits = 0
        self.maximumRelaxationIterations = 100
        self.learningRate = 1.0
        self.learningNoiseMethod = NoiseNone
        self.learningNoiseScale = 0.0
        self.unitsUpdatedPerStep = 1
        self.allowIntensiveDataCollection = False
        self.seed = None     # additive: SetSeed
        self.device = 0      # additive: SetDevice
        self.columnStream = capi.HN_COLUMNS_AUTO  # additive: SetColumnStream
        self.normAlgorithm = 1  # additive: SetNormAlgorithm (hn_set_norm_algorithm: 1 LAPACK-3.10 Dlassq, the default; 0 classic)

    def SetRandMatrixInit(self, f): self.randMatrixInit = f; return self
    def SetNetworkDimension(self, d): self.dimension = d; return self
    def SetNetworkDomain(self, d): self.domain = d; return self
    def SetForceSymmetric(self, f): self.forceSymmetric = f; return self
    def SetForceZeroDiagonal(self, f): self.forceZeroDiagonal = f; return self
    def SetNetworkLearningMethod(self, m): self.learningMethod = m; return self
    def SetNetworkLearningRule(self, r): self.learningRule = r; return self
    def SetEpochs(self, e): self.epochs = e; return self
    def SetMaximumRelaxationUnstableUnits(self, u): self.maximumRelaxationUnstableUnits = u; return self
    def SetMaximumRelaxationIterations(self, i): self.maximumRelaxationIterations = i; return self
    def SetLearningRate(self, r): self.learningRate = r; return self
    def SetLearningNoiseMethod(self, m): self.learningNoiseMethod = m; return self
    def SetLearningNoiseRatio(self, r): self.learningNoiseScale = r; return self
    def SetUnitsUpdatedPerStep(self, u): self.unitsUpdatedPerStep = u; return self
    def SetDataCollector(self, c): return self
    def SetLogger(self, l): return self
    def SetAllowIntensiveDataCollection(self, f): self.allowIntensiveDataCollection = f; return self
    def SetSeed(self, s): self.seed = s; return self
    def SetDevice(self, d): self.device = d; return self
    def SetColumnStream(self, m): self.columnStream = m; return self
    def SetNormAlgorithm(self, a): self.normAlgorithm = a; return self

    def Build(self) -> HopfieldNetwork:
        # the four panics of HopfieldNetworkBuilder.go:215-229, as Python exceptions with the same text
        if self.dimension <= 0:
            raise ValueError("HopfieldNetworkBuilder encountered an error during build! Dimension must be explicitly set to a positive integer!")
        if self.epochs <= 0:
            raise ValueError("HopfieldNetworkBuilder encountered an error during build! Epochs must be a positive integer!")
        if self.learningNoiseScale < 0.0 or self.learningNoiseScale > 1.0:
            raise ValueError("HopfieldNetworkBuilder encountered an error during build! learningNoiseRatio must be in range [0.0, 1.0]!")
        if self.unitsUpdatedPerStep < 0 or self.unitsUpdatedPerStep > self.dimension:
            raise ValueError("HopfieldNetworkBuilder encountered an error during build! unitsUpdatedPerStep must be a positive integer that is smaller than the network dimension!")
        cfg = capi.HnConfig(self.dimension, self.domain, int(self.forceSymmetric), int(self.forceZeroDiagonal),
                            self.maximumRelaxationUnstableUnits, self.maximumRelaxationIterations,
                            self.learningRate, self.device, self.columnStream)
        import time
        seed = self.seed if self.seed is not None else time.time_ns()  # :231 time-seeded unless SetSeed
        net = HopfieldNetwork(cfg, self.epochs, self.learningRule if self.learningRule is not None else 0,
                              self.learningMethod if self.learningMethod is not None else 0,
                              self.learningNoiseMethod, self.learningNoiseScale, self.unitsUpdatedPerStep, seed,
                              self.allowIntensiveDataCollection)
        if self.normAlgorithm != 1:
            check(capi.load().hn_set_norm_algorithm(net._h, int(self.normAlgorithm)))
        if self.randMatrixInit:
            # :235-246 distuv.Normal{0, 0.01, randSrc}.Rand() for every element, row-major
            w = np.empty(self.dimension * self.dimension)
            rg = net.randomGenerator
            for i in range(w.size):
                w[i] = rg.NormFloat64() * 0.01 + 0.0
            net.SetMatrix(w.reshape(self.dimension, self.dimension))
        return net


def NewHopfieldNetworkBuilder() -> HopfieldNetworkBuilder:
    return HopfieldNetworkBuilder()
