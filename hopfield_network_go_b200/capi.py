"""ctypes binding of include/hopfield_b200.h — the same C ABI a cgo binding uses.

There is no CPU fallback: if libhopfield_b200.so is missing this module raises,
and hn_create fails when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

HN_OK = 0
HN_E_INVALID, HN_E_NO_DEVICE, HN_E_CUDA, HN_E_OOM, HN_E_POOL, HN_E_NCCL, HN_E_STATE = -1, -2, -3, -4, -5, -6, -7
HN_RELAX_SERIAL_STREAM, HN_RELAX_DEDUP, HN_RELAX_HISTORY = 0x1, 0x2, 0x4
HN_COLUMNS_F64, HN_COLUMNS_F32, HN_COLUMNS_I16, HN_COLUMNS_AUTO = 0, 1, 2, 3

EXPORTED_SYMBOLS = [
    "hn_default_config", "hn_abi_version", "hn_last_error", "hn_create", "hn_destroy", "hn_set_weights",
    "hn_get_weights", "hn_set_integer_structure", "hn_set_norm_algorithm", "hn_integer_stream_active", "hn_set_targets", "hn_num_targets", "hn_get_targets", "hn_hebbian_epoch",
    "hn_delta_epoch", "hn_thermal_delta_epoch", "hn_rule_epoch", "hn_enforce_constraints", "hn_normalize_frobenius", "hn_energy_profiles",
    "hn_learn_states", "hn_learn_states_sharded", "hn_unit_energy", "hn_relax_batch", "hn_update_states", "hn_upload_perm_pool", "hn_upload_probes",
    "hn_generate_probes", "hn_download_probes", "hn_relax_resident", "hn_download_results", "hn_last_relax_timing", "hn_last_learn_timing", "hn_get_margin", "hn_unique_table_stats", "hn_debug_integer_fields",
    "hn_merge_unique", "hn_export_unique_device", "hn_merge_unique_device", "hn_host_alloc", "hn_host_free", "hn_pack_states_f64", "hn_unpack_states_f64", "hn_comm_unique_id", "hn_comm_attach", "hn_comm_detach",
    "hn_broadcast_weights", "hn_rng_create", "hn_rng_destroy", "hn_rng_uint64", "hn_rng_uint64n",
    "hn_rng_float64", "hn_rng_norm_float64", "hn_rng_advance", "hn_rng_shuffle_u16", "hn_rng_perm_pool", "hn_rng_apply_noise", "hn_rng_generate_states",
]


class HnConfig(C.Structure):
    _fields_ = [("dimension", C.c_int32), ("domain", C.c_int32), ("force_symmetric", C.c_int32),
                ("force_zero_diagonal", C.c_int32), ("max_relax_unstable_units", C.c_int32),
                ("max_relax_iterations", C.c_int32), ("learning_rate", C.c_double), ("device", C.c_int32),
                ("column_stream", C.c_int32)]


class HnRelaxOut(C.Structure):
    _fields_ = [("final_states", C.c_void_p), ("num_steps", C.c_void_p), ("stable", C.c_void_p),
                ("distances", C.c_void_p), ("energies", C.c_void_p), ("flips", C.c_void_p),
                ("margin_hits", C.c_void_p), ("attractor_id", C.c_void_p), ("unique_first", C.c_void_p),
                ("unique_hits", C.c_void_p), ("unique_cap", C.c_int32), ("n_unique", C.c_int32),
                ("history", C.c_void_p), ("state_hash", C.c_void_p), ("total_flips", C.c_int64), ("total_sweeps", C.c_int64),
                ("total_margin_hits", C.c_int64), ("device_ms", C.c_double)]


class HnLearnTiming(C.Structure):
    _fields_ = [("sweep_fields_ms", C.c_double), ("sweep_ms", C.c_double), ("dw_gemm_ms", C.c_double),
                ("allreduce_ms", C.c_double), ("update_ms", C.c_double), ("energy_gemm_ms", C.c_double),
                ("energy_epilogue_ms", C.c_double), ("targets", C.c_int32), ("reserved", C.c_int32),
                ("sweep_flips", C.c_int64)]


class HnRelaxTiming(C.Structure):
    _fields_ = [("fields_ms", C.c_double), ("relax_ms", C.c_double), ("post_ms", C.c_double),
                ("launches", C.c_int32), ("dense_sweeps", C.c_int32), ("dense_ms", C.c_double),
                ("dense_kernel_ms", C.c_double), ("dense_gemm_ms", C.c_double), ("dense_scan_ms", C.c_double),
                ("dense_flips", C.c_int64)]


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load libhopfield_b200.so (must have been built: __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                           "There is no CPU fallback.")
    L = C.CDLL(path)
    L.hn_last_error.restype = C.c_char_p
    L.hn_rng_uint64.restype = C.c_uint64
    L.hn_rng_uint64n.restype = C.c_uint64
    L.hn_rng_uint64n.argtypes = [C.c_void_p, C.c_uint64]
    L.hn_rng_uint64.argtypes = [C.c_void_p]
    L.hn_rng_float64.restype = C.c_double
    L.hn_rng_float64.argtypes = [C.c_void_p]
    L.hn_rng_norm_float64.restype = C.c_double
    L.hn_rng_norm_float64.argtypes = [C.c_void_p]
    L.hn_rng_create.argtypes = [C.c_uint64, C.c_void_p]
    L.hn_rng_destroy.argtypes = [C.c_void_p]
    L.hn_rng_shuffle_u16.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hn_rng_perm_pool.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.hn_rng_apply_noise.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int32]
    L.hn_rng_generate_states.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p]
    L.hn_create.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_destroy.argtypes = [C.c_void_p]
    L.hn_set_weights.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_get_weights.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_set_integer_structure.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_set_norm_algorithm.argtypes = [C.c_void_p, C.c_int32]
    L.hn_integer_stream_active.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_set_targets.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hn_num_targets.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_get_targets.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_hebbian_epoch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hn_delta_epoch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.hn_thermal_delta_epoch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.hn_rule_epoch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.hn_enforce_constraints.argtypes = [C.c_void_p]
    L.hn_normalize_frobenius.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_energy_profiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hn_learn_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    L.hn_learn_states_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
    L.hn_unit_energy.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.hn_relax_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32,
                                 C.c_void_p]
    L.hn_update_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.hn_upload_perm_pool.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hn_upload_probes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hn_rng_advance.argtypes = [C.c_void_p, C.c_uint64]
    L.hn_generate_probes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double]
    L.hn_download_probes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hn_relax_resident.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.hn_download_results.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_last_relax_timing.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_last_learn_timing.argtypes = [C.c_void_p, C.c_void_p]
    L.hn_get_margin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hn_debug_integer_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.hn_unique_table_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hn_merge_unique.argtypes = [C.c_int64] + [C.c_void_p] * 7
    L.hn_export_unique_device.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
    L.hn_merge_unique_device.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_int32] + [C.c_void_p] * 4
    L.hn_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.hn_host_free.argtypes = [C.c_void_p]
    L.hn_pack_states_f64.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.hn_unpack_states_f64.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.hn_comm_unique_id.argtypes = [C.c_void_p]
    L.hn_comm_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    L.hn_comm_detach.argtypes = [C.c_void_p]
    L.hn_broadcast_weights.argtypes = [C.c_void_p, C.c_int32]
    _lib = L
    return L


class HopfieldError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def check(rc: int) -> None:
    if rc != HN_OK:
        raise HopfieldError(rc, load().hn_last_error().decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def words(n: int) -> int:
    return (n + 63) // 64


def pack_states(states_f64: np.ndarray) -> np.ndarray:
    """float64 [count][N] -> packed uint64 [count][words] with the activation function applied."""
    s = np.ascontiguousarray(states_f64, dtype=np.float64)
    s = s.reshape(-1, s.shape[-1])
    out = np.zeros((s.shape[0], words(s.shape[1])), dtype=np.uint64)
    check(load().hn_pack_states_f64(s.shape[1], s.shape[0], ptr(s), ptr(out)))
    return out


def unpack_states(packed: np.ndarray, n: int, domain: int = 0) -> np.ndarray:
    p = np.ascontiguousarray(packed, dtype=np.uint64).reshape(-1, words(n))
    out = np.empty((p.shape[0], n), dtype=np.float64)
    check(load().hn_unpack_states_f64(n, domain, p.shape[0], ptr(p), ptr(out)))
    return out
