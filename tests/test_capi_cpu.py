"""CPU: the C-ABI library loads, exports every symbol include/hopfield_b200.h declares, and fails loudly
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hopfield_network_go_b200 import capi
from hopfield_network_go_b200 import hopfieldnetwork as hn
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "hopfield_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hn_[a-z0-9_]+)\s*\(", header)) - {"hn_words"}  # hn_words is static inline
    assert len(declared) >= 35
    lib = capi.load()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert declared == set(capi.EXPORTED_SYMBOLS)
    assert lib.hn_abi_version() == 1


def test_struct_layouts_match_header(built):
    assert C.sizeof(capi.HnConfig) == 40
    assert C.sizeof(capi.HnRelaxTiming) == 72 and C.sizeof(capi.HnLearnTiming) == 72
    assert capi.HnRelaxOut.unique_cap.offset == 80 and capi.HnRelaxOut.state_hash.offset == 96
    assert C.sizeof(capi.HnRelaxOut) == 136
    cfg = capi.HnConfig()
    capi.load().hn_default_config(C.byref(cfg))
    assert (cfg.domain, cfg.force_symmetric, cfg.force_zero_diagonal, cfg.max_relax_unstable_units,
            cfg.max_relax_iterations, cfg.learning_rate) == (0, 1, 1, 0, 100, 1.0)   # HopfieldNetworkBuilder.go:40-55
    assert cfg.column_stream == capi.HN_COLUMNS_AUTO == 3


def test_no_cpu_fallback(built):
    if _have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.HopfieldError) as ei:
        hn.NewHopfieldNetworkBuilder().SetNetworkDimension(16).SetEpochs(1).Build()
    assert ei.value.code == capi.HN_E_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_builder_validation_messages():
    """The four Build panics of HopfieldNetworkBuilder.go:215-229, same text."""
    b = hn.NewHopfieldNetworkBuilder
    with pytest.raises(ValueError, match="Dimension must be explicitly set to a positive integer"):
        b().SetEpochs(1).Build()
    with pytest.raises(ValueError, match="Epochs must be a positive integer"):
        b().SetNetworkDimension(8).Build()
    with pytest.raises(ValueError, match=r"learningNoiseRatio must be in range \[0.0, 1.0\]"):
        b().SetNetworkDimension(8).SetEpochs(1).SetLearningNoiseRatio(1.5).Build()
    with pytest.raises(ValueError, match="unitsUpdatedPerStep must be a positive integer"):
        b().SetNetworkDimension(8).SetEpochs(1).SetUnitsUpdatedPerStep(9).Build()


def test_create_argument_errors(built):
    lib = capi.load()
    h = C.c_void_p()
    cfg = capi.HnConfig()
    lib.hn_default_config(C.byref(cfg))
    assert lib.hn_create(C.byref(cfg), C.byref(h)) == capi.HN_E_INVALID          # dimension 0
    assert b"Dimension must be explicitly set" in lib.hn_last_error()
    cfg.dimension = 70000
    assert lib.hn_create(C.byref(cfg), C.byref(h)) == capi.HN_E_INVALID
    assert lib.hn_create(None, C.byref(h)) == capi.HN_E_INVALID
    assert lib.hn_destroy(None) == capi.HN_OK


def test_pack_unpack_roundtrip_and_activation(built):
    rs = np.random.default_rng(0)
    for n in (1, 63, 64, 65, 100, 4096):
        x = rs.standard_normal((5, n))
        x[0, 0] = 0.0                                   # x <= 0 -> low value
        p = capi.pack_states(x)
        assert p.shape == (5, (n + 63) // 64)
        bip = capi.unpack_states(p, n, 0)
        assert np.array_equal(bip, np.where(x <= 0, -1.0, 1.0))
        assert np.array_equal(capi.unpack_states(p, n, 1), np.where(x <= 0, 0.0, 1.0))
        assert np.array_equal(capi.pack_states(bip), p)
        if n % 64:
            assert (p[:, -1] >> np.uint64(n % 64)).max() == 0   # padding bits are zero


def test_rng_advance_equals_drawing(built):
    """hn_rng_advance(k) leaves the state k Uint64 draws would leave (the jump-ahead behind hn_generate_probes)."""
    for seed, k in ((1, 0), (1, 1), (7, 2), (42, 1000), (5, 4096 * 3 + 17), (9, (1 << 33) + 12345)):
        a, b = hn.Rand(seed), hn.Rand(seed)
        if k <= 20000:
            for _ in range(k):
                a.Uint64()
            b.Advance(k)
        else:   # composition: advance(k) == advance(k1) then advance(k - k1)
            a.Advance(1 << 33)
            a.Advance(k - (1 << 33))
            b.Advance(k)
        assert [a.Uint64() for _ in range(4)] == [b.Uint64() for _ in range(4)]
    # against the oracle's independent restatement of the generator
    o, g = orc.Rng(77), hn.Rand(77)
    g.Advance(5000)
    for _ in range(5000):
        o.uint64()
    assert o.uint64() == g.Uint64()


def test_hot_kernels_do_not_spill(built):
    """The relaxation kernels are compiled against a register cap per occupancy target (__launch_bounds__ in relax.cuh:
    56 registers at 9 CTAs/SM for the shipped integer kernel).  A change that makes one of them spill compiles and passes
    every parity test but costs tens of percent (measured: 170 -> 252 ms per config-3 step for 148 bytes of spill
    stores), so the built library is checked: no stack frame in any relaxation / dense-sweep / contraction kernel."""
    import re
    import shutil
    import subprocess
    from hopfield_network_go_b200 import build
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([tool, "-res-usage", build.LIB_PATH], capture_output=True, text=True).stdout
    seen = 0
    for name, stack in re.findall(r"Function (\S+):\s*\n\s*REG:\d+ STACK:(\d+)", out):
        if any(k in name for k in ("relax_kernel", "dense_sweep_kernel", "igemm_tc05_kernel")):
            seen += 1
            assert int(stack) == 0, f"{name} spills ({stack} bytes of stack)"
    assert seen >= 20, "resource usage of the hot kernels not found in the library"


def test_tensor_path_sass(built):
    """The Blackwell tensor path is really there and is issued the cheap way: the dense-sweep and contraction kernels
    contain UTCIMMA (tcgen05.mma.kind::i8), UTMALDG (TMA), LDTM (tcgen05.ld) and UTCBAR (tcgen05.commit), and NO
    elect-broadcast-branch loop (BRA.U.ANY) around them — under a divergent `if (lane == 0)` the compiler wraps every
    such instruction in one, which cost 450 cycles per ring slot."""
    import re
    import shutil
    import subprocess
    from hopfield_network_go_b200 import build
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([tool, "-sass", build.LIB_PATH], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    seen = 0
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if "dense_sweep_kernel" in name or "igemm_tc05_kernel" in name:
            seen += 1
            for mnemonic in ("UTCIMMA", "UTMALDG", "LDTM", "UTCBAR"):
                assert re.search(r"\b" + mnemonic, f), f"{mnemonic} missing in {name}"
            assert "BRA.U.ANY" not in f, f"elect-broadcast loop around a tensor / TMA instruction in {name}"
    assert seen >= 4
