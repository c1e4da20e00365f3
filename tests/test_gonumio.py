"""-m "not gpu": gonum binary matrix / vector-collection encoding (hopfield_network_go_b200/gonumio.py).
The byte-level expectations are written out by hand from gonum's storage header (version 1, 'G','F','A', 40 bytes)."""
import struct

import numpy as np
import pytest

from hopfield_network_go_b200 import gonumio


def test_dense_bytes_known_answer(tmp_path):
    m = np.array([[1.0, -2.5], [0.0, 3.0], [4.0, 5.0]])
    raw = gonumio.marshal_dense(m)
    assert len(raw) == 40 + 6 * 8
    assert raw[:8] == bytes([1, 0, 0, 0, ord("G"), ord("F"), ord("A"), 0])
    assert struct.unpack_from("<qqqq", raw, 8) == (3, 2, 0, 0)
    assert struct.unpack_from("<6d", raw, 40) == (1.0, -2.5, 0.0, 3.0, 4.0, 5.0)          # row major
    p = tmp_path / "matrix.bin"
    gonumio.SaveMatrix(m, str(p))
    assert p.read_bytes() == raw
    assert np.array_equal(gonumio.LoadMatrix(str(p)), m)


def test_vector_collection_roundtrip_and_framing(tmp_path):
    rs = np.random.default_rng(3)
    v = np.where(rs.random((7, 100)) < 0.5, -1.0, 1.0)
    p = tmp_path / "targetStates.bin"
    gonumio.SaveVectorCollection(v, str(p))
    raw = p.read_bytes()
    assert len(raw) == 7 * (40 + 100 * 8)
    assert struct.unpack_from("<qq", raw, 8) == (100, 1)                                     # a VecDense is n x 1
    assert struct.unpack_from("<qq", raw, 840 + 8) == (100, 1)                               # second record right behind
    back = gonumio.LoadVectorCollection(str(p))
    assert back.shape == (7, 100) and np.array_equal(back, v)
    (tmp_path / "empty.bin").write_bytes(b"")
    assert gonumio.LoadVectorCollection(str(tmp_path / "empty.bin")).shape == (0, 0)


def test_rejects_foreign_and_truncated_files(tmp_path):
    good = gonumio.marshal_vec(np.arange(5.0))
    for bad in (good[:39], good[:-3], b"\x02" + good[1:], good[:4] + b"S" + good[5:]):
        p = tmp_path / "bad.bin"
        p.write_bytes(bad)
        with pytest.raises(ValueError):
            gonumio.LoadVectorCollection(str(p))
    p = tmp_path / "dense.bin"
    p.write_bytes(gonumio.marshal_dense(np.zeros((2, 2))))
    with pytest.raises(ValueError):
        gonumio.LoadVectorCollection(str(p))      # cols != 1
