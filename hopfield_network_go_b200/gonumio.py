"""gonumio.py — the on-disk formats of main.go:169,187-188,214 (matrix.bin, targetStates.bin, -targetStatesFile,
-probeStatesFile).  SURVEY 8f #2.

The reference calls github.com/hmcalister/gonum-matrix-io/pkg/gonumio (go.mod; source NOT in the reference tree):
SaveMatrix, SaveVectorCollection, LoadVectorCollection.  What is restated here is gonum's own binary encoding
(gonum.org/v1/gonum/mat io.go, v0.12.0: Dense/VecDense.MarshalBinaryTo), which such a wrapper streams:

    header, 40 bytes little endian:  uint32 version = 1 | byte form 'G' | byte packing 'F' | byte uplo 'A' |
                                     bool unit = 0 | int64 rows | int64 cols | int64 ku = 0 | int64 kl = 0
    payload: rows*cols float64, little endian, row major          (a VecDense is rows = n, cols = 1)

A vector collection is taken to be the vectors' encodings back to back until end of file.  UNVERIFIED against a Go
build (none can run here): the header layout is recalled from gonum's sources, the collection framing is the
simplest one consistent with LoadVectorCollection(path) -> []*mat.VecDense.  Host-side only; no device code."""
from __future__ import annotations

import struct

import numpy as np

_HEADER = struct.Struct("<IBBB?qqqq")
assert _HEADER.size == 40
_VERSION = 1


def _pack_header(rows: int, cols: int) -> bytes:
    return _HEADER.pack(_VERSION, ord("G"), ord("F"), ord("A"), False, rows, cols, 0, 0)


def _read_header(buf: bytes, off: int):
    if len(buf) - off < _HEADER.size:
        raise ValueError("gonumio: truncated header")
    version, form, packing, uplo, unit, rows, cols, ku, kl = _HEADER.unpack_from(buf, off)
    if version != _VERSION:
        raise ValueError(f"gonumio: unsupported storage version {version}")
    if (form, packing, uplo) != (ord("G"), ord("F"), ord("A")) or unit or ku or kl:
        raise ValueError("gonumio: not a general dense matrix / vector")
    if rows < 0 or cols < 0:
        raise ValueError("gonumio: negative dimension")
    return rows, cols, off + _HEADER.size


def marshal_dense(m: np.ndarray) -> bytes:
    m = np.ascontiguousarray(m, dtype="<f8")
    if m.ndim != 2:
        raise ValueError("marshal_dense: 2-D array expected")
    return _pack_header(m.shape[0], m.shape[1]) + m.tobytes()


def marshal_vec(v: np.ndarray) -> bytes:
    v = np.ascontiguousarray(v, dtype="<f8").reshape(-1)
    return _pack_header(v.size, 1) + v.tobytes()


def SaveMatrix(matrix: np.ndarray, path: str) -> None:
    """gonumio.SaveMatrix(network.GetMatrix(), "matrix.bin") — main.go:187."""
    with open(path, "wb") as f:
        f.write(marshal_dense(matrix))


def LoadMatrix(path: str) -> np.ndarray:
    buf = open(path, "rb").read()
    rows, cols, off = _read_header(buf, 0)
    need = rows * cols * 8
    if len(buf) - off != need:
        raise ValueError("gonumio: payload size does not match the header")
    return np.frombuffer(buf, dtype="<f8", count=rows * cols, offset=off).reshape(rows, cols).copy()


def SaveVectorCollection(vectors: np.ndarray, path: str) -> None:
    """gonumio.SaveVectorCollection(targetStates, "targetStates.bin") — main.go:188.  `vectors`: [count][n] float64."""
    vectors = np.ascontiguousarray(vectors, dtype="<f8")
    if vectors.ndim != 2:
        raise ValueError("SaveVectorCollection: [count][n] array expected")
    with open(path, "wb") as f:
        for v in vectors:
            f.write(marshal_vec(v))


def LoadVectorCollection(path: str) -> np.ndarray:
    """gonumio.LoadVectorCollection(path) — main.go:169,214.  Returns [count][n] float64 (all vectors one length)."""
    buf = open(path, "rb").read()
    out, off = [], 0
    while off < len(buf):
        rows, cols, off = _read_header(buf, off)
        if cols != 1:
            raise ValueError("gonumio: a vector collection holds column vectors")
        if len(buf) - off < rows * 8:
            raise ValueError("gonumio: truncated vector")
        out.append(np.frombuffer(buf, dtype="<f8", count=rows, offset=off).copy())
        off += rows * 8
    if out and any(v.size != out[0].size for v in out):
        raise ValueError("gonumio: vectors of different lengths")
    return np.stack(out) if out else np.zeros((0, 0))
