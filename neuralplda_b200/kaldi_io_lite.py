"""Readers for the three Kaldi model files the reference initialises from.

The reference shells out to Kaldi binaries (copy-matrix, copy-vector,
ivector-copy-plda: utils/models.py:441-448, Kaldi2NumpyUtils/kaldiPlda2numpydict.py:17)
and parses their text output.  Those binaries are not a dependency here: the
binary ('\\0B' + FM/DM/FV/DV) and text forms are parsed directly.
"""
from __future__ import annotations

import struct

import numpy as np

_DT = {"F": np.float32, "D": np.float64}


class _Cursor:
    def __init__(self, buf, pos=0):
        self.buf, self.pos = buf, pos

    def token(self):
        end = self.buf.index(b" ", self.pos)
        tok = self.buf[self.pos:end].decode()
        self.pos = end + 1
        return tok

    def int32(self):
        if self.buf[self.pos] != 4:
            raise ValueError("malformed Kaldi binary: expected int32 size marker")
        (v,) = struct.unpack_from("<i", self.buf, self.pos + 1)
        self.pos += 5
        return v

    def array(self, dt, count):
        a = np.frombuffer(self.buf, dtype=dt, count=count, offset=self.pos).astype(np.float64)
        self.pos += count * np.dtype(dt).itemsize
        return a

    def vector(self):
        tok = self.token()
        if len(tok) != 2 or tok[1] != "V" or tok[0] not in _DT:
            raise ValueError(f"expected a Kaldi vector, found token {tok!r}")
        return self.array(_DT[tok[0]], self.int32())

    def matrix(self):
        tok = self.token()
        if len(tok) != 2 or tok[1] != "M" or tok[0] not in _DT:
            raise ValueError(f"expected a Kaldi matrix, found token {tok!r} (compressed matrices unsupported)")
        r, c = self.int32(), self.int32()
        return self.array(_DT[tok[0]], r * c).reshape(r, c)


def _text_numbers(txt):
    return txt.replace("[", " ").replace("]", " ")


def read_vector(path):
    buf = open(path, "rb").read()
    if buf[:2] == b"\0B":
        return _Cursor(buf, 2).vector()
    return np.asarray(_text_numbers(buf.decode()).split(), dtype=np.float64)


def read_matrix(path):
    buf = open(path, "rb").read()
    if buf[:2] == b"\0B":
        return _Cursor(buf, 2).matrix()
    rows = [r.split() for r in _text_numbers(buf.decode()).strip().split("\n")]
    return np.asarray([r for r in rows if r], dtype=np.float64)


def read_plda(path):
    """Kaldi <Plda>: mean, diagonalising transform, psi; plus the diagonal P/Q of
    the two-covariance score derived as in kaldiPlda2numpydict.py:34-38."""
    buf = open(path, "rb").read()
    if buf[:2] == b"\0B":
        cur = _Cursor(buf, 2)
        if cur.token() != "<Plda>":
            raise ValueError("not a Kaldi Plda object")
        mean, transform, psi = cur.vector(), cur.matrix(), cur.vector()
    else:
        body = buf.decode().replace("<Plda>", "").replace("</Plda>", "")
        parts = [p for p in body.split("]") if p.strip()]
        mean = np.asarray(parts[0].replace("[", " ").split(), dtype=np.float64)
        rows = [r.split() for r in parts[1].replace("[", " ").strip().split("\n")]
        transform = np.asarray([r for r in rows if r], dtype=np.float64)
        psi = np.asarray(parts[2].replace("[", " ").split(), dtype=np.float64)
    tot = 1.0 + psi
    return {
        "plda_mean": mean,
        "diagonalizing_transform": transform,
        "Psi_across_covar_diag": psi,
        "diagP": psi / (tot * (tot - psi * psi / tot)),
        "diagQ": (1.0 / tot) - 1.0 / (tot - psi * psi / tot),
    }
