"""Trial-list reader and score-file writer of libnplda.so (include/nplda.h, "Host-side text I/O").

The reference reads trial files with np.genfromtxt(dtype='str') and maps ids through dicts row by row
(sv_trials_loaders.py:376-383, 399-406; scorefile_generator.py:26, 45), and writes scores with
ndarray.astype(str) + np.savetxt (scorefile_generator.py:37-38, 54-55).  Same formats here, in one native pass.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib

MODE_ASIS, MODE_SPLITEXT, MODE_BASENAME_SPLITEXT = 0, 1, 2


def format_f32(v):
    """str(np.float32(v)) as the native writer prints it."""
    buf = ctypes.create_string_buffer(32)
    n = lib().nplda_format_f32(float(np.float32(v)), buf)
    return buf.raw[:n].decode()


class TrialFile:
    """A parsed trial list (every row the same number of whitespace-separated fields)."""

    def __init__(self, path):
        h = ctypes.c_void_p()
        rc = lib().nplda_trials_open(str(path).encode(), ctypes.byref(h))
        if rc == -6:
            raise ValueError(f"{path}: some rows have a different number of columns")      # np.genfromtxt's error class
        if rc == -5:
            raise OSError(f"{path} not found or unreadable")
        check(rc, "nplda_trials_open")
        self._h = h
        self.rows = int(lib().nplda_trials_rows(h))
        self.cols = int(lib().nplda_trials_cols(h))

    def close(self):
        if getattr(self, "_h", None):
            lib().nplda_trials_close(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def field(self, row, col):
        p = ctypes.c_void_p()
        n = lib().nplda_trials_field(self._h, row, col, ctypes.byref(p))
        if n < 0:
            raise IndexError((row, col))
        return ctypes.string_at(p, n).decode()

    def row(self, row):
        return [self.field(row, c) for c in range(self.cols)]

    def map_ids(self, col, ids, values=None, mode=MODE_ASIS, first_row=0):
        """int64 array over rows [first_row, rows): values[k] (or k) of the id the field equals, -1 if unknown."""
        ids = list(ids)
        blob = ("\n".join(ids) + "\n").encode() if ids else b""
        if blob.count(b"\n") != len(ids):
            raise ValueError("utterance ids must not contain newlines")
        vals = None if values is None else np.ascontiguousarray(values, dtype=np.int64)
        out = np.empty(max(0, self.rows - first_row), dtype=np.int64)
        check(lib().nplda_trials_map_ids(self._h, col, mode, blob, len(blob), len(ids),
                                         None if vals is None else vals.ctypes.data_as(ctypes.c_void_p), first_row,
                                         out.ctypes.data_as(ctypes.c_void_p)), "nplda_trials_map_ids")
        return out

    def col_float(self, col, first_row=0):
        n = max(0, self.rows - first_row)
        out, ok = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.uint8)
        check(lib().nplda_trials_col_float(self._h, col, first_row, out.ctypes.data_as(ctypes.c_void_p),
                                           ok.ctypes.data_as(ctypes.c_void_p)), "nplda_trials_col_float")
        return out, ok.astype(bool)

    def write_scores(self, path, scores, ncols_keep, header_line=None, first_row=0):
        s = np.ascontiguousarray(scores, dtype=np.float32)
        if s.shape != (max(0, self.rows - first_row),):
            raise ValueError("one score per trial row expected")
        rc = lib().nplda_scores_write(str(path).encode(), self._h, first_row, ncols_keep,
                                      s.ctypes.data_as(ctypes.c_void_p),
                                      None if header_line is None else header_line.encode())
        if rc == -5:
            raise OSError(f"cannot write {path}")
        check(rc, "nplda_scores_write")
