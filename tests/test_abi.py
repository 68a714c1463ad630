"""The C-ABI library: loads, exports every symbol include/nplda.h declares, and rejects bad
arguments before touching a GPU (no compute calls here: this file runs without a device)."""
import ctypes
import os
import re

import pytest

from neuralplda_b200 import _lib
from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nplda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b((?:nplda|dplda)_[a-z0-9_]+)\s*\(", hdr)
    return sorted(set(names))


def test_library_builds_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    L = _lib.lib()
    assert L.nplda_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/nplda.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in neuralplda_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_error_strings_and_sizes():
    L = _lib.lib()
    assert L.nplda_error_string(0) == b"ok"
    assert b"unsupported" in L.nplda_error_string(-2)
    assert L.nplda_pack_bytes(512, 170, 170) > 512 * 192 * 4
    assert L.nplda_pack_bytes(512, 300, 170) == -2          # layer wider than the kernels support
    assert L.nplda_bwd_workspace_bytes(1000, 512, 170, 170) > 3 * 2 * 1024 * 192 * 4
    assert L.nplda_host_scratch_bytes(1024, 512) == 2 * (2 * 1024 * 512 * 4 + 4096)
    assert L.nplda_launch_count() == 0                       # nothing has been launched in this process


def test_bad_arguments_are_rejected_without_a_device():
    L = _lib.lib()
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)
    assert L.nplda_score_fwd(null, null, 5, 512, 170, 170, null, null, 0, null) == -1
    assert L.nplda_score_fwd(null, null, -1, 512, 170, 170, one, null, 0, null) == -1
    assert L.nplda_score_fwd(one, one, 5, 512, 300, 170, one, one, 0, null) == -2
    assert L.nplda_score_fwd(one, one, 5, 512, 170, 170, one, one, 7, null) == -1      # unknown impl
    assert L.nplda_score_fwd(null, null, 0, 512, 170, 170, one, null, 0, null) == 0     # n == 0 is a no-op
    assert L.nplda_loss_accum(null, null, 4, null, 2, 15.0, null, null, null) == -1
    assert L.nplda_loss_accum(null, null, 0, one, 2, 15.0, null, one, null) == 0
    assert L.nplda_loss_accum(one, one, 4, one, 9, 15.0, null, one, null) == -1           # K > NPLDA_MAX_BETAS
    assert L.nplda_minc_sweep(null, 0, null, 0, 1.0, 1.0, _lib.betas_array([99.0]), 1, null, null, null) == -1
    with pytest.raises(RuntimeError):
        _lib.check(-3, "x")


def test_header_constants_match_the_python_binding():
    """Every NPLDA_IMPL_* / NPLDA_PACK_* / NPLDA_LOSS_* constant of include/nplda.h has the same value in _lib.py (the
    ctypes layer passes them through as plain ints), and an impl id the header does not define is rejected."""
    hdr = open(os.path.join(ROOT, "include", "nplda.h")).read()
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+NPLDA_((?:IMPL|PACK|LOSS)_[A-Z0-9_]+)\s+(-?\d+)", hdr)}
    assert {"IMPL_AUTO", "IMPL_SIMT", "IMPL_TC", "IMPL_TC_F8", "IMPL_TC_BF16", "IMPL_TC_PAIR", "IMPL_TC_PAIR_F8",
            "PACK_MIXED", "PACK_EPOCH_ODD", "PACK_PAIR"} <= set(consts)
    for name, value in consts.items():
        assert getattr(_lib, name) == value, name
    impls = sorted(v for k, v in consts.items() if k.startswith("IMPL_"))
    assert impls == list(range(len(impls)))                       # dense ids: the dispatcher's range check covers them all
    L = _lib.lib()
    one = ctypes.c_void_p(16)
    assert L.nplda_score_fwd(one, one, 5, 512, 170, 170, one, one, len(impls), ctypes.c_void_p(None)) == -1
