"""ctypes binding of libnplda.so (the C ABI declared in include/nplda.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C neuralplda_b200/csrc`` and loaded lazily on first use.  There is no
fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NPLDA_LIB") or os.path.join(_HERE, "libnplda.so")    # NPLDA_LIB: an experiment build (csrc/Makefile)
CSRC_DIR = os.path.join(_HERE, "csrc")

IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_TC_F8, IMPL_TC_BF16, IMPL_TC_PAIR, IMPL_TC_PAIR_F8 = 0, 1, 2, 3, 4, 5, 6
LOSS_SOFTCDET, LOSS_CROSSENTROPY = 0, 1
PACK_MIXED, PACK_EPOCH_ODD, PACK_PAIR, PREPARE_IF_CHANGED = 1, 2, 4, 1
ERR_UNSUPPORTED_DIM = -2
MAX_BETAS = 8

_lib = None
_lock = threading.Lock()

c_f32p = ctypes.c_void_p   # device pointers travel as integers
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); one entry per symbol declared in include/nplda.h
SIGNATURES = {
    "nplda_version": (c_int, []),
    "nplda_error_string": (ctypes.c_char_p, [c_int]),
    "nplda_launch_count": (c_i64, []),
    "nplda_pack_bytes": (c_i64, [c_int, c_int, c_int]),
    "nplda_pack_weights": (c_int, [c_vp] * 6 + [c_int] * 3 + [c_vp, c_i64, c_int, c_vp]),
    "dplda_pack_weights": (c_int, [c_vp] * 4 + [c_int] * 2 + [c_vp, c_i64, c_int, c_vp]),
    "nplda_score_fwd": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "dplda_score_fwd": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "nplda_score_fwd_indexed": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp,
                                        c_vp, c_int, c_vp]),
    "dplda_score_fwd_indexed": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp,
                                        c_int, c_vp]),
    "nplda_split_bytes": (c_i64, [c_i64, c_int]),
    "nplda_table_split": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "nplda_score_fwd_split": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "dplda_fwd_workspace_bytes": (c_i64, [c_i64, c_int, c_int]),
    "dplda_score_fwd_ws": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
    "nplda_gather_pairs": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "nplda_rowtab_bytes": (c_i64, [c_i64]),
    "nplda_table_prepare": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_int, c_vp, c_int, c_vp]),
    "nplda_score_pairs": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "nplda_trial_rows": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "nplda_trial_grid_gather": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "nplda_embed_fwd": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "nplda_score_from_embeddings": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "dplda_score_from_embeddings": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    "nplda_loss_accum": (c_int, [c_vp, c_vp, c_i64, c_vp, c_int, ctypes.c_float, c_vp, c_vp, c_vp]),
    "nplda_loss_finalize": (c_int, [c_vp, ctypes.POINTER(ctypes.c_double), c_int, c_vp, c_vp]),
    "nplda_loss_bwd": (c_int, [c_vp, c_vp, c_i64, c_vp, ctypes.POINTER(ctypes.c_double), c_int,
                               ctypes.c_float, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "nplda_bwd_workspace_bytes": (c_i64, [c_i64, c_int, c_int, c_int]),
    "nplda_score_bwd": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int] + [c_vp] * 6 + [c_vp] + [c_vp] * 8
                        + [c_vp, c_i64, c_vp]),
    "dplda_score_bwd": (c_int, [c_vp, c_vp, c_i64, c_int, c_int] + [c_vp] * 3 + [c_vp] + [c_vp] * 6
                        + [c_vp, c_i64, c_vp]),
    "nplda_score_bwd_act": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int] + [c_vp] * 6 + [c_vp] + [c_vp] * 8
                            + [c_vp, c_vp, c_i64, c_vp]),
    "dplda_score_bwd_act": (c_int, [c_vp, c_vp, c_i64, c_int, c_int] + [c_vp] * 3 + [c_vp] + [c_vp] * 6
                            + [c_vp, c_vp, c_i64, c_vp]),
    "nplda_act_floats": (c_i64, [c_i64, c_int]),
    "nplda_score_fwd_train": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "dplda_score_fwd_train": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "dplda_score_fwd_train_u": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "dplda_lr_bwd_workspace_bytes": (c_i64, [c_i64, c_int]),
    "dplda_lr_bwd": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "nplda_minc_sweep": (c_int, [c_vp, c_i64, c_vp, c_i64, ctypes.c_float, ctypes.c_float,
                                 ctypes.POINTER(ctypes.c_double), c_int, c_vp, c_vp, c_vp]),
    "nplda_split_by_label": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "nplda_sort_f32": (c_int, [c_vp, c_i64, c_vp]),
    "nplda_score_grid": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "nplda_score_grid_impl": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_int, c_vp]),
    "nplda_cohort_stats": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "nplda_score_norm": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "nplda_trials_open": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_vp)]),
    "nplda_trials_close": (None, [c_vp]),
    "nplda_trials_rows": (c_i64, [c_vp]),
    "nplda_trials_cols": (c_int, [c_vp]),
    "nplda_trials_field": (c_i64, [c_vp, c_i64, c_int, ctypes.POINTER(c_vp)]),
    "nplda_trials_map_ids": (c_int, [c_vp, c_int, c_int, ctypes.c_char_p, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "nplda_trials_col_float": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp]),
    "nplda_scores_write": (c_int, [ctypes.c_char_p, c_vp, c_i64, c_int, c_vp, ctypes.c_char_p]),
    "nplda_format_f32": (c_int, [ctypes.c_float, ctypes.c_char_p]),
    "nplda_debug_backward_paths": (None, [c_int, c_int, c_int]),
    "nplda_host_scratch_bytes": (c_i64, [c_i64, c_int]),
    "nplda_score_fwd_host": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_i64, c_vp, c_i64,
                                     c_int, c_int]),
}


def build(verbose=False):
    """Compile libnplda.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libnplda.so failed")
    return LIB_PATH


def lib():
    """The loaded library, with argtypes set.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU or PyTorch fallback for the scoring path)")
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().nplda_error_string(int(code)).decode()
        raise RuntimeError(f"libnplda {what} failed ({code}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """The current stream of the current device as a void* (the raw-handle query avoids building a Stream object
    per launch; the small-batch training loop of the reference is bound by such host costs)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def on_device(dev):
    """Context that makes `dev` the current CUDA device; free when it already is."""
    if dev.index is None or torch.cuda.current_device() == dev.index:
        return _NO_GUARD
    return torch.cuda.device(dev)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("neuralplda_b200 scores on the GPU only: got a tensor on "
                               f"'{t.device}' (there is no CPU fallback; move the module and inputs to cuda)")


def launch_count():
    return int(lib().nplda_launch_count())


def betas_array(betas):
    arr = (ctypes.c_double * max(1, len(betas)))(*[float(b) for b in betas])
    return arr
