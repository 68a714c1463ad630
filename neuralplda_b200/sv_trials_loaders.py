"""Trial batching with a device-resident x-vector table.

Same entry points, argument meaning and batch layout as the reference's
utils/sv_trials_loaders.py:371-437: trial TSVs become DataLoaders of
(int64 idx1[B], int64 idx2[B], float32 label[B]); per batch the two sides are
returned as [B, D] fp32 tensors on `device`.  The reference does the gather
with a Python loop over B dict lookups plus a host->device copy of 4 KB per
trial every batch; here the whole x-vector dict is uploaded once and the gather
is an index_select on the GPU (or is fused into the score kernel through
NeuralPlda.forward_indexed).
"""
from __future__ import annotations

import ctypes
import os
import weakref

import numpy as np
import torch
from torch.utils.data import TensorDataset, DataLoader, ConcatDataset, Subset

from ._lib import check, lib, on_device, ptr, stream_ptr
from .lazy import GatheredRows, PairGather

# Under torch.no_grad() the loaders return lazily gathered rows (lazy.py): forward() scores them straight from the table.
LAZY_GATHER = True


class XvectorTable:
    """utt_id -> row of a [n_utts, D] fp32 device tensor; rows follow list(mega_dict)
    (the reference's num_to_id order, xvector_NeuralPlda_pytorch.py:118-119)."""

    def __init__(self, mega_dict, device):
        self.ids = list(mega_dict)
        self.row_of = {u: i for i, u in enumerate(self.ids)}
        host = torch.from_numpy(np.asarray([mega_dict[u] for u in self.ids])).float()
        self.device = torch.device(device)
        self.table = host.to(self.device)

    def rows_for_ids(self, ids):
        return torch.tensor([self.row_of[u] for u in ids], dtype=torch.int64)


# (mega_dict object, {device: XvectorTable}), most recently used last.  Plain dicts cannot be weakly referenced, so an
# entry holds its dict STRONGLY: while an entry exists its dict's id cannot be recycled by another dict (the failure
# mode of a cache keyed on id()).  The cache is small (`MAX_TABLES` dicts; the reference uses one) and
# `release_tables()` empties it.  A cached table is re-uploaded when the dict's length, its first / last key or the
# array OBJECTS stored under them changed; vectors overwritten in place inside the same numpy arrays are not
# detectable without reading every vector -- call `refresh_table(mega_dict)` after doing that.
_tables = []
MAX_TABLES = 4


def _signature(mega_dict):
    n = len(mega_dict)
    if n == 0:
        return (0,)
    first = next(iter(mega_dict))
    last = next(reversed(mega_dict))
    return (n, first, last, id(mega_dict[first]), id(mega_dict[last]))


def get_table(mega_dict, device):
    """One upload per (dict object, device); re-uploaded when the dict's signature changed."""
    device = torch.device(device)
    sig = _signature(mega_dict)
    ent = None
    for i, e in enumerate(_tables):
        if e[0] is mega_dict:
            ent = _tables.pop(i)
            break
    if ent is None or ent[1] != sig:
        ent = (mega_dict, sig, {})
    _tables.append(ent)
    del _tables[:-MAX_TABLES]
    tab = ent[2].get(str(device))
    if tab is None:
        tab = XvectorTable(mega_dict, device)
        ent[2][str(device)] = tab
    return tab


def refresh_table(mega_dict):
    """Forget the device copies of this dict (after its vectors were modified in place)."""
    _tables[:] = [e for e in _tables if e[0] is not mega_dict]


def release_tables():
    """Drop every cached device table (and the references to their dicts)."""
    del _tables[:]


def check_pending_errors():
    """Synchronise and raise the KeyError of any gather that met a row outside its table.  The per-batch path
    reports such a row at the NEXT loader call (no host round trip per batch); call this after the last batch of a
    loop -- `GraphedTrainStep.flush()` and `minc()` do -- so that the final batch cannot slip through.  Until it is
    reported the affected rows are NaN, so nothing computed from them can pass for a result."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    for e in _tables:
        for tab in e[2].values():
            flag = tab.__dict__.get("_bad_flag")
            if flag is not None:
                flag.check()


def _rows_from_nums(tab, num_to_id_dict, data):
    nums = data.detach().to("cpu", torch.int64)
    # fast path: num_to_id_dict enumerates list(mega_dict) as in the reference drivers
    n = len(tab.ids)
    ident = getattr(tab, "_ident", None)
    if ident is None or ident[0] is not num_to_id_dict:
        ok = len(num_to_id_dict) == n and all(num_to_id_dict.get(i) == u for i, u in enumerate(tab.ids))
        remap = None if ok else {k: tab.row_of[v] for k, v in num_to_id_dict.items() if v in tab.row_of}
        tab._ident = (num_to_id_dict, ok, remap)
        ident = tab._ident
    if ident[1]:
        if nums.numel() and (int(nums.min()) < 0 or int(nums.max()) >= n):
            raise KeyError(int(nums.max()))
        return nums
    return torch.tensor([ident[2][int(i)] for i in nums], dtype=torch.int64)


class _BadIndexFlag:
    """int32 flag in pinned host memory that the gather kernel raises for rows outside the table.  The host reads
    it without a synchronising copy: a raised flag is seen at the next loader call after the kernel ran (the
    reference's loop reads loss.item() every step, so at the next batch at the latest) or by check_pending_errors();
    the rows themselves come back as NaN."""

    def __init__(self):
        self.t = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.ptr = ctypes.c_void_p(self.t.data_ptr())

    def check(self):
        if int(self.t[0]) != 0:
            self.t[0] = 0
            raise KeyError("a trial of an earlier batch referred to a row outside the x-vector table "
                           "(the reference raises KeyError from its dict lookup)")


def bad_flag_of(tab):
    flag = tab.__dict__.get("_bad_flag")
    if flag is None:
        flag = tab._bad_flag = _BadIndexFlag()
    return flag


def _gather_now(tab, r1, r2):
    """(X1, X2) = (table[r1], table[r2]) on the device, one launch (nplda_gather_pairs)."""
    flag = bad_flag_of(tab)
    n, d = r1.numel(), tab.table.shape[1]
    x1 = torch.empty(n, d, dtype=torch.float32, device=tab.device)
    x2 = torch.empty(n, d, dtype=torch.float32, device=tab.device)
    with on_device(tab.table.device):
        check(lib().nplda_gather_pairs(ptr(tab.table), tab.table.shape[0], d, ptr(r1), ptr(r2), n, ptr(x1), ptr(x2),
                                       flag.ptr, stream_ptr()), "nplda_gather_pairs")
    return x1, x2


def _gather(tab, r1, r2, lazy=None):
    """The loaders' return value: the materialised pair, or -- under no_grad -- its lazily gathered stand-in."""
    bad_flag_of(tab).check()
    if lazy is None:
        lazy = LAZY_GATHER and not torch.is_grad_enabled()
    if lazy and r1.numel() > 0:
        state = PairGather(tab, r1, r2, _gather_now)
        return GatheredRows(state, 0), GatheredRows(state, 1)
    return _gather_now(tab, r1, r2)


def _device_rows(tab, num_to_id_dict, data, device):
    """Row indices on the device.  Index tensors that already are on the GPU (the reference's loop moves the batch
    there first, xvector_NeuralPlda_pytorch.py:37) stay there when num_to_id_dict enumerates list(mega_dict) --
    no device->host round trip, no synchronisation; out-of-range rows are caught by the gather kernel."""
    if data.is_cuda and data.dtype == torch.int64 and data.dim() == 1:
        ident = getattr(tab, "_ident", None)
        if ident is not None and ident[0] is num_to_id_dict and ident[1]:
            return data.to(device).contiguous()
    return _rows_from_nums(tab, num_to_id_dict, data).to(device, non_blocking=True)


def load_xvec_trials_from_numbatch(mega_dict, num_to_id_dict, data1, data2, device):
    """sv_trials_loaders.py:418-426: (X1, X2) [B, D] fp32 on `device`."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("neuralplda_b200 keeps x-vectors on the GPU; got device '%s'" % device)
    tab = get_table(mega_dict, device)
    r1 = _device_rows(tab, num_to_id_dict, data1, tab.table.device)
    r2 = _device_rows(tab, num_to_id_dict, data2, tab.table.device)
    if r1.shape != r2.shape:
        raise RuntimeError("data1 and data2 must have the same length")
    return _gather(tab, r1, r2)


def strip_id(s):
    return os.path.splitext(os.path.basename(s))[0]


def load_xvec_trials_from_idbatch(mega_dict, trials, device):
    """sv_trials_loaders.py:429-437: ids are basename/splitext-stripped before lookup."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("neuralplda_b200 keeps x-vectors on the GPU; got device '%s'" % device)
    tab = get_table(mega_dict, device)
    trials = np.asarray(trials)
    if trials.size == 0:
        empty = tab.table.new_zeros((0, tab.table.shape[1]))
        return empty, empty.clone()
    r1 = tab.rows_for_ids([strip_id(d) for d in trials[:, 0]]).to(tab.table.device, non_blocking=True)
    r2 = tab.rows_for_ids([strip_id(d) for d in trials[:, 1]]).to(tab.table.device, non_blocking=True)
    return _gather(tab, r1, r2)


def _read_trials(f, id_to_num_dict, strip_ext_col2):
    """Rows with unknown ids or unparsable labels are silently dropped, as the
    reference's try/except does (sv_trials_loaders.py:379-383, 402-406).  Parsing and
    the id lookups run in the native reader (textio.TrialFile), one pass over the file."""
    from .textio import TrialFile, MODE_ASIS, MODE_SPLITEXT
    keys = list(id_to_num_dict)
    if not all(isinstance(k, str) for k in keys):
        raise TypeError("id_to_num_dict must map utterance id strings to row numbers")
    vals = np.asarray([id_to_num_dict[k] for k in keys], dtype=np.int64)
    with TrialFile(f) as tf:
        if tf.rows == 0 or tf.cols < 3:               # tr[2] raises in the reference: every row is dropped
            x1 = x2 = np.zeros(0, np.int64); lab = np.zeros(0, np.float32)
        else:
            p1 = tf.map_ids(0, keys, None, MODE_ASIS)
            p2 = tf.map_ids(1, keys, None, MODE_SPLITEXT if strip_ext_col2 else MODE_ASIS)
            lab, ok = tf.col_float(2)
            keep = (p1 >= 0) & (p2 >= 0) & ok
            x1, x2, lab = vals[p1[keep]], vals[p2[keep]], lab[keep]
    return TensorDataset(torch.from_numpy(np.ascontiguousarray(x1)), torch.from_numpy(np.ascontiguousarray(x2)),
                         torch.from_numpy(np.ascontiguousarray(lab, dtype=np.float32)))


def combine_trials_and_get_loader(trials_key_files_list, id_to_num_dict, subsample_factors=None, batch_size=2048, subset=0):
    """sv_trials_loaders.py:371-392."""
    if subsample_factors is None:
        subsample_factors = [1 for w in trials_key_files_list]
    datasets = []
    for f, sf in zip(trials_key_files_list, subsample_factors):
        tdset = _read_trials(f, id_to_num_dict, strip_ext_col2=False)
        inds = np.arange(len(tdset))[np.random.rand(len(tdset)) < sf]
        datasets.append(Subset(tdset, inds))
    combined_dataset = ConcatDataset(datasets)
    if subset > 0:
        inds = np.arange(len(combined_dataset))[np.random.rand(len(combined_dataset)) < subset]
        combined_dataset = Subset(combined_dataset, inds)
    return DataLoader(combined_dataset, batch_size=batch_size, shuffle=True)


def get_trials_loaders_dict(trials_key_files_list, id_to_num_dict, subsample_factors=None, batch_size=2048, subset=0):
    """sv_trials_loaders.py:394-415 (column 2 loses its extension; keyed by file basename)."""
    trials_loaders_dict = {}
    if subsample_factors is None:
        subsample_factors = [1 for w in trials_key_files_list]
    for f, sf in zip(trials_key_files_list, subsample_factors):
        tdset = _read_trials(f, id_to_num_dict, strip_ext_col2=True)
        inds = np.arange(len(tdset))[np.random.rand(len(tdset)) < sf]
        dataset = Subset(tdset, inds)
        if subset > 0:
            inds = np.arange(len(dataset))[np.random.rand(len(dataset)) < subset]
            dataset = Subset(dataset, inds)
        trials_loaders_dict[os.path.splitext(os.path.basename(f))[0]] = DataLoader(dataset, batch_size=batch_size, shuffle=True)
    return trials_loaders_dict
