"""Rows of a device-resident x-vector table that are gathered on first use.

The reference's loop is  X1, X2 = load_xvec_trials_from_numbatch(...);  S = model(X1, X2)
(xvector_NeuralPlda_pytorch.py:38-39, :62-65): two calls, with the materialised [B, 512] pair as the interface between
them.  On the GPU that interface costs 4 KB per trial written by the gather and read back by the score kernel, three
times the traffic of scoring straight from the table.  Under `torch.no_grad()` (validation and scoring loops) the
loaders therefore return `GatheredRows`: tensors in every respect (shape, dtype, device, any torch op) whose storage is
only produced when something other than this package's `forward` touches them.  `NeuralPlda.forward` / `DPlda.forward`
recognise a pair of them and score from (table, row indices) directly -- the embed-once row table, or the pre-split
tensor-core kernel -- so the gather never runs.  With gradients enabled (the training loop) the loaders return ordinary
tensors.
"""
from __future__ import annotations

import torch
import torch.utils._pytree as pytree


class PairGather:
    """Shared state of the two sides of one loader call: the table, both row-index tensors, and -- once anything asked
    for the data -- the materialised pair (one nplda_gather_pairs launch for both sides)."""

    def __init__(self, tab, r1, r2, gather):
        self.tab, self.r1, self.r2, self._gather = tab, r1, r2, gather
        self._pair = None

    def dense(self):
        if self._pair is None:
            self._pair = self._gather(self.tab, self.r1, self.r2)
        return self._pair


class GatheredRows(torch.Tensor):
    @staticmethod
    def __new__(cls, state, side):
        rows = state.r2 if side else state.r1
        t = state.tab.table
        r = torch.Tensor._make_wrapper_subclass(cls, (rows.numel(), t.shape[1]), dtype=torch.float32, device=t.device,
                                                requires_grad=False)
        r._state, r._side = state, side
        return r

    def materialize(self):
        """The ordinary [B, D] fp32 CUDA tensor with these rows."""
        return self._state.dense()[self._side]

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        def un(t):
            return t.materialize() if isinstance(t, GatheredRows) else t
        return func(*pytree.tree_map(un, args), **pytree.tree_map(un, kwargs or {}))

    def __reduce_ex__(self, proto):
        return self.materialize().__reduce_ex__(proto)


def dense(t):
    return t.materialize() if isinstance(t, GatheredRows) else t


def lazy_pair(x1, x2):
    """(table object, rows1, rows2) when x1 / x2 are the two untouched sides of ONE loader call, else None."""
    if isinstance(x1, GatheredRows) and isinstance(x2, GatheredRows) and x1._state is x2._state \
            and x1._side == 0 and x2._side == 1 and x1._state._pair is None:
        s = x1._state
        return s.tab, s.r1, s.r2
    return None
