"""Autograd bindings of the C-ABI kernels (include/nplda.h).

PyTorch is plumbing here: it owns the device buffers and the autograd tape;
every arithmetic step of the hot path runs in libnplda.so.
"""
from __future__ import annotations

import ctypes
import weakref

import torch

from . import _lib
from ._lib import check, lib, on_device, ptr, require_cuda, stream_ptr


def _f32c(t):
    """fp32 contiguous view/copy on the same device (the reference casts with .float())."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class PackedWeights:
    """Device workspace holding the packed weights of one module.

    The weights are re-packed on EVERY call (three small kernels): parameter changes cannot be detected on the
    host -- the reference itself writes parameters through `.data.copy_()` (models.py:449-457, :420) and fused
    optimisers update them in place, and neither bumps tensor._version.  What is cached beyond the pack (the
    per-utterance row table of the trial-list / grid paths) is validated on the device against the content
    fingerprint the pack kernel computes (csrc/pack.cu, csrc/pairs.cu)."""

    def __init__(self):
        self.buf = None
        self.epoch = 0

    def get(self, kind, params, d_in, d1, d2, mixed=False, pair=False):
        dev = params[0].device
        nbytes = lib().nplda_pack_bytes(d_in, d1, d2)
        if nbytes < 0:
            check(nbytes, "nplda_pack_bytes")
        if self.buf is None or self.buf.numel() != nbytes or self.buf.device != dev:
            self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self.rowtab_key = None
        self.epoch ^= 1
        flags = ((_lib.PACK_MIXED if mixed else 0) | (_lib.PACK_EPOCH_ODD if self.epoch else 0)
                 | (_lib.PACK_PAIR if pair else 0))
        ps = [_f32c(p.detach()) for p in params]
        with on_device(dev):
            if kind == "nplda":
                rc = lib().nplda_pack_weights(*[ptr(p) for p in ps], d_in, d1, d2, ptr(self.buf),
                                              self.buf.numel(), flags, stream_ptr())
            else:
                rc = lib().dplda_pack_weights(*[ptr(p) for p in ps], d_in, d1, ptr(self.buf),
                                              self.buf.numel(), flags, stream_ptr())
        check(rc, "pack_weights")
        return self.buf

    dense_skip = 0            # dense-list probe (score_indexed): calls left to skip / current back-off, per table
    dense_backoff = 0
    dense_table = None
    rowtab = None
    rowtab_key = None
    rowtab_table = None       # strong reference to the table the cached rows were built from

    @staticmethod
    def _table_key(kind, table, d_in, d1, d2):
        """Identity of a caller-owned table: the tensor OBJECT (held strongly in `rowtab_table` while its rows are
        cached, so neither its id nor its storage address can be recycled by another table), its version counter and
        shape.  Writes that bypass the version counter (`table.data.copy_()`) need `invalidate_table_caches()`."""
        return (kind, d_in, d1, d2, id(table), table._version, tuple(table.shape), table.device)

    def rowtab_valid(self, kind, table, d_in, d1, d2):
        """Rows of THIS table are in the cache (whether they match the current parameters is checked on the device)."""
        return (self.rowtab_table is table and self.rowtab_key is not None
                and self.rowtab_key == self._table_key(kind, table, d_in, d1, d2))

    def get_rowtab(self, kind, table, params, d_in, d1, d2):
        """Per-utterance score operands of `table` (nplda_table_prepare).  The table is identified on the host
        (pointer, version, shape -- a table is an input the caller owns; the loaders' table is never written);
        the parameters by the fingerprint of the fresh pack, on the device: the prepare call is a few empty
        launches when nothing changed."""
        pack = self.get(kind, params, d_in, d1, d2)
        src = table
        table = _f32c(table)
        # a converted / compacted copy is a temporary: its rows are never cached (the next temporary may reuse its address)
        key = self._table_key(kind, src, d_in, d1, d2) if table is src else None
        n_rows = table.shape[0]
        nbytes = lib().nplda_rowtab_bytes(n_rows)
        if self.rowtab is None or self.rowtab.numel() * 4 < nbytes or self.rowtab.device != table.device:
            self.rowtab = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=table.device)
            self.rowtab_key = None
        same = key is not None and self.rowtab_table is src and key == self.rowtab_key
        flags = (_lib.PREPARE_IF_CHANGED if same else 0) | (_lib.PACK_EPOCH_ODD if self.epoch else 0)
        with on_device(table.device):
            check(lib().nplda_table_prepare(ptr(table), n_rows, d_in, d1, d2, ptr(pack), 0 if kind == "nplda" else 1,
                                            ptr(self.rowtab), flags, stream_ptr()), "nplda_table_prepare")
        self.rowtab_key = key
        self.rowtab_table = src if key is not None else None
        return self.rowtab


# ---- pre-split tables (csrc/score_tcx.cu) ----------------------------------------------------------------------
# The split image of a table depends on the table only, not on any module's parameters: one cache for the process,
# keyed by the table tensor OBJECT.  Entries die with their table (weakref finaliser), so a recycled id never finds a
# stale entry; the version counter covers in-place updates through tensor ops.
_split_cache = {}


def _drop_split(key):
    _split_cache.pop(key, None)


def invalidate_table_caches(packed=None):
    """Forget every cached per-table product (split images; the row table of `packed` if given).  Needed only after
    writing into a table behind autograd's back (`table.data.copy_()` does not bump the version counter)."""
    _split_cache.clear()
    if packed is not None:
        packed.rowtab_key = None
        packed.rowtab_table = None


def split_table(table):
    """bf16 hi/lo operand image of a [rows, d_in] fp32 CUDA table (nplda_table_split), cached per table object."""
    require_cuda(table)
    src = table
    table = _f32c(table)
    cacheable = table is src
    ent = _split_cache.get(id(src)) if cacheable else None
    if ent is not None and ent[0]() is src and ent[1] == src._version and ent[2].device == src.device:
        return ent[2]
    n_rows, d_in = table.shape
    nbytes = lib().nplda_split_bytes(n_rows, d_in)
    if nbytes < 0:
        check(int(nbytes), "nplda_split_bytes")
    split = torch.empty(max(int(nbytes), 128), dtype=torch.uint8, device=table.device)
    with on_device(table.device):
        check(lib().nplda_table_split(ptr(table), n_rows, d_in, ptr(split), stream_ptr()), "nplda_table_split")
    if cacheable:
        key = id(src)
        _split_cache[key] = (weakref.ref(src), src._version, split)
        weakref.finalize(src, _drop_split, key)
    return split


def split_supported(kind, d_in, d1, d2):
    return kind == "nplda" and d_in % 32 == 0 and d_in >= 64 and max(d1, d2) <= 176


def score_split(table, split, i1, i2, params, dims, packed, flag_ptr=None):
    """NeuralPlda scores of trials (table[i1[k]], table[i2[k]]) from the pre-split image of the table
    (nplda_score_fwd_split: TMA row gather + tcgen05 CTA-pair MMAs)."""
    d_in, d1, d2 = dims
    n = i1.numel()
    dev = split.device
    scores = torch.empty(n, dtype=torch.float32, device=dev)
    fp, flag = _flag(dev, flag_ptr)
    pack = packed.get("nplda", params, d_in, d1, d2, pair=True)
    with on_device(dev):
        check(lib().nplda_score_fwd_split(ptr(split), table.shape[0], ptr(i1), ptr(i2), n, d_in, d1, d2, ptr(pack),
                                          ptr(scores), fp, stream_ptr()), "nplda_score_fwd_split")
    return scores, flag


def _zero_grads(params, need):
    """Zeroed gradient buffers for the parameters that need one: slices of ONE allocation (one fill launch per
    backward instead of one per parameter; the reference trains with 128-pair batches, where launches count)."""
    sizes = [(p.numel() + 3) // 4 * 4 if nd else 0 for p, nd in zip(params, need)]     # 16-byte aligned slices
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=params[0].device)
    out, o = [], 0
    for p, nd, sz in zip(params, need, sizes):
        out.append(flat[o:o + p.numel()].view(p.shape) if nd else None)
        o += sz
    return out


SAVE_ACTIVATIONS_MIN_PAIRS = 64       # one tile of the tensor-core kernel; below, the all-fp32 tile kernel does everything


def _wants_activations(ctx, packed, impl, n, dev, grad_mode):
    """A forward whose backward will run keeps the activations the backward needs (3 KB per pair for NeuralPlda,
    4.6 KB for DPlda -- about the size of the inputs; PyTorch's autograd in the reference keeps far more) instead of
    recomputing them in the backward.  `module.packed.save_activations = False` restores recomputation.
    `grad_mode` is torch.is_grad_enabled() at the CALL (inside Function.forward it is always off, and
    ctx.needs_input_grad ignores torch.no_grad()): evaluation forwards take the plain score kernel."""
    if not (grad_mode and n >= SAVE_ACTIVATIONS_MIN_PAIRS and impl in (_lib.IMPL_AUTO, _lib.IMPL_TC)
            and any(ctx.needs_input_grad) and getattr(packed, "save_activations", True)):
        return False
    nbytes = 6 * n * 176 * 4                          # upper bound (DPlda); only worth a driver query when it is large
    if nbytes > (1 << 30):
        free, _ = torch.cuda.mem_get_info(dev)
        if nbytes > free // 2:                        # not obviously fine: count what the caching allocator holds free, too
            cached = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)     # (memory_stats: ~0.2 ms of host time)
            if nbytes > (free + cached) // 2:         # huge batch: recompute in the backward rather than risk OOM
                return False
    return True


def _check_pair_inputs(x1, x2, d_in):
    require_cuda(x1, x2)
    if x1.dim() != 2 or x2.dim() != 2 or x1.shape != x2.shape:
        raise RuntimeError(f"expected two [N, {d_in}] tensors, got {tuple(x1.shape)} and {tuple(x2.shape)}")
    if x1.shape[1] != d_in:
        # same failure class as the reference's addmm shape error
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({x1.shape[0]}x{x1.shape[1]} and "
                           f"{d_in}x...)")


class NpldaScoreFn(torch.autograd.Function):
    """S = NeuralPlda.forward(x1, x2)  (reference utils/models.py:378-382)."""

    @staticmethod
    def forward(ctx, x1, x2, W1, b1, W2, b2, P_sqrt, Q, packed, impl, grad_mode=True):
        d1, d_in = W1.shape
        d2 = W2.shape[0]
        _check_pair_inputs(x1, x2, d_in)
        x1c, x2c = _f32c(x1), _f32c(x2)
        n = x1c.shape[0]
        pack = packed.get("nplda", (W1, b1, W2, b2, P_sqrt, Q), d_in, d1, d2, mixed=(impl in (_lib.IMPL_TC_F8, _lib.IMPL_TC_PAIR_F8)))
        scores = torch.empty(n, dtype=torch.float32, device=x1c.device)
        ctx.act = None
        with on_device(x1c.device):
            rc = _lib.ERR_UNSUPPORTED_DIM
            if _wants_activations(ctx, packed, impl, n, x1c.device, grad_mode):
                act = torch.empty(int(lib().nplda_act_floats(n, 0)), dtype=torch.float32, device=x1c.device)
                rc = lib().nplda_score_fwd_train(ptr(x1c), ptr(x2c), n, d_in, d1, d2, ptr(pack), ptr(scores), ptr(act),
                                                 stream_ptr())
                if rc == 0:
                    ctx.act = act
                elif rc != _lib.ERR_UNSUPPORTED_DIM:
                    check(rc, "nplda_score_fwd_train")
            if rc != 0:
                check(lib().nplda_score_fwd(ptr(x1c), ptr(x2c), n, d_in, d1, d2, ptr(pack), ptr(scores), impl,
                                            stream_ptr()), "nplda_score_fwd")
        ctx.save_for_backward(x1c, x2c, W1, b1, W2, b2, P_sqrt, Q)
        return scores

    @staticmethod
    def backward(ctx, ds):
        x1, x2, W1, b1, W2, b2, P_sqrt, Q = ctx.saved_tensors
        d1, d_in = W1.shape
        d2 = W2.shape[0]
        n = x1.shape[0]
        dev = x1.device
        ds = _f32c(ds)
        need = ctx.needs_input_grad
        params = [_f32c(p.detach()) for p in (W1, b1, W2, b2, P_sqrt, Q)]
        grads = _zero_grads(params, need[2:8])
        dx1 = torch.zeros_like(x1) if need[0] else None
        dx2 = torch.zeros_like(x2) if need[1] else None
        if n > 0:
            with on_device(dev):
                wsb = lib().nplda_bwd_workspace_bytes(n, d_in, d1, d2)
                if wsb < 0:
                    check(wsb, "nplda_bwd_workspace_bytes")
                ws = torch.empty(max(int(wsb), 16), dtype=torch.uint8, device=dev)
                check(lib().nplda_score_bwd_act(ptr(x1), ptr(x2), n, d_in, d1, d2, *[ptr(p) for p in params],
                                                ptr(ds), *[ptr(g) for g in grads], ptr(dx1), ptr(dx2), ptr(ctx.act),
                                                ptr(ws), ws.numel(), stream_ptr()), "nplda_score_bwd")
        return (dx1, dx2, *grads, None, None, None)


class DpldaScoreFn(torch.autograd.Function):
    """S = DPlda.forward(x1, x2)  (reference utils/models.py:491-495)."""

    @staticmethod
    def forward(ctx, x1, x2, W1, b1, w_lr, c_lr, packed, impl, grad_mode=True):
        d1, d_in = W1.shape
        _check_pair_inputs(x1, x2, d_in)
        if w_lr.numel() != 2 * d1 * d1 + d1:
            raise RuntimeError("logistic_regres.weight has the wrong size for layer1_LDA_dim")
        x1c, x2c = _f32c(x1), _f32c(x2)
        n = x1c.shape[0]
        pack = packed.get("dplda", (W1, b1, w_lr, c_lr), d_in, d1, d1)
        scores = torch.empty(n, dtype=torch.float32, device=x1c.device)
        ctx.act = None
        ctx.lda_frozen = False
        if _wants_activations(ctx, packed, impl, n, x1c.device, grad_mode):
            # LDA frozen (the reference driver's case, xvector_DPlda_pytorch.py:140-147) and no input gradients: the
            # backward is the gradient of logistic_regres alone and needs only the normalised rows u
            frozen = not any(ctx.needs_input_grad[:4])
            with on_device(x1c.device):
                act = torch.empty(int(lib().nplda_act_floats(n, 2 if frozen else 1)), dtype=torch.float32, device=x1c.device)
                fn = lib().dplda_score_fwd_train_u if frozen else lib().dplda_score_fwd_train
                rc = fn(ptr(x1c), ptr(x2c), n, d_in, d1, ptr(pack), ptr(scores), ptr(act), stream_ptr())
            if rc == 0:
                ctx.act = act
                ctx.lda_frozen = frozen
                ctx.save_for_backward(x1c, x2c, W1, b1, w_lr)
                return scores
            if rc != _lib.ERR_UNSUPPORTED_DIM:
                check(rc, "dplda_score_fwd_train")
        with on_device(x1c.device):
            wsb = lib().dplda_fwd_workspace_bytes(n, d_in, d1) if impl in (_lib.IMPL_AUTO, _lib.IMPL_TC) else 0
            if wsb < 0:
                check(wsb, "dplda_fwd_workspace_bytes")
            ws = torch.empty(int(wsb), dtype=torch.uint8, device=x1c.device) if wsb > 0 else None
            check(lib().dplda_score_fwd_ws(ptr(x1c), ptr(x2c), n, d_in, d1, ptr(pack), ptr(scores), impl,
                                           ptr(ws), int(wsb), stream_ptr()), "dplda_score_fwd")
        ctx.save_for_backward(x1c, x2c, W1, b1, w_lr)
        return scores

    @staticmethod
    def backward(ctx, ds):
        x1, x2, W1, b1, w_lr = ctx.saved_tensors
        d1, d_in = W1.shape
        n = x1.shape[0]
        dev = x1.device
        ds = _f32c(ds)
        need = ctx.needs_input_grad
        W1c, b1c, wc = _f32c(W1.detach()), _f32c(b1.detach()), _f32c(w_lr.detach())
        dW1, db1, dw, dc = _zero_grads([W1c, b1c, wc, b1c[:1]], need[2:6])       # dc: [1] like logistic_regres.bias
        dx1 = torch.zeros_like(x1) if need[0] else None
        dx2 = torch.zeros_like(x2) if need[1] else None
        if n > 0 and ctx.lda_frozen:
            with on_device(dev):
                wsb = lib().dplda_lr_bwd_workspace_bytes(n, d1)
                if wsb < 0:
                    check(wsb, "dplda_lr_bwd_workspace_bytes")
                ws = torch.empty(max(int(wsb), 16), dtype=torch.uint8, device=dev)
                check(lib().dplda_lr_bwd(ptr(ctx.act), n, d1, ptr(ds), ptr(dw), ptr(dc), ptr(ws), ws.numel(), stream_ptr()),
                      "dplda_lr_bwd")
            return (None, None, None, None, dw, dc, None, None, None)
        if n > 0:
            with on_device(dev):
                wsb = lib().nplda_bwd_workspace_bytes(n, d_in, d1, d1)
                if wsb < 0:
                    check(wsb, "nplda_bwd_workspace_bytes")
                ws = torch.empty(max(int(wsb), 16), dtype=torch.uint8, device=dev)
                check(lib().dplda_score_bwd_act(ptr(x1), ptr(x2), n, d_in, d1, ptr(W1c), ptr(b1c), ptr(wc), ptr(ds),
                                                ptr(dW1), ptr(db1), ptr(dw), ptr(dc), ptr(dx1), ptr(dx2), ptr(ctx.act),
                                                ptr(ws), ws.numel(), stream_ptr()), "dplda_score_bwd")
        return (dx1, dx2, dW1, db1, dw, dc, None, None, None)


def embed(kind, x, params, dims, packed):
    """extract_plda_embeddings of the reference (models.py:366-370 / 478-481), no autograd."""
    from .lazy import dense
    x = dense(x)
    require_cuda(x)
    d_in, d1, d2 = dims
    if x.dim() != 2 or x.shape[1] != d_in:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({x.shape[0]}x{x.shape[-1]} and {d_in}x{d1})")
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        raise RuntimeError("extract_plda_embeddings is provided without autograd: call it under torch.no_grad() "
                           "(training goes through forward(x1, x2), which is differentiable)")
    x = _f32c(x)
    n = x.shape[0]
    pack = packed.get(kind, params, d_in, d1, d2)
    width = d2 if kind == "nplda" else d1
    out = torch.empty(n, width, dtype=torch.float32, device=x.device)
    with on_device(x.device):
        check(lib().nplda_embed_fwd(ptr(x), n, d_in, d1, d2, ptr(pack), ptr(out), 0 if kind == "nplda" else 1,
                                    stream_ptr()), "nplda_embed_fwd")
    return out


def score_from_embeddings(kind, e1, e2, params, dims, packed):
    """forward_from_plda_embeddings of the reference (models.py:372-376 / 483-489), no autograd."""
    require_cuda(e1, e2)
    d_in, d1, d2 = dims
    width = d2 if kind == "nplda" else d1
    if e1.shape != e2.shape or e1.dim() != 2 or e1.shape[1] != width:
        raise RuntimeError(f"expected two [N, {width}] embedding tensors")
    if torch.is_grad_enabled() and (e1.requires_grad or e2.requires_grad or any(p.requires_grad for p in params)):
        raise RuntimeError("forward_from_plda_embeddings is provided without autograd: call it under torch.no_grad()")
    e1, e2 = _f32c(e1), _f32c(e2)
    n = e1.shape[0]
    scores = torch.empty(n, dtype=torch.float32, device=e1.device)
    with on_device(e1.device):
        if kind == "nplda":
            ps, q = _f32c(params[4].detach()), _f32c(params[5].detach())
            check(lib().nplda_score_from_embeddings(ptr(e1), ptr(e2), n, d2, ptr(ps), ptr(q), ptr(scores),
                                                    stream_ptr()), "nplda_score_from_embeddings")
        else:
            pack = packed.get(kind, params, d_in, d1, d2)
            check(lib().dplda_score_from_embeddings(ptr(e1), ptr(e2), n, d_in, d1, ptr(pack), ptr(scores),
                                                    stream_ptr()), "dplda_score_from_embeddings")
    return scores


SPLIT_MIN_TRIALS = 128       # one CTA-pair tile; below, the fp32 kernels are launch-bound either way
GRID_GATHER_MIN_TRIALS = 1 << 18     # dense-list path: worth its extra launches and the read-back of two row counts
GRID_GATHER_MAX_FILL = 6             # ... when the sub-grid has at most this many entries per trial
GRID_GATHER_MAX_BYTES = 2 << 30      # ... and fits this scratch budget


def _score_dense_list(rowtab, n_rows, i1, i2, scores, flag_ptr):
    """Trial list as a sub-grid + gather (csrc/pairs.cu): True if the list was dense enough and `scores` is filled.
    Reads the two row counts back to the host (one small synchronising copy) to size the grid."""
    dev = rowtab.device
    n = i1.numel()
    flags = torch.empty(2 * n_rows, dtype=torch.int32, device=dev)
    pos = torch.empty(2 * n_rows, dtype=torch.int32, device=dev)
    lst = torch.empty(2 * n_rows, dtype=torch.int64, device=dev)
    counts = torch.empty(2, dtype=torch.int32, device=dev)
    with on_device(dev):
        check(lib().nplda_trial_rows(ptr(i1), ptr(i2), n, n_rows, ptr(flags), ptr(pos), ptr(lst), ptr(counts), flag_ptr,
                                     stream_ptr()), "nplda_trial_rows")
        ne, nt = counts.tolist()
        cells = ne * nt
        if cells == 0 or cells > GRID_GATHER_MAX_FILL * n or cells * 4 > GRID_GATHER_MAX_BYTES:
            return False
        ld = (nt + 3) // 4 * 4
        grid = torch.empty(ne * ld, dtype=torch.float32, device=dev)
        check(lib().nplda_score_grid_impl(ptr(rowtab), n_rows, ptr(lst), ne, ptr(lst[n_rows:]), nt, ptr(grid), ld, flag_ptr,
                                          _lib.IMPL_AUTO, stream_ptr()), "nplda_score_grid")
        check(lib().nplda_trial_grid_gather(ptr(grid), ld, ptr(pos), n_rows, ptr(i1), ptr(i2), n, ptr(scores), stream_ptr()),
              "nplda_trial_grid_gather")
    return True


def _flag(dev, flag_ptr):
    """(ctypes pointer for the kernels, tensor to return): a fresh device int32, or the caller's (pinned host) flag."""
    if flag_ptr is not None:
        return flag_ptr, None
    t = torch.zeros(1, dtype=torch.int32, device=dev)
    return ptr(t), t


def score_indexed(kind, table, i1, i2, params, dims, packed, impl=_lib.IMPL_AUTO, embed_once=None, use_split=None,
                  flag_ptr=None):
    """Scores of trials (table[i1[k]], table[i2[k]]) (replaces sv_trials_loaders.load_xvec_trials_from_numbatch
    + forward).  Three device paths:
      * embed once: every table row is transformed ONCE (nplda_table_prepare, cached per table object / parameter
        fingerprint) and a trial is r[i] + r[j] + A[i].B[j] (nplda_score_pairs) -- default when the row table is
        already valid or the table has at most twice as many rows as this call has trials;
      * split: both sides of every trial go through both layers on the tensor cores, rows gathered by TMA from the
        pre-split image of the table (nplda_score_fwd_split) -- default otherwise (e.g. training-sized batches over
        a large table) for NeuralPlda shapes the kernel takes;
      * the fp32 kernel with the gather fused in (any shape, DPlda)."""
    require_cuda(table, i1, i2)
    d_in, d1, d2 = dims
    if table.dim() != 2 or table.shape[1] != d_in:
        raise RuntimeError(f"table must be [rows, {d_in}]")
    i1 = i1.to(torch.int64).contiguous()
    i2 = i2.to(torch.int64).contiguous()
    if i1.shape != i2.shape or i1.dim() != 1:
        raise RuntimeError("index tensors must be 1-D and the same length")
    n = i1.numel()
    dev = table.device
    if embed_once is None:
        embed_once = (not use_split) and max(d1, d2) < 176 and table.shape[0] > 0 and (
            packed.rowtab_valid(kind, table, d_in, d1, d2) or table.shape[0] <= 2 * n)
    if embed_once:
        scores = torch.empty(n, dtype=torch.float32, device=dev)
        fp, flag = _flag(dev, flag_ptr)
        rowtab = packed.get_rowtab(kind, table, params, d_in, d1, d2)
        if n >= GRID_GATHER_MIN_TRIALS and impl != _lib.IMPL_SIMT:
            # lists over this table that turned out sparse are not probed again for a while (the probe costs one pass
            # over the index lists and a small synchronising read-back): back-off doubles up to 1024 calls
            if packed.dense_table is None or packed.dense_table() is not table:       # another table: forget the back-off
                packed.dense_table = weakref.ref(table)
                packed.dense_skip = packed.dense_backoff = 0
            if packed.dense_skip > 0:
                packed.dense_skip -= 1
            elif _score_dense_list(rowtab, table.shape[0], i1, i2, scores, fp):
                packed.dense_backoff = 0
                return scores, flag
            else:
                packed.dense_backoff = min(1024, max(8, 2 * packed.dense_backoff))
                packed.dense_skip = packed.dense_backoff
        with on_device(dev):
            check(lib().nplda_score_pairs(ptr(rowtab), table.shape[0], ptr(i1), ptr(i2), n, ptr(scores), fp,
                                          stream_ptr()), "nplda_score_pairs")
        return scores, flag
    if use_split is None:
        use_split = (split_supported(kind, d_in, d1, d2) and impl in (_lib.IMPL_AUTO, _lib.IMPL_TC)
                     and n >= SPLIT_MIN_TRIALS and table.shape[0] > 0)
    if use_split:
        if not split_supported(kind, d_in, d1, d2):
            raise RuntimeError("the pre-split tensor-core path takes NeuralPlda shapes with d_in % 32 == 0, widths <= 176")
        return score_split(table, split_table(table), i1, i2, params, dims, packed, flag_ptr)
    table = _f32c(table)
    scores = torch.empty(n, dtype=torch.float32, device=dev)
    fp, flag = _flag(dev, flag_ptr)
    pack = packed.get(kind, params, d_in, d1, d2, mixed=(impl in (_lib.IMPL_TC_F8, _lib.IMPL_TC_PAIR_F8)))
    if impl in (_lib.IMPL_TC, _lib.IMPL_TC_F8, _lib.IMPL_TC_BF16, _lib.IMPL_TC_PAIR):
        impl = _lib.IMPL_AUTO                     # the materialised-pair tcgen05 kernel has no gather; fp32 kernel
    with on_device(dev):
        if kind == "nplda":
            rc = lib().nplda_score_fwd_indexed(ptr(table), table.shape[0], ptr(i1), ptr(i2), n, d_in, d1, d2,
                                               ptr(pack), ptr(scores), fp, impl, stream_ptr())
        else:
            rc = lib().dplda_score_fwd_indexed(ptr(table), table.shape[0], ptr(i1), ptr(i2), n, d_in, d1,
                                               ptr(pack), ptr(scores), fp, impl, stream_ptr())
    check(rc, "score_fwd_indexed")
    return scores, flag


def score_grid(kind, table, enrol_rows, test_rows, params, dims, packed, impl=_lib.IMPL_AUTO):
    """[E, T] scores of every enrol row against every test row of `table` (enrol-major trial order): the row
    table of nplda_table_prepare (cached like in score_indexed) and one fp32 grid product (nplda_score_grid)."""
    require_cuda(table, enrol_rows, test_rows)
    d_in, d1, d2 = dims
    if table.dim() != 2 or table.shape[1] != d_in:
        raise RuntimeError(f"table must be [rows, {d_in}]")
    if max(d1, d2) >= 176:
        raise RuntimeError("grid scoring supports layer widths up to 175")
    er = enrol_rows.to(torch.int64).contiguous()
    tr = test_rows.to(torch.int64).contiguous()
    if er.dim() != 1 or tr.dim() != 1:
        raise RuntimeError("enrol_rows and test_rows must be 1-D")
    dev = table.device
    scores = torch.empty(er.numel(), tr.numel(), dtype=torch.float32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    if scores.numel():
        if table.shape[0] == 0:
            raise RuntimeError("empty x-vector table")
        rowtab = packed.get_rowtab(kind, table, params, d_in, d1, d2)
        with on_device(dev):
            check(lib().nplda_score_grid_impl(ptr(rowtab), table.shape[0], ptr(er), er.numel(), ptr(tr), tr.numel(),
                                              ptr(scores), tr.numel(), ptr(flag),
                                              _lib.IMPL_SIMT if impl == _lib.IMPL_SIMT else _lib.IMPL_AUTO, stream_ptr()),
                  "nplda_score_grid")
    return scores, flag


# ------------------------------------------------------------------------------
# Losses
# ------------------------------------------------------------------------------


def loss_accumulators(scores, target, thresholds, alpha, th_xent, group=None):
    """Raw fp64 sums (layout in include/nplda.h), all-reduced over `group` if given.

    The per-GPU shards of a trial list must be combined as RAW SUMS before any
    division: per-rank label counts differ.
    """
    require_cuda(scores, target)
    s, t = _f32c(scores), _f32c(target)
    if s.shape != t.shape:
        raise RuntimeError("scores and targets must have the same shape")
    K = 0 if thresholds is None else thresholds.numel()
    acc = torch.zeros(4 * K + 4, dtype=torch.float64, device=s.device)
    th = None if K == 0 else _f32c(thresholds.detach())
    thx = None if th_xent is None else _f32c(th_xent.detach())
    with on_device(s.device):
        check(lib().nplda_loss_accum(ptr(s), ptr(t), s.numel(), ptr(th), K, float(alpha), ptr(thx), ptr(acc),
                                     stream_ptr()), "nplda_loss_accum")
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=None if group is True else group)
    return acc


def finalize(acc, betas):
    out = torch.empty(4, dtype=torch.float32, device=acc.device)
    K = len(betas)
    with on_device(acc.device):
        check(lib().nplda_loss_finalize(ptr(acc), _lib.betas_array(betas), K, ptr(out), stream_ptr()),
              "nplda_loss_finalize")
    return out


class LossFn(torch.autograd.Function):
    """softcdet / crossentropy of the reference (models.py:384-393) on the GPU.

    thresholds: [K] fp32 (the Th{beta} parameters concatenated), th_xent: [1] or None.
    """

    @staticmethod
    def forward(ctx, scores, target, thresholds, th_xent, betas, alpha, loss_id, group):
        acc = loss_accumulators(scores, target, thresholds, alpha, th_xent, group)
        out = finalize(acc, betas)
        ctx.save_for_backward(_f32c(scores), _f32c(target), thresholds, th_xent, acc)
        ctx.betas, ctx.alpha, ctx.loss_id, ctx.group = list(betas), float(alpha), loss_id, group
        return out[loss_id].clone()

    @staticmethod
    def backward(ctx, g):
        s, t, thresholds, th_xent, acc = ctx.saved_tensors
        K = len(ctx.betas)
        dev = s.device
        ds = torch.empty_like(s)
        dth = torch.zeros(K + 1, dtype=torch.float64, device=dev)
        g = _f32c(g).reshape(1)
        th = None if thresholds is None else _f32c(thresholds.detach())
        thx = None if th_xent is None else _f32c(th_xent.detach())
        with on_device(dev):
            check(lib().nplda_loss_bwd(ptr(s), ptr(t), s.numel(), ptr(th), _lib.betas_array(ctx.betas), K,
                                       ctx.alpha, ptr(thx), ptr(acc), ctx.loss_id, ptr(g), ptr(ds), ptr(dth),
                                       stream_ptr()), "nplda_loss_bwd")
        # Threshold gradients are sums over this rank's trials; under a process
        # group the caller all-reduces parameter gradients (DDP-style), as for
        # the other parameters.
        # A parameter the loss does not read gets NO gradient (None), as under the reference's autograd: softCdet never
        # touches threshold_Xent and BCE never touches the Th{beta} thresholds (models.py:384-393).  A zero tensor would
        # not be equivalent: Adam(weight_decay=1e-5) -- the reference's optimiser, xvector_NeuralPlda_pytorch.py:139 --
        # skips parameters whose grad is None but decays (and, normalised, moves by lr per step) those with a zero grad.
        soft = ctx.loss_id == _lib.LOSS_SOFTCDET
        dthr = dth[:K].float() if (thresholds is not None and ctx.needs_input_grad[2] and soft) else None
        dthx = dth[K:].float() if (th_xent is not None and ctx.needs_input_grad[3] and not soft) else None
        return ds, None, dthr, dthx, None, None, None, None


def sorted_populations(scores, target):
    """(sorted target scores, sorted non-target scores) of models.py:407-408 -- `torch.sort(output[target > 0.5])`,
    `torch.sort(output[target < 0.5])` -- with the library's own split and sort kernels (csrc/sort.cu).  Reads the two
    population sizes back to the host (the sweep's caller synchronises anyway for the label sums)."""
    require_cuda(scores, target)
    s, t = _f32c(scores.detach()).reshape(-1), _f32c(target.detach()).reshape(-1)
    if s.shape != t.shape:
        raise RuntimeError("scores and targets must have the same shape")
    n = s.numel()
    dev = s.device
    tgt = torch.empty(n, dtype=torch.float32, device=dev)
    non = torch.empty(n, dtype=torch.float32, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    with on_device(dev):
        check(lib().nplda_split_by_label(ptr(s), ptr(t), n, ptr(tgt), ptr(non), ptr(counts), stream_ptr()), "nplda_split_by_label")
        n_t, n_n = counts.tolist()
        check(lib().nplda_sort_f32(ptr(tgt), n_t, stream_ptr()), "nplda_sort_f32")
        check(lib().nplda_sort_f32(ptr(non), n_n, stream_ptr()), "nplda_sort_f32")
    return tgt[:n_t], non[:n_n]


def minc_sweep(tgt_sorted, non_sorted, sum_t, sum_n, betas):
    """(min cost [K] fp32, argmin index [K] int64 into tgt_sorted); models.py:406-421."""
    require_cuda(tgt_sorted, non_sorted)
    K = len(betas)
    dev = tgt_sorted.device
    out_min = torch.empty(K, dtype=torch.float32, device=dev)
    out_arg = torch.empty(K, dtype=torch.int64, device=dev)
    with on_device(dev):
        check(lib().nplda_minc_sweep(ptr(tgt_sorted), tgt_sorted.numel(), ptr(non_sorted), non_sorted.numel(),
                                     ctypes.c_float(sum_t), ctypes.c_float(sum_n), _lib.betas_array(betas), K,
                                     ptr(out_min), ptr(out_arg), stream_ptr()), "nplda_minc_sweep")
    return out_min, out_arg
