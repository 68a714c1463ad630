"""Trial-list sharding across GPUs (one process per GPU, torch.distributed).

The scoring path is embarrassingly parallel over trials: each rank scores a contiguous range of
the trial list with replicated parameters and no data-path collective.  The only coupling is the
loss normalisation: the 4K+4 fp64 accumulators (include/nplda.h) are summed across ranks with ONE
all-reduce (96 B at K=2) before any division; during training the parameter gradients are
all-reduced the same way (466 KB for NeuralPlda).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n, world_size, rank):
    """Contiguous [begin, end) of rank's share of n trials (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_accumulators(acc, group=None):
    """Sum the raw accumulators over ranks (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def losses_from_accumulators(acc, betas):
    """(softcdet, crossentropy, cdet) from the all-reduced sums -- host-side mirror of
    nplda_loss_finalize, used by the CPU tests and for logging."""
    a = [float(v) for v in acc.tolist()]
    K = len(betas)
    nt, nn, sb, n = a[4 * K: 4 * K + 4]
    soft = sum(a[4 * k] / nt + betas[k] * a[4 * k + 1] / nn for k in range(K)) / max(K, 1)
    hard = sum(a[4 * k + 2] / nt + betas[k] * a[4 * k + 3] / nn for k in range(K)) / max(K, 1)
    return soft, sb / n, hard


def allreduce_gradients(module, group=None):
    """DDP-style sum of parameter gradients after backward on each rank's shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off: off + g.numel()].view_as(g))
        off += g.numel()
