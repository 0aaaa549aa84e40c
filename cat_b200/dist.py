"""Minibatch sharding of the CTC-CRF loss across the GPUs of one box.

Utterances are independent in both numerator and denominator and the den graph is read-only, so the only
partitioning is over utterances; the den graph is replicated (each rank's ``CRFContext`` loads it, as the
reference does per process, den_calculate.cu:375-390).  The path has exactly one exchange: an all-reduce
of the 2-vector [sum of per-utterance costs, utterance count] (NCCL over NVLink on the GPU box, gloo in the
CPU tests).  Gradients w.r.t. the logits stay rank-local.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_by_length(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first balancing of total frames per rank (cf. cat/shared/coreutils.py:445-490
    ``weighted_group``).  Returns, per rank, utterance indices in descending length order, so that within a
    rank lanes retire together and every rank's sequential critical path (its longest utterance) and total
    frame count are balanced."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    counts = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    cap = -(-len(lengths) // world_size)
    for i in order:
        cand = [r for r in range(world_size) if counts[r] < cap]
        r = min(cand, key=lambda r: (loads[r], r))
        shards[r].append(i)
        loads[r] += int(lengths[i])
        counts[r] += 1
    return shards


def shard_contiguous(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Fixed-length batches: contiguous N/G blocks."""
    per = -(-n // world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def select_utterances(labels: torch.Tensor, ly: torch.Tensor, idx: Sequence[int]) -> torch.Tensor:
    """Flattened labels of the selected utterances, in the order of ``idx``."""
    off = torch.zeros(ly.numel() + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(ly.to(torch.int64), 0)
    parts = [labels[int(off[i]):int(off[i + 1])] for i in idx]
    return torch.cat(parts) if parts else labels[:0]


def allreduce_cost_tensor(cost_sum: torch.Tensor, count: int, group=None) -> torch.Tensor:
    """One all-reduce(SUM) of [sum cost, count]; returns the global pair as a 2-vector on the device of ``cost_sum``
    (no host synchronisation)."""
    v = torch.stack([cost_sum.reshape(()).float(), torch.full((), float(count), device=cost_sum.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v


def allreduce_cost(cost_sum: torch.Tensor, count: int, group=None) -> Tuple[torch.Tensor, int]:
    """Same, with the count read back to the host."""
    v = allreduce_cost_tensor(cost_sum, count, group)
    return v[0], int(round(float(v[1])))


def local_shard(logits: torch.Tensor, labels: torch.Tensor, lx: torch.Tensor, ly: torch.Tensor, rank: int, world_size: int):
    """This rank's share of a GLOBAL batch: (utterance indices, logits[idx], flattened labels, lx[idx], ly[idx]), longest
    first.  In training the data loader does this (every rank reads its own utterances, cat/shared/data.py:471-585);
    benchmarks call it once, outside the timed region."""
    idx = shard_by_length(lx.tolist(), world_size)[rank]
    sel = torch.tensor(idx, dtype=torch.long)
    return idx, logits[sel.to(logits.device)].contiguous(), select_utterances(labels, ly, idx), lx[sel], ly[sel]


def sharded_step(loss_fn: Callable[..., Tuple[torch.Tensor, torch.Tensor]], shard, size_average: bool = True, group=None):
    """One step on an already sharded batch: the rank-local op, then the path's ONE collective -- all-reduce(SUM) of
    [sum of costs, utterance count] -- and the normalisation by the GLOBAL count (cat/shared/manager.py:546 arrives at the
    same factor from local_bs * world / global_bs).  Returns (global loss, local gradient of the global loss)."""
    idx, logits, labels, lx, ly = shard
    if idx:
        cost, grad = loss_fn(logits, labels, lx, ly)
    else:
        cost, grad = torch.zeros((), device=logits.device), logits[:0]
    v = allreduce_cost_tensor(cost, len(idx), group)          # the only exchange of the path; stays on the device
    scale = 1.0 / v[1].clamp(min=1.0) if size_average else torch.ones((), device=v.device)
    return v[0] * scale, grad * scale.to(grad.dtype)


def sharded_loss(loss_fn: Callable[..., Tuple[torch.Tensor, torch.Tensor]],
                 logits: torch.Tensor, labels: torch.Tensor, lx: torch.Tensor, ly: torch.Tensor,
                 rank: int, world_size: int, size_average: bool = True, group=None):
    """Evaluate the loss of a GLOBAL batch with the minibatch sharded over ``world_size`` ranks.

    ``loss_fn(logits_shard, labels_shard, lx_shard, ly_shard) -> (sum of per-utterance costs, grad of that
    sum w.r.t. logits_shard)`` is the rank-local op (size_average=False).  Returns
    (global loss, local utterance indices, local gradient of the global loss)."""
    shard = local_shard(logits, labels, lx, ly, rank, world_size)
    loss, grad = sharded_step(loss_fn, shard, size_average, group)
    return loss, shard[0], grad


def cuda_loss_fn(lamb: float):
    """The rank-local op for ``sharded_loss`` / ``sharded_step``: this library's fused CUDA loss with size_average=False,
    i.e. (sum of per-utterance costs, d sum / d logits)."""
    from . import _C

    def fn(logits, labels, lx, ly):
        loss, grad, _ = _C.ctc_crf_loss_fwd(logits, labels.to(torch.int32), lx.to(torch.int32), ly.to(torch.int32), lamb, False)
        return loss.reshape(()), grad
    return fn
