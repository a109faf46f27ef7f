"""Batch sharding over the GPUs of one box (one process per GPU, ``torch.distributed``).

The D-ADMM problems of a batch are independent given the shared ``A``, the hyper-parameter table /
hypernetwork weights and the per-problem graphs, so the data path has NO collective:
each rank solves a contiguous slice of the batch.  The only exchanges are one all-reduce(sum) of the
parameter gradients and one of the per-iteration loss sums per step -- NCCL over NVLink on the GPU box,
gloo in the CPU tests.  The reference has no distributed code at all; this is new surface.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of B problems owned by ``rank`` (first B % world ranks get one extra)."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(b: torch.Tensor, label: torch.Tensor, graph_list: Sequence, rank: int, world: int):
    lo, hi = shard_range(len(b), rank, world)
    return b[lo:hi], label[lo:hi], list(graph_list[lo:hi]), (lo, hi)


def sharded_noise(B: int, lo: int, hi: int, P: int, n: int, device, seed: int):
    """The reference's three N(0,1)*1e-2 draws (unfolded_DLASSO.py:49-51) for the FULL batch under one
    seed, sliced to this rank's problems -- so an N-rank run sees the same initial state as a 1-rank run."""
    gen = torch.Generator(device=device).manual_seed(seed)
    out = []
    for _ in range(3):
        full = torch.randn((B, P, n, 1), device=device, generator=gen)
        out.append((full[lo:hi] * 1e-2).contiguous())
        del full
    return tuple(out)


def allreduce_sum_(tensors: Iterable[torch.Tensor], group=None) -> None:
    """In-place all-reduce(sum) of a list of tensors as ONE flat bucket (they are small: [K,P,4] for
    model #1, the hypernetwork for model #3), so the cost is one launch latency."""
    ts = [t for t in tensors if t is not None]
    if not ts or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in ts:
        k = t.numel()
        t.copy_(flat[off:off + k].view_as(t).to(t.dtype))
        off += k


def allreduce_gradients(module: torch.nn.Module, extra=(), group=None) -> None:
    """Sum parameter gradients (and any ``extra`` tensors, e.g. the detached per-rank loss shares) over ranks in
    one bucket.  With ``compute_loss(..., global_batch=B_total)`` each rank's loss is its share of the global mean,
    so the summed gradients equal the single-process gradient and the summed losses equal the global loss."""
    allreduce_sum_([p.grad for p in module.parameters() if p.grad is not None] + list(extra), group)
