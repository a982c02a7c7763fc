"""Host-side data-parallel plumbing (device-agnostic: the same code runs under
``gloo`` on CPU tensors in the tests and under ``nccl`` on the B200s).

The reference has no distributed training (``utils/distributed.py`` is never
called and DDP is commented out, src/xview_lstm/agent.py:144-150); the semantics
here are torch-DDP defaults: replicas start from rank 0's parameters, every rank
processes its own shard of episodes / poses, gradients are summed across ranks
and divided by the world size, BatchNorm statistics stay per-rank.
"""
from __future__ import annotations

import torch


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous, balanced shard ``[lo, hi)`` of ``n_total`` independent units
    (episodes for training, poses for rendering)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _default_cuts():
    import os
    e = os.environ.get("AVDN_BUCKET_CUTS")          # e.g. "0.75,0.45" (experiments)
    if e is None:
        return (0.75, 0.5, 0.3, 0.18)
    return tuple(float(x) for x in e.split(",") if x.strip())      # "" = one bucket, reduced after the backward pass


def bucket_edges(first_offsets, total, cut_fracs=None):
    """Split a flat gradient arena laid out in FORWARD layer order into buckets that
    complete in BACKWARD order.  ``first_offsets[i]`` is the arena offset of the first
    parameter of layer ``i``.  Returns ``[(layer_pos, lo, hi), ...]``: once the backward
    pass has finished layer ``layer_pos``, ``flat[lo:hi]`` is final and can be reduced.

    The default cuts are for the Darknet trunk: its deep blocks hold the parameters (and finish first), its first
    ten blocks hold a third of the backward TIME but < 1 MB of gradients -- so the last bucket, the only one the
    optimiser has to wait for, is a latency-sized all-reduce."""
    if cut_fracs is None:
        cut_fracs = _default_cuts()
    nl = len(first_offsets)
    cuts = sorted({min(nl - 1, max(0, int(nl * f))) for f in cut_fracs} | {0}, reverse=True)
    edges, hi = [], total
    for c in cuts:
        lo = first_offsets[c]
        if hi > lo:
            edges.append((c, lo, hi))
        hi = lo
    return edges


def allreduce_sum_(flat: torch.Tensor, lo: int, hi: int, group=None):
    import torch.distributed as dist
    dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group)


def broadcast_(tensors, src=0, group=None):
    import torch.distributed as dist
    for t in tensors:
        dist.broadcast(t, src, group=group)


# ---- result merging of multi-GPU evaluation (src/utils/distributed.py:90-164) ----
def get_world_size():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def all_gather(data, group=None):
    """``all_gather`` of arbitrary picklable data (the per-rank trajectory dicts of ``agent.test``): returns
    the list of every rank's object, in rank order (src/utils/distributed.py:90-130)."""
    import torch.distributed as dist
    world = get_world_size()
    if world == 1:
        return [data]
    out = [None] * world
    dist.all_gather_object(out, data, group=group)
    return out


def merge_dist_results(results):
    """src/utils/distributed.py:160-164: concatenate the per-rank result lists."""
    outs = []
    for res in results:
        outs.extend(res)
    return outs


def reduce_dict(input_dict, average=True, group=None):
    """src/utils/distributed.py:133-157: sum (or average) a dict of scalar tensors over the ranks."""
    import torch.distributed as dist
    world = get_world_size()
    if world < 2:
        return input_dict
    with torch.no_grad():
        names = sorted(input_dict.keys())
        values = torch.stack([input_dict[k] for k in names], dim=0)
        dist.all_reduce(values, group=group)
        if average:
            values /= world
        return {k: v for k, v in zip(names, values)}
