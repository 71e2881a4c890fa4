"""Data parallelism for the U-Net engine: one process per GPU, batch sharded
across ranks, ONE all-reduce of the flat gradient buffer per step.

The reference has no distributed code (util.py:28-29 hard-wires cuda:0); this is
new.  BatchNorm uses per-rank batch statistics (the DistributedDataParallel
default); running statistics are taken from rank 0 when `sync_buffers` is called.
"""
import torch
import torch.distributed as dist


def allreduce_mean_(flat, group=None):
    """In-place average of a flat gradient buffer over the process group."""
    if not dist.is_initialized():
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


def broadcast_state(net, src=0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(net.parameters()) + list(net.buffers()):
        dist.broadcast(t.data, src=src, group=group)
    if hasattr(net, "mark_weights_changed"):
        net.mark_weights_changed()          # `.data` writes do not bump version counters


def sync_buffers(net, src=0, group=None):
    """BN running statistics from rank `src` (what DDP's broadcast_buffers does)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in net.buffers():
        dist.broadcast(t.data, src=src, group=group)


def shard_batch(batch_size, rank=None, world=None):
    """[start, stop) of this rank's images in a global batch (contiguous shards)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    per = (batch_size + world - 1) // world
    start = min(rank * per, batch_size)
    return start, min(start + per, batch_size)


def data_parallel(net, group=None, broadcast=True):
    """Attach the single gradient all-reduce to `net` (a deepfluorolabeling UNet):
    after the engine's backward fills the flat fp32 gradient buffer, it is averaged
    over ranks once, before autograd hands the per-parameter views to the optimiser."""
    if broadcast:
        broadcast_state(net, 0, group)
    net.grad_hook = lambda flat: allreduce_mean_(flat, group)
    return net
