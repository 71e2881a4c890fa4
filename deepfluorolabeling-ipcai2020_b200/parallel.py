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


def data_parallel(net, group=None, broadcast=True, overlap=True):
    """Attach the gradient all-reduce to `net` (a deepfluorolabeling UNet).

    overlap=False: ONE all-reduce of the flat fp32 gradient buffer after the engine's backward has filled it.
    overlap=True (default on CUDA/NCCL): the buffer is reduced as two buckets.  The engine lays the gradients out so
    that flat[:early] -- heads, decoder, deep encoder levels: 97 % of the paper network's parameters -- is final when
    the backward pass still has the shallow encoder levels to run (its most expensive 40 %), and calls back at that
    point; the early bucket's all-reduce is enqueued on a communication stream right there and runs beside the rest of
    the backward pass.  The small tail bucket follows after backward, and the compute stream waits for the
    communication stream before the gradients are handed to autograd.  Same arithmetic either way (NCCL AVG per
    element).  Under CUDA-graph capture the communication stream becomes a parallel branch of the graph."""
    if broadcast:
        broadcast_state(net, 0, group)
    use_overlap = (overlap and dist.is_initialized() and dist.get_world_size(group) > 1
                   and dist.get_backend(group) == "nccl" and next(net.parameters()).is_cuda)
    if not use_overlap:
        net.grad_bucket_hook = None
        net.grad_hook = lambda flat: allreduce_mean_(flat, group)
        return net
    dev = next(net.parameters()).device
    comm = torch.cuda.Stream(device=dev)
    state = {"early": 0}

    def bucket_hook(flat, offset, numel):          # runs on `comm` (made current by the module), mid-backward
        state["early"] = offset + numel
        allreduce_mean_(flat[offset:offset + numel], group)

    def tail_hook(flat):                           # runs on the compute stream after fu_backward returned
        early = state["early"]
        state["early"] = 0
        if early < flat.numel():
            allreduce_mean_(flat[early:] if early else flat, group)     # tail bucket (or everything, if no callback came)
        torch.cuda.current_stream(dev).wait_stream(comm)                # the early bucket's all-reduce has landed

    net.bucket_stream = comm
    net.grad_bucket_hook = bucket_hook
    net.grad_hook = tail_hook
    return net
