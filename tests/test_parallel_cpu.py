"""Host-side logic of the data-parallel path (one process per GPU, one gradient all-reduce per step),
exercised with world_size=2 over gloo on CPU: the flat-gradient averaging, the state broadcast and the
batch sharding.  The engine itself is not involved (no GPU here)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import importlib
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module(PKG)
    par = pkg.parallel
    # 1. flat gradient averaging (what net.grad_hook does after fu_backward)
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    par.allreduce_mean_(flat)
    ok1 = torch.allclose(flat, torch.arange(10, dtype=torch.float32) * (1 + 2) / 2)
    # 2. state broadcast from rank 0 (parameters and BN buffers)
    torch.manual_seed(100 + rank)
    net = pkg.UNet(n_classes=3, depth=2, wf=2, batch_norm=True, padding=True, max_pool=False, num_lands=2)
    with torch.no_grad():
        net.down_path[0].block[2].running_mean.fill_(float(rank + 1))
    par.data_parallel(net)
    sums = torch.tensor([float(sum(p.double().sum() for p in net.parameters())),
                         float(net.down_path[0].block[2].running_mean.sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    ok2 = all(torch.equal(g, gathered[0]) for g in gathered) and float(gathered[0][1]) == 4.0
    # 3. the hook installed by data_parallel averages in place
    g = torch.full((5,), float(rank))
    net.grad_hook(g)
    ok3 = torch.allclose(g, torch.full((5,), 0.5))
    # 4. contiguous batch shards cover the batch exactly once
    s = par.shard_batch(7)
    all_s = [None] * world
    dist.all_gather_object(all_s, s)
    ok4 = all_s == [(0, 4), (4, 7)]
    q.put((rank, ok1, ok2, ok3, ok4))
    dist.destroy_process_group()


def test_data_parallel_host_logic_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r


def test_shard_batch_without_process_group():
    import importlib
    par = importlib.import_module(PKG).parallel
    assert par.shard_batch(32, rank=0, world=1) == (0, 32)
    assert par.shard_batch(256, rank=3, world=8) == (96, 128)
    assert par.shard_batch(5, rank=2, world=4) == (4, 5)
    assert par.shard_batch(5, rank=3, world=4) == (5, 5)
    t = torch.ones(3)
    assert par.allreduce_mean_(t) is t
