"""Numerical checks of the data-parallel path on real GPUs (NCCL, one process per GPU; skipped below 2 GPUs):
  * after several training steps every rank holds bit-identical parameters (overlapped two-bucket all-reduce,
    eager and CUDA-graph replay);
  * without BatchNorm (whose statistics are per rank by design) the data-parallel run equals a
    single-process run on the concatenated batch;
  * the overlapped two-bucket reduction equals the single all-reduce of the whole buffer;
  * the process group is destroyed cleanly while captured graphs that contain NCCL kernels have been released.
"""
import os
import socket

import pytest
import torch

from conftest import load_pkg

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(pkg, kw, precision, dev, seed=0):
    torch.manual_seed(seed)
    return pkg.UNet(precision=precision, **kw).to(dev).train()


def _batch(B, S, T, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, S, S, generator=g)
    ts = torch.nn.functional.one_hot(torch.randint(0, 7, (B, T, T), generator=g), 7).permute(0, 3, 1, 2).float().contiguous()
    th = torch.rand(B, 14, T, T, generator=g)
    return x, ts, th


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pkg = load_pkg()
    B, S, T = 4, 64, 56
    kw = dict(n_classes=7, depth=4, wf=5, batch_norm=False, padding=True, max_pool=False, num_lands=14)

    def train(net, batches, graph=False, lr=0.05):
        opt = torch.optim.SGD(net.parameters(), lr=lr, momentum=0.9, nesterov=True, fused=True)
        crit = pkg.FusedDiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)

        def step(x, ts, th):
            opt.zero_grad(set_to_none=True)
            seg, heat = net(x)
            loss = crit((seg, heat), (ts, th))
            loss.backward()
            opt.step()
            return loss
        call = pkg.GraphedStep(step, batches[0], warmup=1, allow_distributed=True, modules=[net]) if graph else step
        for b in batches:
            call(*b)
        torch.cuda.synchronize()
        if graph:
            del call
        return torch.cat([p.detach().flatten() for p in net.parameters()])

    # every rank trains on its own shard of a global batch of 2B; 3 steps
    full = [_batch(world * B, S, T, 10 + i) for i in range(3)]
    shard = [tuple(t[rank * B:(rank + 1) * B].to(dev) for t in b) for b in full]

    results = {}
    for name, overlap, graph, precision in (("overlap", True, False, "parity_tc"), ("single", False, False, "parity_tc"),
                                            ("overlap_graph_bf16", True, True, "bf16")):
        net = _make(pkg, kw if precision != "bf16" else dict(kw, batch_norm=True), precision, dev)
        pkg.parallel.data_parallel(net, overlap=overlap)
        assert (net.grad_bucket_hook is not None) == overlap
        flat = train(net, shard, graph=graph)
        # bit-identical parameters on every rank
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, flat), f"{name}: rank {rank} parameters differ from rank 0"
        results[name] = flat
        if overlap:
            assert 0 < net.early_grad_numel < flat.numel() + 64
    # two buckets == one all-reduce (NCCL may chunk the rings differently: rounding-level differences only)
    d = float((results["overlap"] - results["single"]).norm() / results["single"].norm())
    assert d < 1e-6, d
    # data parallel == single process on the concatenated batch (no BatchNorm; mean-reduced losses)
    if rank == 0:
        net = _make(pkg, kw, "parity_tc", dev)
        big = [tuple(t.to(dev) for t in b) for b in full]
        flat1 = train(net, big)
        d1 = float((results["overlap"] - flat1).norm() / flat1.norm())
        assert d1 < 2e-5, d1
        with open(os.path.join(out_dir, "ok"), "w") as f:
            f.write(f"{d} {d1}")
    dist.barrier()
    dist.destroy_process_group()          # must return (graphs holding NCCL kernels were released above)


def test_data_parallel_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(tmp_path, "ok"))
