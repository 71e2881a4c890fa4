"""Host logic of the configs[3] harness (tools/train_loop.py): the SGDR schedule restatement against the reference's
own warm_restarts_lr.WarmRestartLR (oracle/_ref), the per-rank sharding of global batches, the dataset schema."""
import os
import sys

import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import train_loop as TL  # noqa: E402


def test_synthetic_store_has_the_reference_hdf5_schema():
    st = TL.make_synthetic_store(num_specimens=2, projs_per_specimen=5, tile=20)
    assert int(st["land-names/num-lands"]) == 14 and "land-names/land-13" in st          # hdf5_layouts/Readme.md:105-110
    for s in ("01", "02"):
        assert st[f"{s}/projs"].shape == (5, 20, 20) and st[f"{s}/projs"].dtype.name == "float32"
        assert st[f"{s}/segs"].shape == (5, 20, 20) and st[f"{s}/segs"].dtype.name == "uint8"
        assert st[f"{s}/lands"].shape == (5, 2, 14)
    ds = TL.SyntheticTileDataset(st, [1, 2])
    assert len(ds) == 10 and ds[3][0].shape == (20, 20) and ds[3][2].shape == (2, 14)


def test_shard_sampler_partitions_every_global_batch():
    n, b, world = 103, 4, 3
    per_rank = []
    for r in range(world):
        s = TL.ShardSampler(n, b, r, world, seed=1)
        s.set_epoch(2)
        per_rank.append(list(s))
    assert all(len(p) == (n // (b * world)) * b for p in per_rank)
    flat = [i for p in per_rank for i in p]
    assert len(set(flat)) == len(flat)                                   # ranks never see the same item
    # global batch k = the k-th slices of every rank, and together they are a contiguous slice of one permutation
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1 + 2)).tolist()
    for k in range(n // (b * world)):
        got = [i for p in per_rank for i in p[k * b:(k + 1) * b]]
        assert got == perm[k * b * world:(k + 1) * b * world]


def test_warm_restart_schedule_matches_the_reference_class():
    from oracle import build_ref
    ref = build_ref.load()
    if ref is None or "warm_restarts_lr" not in ref:
        pytest.skip("oracle/_ref not built on this box")
    w = torch.nn.Parameter(torch.zeros(3))
    opt_a = torch.optim.SGD([w], lr=0.1)
    opt_b = torch.optim.SGD([w], lr=0.1)
    a = ref["warm_restarts_lr"].WarmRestartLR(opt_a, init_run_period_epochs=2, growth_factor=2)     # train.py:337
    b = TL.WarmRestartLR(opt_b, init_run_period_epochs=2, growth_factor=2)
    for epoch in range(7):
        for i in range(1, 6):
            a.intra_epoch_step(i / 5)
            b.intra_epoch_step(i / 5)
            assert abs(opt_a.param_groups[0]["lr"] - opt_b.param_groups[0]["lr"]) < 1e-12, (epoch, i)
        a.step()
        b.step()
        assert a.just_restarted == b.just_restarted, epoch
        assert abs(opt_a.param_groups[0]["lr"] - opt_b.param_groups[0]["lr"]) < 1e-12, epoch
