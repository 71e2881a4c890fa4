"""BASELINE.json configs[3]: the train.py loop on a synthetic dataset with the reference's HDF5 schema.

The reference's train.py / dataset.py cannot be imported here (h5py is absent from the image and there is no network),
so this harness restates the LOOP -- not the model -- around the drop-in network, line by line:

  * dataset: an in-memory store with the layout of hdf5_layouts/Readme.md:105-117 (`land-names/num-lands`,
    `land-XX`, per specimen `NN/projs` f32 N x 180 x 180, `NN/segs` u1, `NN/lands` f32 N x 2 x 14) behind a
    torch.utils.data.Dataset whose items are what dataset.py:91-109 starts from: the raw tile, its label map and its
    landmark coordinates.  Reflect padding + z-score (dataset.py:287-293), heat-map targets (:295-325) and one-hot
    masks (:448-452) are finished on the device by prepost.py, per batch;
  * DataLoader(batch_size, shuffle=True, num_workers=0): train.py:293-296 without --data-aug;
  * per iteration (train.py:391-443): host -> device copies, zero_grad, forward, centre crop + Dice/NCC loss, backward,
    SGD(momentum 0.9, nesterov, weight decay) step, WarmRestartLR.intra_epoch_step, loss.item();
  * per epoch: WarmRestartLR.step (train.py:456-464) and a save_net-format checkpoint every `checkpoint_every`
    epochs (train.py:473-515), written to a temporary file and moved into place.

Under torch.distributed every rank owns 1/world of every global batch (a DistributedSampler-style strided shard of the
shuffled index list) and the gradient all-reduce of parallel.data_parallel runs inside the step.

    torchrun ... bench.py --train-loop --gpus 8      (one JSON line like every other bench run)
    python tools/train_loop.py                       (a short single-GPU run that prints the loss trajectory)
"""
import math
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ---------------------------------------------------------------------------------------------------------------
# synthetic dataset with the reference's HDF5 layout (hdf5_layouts/Readme.md:105-117)
# ---------------------------------------------------------------------------------------------------------------
def make_synthetic_store(num_specimens=2, projs_per_specimen=256, tile=180, num_lands=14, num_classes=7, seed=0):
    """{hdf5 path: array}: what h5py.File(...)[path][:] would return for a file written by the reference's tools."""
    rng = np.random.default_rng(seed)
    store = {"land-names/num-lands": np.int64(num_lands)}
    for l in range(num_lands):
        store[f"land-names/land-{l:02d}"] = f"synthetic-{l:02d}"
    for s in range(num_specimens):
        n = projs_per_specimen
        store[f"{s + 1:02d}/projs"] = (rng.random((n, tile, tile), dtype=np.float32) * 4000.0)
        store[f"{s + 1:02d}/segs"] = rng.integers(0, num_classes, (n, tile, tile), dtype=np.uint8)
        store[f"{s + 1:02d}/lands"] = (rng.random((n, 2, num_lands), dtype=np.float32) * (tile - 1))
    return store


class SyntheticTileDataset(torch.utils.data.Dataset):
    """Items as dataset.py:91-109 reads them from the file, before padding / normalisation / target synthesis."""

    def __init__(self, store, specimens):
        self.projs = torch.from_numpy(np.concatenate([store[f"{s:02d}/projs"] for s in specimens]))
        self.segs = torch.from_numpy(np.concatenate([store[f"{s:02d}/segs"] for s in specimens]))
        self.lands = torch.from_numpy(np.concatenate([store[f"{s:02d}/lands"] for s in specimens]))

    def __len__(self):
        return self.projs.shape[0]

    def __getitem__(self, i):
        return self.projs[i], self.segs[i], self.lands[i]


class ShardSampler(torch.utils.data.Sampler):
    """Shuffled indices, the same permutation on every rank (seeded per epoch), of which rank r takes every
    world-th GLOBAL batch slot: global batch b consists of perm[b*G:(b+1)*G] and rank r owns its r-th per-rank slice."""

    def __init__(self, n, per_rank_batch, rank, world, seed=0):
        self.n, self.b, self.rank, self.world, self.seed, self.epoch = n, per_rank_batch, rank, world, seed, 0

    def set_epoch(self, e):
        self.epoch = e

    def __len__(self):
        g = self.b * self.world
        return (self.n // g) * self.b

    def __iter__(self):
        gen = torch.Generator().manual_seed(self.seed + self.epoch)
        perm = torch.randperm(self.n, generator=gen).tolist()
        g = self.b * self.world
        for s in range(0, (self.n // g) * g, g):
            yield from perm[s + self.rank * self.b: s + (self.rank + 1) * self.b]


# ---------------------------------------------------------------------------------------------------------------
# SGDR schedule of warm_restarts_lr.py:14-63, restated for a learning rate that lives in a device tensor (a captured
# optimizer step reads the tensor; a Python float would be frozen into the graph)
# ---------------------------------------------------------------------------------------------------------------
class WarmRestartLR:
    def __init__(self, optimizer, init_run_period_epochs=10, lr_min=0.0, growth_factor=2):
        self.opt = optimizer
        self.base_lrs = [float(g["lr"]) for g in optimizer.param_groups]
        self.cur_run_period_epochs = init_run_period_epochs          # warm_restarts_lr.py:16
        self.lr_min = lr_min
        self.next_restart_epoch = init_run_period_epochs             # :20
        self.last_restart_epoch = 0                                  # :22
        self.period_growth_factor = growth_factor
        self.cur_epoch_ratio = 0.0
        self.last_epoch = 0
        self.just_restarted = False

    def get_lr(self):                                                # :55-62
        shift_cos = 1 + math.cos(math.pi * (self.last_epoch - self.last_restart_epoch + self.cur_epoch_ratio)
                                 / self.cur_run_period_epochs)
        return [self.lr_min + ((b - self.lr_min) / 2) * shift_cos for b in self.base_lrs]

    def _apply(self):
        for g, lr in zip(self.opt.param_groups, self.get_lr()):
            if isinstance(g["lr"], torch.Tensor):
                g["lr"].fill_(lr)
            else:
                g["lr"] = lr

    def intra_epoch_step(self, epoch_ratio):                         # :32-36
        self.cur_epoch_ratio = epoch_ratio
        self._apply()

    def step(self):                                                  # :38-53
        self.cur_epoch_ratio = 0.0
        self.last_epoch += 1
        self._apply()          # (the base class applies get_lr() BEFORE the restart bookkeeping below, :41)
        if self.last_epoch >= self.next_restart_epoch:
            self.last_restart_epoch = self.next_restart_epoch
            self.cur_run_period_epochs *= self.period_growth_factor
            self.next_restart_epoch += self.cur_run_period_epochs
            self.just_restarted = True
        else:
            self.just_restarted = False

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "opt"}


def save_net(path, epoch, net, optimizer, lr_sched, loss, cfg):
    """train.py:473-515: the same dictionary keys, written to a temporary name and moved into place."""
    tmp = f"{path}.tmp"
    torch.save({'epoch': epoch, 'model-state-dict': net.state_dict(), 'optim-type': 'sgd',
                'optimizer-state-dict': optimizer.state_dict(), 'scheduler-state-dict': lr_sched.state_dict(),
                'loss': loss, 'best-valid-loss': None, 'save-best-valid': False, 'num-classes': 7, 'depth': 6,
                'init-feats-exp': 5, 'batch-norm': True, 'padding': True, 'no-max-pool': True, 'pad-img-size': cfg["size"],
                'batch-size': cfg["batch"], 'data-aug': False, 'opt-nesterov': True, 'opt-momentum': 0.9,
                'opt-wgt-decay': 1e-4, 'num-lands': 14, 'heat-coeff': 0.5, 'use-dice-valid': False, 'unet-use-res': True,
                'unet-block-depth': 2, 'lrs-meth': 'cos', 'lrs-num-epochs': cfg["lrs_epochs"], 'lrs-growth-factor': 2,
                'lrs-max-num-restarts': -1, 'lrs-save-restart-net-prefix': '', 'lrs-save-after-n-restarts': 0,
                'lrs-num-restarts': 0, 'lrs-patience': 10, 'lrs-cooldown': 10, 'checkpoint-freq': cfg["checkpoint_every"],
                'train-idx': None, 'valid-idx': None}, tmp)
    shutil.move(tmp, path)


class TrainLoop:
    """The loop of train.py:376-443 around the drop-in network."""

    def __init__(self, pkg, dev, rank=0, world=1, batch=32, size=192, tile=180, precision="bf16", graph=True,
                 projs_per_specimen=256, lrs_epochs=4, checkpoint_every=1, ckpt_dir=None):
        self.pkg, self.dev, self.rank, self.world = pkg, dev, rank, world
        self.cfg = dict(batch=batch, size=size, tile=tile, lrs_epochs=lrs_epochs, checkpoint_every=checkpoint_every)
        store = make_synthetic_store(2, projs_per_specimen, tile)
        self.ds = SyntheticTileDataset(store, [1, 2])                                   # train.py:277-290 (two specimens)
        self.sampler = ShardSampler(len(self.ds), batch, rank, world)
        self.dl = torch.utils.data.DataLoader(self.ds, batch_size=batch, sampler=self.sampler, num_workers=0,
                                              pin_memory=True, drop_last=True)          # train.py:293-296
        torch.manual_seed(0)
        paper = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14)
        self.net = pkg.UNet(precision=precision, **paper).to(dev)                       # train.py:313-319
        if world > 1:
            pkg.parallel.data_parallel(self.net)
        self.crit = pkg.FusedDiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)       # train.py:324
        self.opt = torch.optim.SGD(self.net.parameters(), lr=torch.tensor(0.1, device=dev), momentum=0.9, weight_decay=1e-4,
                                   nesterov=True, fused=True)                           # train.py:333-334
        self.sched = WarmRestartLR(self.opt, init_run_period_epochs=lrs_epochs)         # train.py:337
        self.mask_buf = torch.empty(batch, 7, tile, tile, device=dev)
        self.epoch, self.graph, self.gstep = 0, graph, None
        self.ckpt_dir = ckpt_dir
        self.h2d_bytes = 0

    def _step(self, x, mask, heat):
        self.opt.zero_grad(set_to_none=True)                                            # train.py:405
        seg, hm = self.net(x)                                                           # train.py:407
        loss = self.crit((seg, hm), (mask, heat))                                       # train.py:414-420 (crop folded in)
        loss.backward()                                                                 # train.py:422
        self.opt.step()                                                                 # train.py:424
        return loss

    def iteration(self, raw, segs, lands):
        """One pass of the body of train.py:391-443 on a batch the DataLoader produced (pinned host tensors)."""
        pp, T, S = self.pkg.prepost, self.cfg["tile"], self.cfg["size"]
        raw, segs, lands = (t.to(self.dev, non_blocking=True) for t in (raw, segs, lands))     # train.py:395-403
        x = pp.prep_tiles(raw, pad_img_dim=S)                                           # dataset.py:287-293
        heat = pp.heatmap_targets(lands, (T, T))                                        # dataset.py:295-325
        self.mask_buf.zero_().scatter_(1, segs.long().unsqueeze(1), 1.0)                # dataset.py:448-452
        if self.graph and self.gstep is None:
            self.gstep = self.pkg.GraphedStep(self._step, (x, self.mask_buf, heat), warmup=2,
                                              allow_distributed=self.world > 1, modules=[self.net])
        loss = (self.gstep or self._step)(x, self.mask_buf, heat)
        return loss

    def run_epoch(self, max_iters=None, on_iter=None):
        self.net.train()                                                                # train.py:381
        self.sampler.set_epoch(self.epoch)
        n_ds, seen, losses = len(self.sampler), 0, []
        for i, (raw, segs, lands) in enumerate(self.dl):                                # train.py:391
            if max_iters is not None and i >= max_iters:
                break
            loss = self.iteration(raw, segs, lands)
            seen += raw.shape[0]
            self.sched.intra_epoch_step(min(1.0, seen / n_ds))                          # train.py:427-428
            losses.append(loss.item())                                                  # train.py:430
            if on_iter is not None:
                on_iter(i, losses[-1])
        self.sched.step()                                                               # train.py:456-464
        self.epoch += 1
        if self.ckpt_dir and self.rank == 0 and self.epoch % self.cfg["checkpoint_every"] == 0:
            save_net(os.path.join(self.ckpt_dir, "checkpoint.pt"), self.epoch, self.net, self.opt, self.sched,
                     torch.tensor(losses[-1] if losses else 0.0), self.cfg)             # train.py:517-520
        return losses

    def close(self):
        self.gstep = None
        self.net._destroy_engine()


def bench_train_loop(args, torch_mod, dist, pkg, dev, rank, world, local, config_dict, metric):
    """bench.py --train-loop: W warm-up iterations, then K timed iterations of the loop (DataLoader fetch + H2D + device
    sample preparation + graph-replayed step + scheduler + loss.item()), with a save_net checkpoint written inside the
    timed region when an epoch ends.  Returns the JSON line (rank 0) -- `value` and `e2e` are the same number here:
    every iteration starts from host data and ends with the loss on the host."""
    ckpt_dir = tempfile.mkdtemp(prefix="fu_ckpt_") if rank == 0 else None
    per_spec = max(256, (args.steps + args.warmup + 4) * args.batch * world // 2 + args.batch * world)
    loop = TrainLoop(pkg, dev, rank, world, batch=args.batch, size=args.size, tile=args.tile, precision=args.precision,
                     graph=not args.no_graph, projs_per_specimen=per_spec, ckpt_dir=ckpt_dir)
    it = iter(loop.dl)
    loop.net.train()
    n_ds, seen = len(loop.sampler), 0
    h2d = None

    def one():
        nonlocal seen, h2d
        raw, segs, lands = next(it)
        if h2d is None:
            h2d = sum(t.numel() * t.element_size() for t in (raw, segs, lands))
        loss = loop.iteration(raw, segs, lands)
        seen += raw.shape[0]
        loop.sched.intra_epoch_step(min(1.0, seen / n_ds))
        return loss.item()

    for _ in range(args.warmup):
        one()

    def barrier():
        if world > 1:
            dist.barrier()
        torch_mod.cuda.synchronize()
    barrier()
    e0, e1 = torch_mod.cuda.Event(enable_timing=True), torch_mod.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    losses = [one() for _ in range(args.steps)]
    loop.sched.step()
    ck0 = time.perf_counter()
    if rank == 0:                                   # the epoch's checkpoint (train.py:517-520), inside the timed region
        save_net(os.path.join(ckpt_dir, "checkpoint.pt"), 1, loop.net, loop.opt, loop.sched, torch_mod.tensor(losses[-1]), loop.cfg)
    ckpt_ms = (time.perf_counter() - ck0) * 1e3
    e1.record()
    torch_mod.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch_mod.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    barrier()
    ckpt_bytes = os.path.getsize(os.path.join(ckpt_dir, "checkpoint.pt")) if rank == 0 else 0
    if ckpt_dir:
        shutil.rmtree(ckpt_dir, ignore_errors=True)
    cnt = loop.net.engine_counters()
    # engine launches of one step (counted on the last eagerly launched step, which the graph replays) + fused loss (3)
    # + sample preparation (prep_stats, prep_apply, heat-map targets)
    launches = (cnt["last_fwd_launches"] + cnt["last_bwd_launches"] + 6) * args.steps
    loop.close()
    if rank != 0:
        return None
    ms_step = ms / args.steps
    value = args.batch * world / (ms_step * 1e-3)
    cfg = config_dict(args, world, loop.graph, "raw tiles + u1 labels + landmark coordinates from a DataLoader over a synthetic "
                                               "HDF5-schema dataset; pad / z-score / heat-maps / one-hot on the device")
    cfg["workload"] = ("train.py loop (train.py:376-443) on a synthetic dataset with the HDF5 schema of hdf5_layouts/Readme.md:105-117: "
                       "DataLoader(shuffle, num_workers=0) -> H2D -> device sample prep -> " + cfg["workload"] +
                       " -> WarmRestartLR.intra_epoch_step -> loss.item(); one save_net checkpoint (%.0f MB) inside the timed region"
                       % (ckpt_bytes / 1e6))
    return {"metric": metric, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": int(h2d or 0), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_step},
            "checkpoint_ms": ckpt_ms, "images_per_s_without_checkpoint": args.batch * world * args.steps / max(1e-9, (ms - ckpt_ms) * 1e-3),
            "gpu_launches": int(launches), "loss_first_last": [losses[0], losses[-1]], "host_wall_ms_per_step": wall / args.steps,
            "build": pkg._capi.lib().fu_build_info().decode()}


if __name__ == "__main__":
    import importlib
    pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
    dev = torch.device("cuda:0")
    loop = TrainLoop(pkg, dev, projs_per_specimen=128, ckpt_dir=tempfile.mkdtemp(prefix="fu_ckpt_"))
    for ep in range(2):
        ls = loop.run_epoch()
        print(f"epoch {ep}: {len(ls)} iterations, loss {ls[0]:.4f} -> {ls[-1]:.4f}, lr {float(loop.opt.param_groups[0]['lr']):.4f}")
    print("checkpoint:", os.listdir(loop.ckpt_dir))
