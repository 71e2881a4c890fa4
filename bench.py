#!/usr/bin/env python
"""Benchmark of the U-Net hot path: images/s of one training step (fwd + Dice/NCC loss + bwd + SGD)
of the paper's dual-head network on synthetic 1x180x180 fluoroscopy tiles (reflect-padded to 192),
per-GPU batch 32 (BASELINE.json configs[1]); one process per GPU, NCCL gradient all-reduce for N>1.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Besides the headline workload the line carries
  * `parity_tc`: the same step in the tensor-core parity mode (fp32 storage, split-bf16 x3 contractions);
  * `other_configs`: BASELINE.json configs[2] (736^2, B=8, seg-only) and configs[4] (1440^2, 2 tiles per GPU,
    heat-map weight 1.0), each with its own images/s, e2e, roofline fraction and clocks (N=1; at N=2 the
    1440^2 configuration runs as the 2 x 2 split the baseline names).
`--impl reference` times the reference's own CPU path on the host cores: the unmodified train_test_code/unet.py
+ dice.py byte-compiled into oracle/_ref by oracle/build_ref.py (kind "reference"), or the oracle port when
oracle/_ref is absent (kind "port"), on the SAME configuration (32 tiles per step).
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "deepfluorolabeling-ipcai2020_b200"

PAPER = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14,
             do_res=True, block_depth=2)
# algorithmic conv/convT FLOPs per image (MACs x 2), SURVEY.md 8d
GF_PER_IMG = {192: 54.475, 736: 796.856, 1440: 3064.196}
METRIC = "images/sec fwd+bwd U-Net (1x180x180, 7+14 heads)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=192, help="network input size (tile 180 padded to 192)")
    ap.add_argument("--tile", type=int, default=180)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "parity_tc"])
    ap.add_argument("--cpu-sample-batch", type=int, default=32,
                    help="tiles per CPU-baseline step (default: the same 32 as the GPU arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity_tc and other_configs measurements")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch every step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--torch-loss", action="store_true",
                    help="use the PyTorch DiceAndHeatMapLoss2D on cropped views instead of the fused device loss")
    ap.add_argument("--loss-in-heads", action="store_true",
                    help="the loss inside the heads kernels (UNet.forward_loss; bf16 only).  Default: the fused device loss as separate "
                         "kernels after the network, which measures faster (DESIGN.md section 5)")
    ap.add_argument("--seg-only", action="store_true",
                    help="num_lands=0 network with DiceLoss2D (train.py:327; BASELINE configs[2]: --batch 8 --size 736 --tile 718)")
    ap.add_argument("--heatmap-wgt", type=float, default=0.5,
                    help="train.py --heat-coeff (BASELINE configs[4] uses 1.0: --batch 2 --size 1440 --tile 1436)")
    ap.add_argument("--host-prep", action="store_true",
                    help="e2e leg ships finished fp32 tiles / masks / heat-maps from the host (91.8 MB per step) instead of raw "
                         "tiles, landmark coordinates and u1 labels finished on the device (5.2 MB per step, the default)")
    ap.add_argument("--train-loop", action="store_true",
                    help="run the step inside the train.py:376-443 loop shape: DataLoader over a synthetic HDF5-schema dataset, "
                         "WarmRestartLR cosine schedule, loss.item() every step (BASELINE configs[3]; tools/train_loop.py)")
    return ap.parse_args()


def make_targets(B, tile, n_classes, n_lands, gen, torch):
    """Targets shaped like dataset.py's: one-hot float masks (dataset.py:448-452) and Gaussian
    heat-maps, sigma 2.5, peak 1/(2 pi sigma^2) at random in-bounds pixels (dataset.py:295-325)."""
    labels = torch.randint(0, n_classes, (B, tile, tile), generator=gen)
    mask = torch.nn.functional.one_hot(labels, n_classes).permute(0, 3, 1, 2).contiguous().float()
    sigma = 2.5
    ys = torch.arange(tile, dtype=torch.float32).view(1, 1, tile, 1)
    xs = torch.arange(tile, dtype=torch.float32).view(1, 1, 1, tile)
    cy = torch.randint(0, tile, (B, n_lands, 1, 1), generator=gen).float()
    cx = torch.randint(0, tile, (B, n_lands, 1, 1), generator=gen).float()
    heat = torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * sigma * sigma)) / (2 * math.pi * sigma * sigma)
    return mask, heat.contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        # median over the busiest half of the samples (the sampler also sees the idle edges)
        busy = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


# ---------------------------------------------------------------------------------------------------------------
# the reference's CPU path (the reference arm and the cpu_baseline leg; the only places bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(torch, steps, warmup, batch, size, tile, threads, seg_only=False, heatmap_wgt=0.5, budget_s=None):
    """One training step of train.py:405-430 on the host cores.  kind "reference": the unmodified reference modules
    (oracle/_ref, byte-compiled unet.py / dice.py / util.py); kind "port": the oracle restatement of unet.py over the
    same ATen primitives.  Returns (images/s, ms/step, kind, steps actually timed)."""
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    from oracle import build_ref
    ref = build_ref.load()
    cfg_kw = dict(PAPER, num_lands=0) if seg_only else PAPER
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 1, size, size, generator=g)
    mask, heat = make_targets(batch, tile, 7, 14, g, torch)
    if ref is not None:
        kind = "reference"
        net = ref["unet"].UNet(**cfg_kw)                                                     # train.py:313
        net.train()                                                                          # train.py:381
        crit = (ref["dice"].DiceLoss2D(skip_bg=False) if seg_only else                       # train.py:324-327
                ref["dice"].DiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=heatmap_wgt))
        crop = ref["util"].center_crop
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True)   # train.py:333-334

        def step():
            opt.zero_grad()                                                                  # train.py:405
            out = net(x)                                                                     # train.py:407
            if seg_only:
                loss = crit(crop(out, mask.shape), mask)                                     # train.py:414-420
            else:
                loss = crit((crop(out[0], mask.shape), crop(out[1], heat.shape)), (mask, heat))
            loss.backward()                                                                  # train.py:422
            opt.step()                                                                       # train.py:424
            return loss.item()                                                               # train.py:430
    else:
        kind = "port"
        from oracle import unet_oracle as O
        pkg = importlib.import_module(PKG)
        net = pkg.UNet(precision="fp32", **cfg_kw)          # parameter container only (CPU); same init as the reference
        cfg = O.UNetConfig(**cfg_kw)
        O.NATIVE_BN = True      # same fused ATen batch_norm the reference's nn.BatchNorm2d calls
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        leaves = []
        for k, _, kd in O.param_schema(cfg):
            if kd == "param":
                sd[k].requires_grad_(True)
                leaves.append(sd[k])
        opt = torch.optim.SGD(leaves, lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True)

        def step():
            opt.zero_grad()
            out = O.forward(sd, cfg, x, training=True)
            if seg_only:
                loss = O.dice_loss(O.center_crop(out["seg"], mask.shape), mask)
            else:
                loss = O.dice_and_heatmap_loss(O.center_crop(out["seg"], mask.shape), O.center_crop(out["heat"], heat.shape),
                                               mask, heat, skip_bg=False, heatmap_wgt=heatmap_wgt)
            loss.backward()
            opt.step()
            return loss.item()

    t0 = time.perf_counter()
    for _ in range(max(1, warmup)):
        step()
    warm_s = (time.perf_counter() - t0) / max(1, warmup)
    if budget_s is not None:        # bounded sample: as many steps as fit the budget (at least 2)
        steps = max(2, min(steps, int(budget_s / max(warm_s, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3, kind, steps


def workload_text(batch, size, tile, seg_only, heatmap_wgt):
    if seg_only:
        heads, loss_name = "7-class seg only, num_lands=0", "DiceLoss2D"
    else:
        heads, loss_name = "7-class seg + 14 heat-maps", "DiceAndHeatMapLoss2D"
        if heatmap_wgt != 0.5:
            loss_name += f"(heatmap_wgt={heatmap_wgt})"
    return (f"paper dual-head U-Net (depth 6, wf 5, BN, learned 2x2/s2 downsample, res 1x1; {heads}), "
            f"{batch} tiles/GPU of 1x{tile}x{tile} reflect-padded to {size}x{size}, "
            f"train step = fwd + {loss_name} + bwd + SGD(nesterov)")


def config_dict(args, world, graphed=False, e2e_inputs=""):
    return {"workload": workload_text(args.batch, args.size, args.tile, args.seg_only, args.heatmap_wgt),
            "per_gpu_batch": args.batch, "global_batch": args.batch * world, "net_input": args.size, "tile": args.tile,
            "parallelism": f"dp{world}", "precision": args.precision,
            "loss": "torch DiceAndHeatMapLoss2D on cropped views" if args.torch_loss
                    else ("fused device DiceAndHeatMapLoss2D (crop folded in)" if (not args.loss_in_heads or args.precision != "bf16")
                          else "DiceAndHeatMapLoss2D inside the heads kernels (UNet.forward_loss: sums from the head kernel, "
                               "gradient formed in the backward head kernel; crop folded in)"),
            "optimizer": "torch.optim.SGD(momentum 0.9, nesterov, wd 1e-4, fused=True)",
            "launch": "one CUDA-graph replay per step (GraphedStep)" if graphed else "eager launches",
            "e2e_inputs": e2e_inputs,
            "l2": "per-step working set (~1.5 GB of NHWC activations + 300 MB of weights/grads) >> 126 MB L2; no flush needed"}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b = args.batch                                   # the SAME per-step sample as our arm's config
    ips, ms, kind, steps = cpu_reference_run(torch, min(args.steps, 10), 1, b, args.size, args.tile, threads,
                                             args.seg_only, args.heatmap_wgt, budget_s=90.0)
    what = ("the unmodified reference (train_test_code/unet.py + dice.py, byte-compiled into oracle/_ref)" if kind == "reference"
            else "the oracle port of unet.py (oracle/_ref not built on this box)")
    line = {"metric": METRIC, "value": ips, "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world), "impl": "reference",
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
                             "sample": f"{steps} train steps of {b} tiles (fwd+loss+bwd+SGD) through {what} on torch-CPU "
                                       f"({threads} threads)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["config"]["precision"] = "fp32 (torch CPU)"
    line["config"]["launch"] = "torch eager on the host"
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def measure(torch, dist, pkg, dev, rank, world, local, *, batch, size, tile, precision, seg_only, heatmap_wgt, steps, warmup,
            graph=True, torch_loss=False, host_prep=False, profile=True, e2e=True, clocks=True, loss_in_heads=False):
    """Times one configuration: device-resident `value`, end-to-end `e2e`, and the roofline of the 3x3 conv family
    from per-kernel CUDA events.  Returns a dict; every rank must call it with the same arguments."""
    torch.manual_seed(0)
    net_cfg = dict(PAPER, num_lands=0) if seg_only else PAPER
    net = pkg.UNet(precision=precision, **net_cfg).to(dev)
    net.train()
    if world > 1:
        pkg.parallel.data_parallel(net)
    # train.py:333-334's optimiser; fused=True is the same update in one multi-tensor kernel
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True, fused=True)
    # train.py:324's loss.  Default: the fused device version (same value and gradient, tests/test_loss_gpu.py),
    # which folds the output crop of train.py:414-417 into its indexing.
    fused_loss = not torch_loss
    in_heads = fused_loss and loss_in_heads and precision == "bf16"      # the loss inside the heads kernels (opt-in)
    if seg_only:
        crit = (pkg.FusedDiceLoss2D if fused_loss else pkg.DiceLoss2D)(skip_bg=False)                     # train.py:327
    else:
        crit = (pkg.FusedDiceAndHeatMapLoss2D if fused_loss else pkg.DiceAndHeatMapLoss2D)(skip_bg=False, heatmap_wgt=heatmap_wgt)
    B, S, T = batch, size, tile
    g = torch.Generator().manual_seed(100 + rank)
    n_host = 2
    host = []
    for _ in range(n_host):
        x = torch.randn(B, 1, S, S, generator=g)
        mask, heat = make_targets(B, T, 7, 14, g, torch)
        host.append(tuple(t.pin_memory() for t in ((x, mask) if seg_only else (x, mask, heat))))
    resident = tuple(t.to(dev) for t in host[0])

    def train_step(x, mask, heat=None):
        opt.zero_grad(set_to_none=True)
        if in_heads:
            loss = net.forward_loss(x, mask if seg_only else (mask, heat), crit)
            loss.backward()
            opt.step()
            return loss
        if seg_only:
            seg = net(x)
            loss = crit(seg, mask) if fused_loss else crit(pkg.center_crop(seg, mask.shape), mask)
            loss.backward()
            opt.step()
            return loss
        seg, hm = net(x)
        if fused_loss:
            loss = crit((seg, hm), (mask, heat))
        else:
            loss = crit((pkg.center_crop(seg, mask.shape), pkg.center_crop(hm, heat.shape)), (mask, heat))
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        barrier()
        return ms

    # ---- one step = one CUDA-graph replay (pkg.GraphedStep); eager when capture fails on any rank ----
    step_call, graphed, launches_per_step, gstep = train_step, False, None, None
    for _ in range(2):
        train_step(*resident)
    ca = net.engine_counters()["kernel_launches"]
    train_step(*resident)
    launches_per_step = net.engine_counters()["kernel_launches"] - ca     # the captured step launches the same kernels
    if graph:
        try:
            gstep = pkg.GraphedStep(train_step, resident, warmup=2, allow_distributed=world > 1, modules=[net])
            step_call, graphed = gstep, True
        except Exception as ex:        # stay correct, say so in the output
            print(f"bench: CUDA-graph capture failed ({type(ex).__name__}: {ex}); running eagerly", file=sys.stderr)
            torch.cuda.synchronize()
        if world > 1:                  # all ranks replay, or none does
            ok = torch.tensor([1.0 if graphed else 0.0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok) < 1.0:
                step_call, graphed = train_step, False

    # ---- device-resident inputs: `value` ----
    for _ in range(warmup):
        step_call(*resident)
    c0 = net.engine_counters()
    sampler = ClockSampler(local)
    if rank == 0 and clocks:
        sampler.start()
    ms = timed(lambda i: step_call(*resident), steps)
    clk = sampler.stop() if (rank == 0 and clocks) else None
    c1 = net.engine_counters()
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    if graphed:
        launches = launches_per_step * steps      # replayed kernels do not pass through the engine's host counters
    if fused_loss:
        launches += 3 * steps       # loss_sums, loss_finalize, loss_backward (stateless entry points, not in the engine's counter)
    ms_step = ms / steps
    res = {"value": B * world / (ms_step * 1e-3), "ms_per_step": ms_step, "graphed": graphed, "clocks": clk,
           "gpu_launches": int(launches), "launches_per_step": int(launches_per_step + (3 if fused_loss else 0))}

    # ---- end to end through the public API: host -> device every step, loss read back every step (train.py:395-430) ----
    if e2e:
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        device_prep = (not host_prep) and (not seg_only)
        if device_prep:
            # the host ships what dataset.py starts from -- raw tiles, landmark coordinates, u1 label maps -- and the
            # device finishes them (prepost.py: reflect pad + z-score dataset.py:287-293, Gaussian heat-map targets
            # dataset.py:295-325, one-hot masks dataset.py:448-452)
            pp = pkg.prepost
            src = []
            for k in range(n_host):
                raw = (torch.rand(B, T, T, generator=g) * 4000.0).pin_memory()
                lands = (torch.rand(B, 2, 14, generator=g) * (T - 1)).pin_memory()
                labels = torch.randint(0, 7, (B, T, T), generator=g).to(torch.uint8).pin_memory()
                src.append((raw, lands, labels))
            mask_buf = torch.empty(B, 7, T, T, device=dev)
            note = ("host ships raw tiles + landmark coordinates + u1 labels; pad/z-score, heat-map targets and one-hot "
                    "masks are made on the device (prepost.py)")
        else:
            src = host
            note = "host ships finished fp32 tiles, one-hot masks and heat-map targets"
        slots = [tuple(torch.empty_like(t, device=dev) for t in src[0]) for _ in range(2)]
        h2d = sum(t.numel() * t.element_size() for t in src[0])

        # device prep + train step as ONE graph (raw tiles in, loss out): the prep kernels ride in the replay and the graph's
        # static inputs are the 5 MB of raw data instead of 92 MB of finished tensors
        full_call = None
        if device_prep:
            def full_step(raw, lands, labels):
                xx = pp.prep_tiles(raw, pad_img_dim=S)
                hh = pp.heatmap_targets(lands, (T, T))
                mask_buf.zero_().scatter_(1, labels.long().unsqueeze(1), 1.0)
                return train_step(xx, mask_buf, hh)
            full_call = full_step
            if graphed:
                try:
                    ex_in = tuple(t.to(dev) for t in src[0])
                    full_call = pkg.GraphedStep(full_step, ex_in, warmup=2, allow_distributed=world > 1, modules=[net])
                    note += "; prep + step replayed as one CUDA graph"
                except Exception as ex:
                    print(f"bench: capture of prep + step failed ({type(ex).__name__}: {ex}); prep runs eagerly", file=sys.stderr)
                    torch.cuda.synchronize()
                    full_call = full_step
                if world > 1:
                    ok = torch.tensor([1.0 if isinstance(full_call, pkg.GraphedStep) else 0.0], device=dev)
                    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                    if float(ok) < 1.0:
                        full_call = full_step
        # the loss of step i travels to a pinned host slot behind the step and is READ while step i + 1 runs: the reference's
        # per-iteration loss.item() (train.py:430) one step late, so the host never idles the GPU between steps
        loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_done = [torch.cuda.Event() for _ in range(2)]
        pending = []

        def prefetch(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                for d, h in zip(slots[s], src[i % n_host]):
                    d.copy_(h, non_blocking=True)
                ready[s].record(copy_stream)

        def launch(i):
            s = i % 2
            if i == 0:
                prefetch(0)
            prefetch(i + 1)                      # the next step's inputs travel while this step computes
            torch.cuda.current_stream().wait_event(ready[s])
            loss = full_call(*slots[s]) if device_prep else step_call(*slots[s])
            freed[s].record(torch.cuda.current_stream())
            return loss

        def read_pending():
            s = pending.pop(0)
            loss_done[s].synchronize()
            return float(loss_host[s])           # device -> host read of a step's result (pinned slot)

        def e2e_step(i, last=None):
            s = i % 2
            loss = launch(i)
            loss_host[s].copy_(loss.detach(), non_blocking=True)
            loss_done[s].record(torch.cuda.current_stream())
            pending.append(s)
            if len(pending) > 1:
                read_pending()                   # the previous step's loss, while this step runs
            if last is not None and i == last:
                while pending:
                    read_pending()               # the final step's loss is read inside the timed region too

        def e2e_step_sync(i):
            return launch(i).item()              # strictly sequential: this step's loss before the next step is issued

        for s in range(2):
            freed[s].record(torch.cuda.current_stream())
        for i in range(2):
            e2e_step(i, last=1)
        e2e_ms = timed(lambda i: e2e_step(i, last=steps - 1), steps) / steps
        for i in range(2):
            e2e_step_sync(i)
        sync_ms = timed(e2e_step_sync, steps) / steps
        # headline: the reference's own sequence (train.py:430: loss.item() of step i before step i + 1 is issued); the pipelined
        # read (every step's loss still read inside the timed region, one step behind the launch) is reported beside it
        res["e2e"] = {"value": B * world / (sync_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": 4, "ms_per_step": sync_ms, "inputs": note,
                      "loss_read": "loss.item() of step i before step i + 1 is issued",
                      "pipelined": {"value": B * world / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                                    "loss_read": "every step, from a pinned host slot, one step behind the launch "
                                                 "(the last one before the timed region ends)"}}

    # ---- per-kernel device time (CUDA events inside the engine) -> roofline of the 3x3 conv family ----
    if profile:
        net.profile(True)
        nprof = 3
        for _ in range(nprof):
            train_step(*resident)
        torch.cuda.synchronize()
        rep = net.profile_report()
        net.profile(False)
        fam = {}
        for r in rep:
            key = r["tag"].split(" ")[0]
            a = fam.setdefault(key, {"ms": 0.0, "flops": 0.0, "launches": 0})
            a["ms"] += r["ms"] / nprof; a["flops"] += r["flops"] / nprof; a["launches"] += r["launches"] / nprof
        conv_ms = sum(v["ms"] for k, v in fam.items() if k.startswith("conv3_"))
        conv_fl = sum(v["flops"] for k, v in fam.items() if k.startswith("conv3_"))
        peak, _, how = peaks()
        ach = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        # DRAM bytes per step of the tensor-core conv kernels from the committed ncu launch list of this workload
        # (tools/ncu_traffic.py; cold-cache per launch, and a superset of the family: it includes the 1x1 / 2x2 layers)
        traffic, traffic_note = None, "no ncu summary committed for this configuration"
        if S == 192 and B == 32 and precision == "bf16":
            for fn in ("r02_ncu_traffic_final.json", "r01_ncu_traffic_final.json"):
                tp = os.path.join(ROOT, "profiles", fn)
                if os.path.exists(tp):
                    try:
                        tj = json.load(open(tp))
                        traffic = tj["tensor_core_conv_kernels_per_step"]["dram_bytes"]
                        traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum per step over tc_conv*/tc_wgrad* launches "
                                        f"(profiles/{fn}; ncu flushes caches per launch; includes the 1x1/2x2 layers)")
                        break
                    except Exception:
                        pass
        res["roofline"] = {"bound": "tensor", "kernel": "3x3 conv family (fwd + dgrad + wgrad, 22 layers x 3)",
                           "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                           "traffic_note": traffic_note, "peak_source": how,
                           "algorithmic_gflop_per_step": conv_fl / 1e9, "kernel_ms_per_step": conv_ms}
        tot = sum(v["ms"] for v in fam.values())
        bd = {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        bd["_engine_total_ms"] = round(tot, 4)
        res["engine_ms_by_family"] = bd
    res["model_tflops"] = res["value"] * GF_PER_IMG.get(S, 0.0) / 1e3
    # release everything this configuration holds before the next one is built
    del gstep, step_call
    net._destroy_engine()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a GPU (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.train_loop:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_loop
        line = train_loop.bench_train_loop(args, torch, dist, pkg, dev, rank, world, local, config_dict, METRIC)
        if rank == 0:
            print(json.dumps(line), flush=True)
        finish(torch, dist, world)
        return

    common = dict(graph=not args.no_graph, torch_loss=args.torch_loss, host_prep=args.host_prep, loss_in_heads=args.loss_in_heads)
    main = measure(torch, dist, pkg, dev, rank, world, local, batch=args.batch, size=args.size, tile=args.tile,
                   precision=args.precision, seg_only=args.seg_only, heatmap_wgt=args.heatmap_wgt, steps=args.steps,
                   warmup=args.warmup, profile=not args.no_profile, **common)
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": config_dict(args, world, main["graphed"], main["e2e"]["inputs"]), "clocks": main["clocks"],
                "gpu_launches": main["gpu_launches"], "e2e": {k: v for k, v in main["e2e"].items() if k != "inputs"},
                "roofline": main.get("roofline"), "engine_ms_by_family": main.get("engine_ms_by_family"),
                "model_tflops": main["model_tflops"],
                "build": pkg._capi.lib().fu_build_info().decode()}

    # ---- the same step in the tensor-core parity mode, and the other BASELINE configurations ----
    default_workload = (args.batch, args.size, args.tile, args.seg_only, args.precision) == (32, 192, 180, False, "bf16")
    if not args.no_extras and default_workload:
        ex_steps, ex_warm = max(3, min(args.steps, 10)), max(3, min(args.warmup, 3))
        if world == 1:
            pt = measure(torch, dist, pkg, dev, rank, world, local, batch=32, size=192, tile=180, precision="parity_tc",
                         seg_only=False, heatmap_wgt=0.5, steps=ex_steps, warmup=ex_warm, profile=not args.no_profile,
                         e2e=False, clocks=False, **common)
            if rank == 0:
                line["parity_tc"] = {"value": pt["value"], "unit": "images/s", "ms_per_step": pt["ms_per_step"],
                                     "precision": "fp32 storage, split-bf16 x3 tcgen05 contractions (1e-5 from the reference: "
                                                  "tests/test_large_goldens_gpu.py)",
                                     "roofline": pt.get("roofline"), "gpu_launches": pt["gpu_launches"], "steps": ex_steps}
        others = []
        if world == 1:
            others.append(dict(name="configs[2]", batch=8, size=736, tile=718, seg_only=True, heatmap_wgt=0.5))
        if world <= 2:
            others.append(dict(name="configs[4]", batch=2, size=1440, tile=1436, seg_only=False, heatmap_wgt=1.0))
        out = []
        for oc in others:
            r = measure(torch, dist, pkg, dev, rank, world, local, batch=oc["batch"], size=oc["size"], tile=oc["tile"],
                        precision="bf16", seg_only=oc["seg_only"], heatmap_wgt=oc["heatmap_wgt"], steps=ex_steps, warmup=ex_warm,
                        profile=not args.no_profile, **common)
            if rank == 0:
                out.append({"baseline_config": oc["name"],
                            "workload": workload_text(oc["batch"], oc["size"], oc["tile"], oc["seg_only"], oc["heatmap_wgt"]),
                            "global_batch": oc["batch"] * world, "n_gpus": world,
                            "value": r["value"], "unit": "images/s", "ms_per_step": r["ms_per_step"], "steps": ex_steps,
                            "e2e": {k: v for k, v in r["e2e"].items() if k != "inputs"}, "clocks": r["clocks"],
                            "roofline": r.get("roofline"), "model_tflops": r["model_tflops"], "gpu_launches": r["gpu_launches"]})
        if rank == 0:
            line["other_configs"] = out

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ips, cms, kind, n = cpu_reference_run(torch, 3, 1, args.cpu_sample_batch, args.size, args.tile, threads,
                                                  args.seg_only, args.heatmap_wgt, budget_s=20.0)
            what = ("unmodified reference unet.py + dice.py (oracle/_ref)" if kind == "reference" else "oracle port of unet.py")
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
                                    "sample": f"{n} train steps of {args.cpu_sample_batch} tiles (fwd+loss+bwd+SGD), {what} on "
                                              f"torch-CPU, {threads} threads, {cms:.0f} ms/step"}
        print(json.dumps(line), flush=True)
    finish(torch, dist, world)


def finish(torch, dist, world):
    """Orderly teardown: every captured graph and engine has been released by measure(); drain, then destroy."""
    if world > 1:
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
