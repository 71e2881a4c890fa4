"""Golden fixtures at BASELINE.json's real sizes, generated from the REAL reference (authoring container only):

    python tests/golden/make_golden_large.py [name ...]

  paper_b32_192   configs[1]: paper dual-head net, 32 tiles of 1x180x180 padded to 192, one training step
                  (train.py:405-422: forward, DiceAndHeatMapLoss2D on the centre-cropped outputs, backward)
  seg_b8_736      configs[2]: seg-only net (num_lands=0), 8 tiles of 1x718x718 padded to 736, DiceLoss2D (train.py:327)
  dual_b2_1440    configs[4]: dual-head net, 2 tiles (one GPU's share of the batch of 4) of 1x1436x1436 padded to
                  1440, heatmap_wgt = 1.0

Weights and inputs are NOT stored: they are functions of the seeds below (the engine's UNet builds the same torch
modules in the same order, so torch.manual_seed reproduces the reference's init; checksums are stored to detect
drift).  Stored: outputs on a strided grid, the loss, BN running statistics after the step and, for every parameter
gradient, its L2 norm, its sum and a strided sample of <= 2048 elements -- from the fp32 reference run AND from the
same module in fp64 (the truth; the fp32 run's distance to it is the configuration's noise floor).
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch

REF = "/root/reference/train_test_code"
HERE = os.path.dirname(os.path.abspath(__file__))

PAPER = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False,
             num_lands=14, do_res=True, block_depth=2)
CASES = {
    "paper_b32_192": dict(kwargs=PAPER, B=32, S=192, T=180, heatmap_wgt=0.5, out_stride=12),
    "seg_b8_736": dict(kwargs=dict(PAPER, num_lands=0), B=8, S=736, T=718, heatmap_wgt=None, out_stride=32),
    "dual_b2_1440": dict(kwargs=PAPER, B=2, S=1440, T=1436, heatmap_wgt=1.0, out_stride=48),
}
GRAD_SAMPLE = 2048


def make_inputs(case, seed=21):
    """Inputs and targets of one step (shared with tests/test_large_goldens_gpu.py): z-scored N(0,1) tiles
    (dataset.py:292-293), one-hot float masks (dataset.py:448-452), Gaussian heat-maps with sigma 2.5 and peak
    1/(2 pi sigma^2) at random in-bounds pixels (dataset.py:295-325)."""
    B, S, T = case["B"], case["S"], case["T"]
    nc, nl = case["kwargs"]["n_classes"], case["kwargs"]["num_lands"]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, S, S, generator=g)
    labels = torch.randint(0, nc, (B, T, T), generator=g)
    mask = torch.nn.functional.one_hot(labels, nc).permute(0, 3, 1, 2).contiguous().float()
    heat = None
    if nl > 0:
        sigma = 2.5
        ys = torch.arange(T, dtype=torch.float32).view(1, 1, T, 1)
        xs = torch.arange(T, dtype=torch.float32).view(1, 1, 1, T)
        cy = torch.randint(0, T, (B, nl, 1, 1), generator=g).float()
        cx = torch.randint(0, T, (B, nl, 1, 1), generator=g).float()
        heat = (torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * sigma * sigma)) / (2 * math.pi * sigma * sigma)).contiguous()
    return x, mask, heat


def grad_sample(t):
    f = t.detach().flatten()
    stride = max(1, (f.numel() + GRAD_SAMPLE - 1) // GRAD_SAMPLE)
    return f[::stride].clone(), stride


def run(name):
    sys.path.insert(0, REF)
    import unet    # the reference network, unmodified
    import dice    # the reference losses
    import util    # center_crop
    case = CASES[name]
    t0 = time.time()
    torch.manual_seed(0)
    net = unet.UNet(**case["kwargs"])
    sums = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in net.state_dict().items()}
    x, mask, heat_t = make_inputs(case)
    net.train()                                                             # train.py:381
    out = net(x)                                                            # train.py:407
    if case["heatmap_wgt"] is None:
        seg, heat = out, None
        crit = dice.DiceLoss2D(skip_bg=False)                               # train.py:327
        loss = crit(util.center_crop(seg, mask.shape), mask)                # train.py:414-420
    else:
        seg, heat = out
        crit = dice.DiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=case["heatmap_wgt"])     # train.py:324
        loss = crit((util.center_crop(seg, mask.shape), util.center_crop(heat, heat_t.shape)), (mask, heat_t))
    loss.backward()                                                         # train.py:422
    st = case["out_stride"]
    rec = {"loss": np.array(float(loss.detach()), dtype=np.float64),
           "seg_s": seg.detach()[:, :, ::st, ::st].numpy(),
           "x_sums": np.array([float(x.double().sum()), float(x.double().abs().sum())]),
           "mask_sum": np.array(float(mask.double().sum()))}
    if heat is not None:
        rec["heat_s"] = heat.detach()[:, :, ::st, ::st].numpy()
        rec["heat_t_sum"] = np.array(float(heat_t.double().sum()))
    for k, v in net.state_dict().items():
        if "running_" in k:
            rec["state_after/" + k] = v.numpy()
    none_grads, strides = [], {}
    for k, p in net.named_parameters():
        if p.grad is None:
            none_grads.append(k)
            continue
        smp, stride = grad_sample(p.grad)
        rec["grad_sample/" + k] = smp.numpy()
        rec["grad_norm/" + k] = np.array([float(p.grad.double().norm()), float(p.grad.double().sum())])
        strides[k] = stride
    # the same step in fp64 (reference module .double()): the truth both fp32 implementations are measured against;
    # the fp32 reference's own distance to it is the noise floor of this configuration (ReLU masks that flip within
    # rounding of zero, cancelling sums in the bias / BN-affine gradients)
    del out, loss
    for p in net.parameters():
        p.grad = None
    torch.manual_seed(0)
    net64 = unet.UNet(**case["kwargs"]).double()
    net64.train()
    out = net64(x.double())
    if case["heatmap_wgt"] is None:
        loss64 = crit(util.center_crop(out, mask.shape), mask.double())
    else:
        loss64 = crit((util.center_crop(out[0], mask.shape), util.center_crop(out[1], heat_t.shape)), (mask.double(), heat_t.double()))
    loss64.backward()
    rec["loss64"] = np.array(float(loss64.detach()), dtype=np.float64)
    for k, p in net64.named_parameters():
        if p.grad is not None:
            rec["grad_sample64/" + k] = grad_sample(p.grad)[0].numpy()
    del out, loss64, net64
    meta = {"case": {k: v for k, v in case.items()}, "param_sums": sums, "none_grads": none_grads, "grad_strides": strides,
            "torch": torch.__version__, "input_seed": 21, "init_seed": 0, "seconds": time.time() - t0}
    rec["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"large_{name}.npz"), **rec)
    print(f"{name}: loss {float(rec['loss']):.6f}, {sum(v.nbytes for v in rec.values()) / 1e6:.2f} MB, {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    for n in (sys.argv[1:] or list(CASES)):
        run(n)
