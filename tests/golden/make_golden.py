"""Generate the golden fixtures in this directory from the REAL reference.

Run once in the authoring container (the only place /root/reference exists):

    python tests/golden/make_golden.py

It imports ``unet.UNet`` from /root/reference/train_test_code unmodified, runs it
on seeded synthetic inputs (fp32, CPU) and stores inputs, weights, outputs,
upstream gradients and autograd parameter gradients as ``.npz`` files.  Nothing
at test/bench time reads /root/reference; the tests read only these files.
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/train_test_code"
HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_unet():
    sys.path.insert(0, REF)
    import unet  # noqa: E402  (the reference module)
    return unet


SMALL_CASES = {
    # name: (kwargs, (B, H, W), training)
    "dual_conv_down_train": (dict(n_classes=7, depth=3, wf=3, batch_norm=True, padding=True, max_pool=False,
                                  num_lands=14, do_res=True, block_depth=2), (2, 32, 32), True),
    "dual_conv_down_eval": (dict(n_classes=7, depth=3, wf=3, batch_norm=True, padding=True, max_pool=False,
                                 num_lands=14, do_res=True, block_depth=2), (2, 32, 32), False),
    "dual_maxpool_train": (dict(n_classes=7, depth=3, wf=3, batch_norm=True, padding=True, max_pool=True,
                                num_lands=14, do_res=True, block_depth=2), (3, 16, 24), True),
    "seg_only_plain_train": (dict(n_classes=5, depth=3, wf=2, batch_norm=False, padding=True, max_pool=True,
                                  num_lands=0, do_res=False, block_depth=1), (2, 16, 16), True),
    "seg_only_bn_nores_train": (dict(n_classes=3, depth=2, wf=3, batch_norm=True, padding=True, max_pool=False,
                                     num_lands=0, do_res=False, block_depth=3), (2, 8, 12), True),
    "lands1_nosoftmax_train": (dict(n_classes=4, depth=2, wf=2, batch_norm=True, padding=True, max_pool=False,
                                    num_lands=6, do_res=True, block_depth=2, lands_num_1x1=1,
                                    do_soft_max=False), (1, 8, 8), True),
    "deep4_wf3_train": (dict(n_classes=7, depth=4, wf=3, batch_norm=True, padding=True, max_pool=False,
                             num_lands=14, do_res=True, block_depth=2), (4, 32, 32), True),
}

PAPER = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False,
             num_lands=14, do_res=True, block_depth=2)


def warm_bn(net, shape, n_iter=3, seed=123):
    g = torch.Generator().manual_seed(seed)
    net.train()
    with torch.no_grad():
        for _ in range(n_iter):
            net(torch.randn(*shape, generator=g))


def run_case(unet, name, kwargs, shape, training, seed):
    torch.manual_seed(seed)
    net = unet.UNet(**kwargs)
    B, H, W = shape
    if kwargs.get("batch_norm"):
        warm_bn(net, (B, 1, H, W))
    # perturb BN affine so gamma/beta gradients and scale paths are non-trivial
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if kwargs.get("batch_norm") and p.dim() == 1 and ".block." in n \
                    and int(n.split(".")[-2]) % 3 == 2:        # BatchNorm2d sits at block.{2,5,8}
                if n.endswith("weight"):
                    p.mul_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
                else:
                    p.add_(0.1 * torch.randn(p.shape, generator=g))
    x = torch.randn(B, 1, H, W, generator=g)
    if kwargs.get("max_pool"):
        # exact ties in the pool windows are common in the real net (ReLU zeros); force some
        x[:, :, ::4, ::4] = 0.0
    state_before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    logits = {}
    hook = net.seg_conv.register_forward_hook(lambda m, i, o: logits.__setitem__("v", o.detach().clone()))
    net.train(training)
    out = net(x)
    hook.remove()
    seg, heat = (out if isinstance(out, tuple) else (out, None))
    d_seg = torch.randn(seg.shape, generator=g)
    d_heat = torch.randn(heat.shape, generator=g) if heat is not None else None
    loss = (seg * d_seg).sum() + ((heat * d_heat).sum() if heat is not None else 0.0)
    loss.backward()
    rec = {"x": x.numpy(), "seg": seg.detach().numpy(), "logits": logits["v"].numpy(),
           "d_seg": d_seg.numpy()}
    if heat is not None:
        rec["heat"] = heat.detach().numpy()
        rec["d_heat"] = d_heat.numpy()
    for k, v in state_before.items():
        rec["state/" + k] = v.numpy()
    for k, v in net.state_dict().items():
        if "running_" in k or "num_batches" in k:
            rec["state_after/" + k] = v.numpy()
    none_grads = []
    for k, p in net.named_parameters():
        if p.grad is None:
            none_grads.append(k)
        else:
            rec["grad/" + k] = p.grad.numpy()
    meta = {"kwargs": kwargs, "shape": list(shape), "training": training, "seed": seed,
            "none_grads": none_grads, "torch": torch.__version__}
    rec["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **rec)
    print(f"{name}: {sum(v.nbytes for v in rec.values()) / 1e6:.2f} MB, none_grads={none_grads}")


def run_paper(unet):
    """Config 1 of BASELINE.json: paper network, eval, 1x1x192x192 synthetic z-scored tile.
    Weights are NOT stored (152 MB); they are reproducible from the seed because the
    engine's UNet constructs the same torch modules in the same order.  Stored:
    per-parameter checksums (to detect init drift), the input, and the outputs
    sub-sampled on a stride-4 grid plus global moments."""
    torch.manual_seed(0)
    net = unet.UNet(**PAPER)
    sums = {k: [float(v.double().sum()), float(v.double().abs().sum())]
            for k, v in net.state_dict().items()}
    warm_bn(net, (2, 1, 192, 192), n_iter=3, seed=7)
    bn_after = {k: v.numpy() for k, v in net.state_dict().items() if "running_" in k}
    net.eval()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 1, 192, 192, generator=g)
    logits = {}
    net.seg_conv.register_forward_hook(lambda m, i, o: logits.__setitem__("v", o.detach().clone()))
    with torch.no_grad():
        seg, heat = net(x)
        logits32 = logits["v"].clone()
        seg64, heat64 = net.double()(x.double())
    rec = {"x": x.numpy(),
           "seg_s4": seg[:, :, ::4, ::4].numpy(), "heat_s4": heat[:, :, ::4, ::4].numpy(),
           "logits_s4": logits32[:, :, ::4, ::4].numpy(),
           "seg64_s4": seg64[:, :, ::4, ::4].numpy(), "heat64_s4": heat64[:, :, ::4, ::4].numpy(),
           "heat_moments": np.array([float(heat.double().sum()), float((heat.double() ** 2).sum())]),
           "seg_moments": np.array([float(seg.double().sum()), float((seg.double() ** 2).sum())])}
    for k, v in bn_after.items():
        rec["bn_after/" + k] = v
    meta = {"kwargs": PAPER, "param_sums": sums, "torch": torch.__version__,
            "warm": {"shape": [2, 1, 192, 192], "n_iter": 3, "seed": 7}, "x_seed": 11}
    rec["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "paper_eval_192.npz"), **rec)
    print("paper_eval_192 written")


def run_losses():
    """dice.py / ncc.py of the reference on seeded inputs (pins oracle + product loss mirrors)."""
    sys.path.insert(0, REF)
    import dice  # noqa: E402
    g = torch.Generator().manual_seed(42)
    seg = torch.softmax(torch.randn(3, 7, 24, 20, generator=g), dim=1).requires_grad_(True)
    heat = torch.randn(3, 14, 24, 20, generator=g).requires_grad_(True)
    tgt_seg = torch.nn.functional.one_hot(torch.randint(0, 7, (3, 24, 20), generator=g), 7).permute(0, 3, 1, 2).float()
    tgt_heat = torch.rand(3, 14, 24, 20, generator=g)
    l_dual = dice.DiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)((seg, heat), (tgt_seg, tgt_heat))
    l_dual.backward()
    l_dice_bg = dice.DiceLoss2D(skip_bg=True)(seg.detach(), tgt_seg)
    l_dice = dice.DiceLoss2D(skip_bg=False)(seg.detach(), tgt_seg)
    np.savez_compressed(os.path.join(HERE, "losses.npz"), seg=seg.detach().numpy(), heat=heat.detach().numpy(),
                        tgt_seg=tgt_seg.numpy(), tgt_heat=tgt_heat.numpy(), l_dual=l_dual.detach().numpy(),
                        l_dice_bg=l_dice_bg.numpy(), l_dice=l_dice.numpy(), d_seg=seg.grad.numpy(),
                        d_heat=heat.grad.numpy())
    print("losses written")


if __name__ == "__main__":
    if "--losses-only" in sys.argv:
        run_losses()
        sys.exit(0)
    torch.set_num_threads(8)
    unet = _ref_unet()
    for i, (name, (kw, shape, training)) in enumerate(SMALL_CASES.items()):
        run_case(unet, name, kw, shape, training, seed=100 + i)
    run_paper(unet)
    run_losses()
