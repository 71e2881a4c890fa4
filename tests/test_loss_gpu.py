"""Fused device loss (csrc/kernels_loss.cuh, fu_loss_forward / fu_loss_backward) against the PyTorch
mirrors of dice.py / ncc.py (losses.py; themselves pinned on the reference's values by
tests/test_oracle_golden.py::test_losses_match_reference_dice_and_ncc) and against the committed golden
loss values generated from the reference.  GPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    assert torch.cuda.is_available()
    p = load_pkg()
    p._capi.lib()
    return p


def _inputs(B, NC, NL, H, Ht, seed, dev):
    g = torch.Generator().manual_seed(seed)
    seg = torch.softmax(torch.randn(B, NC, H, H, generator=g), dim=1)
    heat = torch.randn(B, NL, H, H, generator=g) * 0.05 + 0.01
    tgt_seg = torch.nn.functional.one_hot(torch.randint(0, NC, (B, Ht, Ht), generator=g), NC).permute(0, 3, 1, 2).float().contiguous()
    tgt_heat = torch.rand(B, NL, Ht, Ht, generator=g) * 0.02
    return [t.to(dev) for t in (seg, heat, tgt_seg, tgt_heat)]


@pytest.mark.parametrize("skip_bg", [False, True])
@pytest.mark.parametrize("H,Ht", [(32, 20), (24, 24), (48, 37)])
def test_fused_dice_heatmap_loss_matches_pytorch_mirror(pkg, skip_bg, H, Ht):
    dev = torch.device("cuda:0")
    seg, heat, tgt_seg, tgt_heat = _inputs(3, 7, 14, H, Ht, 5 + H, dev)
    ref_crit = pkg.DiceAndHeatMapLoss2D(skip_bg=skip_bg, heatmap_wgt=0.3)
    fus_crit = pkg.FusedDiceAndHeatMapLoss2D(skip_bg=skip_bg, heatmap_wgt=0.3)
    a_seg, a_heat = seg.clone().requires_grad_(True), heat.clone().requires_grad_(True)
    l_ref = ref_crit((pkg.center_crop(a_seg, tgt_seg.shape), pkg.center_crop(a_heat, tgt_heat.shape)), (tgt_seg, tgt_heat))
    (l_ref * 1.7).backward()
    # (1) uncropped outputs: the crop happens inside the kernel, gradients come back full size
    b_seg, b_heat = seg.clone().requires_grad_(True), heat.clone().requires_grad_(True)
    l_fus = fus_crit((b_seg, b_heat), (tgt_seg, tgt_heat))
    (l_fus * 1.7).backward()
    assert abs(float(l_fus) - float(l_ref)) < 2e-6 * max(1.0, abs(float(l_ref)))
    for got, want in ((b_seg.grad, a_seg.grad), (b_heat.grad, a_heat.grad)):
        err = float((got - want).double().norm() / (want.double().norm() + 1e-30))
        assert err < 2e-5, err
    # (2) the train.py:414-418 call pattern: cropped views in, autograd pads the gradient
    c_seg, c_heat = seg.clone().requires_grad_(True), heat.clone().requires_grad_(True)
    l_c = fus_crit((pkg.center_crop(c_seg, tgt_seg.shape), pkg.center_crop(c_heat, tgt_heat.shape)), (tgt_seg, tgt_heat))
    (l_c * 1.7).backward()
    assert abs(float(l_c) - float(l_ref)) < 2e-6 * max(1.0, abs(float(l_ref)))
    assert float((c_seg.grad - a_seg.grad).abs().max()) < 1e-7 + 2e-5 * float(a_seg.grad.abs().max())
    assert float((c_heat.grad - a_heat.grad).abs().max()) < 1e-7 + 2e-5 * float(a_heat.grad.abs().max())


@pytest.mark.parametrize("skip_bg", [False, True])
def test_fused_dice_loss_seg_only(pkg, skip_bg):
    dev = torch.device("cuda:0")
    seg, _, tgt_seg, _ = _inputs(2, 7, 1, 40, 28, 11, dev)
    a = seg.clone().requires_grad_(True)
    l_ref = pkg.DiceLoss2D(skip_bg=skip_bg)(pkg.center_crop(a, tgt_seg.shape), tgt_seg)
    l_ref.backward()
    b = seg.clone().requires_grad_(True)
    l_fus = pkg.FusedDiceLoss2D(skip_bg=skip_bg)(b, tgt_seg)
    l_fus.backward()
    assert abs(float(l_fus) - float(l_ref)) < 2e-6
    assert float((b.grad - a.grad).double().norm() / a.grad.double().norm()) < 2e-5


def test_fused_loss_matches_reference_golden_values(pkg):
    """losses.npz holds inputs and the loss values the REFERENCE's dice.py / ncc.py returned for them."""
    z = np.load(os.path.join(GOLDEN, "losses.npz"))
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    seg, heat = t["seg"].to(dev), t["heat"].to(dev)
    tgt_seg, tgt_heat = t["tgt_seg"].to(dev), t["tgt_heat"].to(dev)
    seg.requires_grad_(True); heat.requires_grad_(True)
    l = pkg.FusedDiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)((seg, heat), (tgt_seg, tgt_heat))
    l.backward()
    assert abs(float(l) - float(t["l_dual"])) < 2e-6
    # gradients the reference's autograd produced for the same inputs
    for got, want in ((seg.grad.cpu(), t["d_seg"]), (heat.grad.cpu(), t["d_heat"])):
        assert float((got - want).double().norm() / want.double().norm()) < 2e-5
    seg, heat = seg.detach(), heat.detach()
    assert abs(float(pkg.FusedDiceLoss2D(skip_bg=True)(seg, tgt_seg)) - float(t["l_dice_bg"])) < 2e-6
    assert abs(float(pkg.FusedDiceLoss2D(skip_bg=False)(seg, tgt_seg)) - float(t["l_dice"])) < 2e-6


def test_fused_loss_rejects_cpu_tensors(pkg):
    seg = torch.rand(1, 7, 8, 8)
    with pytest.raises(RuntimeError):
        pkg.FusedDiceLoss2D()(seg, seg)


@pytest.mark.parametrize("case", [
    dict(B=3, H=48, W=64, crop=(6, 6), nl=14, skip_bg=False, softmax=True),     # dual head, centre crop as train.py:414-417
    dict(B=2, H=32, W=48, crop=(0, 0), nl=14, skip_bg=True, softmax=True),      # no crop, background class skipped
    dict(B=5, H=40, W=24, crop=(4, 2), nl=0, skip_bg=False, softmax=True),      # seg-only (train.py:327), uneven crop
    dict(B=2, H=48, W=48, crop=(6, 6), nl=14, skip_bg=False, softmax=False),    # logits as segmentation output
])
def test_loss_inside_the_heads_kernels_matches_the_separate_loss(pkg, case):
    """UNet.forward_loss (fu_forward_loss / fu_backward_loss: Dice / NCC sums from the head kernel, loss gradient formed in the
    backward head kernel) against criterion(net(x), target) -- the same network, the same arithmetic as separate kernels.  The
    loss must agree to fp32 rounding and every parameter gradient to the bf16 rounding of the head's feature gradient."""
    dev = torch.device("cuda:0")
    B, H, W, nl = case["B"], case["H"], case["W"], case["nl"]
    kw = dict(n_classes=7, depth=3, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=nl, do_soft_max=case["softmax"])
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 1, H, W, generator=g).to(dev)
    Ht, Wt = H - 2 * case["crop"][0], W - 2 * case["crop"][1]
    mask = torch.nn.functional.one_hot(torch.randint(0, 7, (B, Ht, Wt), generator=g), 7).permute(0, 3, 1, 2).float().contiguous().to(dev)
    heat = torch.rand(B, nl, Ht, Wt, generator=g).to(dev) if nl else None
    crit = (pkg.FusedDiceAndHeatMapLoss2D(skip_bg=case["skip_bg"], heatmap_wgt=0.5) if nl else pkg.FusedDiceLoss2D(skip_bg=case["skip_bg"]))
    target = (mask, heat) if nl else mask
    res = {}
    for mode in ("fused", "separate"):
        torch.manual_seed(0)
        net = pkg.UNet(precision="bf16", **kw).to(dev).train()
        if mode == "fused":
            loss = net.forward_loss(x, target, crit)
        else:
            loss = crit(net(x), target)
        (loss * 3.0).backward()                 # a non-trivial upstream gradient
        torch.cuda.synchronize()
        res[mode] = (float(loss), {n: p.grad.cpu() for n, p in net.named_parameters() if p.grad is not None},
                     {n: b.cpu() for n, b in net.named_buffers()})
    lf, ls = res["fused"][0], res["separate"][0]
    assert abs(lf - ls) < 2e-6 * max(1.0, abs(ls)), (lf, ls)
    assert set(res["fused"][1]) == set(res["separate"][1])
    for k, v in res["fused"][1].items():
        w = res["separate"][1][k]
        e = float((v.double() - w.double()).norm() / (w.double().norm() + 1e-30))
        assert e < (5e-3 if k.startswith(("seg", "lands")) else 3e-2), (k, e)
    for k, v in res["fused"][2].items():        # BN running statistics: the forward is the same forward
        assert torch.allclose(v.double(), res["separate"][2][k].double(), rtol=1e-6, atol=1e-9), k


def test_forward_loss_rejects_what_it_cannot_fuse(pkg):
    dev = torch.device("cuda:0")
    kw = dict(n_classes=7, depth=2, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=0)
    x = torch.randn(1, 1, 16, 16).to(dev)
    mask = torch.zeros(1, 7, 16, 16, device=dev)
    net32 = pkg.UNet(precision="fp32", **kw).to(dev).train()
    with pytest.raises(ValueError):
        net32.forward_loss(x, mask, pkg.FusedDiceLoss2D(skip_bg=False))          # parity modes keep the separate kernels
    net = pkg.UNet(precision="bf16", **kw).to(dev).train()
    with pytest.raises(TypeError):
        net.forward_loss(x, mask, pkg.DiceLoss2D(skip_bg=False))
    with pytest.raises(RuntimeError):
        net.eval().forward_loss(x, mask, pkg.FusedDiceLoss2D(skip_bg=False))
