"""CPU oracle for the dual-head U-Net hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

A functional restatement of /root/reference/train_test_code/unet.py (forward)
and of the autograd backward PyTorch derives from it, written out explicitly so
that every CUDA kernel of the engine has a formula to be checked against.

* Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
  legs may import this module.  The product path (the package
  ``deepfluorolabeling-ipcai2020_b200``) never does.
* The arithmetic of the reference lives in a third-party dependency, PyTorch
  ATen (the reference pins nothing; torch 2.11.0+cu128 is what is installed
  here and on the GPU box).  The contractions below therefore call the same
  ATen CPU primitives (``F.conv2d`` ...) the reference's ``nn.Conv2d`` modules
  resolve to; everything *around* them (block order, BN statistics, residual,
  concat order, heads, and the complete backward) is restated by hand.
* Pinned: ``tests/golden/make_golden.py`` imports the real reference
  ``unet.UNet`` from /root/reference in the authoring container and stores its
  outputs / autograd gradients as fixtures; ``tests/test_oracle_golden.py``
  checks this file against them.  The reference itself ships no tests or
  golden vectors, so these fixtures are the pin.

All tensors are NCHW; dtype follows the inputs (fp32 or fp64).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default, unet.py:215,222
BN_MOMENTUM = 0.1    # nn.BatchNorm2d default


@dataclass
class UNetConfig:
    """Constructor arguments of the reference UNet (unet.py:41-45)."""
    in_channels: int = 1
    n_classes: int = 2
    depth: int = 5
    wf: int = 6
    padding: bool = False
    pad_mode: str = "zeros"
    batch_norm: bool = False
    up_mode: str = "upconv"
    max_pool: bool = True
    num_lands: int = 0
    do_res: bool = True
    block_depth: int = 2
    lands_block_depth: int = 0
    lands_num_1x1: int = 2
    do_soft_max: bool = True

    def as_kwargs(self) -> dict:
        return asdict(self)


def paper_config() -> UNetConfig:
    """train_test_code/Readme.md:16 -> train.py:313."""
    return UNetConfig(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True,
                      max_pool=False, num_lands=14, do_res=True, block_depth=2)


def _check_supported(cfg: UNetConfig) -> None:
    if not cfg.padding or cfg.pad_mode != "zeros" or cfg.up_mode != "upconv" \
            or cfg.lands_block_depth != 0:
        raise ValueError("oracle covers padding=True, zeros, upconv, lands_block_depth=0 "
                         "(SURVEY.md section 8b)")


# --------------------------------------------------------------------------
# state_dict schema
# --------------------------------------------------------------------------
def param_schema(cfg: UNetConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) in the reference's state_dict order.
    kind in {'param', 'buffer'}.  Follows unet.py:80-159 / :196-224 / :236-246."""
    _check_supported(cfg)
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv_block(prefix: str, cin: int, cout: int):
        if cfg.do_res:
            out.append((f"{prefix}.res_conv1x1.weight", (cout, cin, 1, 1), "param"))
            out.append((f"{prefix}.res_conv1x1.bias", (cout,), "param"))
        idx = 0
        c = cin
        for _ in range(cfg.block_depth):
            out.append((f"{prefix}.block.{idx}.weight", (cout, c, 3, 3), "param"))
            out.append((f"{prefix}.block.{idx}.bias", (cout,), "param"))
            idx += 2  # conv, relu
            if cfg.batch_norm:
                out.append((f"{prefix}.block.{idx}.weight", (cout,), "param"))
                out.append((f"{prefix}.block.{idx}.bias", (cout,), "param"))
                out.append((f"{prefix}.block.{idx}.running_mean", (cout,), "buffer"))
                out.append((f"{prefix}.block.{idx}.running_var", (cout,), "buffer"))
                out.append((f"{prefix}.block.{idx}.num_batches_tracked", (), "buffer"))
                idx += 1
            c = cout

    chans = [2 ** (cfg.wf + i) for i in range(cfg.depth)]
    # nn.Module registers downsample_convs before down_path (unet.py:78-84)
    if not cfg.max_pool:
        for i, c in enumerate(chans):
            out.append((f"downsample_convs.{i}.weight", (c, c, 2, 2), "param"))
            out.append((f"downsample_convs.{i}.bias", (c,), "param"))
    prev = cfg.in_channels
    for i, c in enumerate(chans):
        conv_block(f"down_path.{i}", prev, c)
        prev = c
    for j, i in enumerate(reversed(range(cfg.depth - 1))):
        c = chans[i]
        out.append((f"up_path.{j}.up.weight", (prev, c, 2, 2), "param"))
        out.append((f"up_path.{j}.up.bias", (c,), "param"))
        conv_block(f"up_path.{j}.conv_block", prev, c)
        prev = c
    out.append(("seg_conv.weight", (cfg.n_classes, prev, 1, 1), "param"))
    if cfg.num_lands > 0:
        nf = cfg.num_lands + cfg.n_classes if cfg.lands_num_1x1 > 1 else cfg.num_lands
        out.append(("lands_1x1.0.weight", (nf, prev + cfg.n_classes, 1, 1), "param"))
        for k in range(cfg.lands_num_1x1 - 1):
            out.append((f"lands_1x1.{k + 1}.weight", (cfg.num_lands, nf, 1, 1), "param"))
            nf = cfg.num_lands
    return out


# --------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------
NATIVE_BN = False   # timing baseline only: call ATen's fused batch_norm like nn.BatchNorm2d does


def _bn_forward(x, sd, prefix, training, tape, new_stats):
    """nn.BatchNorm2d, unet.py:214-215,221-222: batch statistics with biased
    variance for the normalisation, unbiased variance into running_var."""
    gamma, beta = sd[f"{prefix}.weight"], sd[f"{prefix}.bias"]
    if NATIVE_BN:
        rm, rv = sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"]   # updated in place, like the module
        if training:
            sd[f"{prefix}.num_batches_tracked"] += 1
        return F.batch_norm(x, rm, rv, gamma, beta, training, BN_MOMENTUM, BN_EPS)
    if training:
        n = x.shape[0] * x.shape[2] * x.shape[3]
        mean = x.mean(dim=(0, 2, 3))
        var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
        if new_stats is not None:
            unbiased = var * (n / max(n - 1, 1))
            new_stats[f"{prefix}.running_mean"] = \
                (1 - BN_MOMENTUM) * sd[f"{prefix}.running_mean"] + BN_MOMENTUM * mean
            new_stats[f"{prefix}.running_var"] = \
                (1 - BN_MOMENTUM) * sd[f"{prefix}.running_var"] + BN_MOMENTUM * unbiased
            new_stats[f"{prefix}.num_batches_tracked"] = \
                sd[f"{prefix}.num_batches_tracked"] + 1
    else:
        mean, var = sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"]
    invstd = torch.rsqrt(var + BN_EPS)
    xhat = (x - mean[None, :, None, None]) * invstd[None, :, None, None]
    y = xhat * gamma[None, :, None, None] + beta[None, :, None, None]
    tape.append(("bn", prefix, xhat, invstd, training))
    return y


def _conv_block_forward(x, sd, cfg, prefix, training, tape, new_stats):
    """UNetConvBlock.forward, unet.py:226-233: [conv3x3 -> ReLU -> BN] x
    block_depth, then ``out += res_conv1x1(x)``."""
    x_in = x
    idx = 0
    for _ in range(cfg.block_depth):
        w, b = sd[f"{prefix}.block.{idx}.weight"], sd[f"{prefix}.block.{idx}.bias"]
        tape.append(("conv", f"{prefix}.block.{idx}", x, 1, 1))
        x = F.conv2d(x, w, b, stride=1, padding=1)
        x = torch.relu(x)
        tape.append(("relu", x))
        idx += 2
        if cfg.batch_norm:
            x = _bn_forward(x, sd, f"{prefix}.block.{idx}", training, tape, new_stats)
            idx += 1
    if cfg.do_res:
        tape.append(("res", f"{prefix}.res_conv1x1", x_in))
        x = x + F.conv2d(x_in, sd[f"{prefix}.res_conv1x1.weight"], sd[f"{prefix}.res_conv1x1.bias"])
    return x


def forward(sd: Dict[str, torch.Tensor], cfg: UNetConfig, x: torch.Tensor,
            training: bool = False, want_tape: bool = False):
    """UNet.forward, unet.py:161-193.

    Returns dict with 'seg' (softmax probabilities or logits if
    do_soft_max=False), 'logits' (seg_x, unet.py:176), 'heat' (or None),
    'new_stats' (BN buffers after this step, training only) and 'tape'."""
    _check_supported(cfg)
    if x.shape[2] % (1 << (cfg.depth - 1)) or x.shape[3] % (1 << (cfg.depth - 1)):
        raise ValueError("H and W must be multiples of 2**(depth-1)")
    tape: list = []
    new_stats: Optional[dict] = {} if training else None
    bridges = []
    for i in range(cfg.depth):
        x = _conv_block_forward(x, sd, cfg, f"down_path.{i}", training, tape, new_stats)
        if i != cfg.depth - 1:
            bridges.append(x)
            tape.append(("bridge_push",))
            if cfg.max_pool:
                tape.append(("maxpool", x))
                x = F.max_pool2d(x, 2)
            else:
                tape.append(("conv", f"downsample_convs.{i}", x, 2, 0))
                x = F.conv2d(x, sd[f"downsample_convs.{i}.weight"],
                             sd[f"downsample_convs.{i}.bias"], stride=2)
    for j in range(cfg.depth - 1):
        # UNetUpBlock.forward, unet.py:254-260: ConvT 2x2/s2, cat([up, bridge]).
        tape.append(("convT", f"up_path.{j}.up", x))
        up = F.conv_transpose2d(x, sd[f"up_path.{j}.up.weight"], sd[f"up_path.{j}.up.bias"], stride=2)
        bridge = bridges[-j - 1]      # crop is the identity when padding=True
        x = torch.cat([up, bridge], dim=1)
        tape.append(("cat", up.shape[1]))
        x = _conv_block_forward(x, sd, cfg, f"up_path.{j}.conv_block", training, tape, new_stats)
    feat = x
    logits = F.conv2d(feat, sd["seg_conv.weight"])               # unet.py:176, no bias
    seg = torch.softmax(logits, dim=1) if cfg.do_soft_max else logits   # Softmax2d, unet.py:179
    heat = None
    h_mid = None
    if cfg.num_lands > 0:
        cat = torch.cat([feat, logits], dim=1)                   # unet.py:187, features first
        h = cat
        mids = [cat]
        for k in range(cfg.lands_num_1x1):                       # no non-linearity between, unet.py:141-159
            h = F.conv2d(h, sd[f"lands_1x1.{k}.weight"])
            mids.append(h)
        heat = h
        h_mid = mids
    tape.append(("heads", feat, logits, seg, h_mid))
    return {"seg": seg, "logits": logits, "heat": heat, "new_stats": new_stats,
            "tape": tape if want_tape else None}


# --------------------------------------------------------------------------
# explicit backward (what loss.backward() at train.py:422 computes)
# --------------------------------------------------------------------------
def _conv_backward(x, w, dy, stride, padding, need_dx=True):
    dw = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=stride, padding=padding)
    db = dy.sum(dim=(0, 2, 3))
    dx = None
    if need_dx:
        dx = torch.nn.grad.conv2d_input(x.shape, w, dy, stride=stride, padding=padding)
    return dx, dw, db


def _maxpool2_backward(x, dy):
    """F.max_pool2d(x, 2) backward: the FIRST maximum in row-major window order
    receives the gradient (measured)."""
    n, c, h, w = x.shape
    win = torch.stack([x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2],
                       x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]], dim=-1)
    mx = win.max(dim=-1).values
    first = torch.full(mx.shape, 3, dtype=torch.int64)
    for k in (3, 2, 1, 0):
        first = torch.where(win[..., k] == mx, torch.full_like(first, k), first)
    dx = torch.zeros_like(x)
    for k, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        dx[:, :, a::2, b::2] = torch.where(first == k, dy, torch.zeros_like(dy))
    return dx


def backward(sd: Dict[str, torch.Tensor], cfg: UNetConfig, tape: list,
             d_seg: Optional[torch.Tensor], d_heat: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Gradients of every reachable parameter given dL/dseg and dL/dheat.
    Unreachable parameters (downsample_convs.{depth-1}) are absent."""
    grads: Dict[str, torch.Tensor] = {}
    tape = list(tape)
    kind, feat, logits, seg, mids = tape.pop()
    assert kind == "heads"
    d_feat = torch.zeros_like(feat)
    d_logits = torch.zeros_like(logits)
    if cfg.num_lands > 0 and d_heat is not None:
        dh = d_heat
        for k in reversed(range(cfg.lands_num_1x1)):
            w = sd[f"lands_1x1.{k}.weight"]
            grads[f"lands_1x1.{k}.weight"] = torch.einsum("nohw,nihw->oi", dh, mids[k])[:, :, None, None]
            dh = torch.einsum("nohw,oi->nihw", dh, w[:, :, 0, 0])
        d_feat = d_feat + dh[:, :feat.shape[1]]
        d_logits = d_logits + dh[:, feat.shape[1]:]
    elif cfg.num_lands > 0:
        for k in range(cfg.lands_num_1x1):
            grads[f"lands_1x1.{k}.weight"] = torch.zeros_like(sd[f"lands_1x1.{k}.weight"])
    if d_seg is not None:
        if cfg.do_soft_max:
            d_logits = d_logits + seg * (d_seg - (d_seg * seg).sum(dim=1, keepdim=True))
        else:
            d_logits = d_logits + d_seg
    w = sd["seg_conv.weight"]
    grads["seg_conv.weight"] = torch.einsum("nohw,nihw->oi", d_logits, feat)[:, :, None, None]
    dx = d_feat + torch.einsum("nohw,oi->nihw", d_logits, w[:, :, 0, 0])

    bridge_grads: list = []          # stack of gradients flowing into encoder outputs
    d_res_in = None                  # pending residual-branch contribution to the block input

    while tape:
        ent = tape.pop()
        kind = ent[0]
        if kind == "res":
            _, prefix, x_in = ent
            w = sd[f"{prefix}.weight"]
            need_dx = prefix != "down_path.0.res_conv1x1"
            d_in, dw, db = _conv_backward(x_in, w, dx, 1, 0, need_dx=need_dx)
            grads[f"{prefix}.weight"], grads[f"{prefix}.bias"] = dw, db
            d_res_in = d_in
        elif kind == "bn":
            _, prefix, xhat, invstd, training = ent
            gamma = sd[f"{prefix}.weight"]
            grads[f"{prefix}.weight"] = (dx * xhat).sum(dim=(0, 2, 3))
            grads[f"{prefix}.bias"] = dx.sum(dim=(0, 2, 3))
            g = (gamma * invstd)[None, :, None, None]
            if training:
                m1 = dx.mean(dim=(0, 2, 3), keepdim=True)
                m2 = (dx * xhat).mean(dim=(0, 2, 3), keepdim=True)
                dx = g * (dx - m1 - xhat * m2)
            else:
                dx = g * dx
        elif kind == "relu":
            dx = dx * (ent[1] > 0).to(dx.dtype)
        elif kind == "conv":
            _, prefix, x_in, stride, padding = ent
            w = sd[f"{prefix}.weight"]
            first = prefix == "down_path.0.block.0"
            d_in, dw, db = _conv_backward(x_in, w, dx, stride, padding, need_dx=not first)
            grads[f"{prefix}.weight"], grads[f"{prefix}.bias"] = dw, db
            dx = d_in
            # the first conv of a block closes the block: add the residual branch
            if prefix.endswith(".block.0"):
                if d_res_in is not None and dx is not None:
                    dx = dx + d_res_in
                d_res_in = None
        elif kind == "cat":
            c_up = ent[1]
            bridge_grads.append(dx[:, c_up:])
            dx = dx[:, :c_up]
        elif kind == "convT":
            _, prefix, x_in = ent
            w = sd[f"{prefix}.weight"]     # (Cin, Cout, 2, 2)
            grads[f"{prefix}.bias"] = dx.sum(dim=(0, 2, 3))
            n, ci, h, wd = x_in.shape
            dyr = dx.reshape(n, w.shape[1], h, 2, wd, 2)
            grads[f"{prefix}.weight"] = torch.einsum("nihw,nohawb->ioab", x_in, dyr)
            dx = torch.einsum("nohawb,ioab->nihw", dyr, w)
        elif kind == "maxpool":
            dx = _maxpool2_backward(ent[1], dx)
        elif kind == "bridge_push":
            # encoder output feeds both the downsample path (dx) and the skip
            dx = dx + bridge_grads.pop()
        else:
            raise AssertionError(kind)
    return grads


# --------------------------------------------------------------------------
# losses (dice.py:20-55, :67-86; ncc.py:12-38) restated for the bench harness
# --------------------------------------------------------------------------
def center_crop(img, dst_shape):
    """util.py:92-114."""
    sr, sc = img.shape[-2], img.shape[-1]
    dr, dc = dst_shape[-2], dst_shape[-1]
    if (dr != sr) or (dc != sc):
        r0 = int((sr - dr) / 2)
        c0 = int((sc - dc) / 2)
        return img[..., r0:r0 + dr, c0:c0 + dc]
    return img


def dice_loss(inp, tgt, skip_bg=False):
    eps = 1.0e-4
    if skip_bg:
        inp, tgt = inp[:, 1:], tgt[:, 1:]
    num = -2 * (tgt * inp).sum(dim=(2, 3)) + eps
    den = (tgt * tgt).sum(dim=(2, 3)) + (inp * inp).sum(dim=(2, 3)) + eps
    return ((num / den).sum(dim=1) / inp.shape[1]).mean()


def ncc_2d(X, Y):
    N = X.shape[-1] * X.shape[-2]
    Xz = X - X.mean(dim=(-2, -1), keepdim=True)
    Yz = Y - Y.mean(dim=(-2, -1), keepdim=True)
    Xs = torch.sqrt((Xz * Xz).sum(dim=(-2, -1)) / (N - 1))
    Ys = torch.sqrt((Yz * Yz).sum(dim=(-2, -1)) / (N - 1))
    return (Xz * Yz).sum(dim=(-2, -1)) / ((N * (Xs * Ys)) + 1.0e-8)


def dice_and_heatmap_loss(seg, heat, tgt_seg, tgt_heat, skip_bg=False, heatmap_wgt=0.5):
    ncc = (ncc_2d(heat, tgt_heat) + 1) * -0.5
    return (1 - heatmap_wgt) * dice_loss(seg, tgt_seg, skip_bg) + heatmap_wgt * ncc.mean()


def loss_and_output_grads(out, cfg: UNetConfig, tgt_seg, tgt_heat, heatmap_wgt=0.5):
    """Loss of train.py:414-420 and its gradient w.r.t. the (uncropped) net
    outputs, via autograd on the tiny loss graph only."""
    seg = out["seg"].detach().requires_grad_(True)
    heat = out["heat"].detach().requires_grad_(True) if out["heat"] is not None else None
    seg_c = center_crop(seg, tgt_seg.shape)
    if heat is not None:
        loss = dice_and_heatmap_loss(seg_c, center_crop(heat, tgt_heat.shape), tgt_seg, tgt_heat,
                                     heatmap_wgt=heatmap_wgt)
    else:
        loss = dice_loss(seg_c, tgt_seg)
    loss.backward()
    return loss.detach(), seg.grad, (heat.grad if heat is not None else None)
