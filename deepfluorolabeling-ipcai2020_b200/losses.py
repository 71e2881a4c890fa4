"""Host-side (PyTorch) mirrors of the reference's losses, which north_star keeps in PyTorch:
``DiceLoss2D`` / ``DiceAndHeatMapLoss2D`` (dice.py:14-86) and ``ncc_2d`` (ncc.py:12-38).
Same class names, arguments and arithmetic; they run on whatever device their inputs are on."""
import torch
import torch.nn.modules.loss


def ncc_2d(X, Y):
    """Normalised cross-correlation of each 2-D image over the last two dims (ncc.py:12-38)."""
    N = X.shape[-1] * X.shape[-2]
    assert N > 1
    d1, d2 = X.dim() - 2, X.dim() - 1
    X_zm = X - torch.mean(X, dim=[d1, d2], keepdim=True)
    X_sd = torch.sqrt(torch.sum(X_zm * X_zm, dim=[d1, d2]) / (N - 1))
    Y_zm = Y - torch.mean(Y, dim=[d1, d2], keepdim=True)
    Y_sd = torch.sqrt(torch.sum(Y_zm * Y_zm, dim=[d1, d2]) / (N - 1))
    return torch.sum(X_zm * Y_zm, dim=[d1, d2]) / ((N * (X_sd * Y_sd)) + 1.0e-8)


class DiceLoss2D(torch.nn.modules.loss._Loss):
    """Negated soft Dice, mean over classes then batch (dice.py:14-55)."""

    def __init__(self, skip_bg=True):
        super().__init__()
        self.skip_bg = skip_bg

    def forward(self, input, target):
        eps = 1.0e-4
        if self.skip_bg:
            input, target = input[:, 1:, :, :], target[:, 1:, :, :]
        numerators = -2 * torch.sum(target * input, dim=(2, 3)) + eps
        denominators = torch.sum(target * target, dim=(2, 3)) + torch.sum(input * input, dim=(2, 3)) + eps
        dices = numerators / denominators
        avg_dices = torch.sum(dices, dim=1) / input.shape[1]
        return torch.mean(avg_dices)


class DiceAndHeatMapLoss2D(torch.nn.modules.loss._Loss):
    """dice_wgt * Dice + heatmap_wgt * mean((ncc+1) * -0.5)  (dice.py:57-86)."""

    def __init__(self, skip_bg=True, heatmap_wgt=0.5):
        super().__init__()
        self.dice_loss = DiceLoss2D(skip_bg=skip_bg)
        assert (heatmap_wgt > 1.0e-8) and (heatmap_wgt < (1 + 1.0e-8))
        self.heatmap_wgt = heatmap_wgt
        self.dice_wgt = 1 - heatmap_wgt

    def forward(self, input, target):
        in_seg, in_heatmaps = input[0], input[1]
        tgt_seg, tgt_heatmaps = target[0], target[1]
        ncc_losses = (ncc_2d(in_heatmaps, tgt_heatmaps) + 1) * -0.5
        return (self.dice_wgt * self.dice_loss(in_seg, tgt_seg)) + (self.heatmap_wgt * torch.mean(ncc_losses))


# ----------------------------------------------------------------------------------------------
# Fused device versions: same arguments and value as the classes above, computed by
# the loss kernels of libfluorounet.so (csrc/kernels_loss.cuh) in one reduction pass + one gradient pass.
# CUDA fp32 tensors only; there is no CPU path here -- use the PyTorch classes above on the CPU.
# ----------------------------------------------------------------------------------------------
def _crop_origin(src, dst):
    """Window origin of util.center_crop (util.py:99-103)."""
    return int((src[-2] - dst[-2]) / 2), int((src[-1] - dst[-1]) / 2)


class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seg, heat, tgt_seg, tgt_heat, skip_bg, dice_wgt, heat_wgt):
        import ctypes as C
        from . import _capi
        L = _capi.lib()
        for name, t in (("seg", seg), ("heat", heat), ("tgt_seg", tgt_seg), ("tgt_heat", tgt_heat)):
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError(f"fused loss: {name} must be a CUDA tensor (no CPU fallback; use DiceAndHeatMapLoss2D)")
            if t.dtype != torch.float32 or t.dim() != 4:
                raise TypeError(f"fused loss: {name} must be a 4-D float32 tensor")
        if (heat is None) != (tgt_heat is None):
            raise ValueError("fused loss: heat-map prediction and target must come together")
        seg_full = seg.contiguous()
        heat_full = heat.contiguous() if heat is not None else None
        tgt_seg = tgt_seg if tgt_seg.stride(-1) == 1 else tgt_seg.contiguous()
        if tgt_heat is not None and tgt_heat.stride(-1) != 1:
            tgt_heat = tgt_heat.contiguous()
        B, NC, H, W = seg_full.shape
        Ht, Wt = tgt_seg.shape[-2:]
        if tuple(tgt_seg.shape[:2]) != (B, NC) or Ht > H or Wt > W:
            raise ValueError(f"fused loss: target {tuple(tgt_seg.shape)} does not fit prediction {tuple(seg_full.shape)}")
        NL = 0
        if heat_full is not None:
            NL = heat_full.shape[1]
            if tuple(heat_full.shape) != (B, NL, H, W) or tuple(tgt_heat.shape) != (B, NL, Ht, Wt):
                raise ValueError("fused loss: heat-map shapes disagree with the segmentation shapes")
        r0, c0 = _crop_origin((H, W), (Ht, Wt))
        d = _capi.FuLossDesc()
        esz = 4

        def put(prefix, t, off):
            setattr(d, prefix, t.data_ptr() + off * esz)
            getattr(d, prefix + "_stride")[:] = [t.stride(0), t.stride(1), t.stride(2)]
        put("seg", seg_full, r0 * W + c0)
        put("mask", tgt_seg, 0)
        if heat_full is not None:
            put("heat", heat_full, r0 * W + c0)
            put("heat_t", tgt_heat, 0)
        d.B, d.n_classes, d.num_lands, d.Ht, d.Wt = B, NC, NL, Ht, Wt
        d.skip_bg, d.dice_wgt, d.heat_wgt = int(skip_bg), float(dice_wgt), float(heat_wgt)
        sums = torch.empty(int(L.fu_loss_workspace_doubles(B, NC, NL)), device=seg.device, dtype=torch.float64)
        loss = torch.empty((), device=seg.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(seg.device).cuda_stream
        with torch.cuda.device(seg.device):       # the loss kernels launch on the current device
            rc = L.fu_loss_forward(C.byref(d), sums.data_ptr(), loss.data_ptr(), stream)
        if rc != 0:
            raise RuntimeError(f"fu_loss_forward failed ({rc}): {_capi.last_error(None)}")
        ctx.desc, ctx.geom = d, (H, W, r0, c0, NL)
        ctx.keep = (seg_full, heat_full, tgt_seg, tgt_heat, sums)     # the descriptor holds raw addresses of these
        return loss

    @staticmethod
    def backward(ctx, dloss):
        import ctypes as C
        from . import _capi
        L = _capi.lib()
        seg_full, heat_full, _, _, sums = ctx.keep
        H, W, r0, c0, NL = ctx.geom
        d_seg = torch.empty_like(seg_full)
        d_heat = torch.empty_like(heat_full) if NL > 0 else None
        dloss = dloss.contiguous().float()
        stream = torch.cuda.current_stream(seg_full.device).cuda_stream
        with torch.cuda.device(seg_full.device):
            rc = L.fu_loss_backward(C.byref(ctx.desc), sums.data_ptr(), dloss.data_ptr(), H, W, r0, c0, d_seg.data_ptr(),
                                    d_heat.data_ptr() if d_heat is not None else None, stream)
        if rc != 0:
            raise RuntimeError(f"fu_loss_backward failed ({rc}): {_capi.last_error(None)}")
        return d_seg, d_heat, None, None, None, None, None


class FusedDiceLoss2D(torch.nn.modules.loss._Loss):
    """DiceLoss2D (dice.py:14-55) on the device in one pass.  ``input`` may be the UNCROPPED network output:
    when it is larger than ``target`` the centre crop of train.py:414-417 (util.center_crop) is applied inside
    the kernel, and the gradient comes back full-size with zeros outside the window."""

    def __init__(self, skip_bg=True):
        super().__init__()
        self.skip_bg = skip_bg

    def forward(self, input, target):
        return _FusedLossFn.apply(input, None, target, None, self.skip_bg, 1.0, 0.0)


class FusedDiceAndHeatMapLoss2D(torch.nn.modules.loss._Loss):
    """DiceAndHeatMapLoss2D (dice.py:57-86) on the device; see FusedDiceLoss2D for the crop convention."""

    def __init__(self, skip_bg=True, heatmap_wgt=0.5):
        super().__init__()
        assert (heatmap_wgt > 1.0e-8) and (heatmap_wgt < (1 + 1.0e-8))
        self.skip_bg = skip_bg
        self.heatmap_wgt = heatmap_wgt
        self.dice_wgt = 1 - heatmap_wgt

    def forward(self, input, target):
        return _FusedLossFn.apply(input[0], input[1], target[0], target[1], self.skip_bg, self.dice_wgt, self.heatmap_wgt)
