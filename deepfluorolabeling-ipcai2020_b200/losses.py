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
