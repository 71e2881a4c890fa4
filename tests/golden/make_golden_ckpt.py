"""A checkpoint written the way train.py:473-515 (save_net) writes it, by the REAL reference modules
(authoring container only):

    python tests/golden/make_golden_ckpt.py

A small dual-head network (same constructor path as train.py:313) is trained for two SGD steps on the CPU with the
reference's unet.UNet, dice.DiceAndHeatMapLoss2D and warm_restarts_lr.WarmRestartLR, then saved with the same
dictionary keys.  tests/test_checkpoint_cpu.py loads it into the drop-in module and back."""
import os
import sys

import torch

REF = "/root/reference/train_test_code"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    import unet
    import dice
    import util
    import warm_restarts_lr
    torch.manual_seed(5)
    kw = dict(n_classes=7, depth=3, wf=3, batch_norm=True, padding=True, max_pool=False, num_lands=14,
              do_res=True, block_depth=2)
    net = unet.UNet(**kw)                                                                  # train.py:313
    criterion = dice.DiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)                  # train.py:324
    optimizer = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=True)   # :333
    lr_sched = warm_restarts_lr.WarmRestartLR(optimizer, init_run_period_epochs=3, growth_factor=2)          # :337
    g = torch.Generator().manual_seed(6)
    net.train()
    loss = None
    for it in range(2):
        x = torch.randn(2, 1, 32, 32, generator=g)
        mask = torch.nn.functional.one_hot(torch.randint(0, 7, (2, 28, 28), generator=g), 7).permute(0, 3, 1, 2).float()
        heat = torch.rand(2, 14, 28, 28, generator=g)
        optimizer.zero_grad()
        seg, hm = net(x)
        loss = criterion((util.center_crop(seg, mask.shape), util.center_crop(hm, heat.shape)), (mask, heat))
        loss.backward()
        optimizer.step()
        lr_sched.intra_epoch_step((it + 1) / 2)
    lr_sched.step()
    # one more forward of the TRAINED network, stored so that the loaded weights can be checked functionally
    net.eval()
    xv = torch.randn(1, 1, 32, 32, generator=g)
    with torch.no_grad():
        seg_v, heat_v = net(xv)
    state = {'epoch': 1, 'model-state-dict': net.state_dict(), 'optim-type': 'sgd',
             'optimizer-state-dict': optimizer.state_dict(), 'scheduler-state-dict': lr_sched.state_dict(),
             'loss': loss.detach(), 'best-valid-loss': 0.5, 'save-best-valid': True, 'num-classes': 7, 'depth': 3,
             'init-feats-exp': 3, 'batch-norm': True, 'padding': True, 'no-max-pool': True, 'pad-img-size': 32,
             'batch-size': 2, 'data-aug': False, 'opt-nesterov': True, 'opt-momentum': 0.9, 'opt-wgt-decay': 1e-4,
             'num-lands': 14, 'heat-coeff': 0.5, 'use-dice-valid': False, 'unet-use-res': True, 'unet-block-depth': 2,
             'lrs-meth': 'cos', 'lrs-num-epochs': 3, 'lrs-growth-factor': 2, 'lrs-max-num-restarts': -1,
             'lrs-save-restart-net-prefix': '', 'lrs-save-after-n-restarts': 0, 'lrs-num-restarts': 0, 'lrs-patience': 10,
             'lrs-cooldown': 10, 'checkpoint-freq': 1, 'train-idx': None, 'valid-idx': None,
             # not part of save_net: a probe input and the reference's outputs on it
             'x-probe': xv, 'seg-probe': seg_v, 'heat-probe': heat_v}
    torch.save(state, os.path.join(HERE, "ckpt_reference_small.pt"))
    print("written", os.path.getsize(os.path.join(HERE, "ckpt_reference_small.pt")), "bytes")


if __name__ == "__main__":
    main()
