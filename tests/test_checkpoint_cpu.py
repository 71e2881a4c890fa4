"""Checkpoint compatibility with train.py (train.py:191-273 resume, :473-515 save_net): a checkpoint
dictionary written by the REAL reference modules (tests/golden/make_golden_ckpt.py) loads into the drop-in module, its
optimizer and schedule states load beside it, and what the drop-in writes back is key-for-key, value-for-value what
the reference wrote.  CPU only (the module is a parameter container until it sees a CUDA input)."""
import os

import torch

from conftest import GOLDEN, load_pkg, rel_l2
from oracle import unet_oracle as O

CKPT = os.path.join(GOLDEN, "ckpt_reference_small.pt")


def _net_from(state, pkg):
    # the constructor call of train.py:313 from the checkpoint's own fields (train.py:198-214)
    return pkg.UNet(n_classes=state['num-classes'], depth=state['depth'], wf=state['init-feats-exp'],
                    batch_norm=state['batch-norm'], padding=state['padding'], max_pool=not state['no-max-pool'],
                    num_lands=state['num-lands'], do_res=state['unet-use-res'], block_depth=state['unet-block-depth'])


def test_reference_checkpoint_loads_and_round_trips(tmp_path):
    pkg = load_pkg()
    state = torch.load(CKPT, weights_only=False)
    net = _net_from(state, pkg)
    missing = net.load_state_dict(state['model-state-dict'], strict=True)            # train.py:316
    assert not missing.missing_keys and not missing.unexpected_keys
    # optimizer / schedule resume exactly as train.py:333-347 does
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=state['opt-momentum'], weight_decay=state['opt-wgt-decay'],
                          nesterov=state['opt-nesterov'])
    opt.load_state_dict(state['optimizer-state-dict'])
    ref_bufs = state['optimizer-state-dict']['state']
    for i, p in enumerate(net.parameters()):
        if i in ref_bufs:                                                           # momentum buffers line up by index
            assert opt.state[p]['momentum_buffer'].shape == p.shape
            assert torch.equal(opt.state[p]['momentum_buffer'], ref_bufs[i]['momentum_buffer'])
    # the loaded weights are the trained ones: the oracle's forward on them reproduces the reference's probe outputs
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.UNetConfig(n_classes=7, depth=3, wf=3, batch_norm=True, padding=True, max_pool=False, num_lands=14)
    out = O.forward(sd, cfg, state['x-probe'], training=False)
    assert rel_l2(out["seg"], state['seg-probe']) < 1e-5 and rel_l2(out["heat"], state['heat-probe']) < 1e-5
    # save_net's dictionary written from the drop-in: same keys, same tensors (train.py:473-515)
    out_state = dict(state)
    out_state['model-state-dict'] = net.state_dict()
    out_state['optimizer-state-dict'] = opt.state_dict()
    path = os.path.join(tmp_path, "ckpt.pt")
    torch.save(out_state, path)
    back = torch.load(path, weights_only=False)
    assert list(back['model-state-dict'].keys()) == list(state['model-state-dict'].keys())
    for k, v in state['model-state-dict'].items():
        assert back['model-state-dict'][k].dtype == v.dtype and torch.equal(back['model-state-dict'][k], v), k
    assert back['optimizer-state-dict']['param_groups'][0]['params'] == state['optimizer-state-dict']['param_groups'][0]['params']
    # and, where the reference is available (authoring container), its own module accepts what we wrote
    ref_dir = "/root/reference/train_test_code"
    if os.path.isdir(ref_dir):
        import sys
        sys.path.insert(0, ref_dir)
        import unet as ref_unet
        rnet = ref_unet.UNet(n_classes=7, depth=3, wf=3, batch_norm=True, padding=True, max_pool=False, num_lands=14)
        rnet.load_state_dict(back['model-state-dict'])
        rnet.eval()
        with torch.no_grad():
            seg, heat = rnet(state['x-probe'])
        assert torch.equal(seg, state['seg-probe']) and torch.equal(heat, state['heat-probe'])
