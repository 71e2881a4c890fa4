"""smoke()-shaped error measurement: engine vs oracle on the tiny depth-3 wf-4 network, several seeds."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
from oracle import unet_oracle as O
dev = torch.device("cuda:0")
def rel(a, b): return float((a.double().cpu() - b.double()).norm() / (b.double().norm() + 1e-30))
for wf in (4, 5):
    kw = dict(n_classes=7, depth=3, wf=wf, batch_norm=True, padding=True, max_pool=False, num_lands=14, do_res=True, block_depth=2)
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        net = pkg.UNet(precision="bf16", **kw).to(dev); net.train()
        x = torch.randn(2, 1, 32, 32)
        sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
        seg, heat = net(x.to(dev))
        g = torch.Generator().manual_seed(1)
        d_seg, d_heat = torch.randn(seg.shape, generator=g), torch.randn(heat.shape, generator=g)
        ((seg * d_seg.to(dev)).sum() + (heat * d_heat.to(dev)).sum()).backward()
        torch.cuda.synchronize()
        cfg = O.UNetConfig(**kw)
        ref = O.forward(sd, cfg, x, training=True, want_tape=True)
        rg = O.backward(sd, cfg, ref["tape"], d_seg, d_heat)
        names = [n for n, p in net.named_parameters() if p.grad is not None]
        e_g = rel(torch.cat([dict(net.named_parameters())[n].grad.flatten() for n in names]), torch.cat([rg[n].flatten() for n in names]))
        worst = sorted(((rel(dict(net.named_parameters())[n].grad, rg[n]), n) for n in names), reverse=True)[:3]
        print(f"env={ {k:v for k,v in os.environ.items() if k.startswith('FU_')} } wf={wf} seed={seed} seg {rel(seg.detach(), ref['seg']):.2e} heat {rel(heat.detach(), ref['heat']):.2e} flat grad {e_g:.3f} worst {[(round(a,3),b) for a,b in worst]}", flush=True)
