"""fp32 engine vs fp64 oracle, per-tensor gradient error, small sweep (diagnostic)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle as O
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
def run(kw, B, S, show=False):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 1, S, S, generator=g)
    torch.manual_seed(0)
    net = pkg.UNet(precision="fp32", **kw).to(dev).train()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    out = net(x.to(dev))
    seg, heat = out if isinstance(out, tuple) else (out, None)
    d_seg = torch.randn(seg.shape, generator=g)
    d_heat = torch.randn(heat.shape, generator=g) if heat is not None else None
    l = (seg * d_seg.to(dev)).sum()
    if heat is not None: l = l + (heat * d_heat.to(dev)).sum()
    l.backward(); torch.cuda.synchronize()
    cfg = O.UNetConfig(**kw)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    ref = O.forward(sd64, cfg, x.double(), training=True, want_tape=True)
    rg = O.backward(sd64, cfg, ref["tape"], d_seg.double(), d_heat.double() if d_heat is not None else None)
    bad = [(n, rel(p.grad.cpu(), rg[n])) for n, p in net.named_parameters() if p.grad is not None]
    worst = max(bad, key=lambda t: t[1])
    print(kw.get("depth"), kw.get("wf"), "mp" if kw.get("max_pool") else "cv", "res" if kw.get("do_res", True) else "nores", B, S, "fwd %.1e" % rel(seg.detach().cpu(), ref["seg"]), "worst", worst, "n_bad", sum(1 for _, e in bad if e > 1e-4))
    if show:
        for n, e in bad:
            if e > 1e-4: print("    ", n, "%.2e" % e)
base = dict(n_classes=7, batch_norm=True, padding=True, max_pool=False, num_lands=14)
import itertools
for depth, wf, B, S in [(3,5,3,48),(3,4,3,48),(4,4,3,48),(4,5,1,48),(4,5,3,32),(3,5,1,16),(3,6,2,32),(3,5,3,24)]:
    run(dict(base, depth=depth, wf=wf), B, S, True)
run(dict(base, depth=3, wf=5, max_pool=True), 3, 48, True)
run(dict(base, depth=3, wf=5, do_res=False), 3, 48, True)
run(dict(base, depth=3, wf=5, num_lands=0), 3, 48, True)
