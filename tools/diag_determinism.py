import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
def run(S, B, lands, precision="bf16"):
    kw = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=lands)
    torch.manual_seed(0)
    net = pkg.UNet(precision=precision, **kw).to(dev).train()
    x = torch.randn(B, 1, S, S, generator=torch.Generator().manual_seed(4)).to(dev)
    gs = []
    for rep, scale in enumerate((1.0, 1.0, 2.0)):
        net.zero_grad()
        o = net(x); o = o if isinstance(o, tuple) else (o,)
        ups = [torch.randn(t.shape, generator=torch.Generator().manual_seed(7 + i)).to(dev) * scale for i, t in enumerate(o)]
        sum((t * u).sum() for t, u in zip(o, ups)).backward()
        torch.cuda.synchronize()
        gs.append({n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None})
    bad = [(n, rel(gs[1][n], gs[0][n]), rel(gs[2][n], 2 * gs[0][n])) for n in gs[0]]
    worst_rep = max(bad, key=lambda t: t[1]); worst_lin = max(bad, key=lambda t: t[2])
    print(f"S={S} B={B} lands={lands} {precision}: repeatability worst {worst_rep[0]} {worst_rep[1]:.2e} | linearity worst {worst_lin[0]} {worst_lin[2]:.2e}")
    for n, a, b in bad:
        if a > 1e-3 or b > 1e-3: print("    ", n, "%.2e %.2e" % (a, b))
for S, B, lands in [(192, 2, 14), (384, 1, 14), (736, 2, 0), (1440, 1, 14)]:
    try:
        run(S, B, lands)
    except Exception as e:
        print("S", S, "FAILED:", repr(e)[:300])
