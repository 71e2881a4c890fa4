import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
dev = torch.device("cuda:0")
S, B = int(sys.argv[1]), int(sys.argv[2])
kw = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14)
torch.manual_seed(0)
net = pkg.UNet(precision="bf16", **kw).to(dev).train()
x = torch.randn(B, 1, S, S, generator=torch.Generator().manual_seed(4)).to(dev)
names = []
for l in range(6):
    for i in range(2):
        names += [f"enc{l}.r{i}", f"enc{l}.dy{i}"]
    names += [f"enc{l}.z0", f"enc{l}.dz1"]
for j in range(5):
    for i in range(2):
        names += [f"dec{j}.r{i}", f"dec{j}.dy{i}"]
    names += [f"dec{j}.z0", f"dec{j}.dz1"]
names += [f"cat{l}" for l in range(5)] + [f"d_cat{l}" for l in range(5)] + [f"down{l}" for l in range(1, 6)] + [f"d_down{l}" for l in range(1, 6)] + ["bott", "d_bott", "hcat", "d_hcat"]
snaps = []
for rep in range(2):
    net.zero_grad()
    seg, heat = net(x)
    ups = [torch.randn(t.shape, generator=torch.Generator().manual_seed(7 + i)).to(dev) for i, t in enumerate((seg, heat))]
    ((seg * ups[0]).sum() + (heat * ups[1]).sum()).backward()
    torch.cuda.synchronize()
    snaps.append({n: net.debug_tensor(n).cpu() for n in names})
    snaps[-1]["seg"] = seg.detach().cpu(); snaps[-1]["heat"] = heat.detach().cpu()
for n in ["seg", "heat"] + names:
    a, b = snaps[0][n], snaps[1][n]
    d = (a - b).abs()
    if float(d.max()) > 0:
        idx = torch.nonzero(d > 0)
        print("%-10s differs: n_diff %d of %d, max %.3e, first %s" % (n, idx.shape[0], d.numel(), float(d.max()), idx[:4].tolist()))
print("done")
