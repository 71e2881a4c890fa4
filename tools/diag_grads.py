"""Per-tensor gradient error of the bf16 tensor-core path and the bf16 CUDA-core path against the
fp64 oracle (diagnostic; run on the GPU box)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle as O
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
dev = torch.device("cuda:0")
kw = dict(n_classes=7, depth=4, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14)
g = torch.Generator().manual_seed(2)
x = torch.randn(3, 1, 48, 48, generator=g)
d_seg = torch.randn(3, 7, 48, 48, generator=torch.Generator().manual_seed(3))
d_heat = torch.randn(3, 14, 48, 48, generator=torch.Generator().manual_seed(4))
res = {}
for mode in ("tc", "simt", "fp32"):
    os.environ.pop("FU_TC_DISABLE", None)
    if mode == "simt":
        os.environ["FU_TC_DISABLE"] = "1"
    torch.manual_seed(0)
    net = pkg.UNet(precision="fp32" if mode == "fp32" else "bf16", **kw).to(dev).train()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    seg, heat = net(x.to(dev))
    ((seg * d_seg.to(dev)).sum() + (heat * d_heat.to(dev)).sum()).backward()
    torch.cuda.synchronize()
    res[mode] = (seg.detach().cpu(), heat.detach().cpu(), {n: p.grad.cpu() for n, p in net.named_parameters() if p.grad is not None})
cfg = O.UNetConfig(**kw)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
ref = O.forward(sd64, cfg, x.double(), training=True, want_tape=True)
rg = O.backward(sd64, cfg, ref["tape"], d_seg.double(), d_heat.double())
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
print("seg/heat err  tc: %.3e %.3e  simt: %.3e %.3e  fp32: %.3e %.3e" % (
    rel(res["tc"][0], ref["seg"]), rel(res["tc"][1], ref["heat"]), rel(res["simt"][0], ref["seg"]),
    rel(res["simt"][1], ref["heat"]), rel(res["fp32"][0], ref["seg"]), rel(res["fp32"][1], ref["heat"])))
names = list(rg.keys())
flat = lambda d: torch.cat([d[n].flatten().double() for n in names])
fr = flat(rg)
for m in ("tc", "simt", "fp32"):
    f = flat(res[m][2])
    print(m, "flat rel err %.3e  cosine %.6f" % (rel(f, fr), float(torch.dot(f, fr) / (f.norm() * fr.norm()))))
print("%-45s %10s %10s %10s %12s" % ("tensor", "tc", "simt", "fp32", "|g|"))
for n in names:
    print("%-45s %10.3e %10.3e %10.3e %12.3e" % (n, rel(res["tc"][2][n], rg[n]), rel(res["simt"][2][n], rg[n]),
                                                 rel(res["fp32"][2][n], rg[n]), float(rg[n].norm())))
