"""Per-layer forward/backward intermediates of the engine vs the fp64 oracle (autograd on the oracle's
own forward).  Diagnostic; run on the GPU box.  usage: diag_layers.py depth wf B S [precision]"""
import importlib, os, re, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle as O
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
dev = torch.device("cuda:0")
depth, wf, B, S = [int(a) for a in sys.argv[1:5]]
precision = sys.argv[5] if len(sys.argv) > 5 else "fp32"
rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / (b.double().norm() + 1e-30))
kw = dict(n_classes=7, batch_norm=True, padding=True, max_pool=False, num_lands=14, depth=depth, wf=wf)
g = torch.Generator().manual_seed(2)
x = torch.randn(B, 1, S, S, generator=g)
torch.manual_seed(0)
net = pkg.UNet(precision=precision, **kw).to(dev).train()
sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
seg, heat = net(x.to(dev))
d_seg = torch.randn(seg.shape, generator=g); d_heat = torch.randn(heat.shape, generator=g)
((seg * d_seg.to(dev)).sum() + (heat * d_heat.to(dev)).sum()).backward(); torch.cuda.synchronize()
cfg = O.UNetConfig(**kw)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
for k, v in sd64.items():
    if v.is_floating_point() and "running" not in k: v.requires_grad_(True)
ref = O.forward(sd64, cfg, x.double(), training=True, want_tape=True)
tape = ref["tape"]
named = []   # (engine name, oracle tensor, is_grad)
cur = None
for ent in tape:
    if ent[0] == "conv" and ".block." in ent[1]:
        m = re.match(r"(down_path|up_path)\.(\d+)\.(?:conv_block\.)?block\.(\d+)", ent[1])
        blk = ("enc" if m.group(1) == "down_path" else "dec") + m.group(2)
        i = int(m.group(3)) // 3
        cur = (blk, i)
        if i > 0 and ent[2].requires_grad:
            ent[2].retain_grad(); named.append((f"{blk}.z{i-1}", ent[2], False)); named.append((f"{blk}.dz{i}", ent[2], True))
    elif ent[0] == "relu" and cur is not None:
        ent[1].retain_grad(); named.append((f"{cur[0]}.r{cur[1]}", ent[1], False)); named.append((f"{cur[0]}.dy{cur[1]}", ent[1], "relu"))
(ref["seg"] * d_seg.double()).sum().add((ref["heat"] * d_heat.double()).sum()).backward()
print("config", kw, B, S, precision)
for name, t, isg in named:
    try:
        v = net.debug_tensor(name)
    except KeyError as e:
        continue
    if isg is False: r = t.detach()
    elif isg is True: r = t.grad
    else: r = t.grad * (t.detach() > 0)
    e = rel(v, r)
    flag = "  <<<<" if e > (1e-4 if precision == "fp32" else 5e-2) else ""
    print("%-12s %-22s err %.3e%s" % (name, tuple(v.shape), e, flag))
    if flag and isg is not False and precision == "fp32":
        dlt = (v.cpu().double() - r).abs()
        idx = torch.nonzero(dlt > 1e-3 * r.abs().max())
        print("      n_bad_elems", idx.shape[0], "of", dlt.numel(), "first", idx[:6].tolist())
