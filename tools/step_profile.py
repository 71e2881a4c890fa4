"""Device time of ONE full training step of the bench workload, split into engine kernels and torch (aten)
kernels (loss, autograd of the crop, SGD), via torch.profiler.  usage: step_profile.py [batch] [size]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
dev = torch.device("cuda:0")
kw = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14)
torch.manual_seed(0)
net = pkg.UNet(precision="bf16", **kw).to(dev).train()
crit = pkg.FusedDiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True, fused=True)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 1, S, S, generator=g).to(dev)
T = S - 12
mask = torch.nn.functional.one_hot(torch.randint(0, 7, (B, T, T), generator=g), 7).permute(0, 3, 1, 2).float().contiguous().to(dev)
heat = torch.rand(B, 14, T, T, generator=g).to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    if os.environ.get("STEP_IN_HEADS", "0") == "1":
        loss = net.forward_loss(x, (mask, heat), crit)      # the loss inside the heads kernels
    else:
        seg, hm = net(x)
        loss = crit((seg, hm), (mask, heat))
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
n = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(n): step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    dt = getattr(e, "device_time_total", 0) or 0
    if dt > 0 and e.device_type.name == "CUDA":
        rows.append((dt / n, e.count / n, e.key))
rows.sort(reverse=True)
eng = sum(r[0] for r in rows if "fu::" in r[2])
oth = sum(r[0] for r in rows if "fu::" not in r[2])
print(f"# per step: engine kernels {eng/1e3:.3f} ms, other device work {oth/1e3:.3f} ms")
for dt, cnt, key in rows[:60]:
    print("%9.1f us  x%-6.1f %s" % (dt, cnt, key[:110]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
print(f"# wall (events) {e0.elapsed_time(e1)/20:.3f} ms/step")
