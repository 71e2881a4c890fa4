"""Per-layer device time of one training step (CUDA events inside the engine).
usage: profile_layers.py [batch] [size] [precision] > profiles/xxx.txt"""
import importlib, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda:0")
kw = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14)
torch.manual_seed(0)
net = pkg.UNet(precision=precision, **kw).to(dev).train()
x = torch.randn(B, 1, S, S, device=dev)
def step():
    net.zero_grad(set_to_none=True)
    seg, heat = net(x)
    (seg.square().mean() + heat.square().mean()).backward()
for _ in range(3): step()
torch.cuda.synchronize()
net.profile(True)
n = 3
for _ in range(n): step()
torch.cuda.synchronize()
rep = net.profile_report(); net.profile(False)
tot = sum(r["ms"] for r in rep) / n
print(f"# B={B} S={S} {precision}: engine kernel time per step {tot:.3f} ms, {len(rep)} (tag,kernel) rows")
print("%-34s %-24s %5s %9s %9s %9s" % ("tag", "kernel", "n", "ms", "TFLOP/s", "GB/s"))
for r in sorted(rep, key=lambda r: -r["ms"]):
    ms = r["ms"] / n
    tf = r["flops"] / n / (ms * 1e-3) / 1e12 if ms > 0 else 0
    gb = r["bytes"] / n / (ms * 1e-3) / 1e9 if ms > 0 else 0
    print("%-34s %-24s %5d %9.4f %9.1f %9.1f" % (r["tag"], r["kernel"][:24], r["launches"] // n, ms, tf, gb))
