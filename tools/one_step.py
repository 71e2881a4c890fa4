"""Two training steps of the bench workload (for ncu captures).  usage: one_step.py [batch] [size] [steps]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
kw = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14)
torch.manual_seed(0)
net = pkg.UNet(precision="bf16", **kw).to(dev).train()
crit = pkg.FusedDiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True, fused=True)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 1, S, S, generator=g).to(dev)
T = S - 12
mask = torch.nn.functional.one_hot(torch.randint(0, 7, (B, T, T), generator=g), 7).permute(0, 3, 1, 2).float().to(dev)
heat = torch.rand(B, 14, T, T, generator=g).to(dev)
for i in range(steps):
    opt.zero_grad(set_to_none=True)
    seg, hm = net(x)
    loss = crit((seg, hm), (mask, heat))
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("done", float(loss))
