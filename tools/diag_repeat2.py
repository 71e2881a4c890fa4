import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
dev = torch.device("cuda:0")
S, B = 192, 2
kw = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14)
x = torch.randn(B, 1, S, S, generator=torch.Generator().manual_seed(4)).to(dev)
def make(env):
    for k in ("FU_TC_V2",): os.environ.pop(k, None)
    os.environ.update(env)
    torch.manual_seed(0)
    return pkg.UNet(precision="bf16", **kw).to(dev).train()
ref_net = make({"FU_TC_V2": "0"})
with torch.no_grad(): ref_net(x)
names = ["enc0.r0", "enc0.r1", "down1", "enc1.r0", "enc1.z0", "enc1.r1", "down2", "enc2.r0"]
ref = {n: ref_net.debug_tensor(n).cpu() for n in names}
net = make({})
for rep in range(6):
    with torch.no_grad(): net(x)
    torch.cuda.synchronize()
    line = []
    for n in names:
        t = net.debug_tensor(n).cpu()
        d = (t - ref[n]).abs()
        nz = torch.nonzero(d > 0)
        line.append("%s:%d" % (n, nz.shape[0]))
        if n == "enc1.r0" and nz.shape[0]:
            pix = sorted(set((int(a), int(c), int(e)) for a, b_, c, e in nz.tolist()))
            line.append("pix(n,h,w)=%s" % pix[:12])
    print("rep", rep, " ".join(line))
