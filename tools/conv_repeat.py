"""Bitwise repeatability of the tensor-core conv kernels on one layer + agreement between halo and v1 kernels."""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
L = pkg._capi.lib(); dev = torch.device("cuda:0")
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
def run(B, Cin, Cout, H, W, mode, env):
    old = {k: os.environ.get(k) for k in env}; os.environ.update(env)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H, W, Cin, generator=g).bfloat16().to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    dy = torch.randn(B, H, W, Cout, generator=g).bfloat16().to(dev)
    y = torch.zeros(B, H, W, Cout if mode == 0 else Cin, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    rc = L.fu_test_conv(1, 1, mode, B, H, W, Cin, Cout, 3, 1, 1, 1, p(x), p(w), p(b), p(y), p(dy), None, p(st) if mode == 0 else None, None)
    assert rc == 0, pkg._capi.last_error(None)
    torch.cuda.synchronize()
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    return y.float().cpu(), st.cpu()
for shp in [(32, 32, 32, 192, 192), (32, 64, 64, 96, 96), (32, 128, 128, 48, 48), (8, 32, 32, 192, 192), (4, 128, 128, 48, 48)]:
    for mode in (0, 1):
        ys = [run(*shp, mode, {}) for _ in range(3)]
        yv1 = run(*shp, mode, {"FU_TC_V2": "0"})
        d01 = float((ys[0][0] - ys[1][0]).abs().max()); d02 = float((ys[0][0] - ys[2][0]).abs().max())
        dv = float((ys[0][0] - yv1[0]).abs().max())
        nbad = int(((ys[0][0] - yv1[0]).abs() > 0.05).sum())
        sd = float((ys[0][1] - ys[1][1]).abs().max() / (ys[0][1].abs().max() + 1e-30)) if mode == 0 else 0.0
        sv = float((ys[0][1] - yv1[1]).abs().max() / (yv1[1].abs().max() + 1e-30)) if mode == 0 else 0.0
        print(shp, "mode", mode, "run-to-run max|d|", d01, d02, "| vs v1 max|d|", dv, "n>0.05:", nbad, "| stats rel r2r", "%.1e" % sd, "vs v1 %.1e" % sv)
        if nbad:
            idx = torch.nonzero((ys[0][0] - yv1[0]).abs() > 0.05)
            print("    first bad idx (n,h,w,c):", idx[:8].tolist())
