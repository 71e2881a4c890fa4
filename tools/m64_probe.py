"""Which TMEM row layout does an M = 64 tcgen05.mma use?  Runs the weight-gradient hook on thin layers and prints the
error against a torch-CPU reference; run with FU_TC_M64 = 0 (M = 128 reference), 1 and 2."""
import ctypes, importlib, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
L = pkg._capi.lib(); dev = torch.device("cuda:0")
p = lambda t: ctypes.c_void_p(t.data_ptr())
torch.manual_seed(0)
for (B, Cin, Cout, H, W, k) in [(2, 32, 32, 64, 64, 3), (2, 64, 32, 64, 64, 3), (2, 64, 64, 48, 48, 3), (2, 128, 64, 48, 48, 3),
                                (2, 64, 32, 64, 64, 1), (4, 256, 64, 16, 16, 3)]:
    x = torch.randn(B, Cin, H, W).bfloat16().float()
    dy = torch.randn(B, Cout, H, W).bfloat16().float()
    xr = x.clone().requires_grad_(False)
    w = torch.zeros(Cout, Cin, k, k, requires_grad=True)
    y = F.conv2d(x, w, padding=k // 2)
    (y * dy).sum().backward()
    ref = w.grad
    xg = x.permute(0, 2, 3, 1).contiguous().to(dev).bfloat16()
    dyg = dy.permute(0, 2, 3, 1).contiguous().to(dev).bfloat16()
    dw = torch.empty(Cout, Cin, k, k, device=dev)
    wz = torch.zeros(Cout, Cin, k, k, device=dev)
    rc = L.fu_test_conv(1, 1, 2, B, H, W, Cin, Cout, k, 1, k // 2, 0, p(xg), p(wz), None, None, p(dyg), p(dw), None, None)
    assert rc == 0, pkg._capi.last_error(None)
    torch.cuda.synchronize()
    err = float((dw.cpu() - ref).norm() / ref.norm())
    print(f"FU_TC_M64={os.environ.get('FU_TC_M64', '0')} shape {(B, Cin, Cout, H, W, k)} rel err {err:.3e}", flush=True)
