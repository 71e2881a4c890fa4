"""Run one conv layer through the tensor-core kernels via fu_test_conv (for ncu).  usage:
conv_probe.py B Cin Cout H W k mode(0 fwd,1 dgrad,2 wgrad) [reps]"""
import ctypes, importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
B, Cin, Cout, H, W, k, mode = [int(a) for a in sys.argv[1:8]]
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
L = pkg._capi.lib(); dev = torch.device("cuda:0")
x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5)
b = torch.randn(Cout, device=dev)
dy = torch.randn(B, H, W, Cout, device=dev).bfloat16()
y = torch.empty(B, H, W, Cout if mode == 0 else Cin, device=dev, dtype=torch.bfloat16)
dw = torch.empty_like(w)
stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for i in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = L.fu_test_conv(1, 1, mode, B, H, W, Cin, Cout, k, 1, k // 2, 1, p(x), p(w), p(b), p(y), p(dy), p(dw), p(stats) if mode == 0 else None, None)
    assert rc == 0, pkg._capi.last_error(None)
    torch.cuda.synchronize()
    print("rep", i, "host ms (incl. pack/alloc)", (time.perf_counter() - t0) * 1e3)
