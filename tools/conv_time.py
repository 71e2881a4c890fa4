"""Time the tensor-core conv kernels on one layer shape with CUDA events (kernel + weight pack + alloc
excluded by timing many reps of the engine-free hook is not possible; so this uses a tiny network-free
loop: fu_test_conv called once per rep, minus a baseline).  usage: conv_time.py B Cin Cout H W k mode"""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
B, Cin, Cout, H, W, k, mode = [int(a) for a in sys.argv[1:8]]
L = pkg._capi.lib(); dev = torch.device("cuda:0")
x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5)
b = torch.randn(Cout, device=dev)
dy = torch.randn(B, H, W, Cout, device=dev).bfloat16()
y = torch.empty(B, H, W, Cout if mode == 0 else Cin, device=dev, dtype=torch.bfloat16)
dw = torch.empty_like(w)
p = lambda t: ctypes.c_void_p(t.data_ptr())
from torch.profiler import profile, ProfilerActivity
def run():
    rc = L.fu_test_conv(1, 1, mode, B, H, W, Cin, Cout, k, 1, k // 2, 1, p(x), p(w), p(b), p(y), p(dy), p(dw), None, None)
    assert rc == 0, pkg._capi.last_error(None)
for _ in range(2): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): run()
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "tc_" in e.key and "pack" not in e.key and "unpack" not in e.key:
        fl = 2.0 * B * H * W * Cin * Cout * k * k
        us = e.device_time_total / e.count
        print("%s env=%s %-28s %8.1f us  %7.1f TFLOP/s" % (sys.argv[1:8], {k2: v for k2, v in os.environ.items() if k2.startswith("FU_TC")}, e.key[:28], us, fl / us / 1e6))
