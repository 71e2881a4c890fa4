"""Run a list of tensor-core conv layer shapes through the kernel test hook in ONE process (for ncu captures
and for CUDA-event timing without paying the interpreter start-up per shape).
usage: conv_shapes.py [--time] "B Cin Cout H W k mode" ...      (mode 0 fwd, 1 dgrad, 2 wgrad)"""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
L = pkg._capi.lib(); dev = torch.device("cuda:0")
args = sys.argv[1:]
timed = False
if args and args[0] == "--time":
    timed = True; args = args[1:]
p = lambda t: ctypes.c_void_p(t.data_ptr())
for spec in args:
    f = [int(a) for a in spec.split()]
    B, Cin, Cout, H, W, k, mode = f[:7]
    stride = f[7] if len(f) > 7 else 1          # 2: Conv2d(k=2,s=2) on x (B,H,W,Cin); -2: ConvTranspose2d(k=2,s=2)
    Ho, Wo = (H // 2, W // 2) if stride == 2 else ((2 * H, 2 * W) if stride == -2 else (H, W))
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(*((Cin, Cout, k, k) if stride == -2 else (Cout, Cin, k, k)), device=dev) / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, device=dev)
    dy = torch.randn(B, Ho, Wo, Cout, device=dev).bfloat16()
    y = torch.empty(*((B, Ho, Wo, Cout) if mode == 0 else (B, H, W, Cin)), device=dev, dtype=torch.bfloat16)
    dw = torch.empty_like(w)
    pad = k // 2 if stride == 1 else 0
    # CONV_STATS=1: forward launches also accumulate the BatchNorm statistics in their epilogue (as in a training step)
    st = torch.zeros(2 * Cout, device=dev, dtype=torch.float64) if (os.environ.get("CONV_STATS") == "1" and mode == 0) else None
    def run():
        rc = L.fu_test_conv(1, 1, mode, B, H, W, Cin, Cout, k, stride, pad, 1 if stride == 1 else 0, p(x), p(w), p(b), p(y), p(dy), p(dw), p(st) if st is not None else None, None)
        assert rc == 0, pkg._capi.last_error(None)
    run()
    if timed:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5): run()
            torch.cuda.synchronize()
        for e in prof.key_averages():
            if "tc_" in e.key and "batched" not in e.key:
                fl = 2.0 * B * (H * W if stride != 2 else Ho * Wo) * Cin * Cout * k * k
                us = e.device_time_total / e.count
                print("%-26s %-30s %8.1f us %7.1f TFLOP/s" % (spec, e.key[:30], us, fl / us / 1e6), flush=True)
    else:
        run()
torch.cuda.synchronize()
print("done")
