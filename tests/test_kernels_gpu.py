"""Kernel-level parity through the C ABI test hook (fu_test_conv): each convolution kernel of the
engine against torch-CPU fp32 convolutions on odd shapes (ragged tiles, channel tails)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from conftest import load_pkg, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    assert torch.cuda.is_available()
    return load_pkg()


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def run_conv(pkg, precision, impl, mode, x, w, bias, dy, k, stride, pad, relu=0, want_stats=False):
    """x: (B,Cin,H,W) cpu fp32; returns torch cpu fp32 result in NCHW."""
    L = pkg._capi.lib()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if precision == 1 else torch.float32
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    xg = x.permute(0, 2, 3, 1).contiguous().to(dev, dt)
    wg = w.contiguous().to(dev)
    bg = bias.to(dev) if bias is not None else None
    dyg = dy.permute(0, 2, 3, 1).contiguous().to(dev, dt) if dy is not None else None
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev) if want_stats else None
    if mode == 0:
        out = torch.empty(B, Ho, Wo, Cout, device=dev, dtype=dt)
        dw = None
    elif mode == 1:
        out = torch.empty(B, H, W, Cin, device=dev, dtype=dt)
        dw = None
    else:
        out = None
        dw = torch.empty_like(wg)
    rc = L.fu_test_conv(precision, impl, mode, B, H, W, Cin, Cout, k, stride, pad, relu, _p(xg), _p(wg), _p(bg),
                        _p(out), _p(dyg), _p(dw), _p(stats), None)
    assert rc == 0, pkg._capi.last_error(None)
    torch.cuda.synchronize()
    if mode == 2:
        return dw.cpu()
    res = out.float().cpu().permute(0, 3, 1, 2).contiguous()
    return (res, stats.cpu()) if want_stats else res


SHAPES = [  # B, Cin, Cout, H, W, k, stride, pad
    (2, 1, 8, 12, 20, 3, 1, 1),
    (3, 16, 24, 9, 7, 3, 1, 1),
    (1, 7, 5, 6, 6, 1, 1, 0),
    (2, 8, 8, 8, 12, 2, 2, 0),
    (2, 64, 96, 24, 24, 3, 1, 1),
    (1, 39, 21, 10, 10, 1, 1, 0),
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("precision", [0, 1])
def test_simt_conv_fwd_dgrad_wgrad(pkg, shape, precision):
    B, Cin, Cout, H, W, k, stride, pad = shape
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    if precision == 1:
        x = x.bfloat16().float()
    tol = 1e-5 if precision == 0 else 1e-2
    y_ref = F.conv2d(x, w, b, stride=stride, padding=pad)
    y, stats = run_conv(pkg, precision, 0, 0, x, w, b, None, k, stride, pad, relu=1, want_stats=True)
    assert rel_l2(y, torch.relu(y_ref)) < tol
    assert rel_l2(stats[:Cout], y.double().sum(dim=(0, 2, 3))) < 1e-4
    assert rel_l2(stats[Cout:], (y.double() ** 2).sum(dim=(0, 2, 3))) < 1e-4
    dy = torch.randn(y_ref.shape, generator=g)
    if precision == 1:
        dy = dy.bfloat16().float()
    dw = run_conv(pkg, precision, 0, 2, x, w, None, dy, k, stride, pad)
    dw_ref = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=stride, padding=pad)
    assert rel_l2(dw, dw_ref) < tol
    if stride == 1:
        dx = run_conv(pkg, precision, 0, 1, x, w, None, dy, k, stride, pad)
        dx_ref = torch.nn.grad.conv2d_input(x.shape, w, dy, stride=stride, padding=pad)
        assert rel_l2(dx, dx_ref) < tol


TC_SHAPES = [  # B, Cin, Cout, H, W, k
    (2, 64, 64, 16, 16, 3),      # KC=64 (128B swizzle), BN=64
    (2, 32, 32, 16, 16, 3),      # KC=32 (64B swizzle), BN=32, 64-byte store boxes
    (1, 64, 128, 24, 24, 3),     # BN=128: two 64-channel store boxes
    (3, 128, 256, 12, 12, 3),    # two N tiles, several images per pixel tile
    (32, 64, 64, 6, 6, 3),       # 6x6 bottom level: (2,2,32) pixel tiles
    (2, 64, 32, 20, 12, 1),      # 1x1
    (2, 96, 64, 10, 14, 3),      # Cin not a multiple of 64 -> 3 chunks of 32
    (3, 64, 64, 9, 7, 3),        # ragged: tiles overhang the image and the batch
    (2, 256, 128, 24, 24, 3),    # long K loop (36 iterations) through the ring twice
]


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tc_conv_fwd_dgrad(pkg, shape):
    """tcgen05 implicit-GEMM conv (forward + data gradient) against torch-CPU fp32 on bf16-representable
    operands: only fp32 accumulation order and the bf16 rounding of the output differ."""
    B, Cin, Cout, H, W, k = shape
    pad = k // 2
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(Cout, generator=g)
    y_ref = torch.relu(F.conv2d(x, w, b, padding=pad))
    y, stats = run_conv(pkg, 1, 1, 0, x, w, b, None, k, 1, pad, relu=1, want_stats=True)
    err = rel_l2(y, y_ref)
    assert err < 6e-3, err                      # bf16 output rounding: ~2^-9 relative
    # exactness check modulo output rounding: compare against the reference rounded the same way
    assert rel_l2(y, y_ref.bfloat16().float()) < 2e-3
    assert rel_l2(stats[:Cout], y.double().sum(dim=(0, 2, 3))) < 1e-4
    assert rel_l2(stats[Cout:], (y.double() ** 2).sum(dim=(0, 2, 3))) < 1e-4
    dy = torch.randn(B, Cout, H, W, generator=g).bfloat16().float()
    dx = run_conv(pkg, 1, 1, 1, x, w, None, dy, k, 1, pad)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, w, dy, stride=1, padding=pad)
    assert rel_l2(dx, dx_ref) < 6e-3
    # weight gradient: bf16 operands, fp32 accumulation over all pixels (split-K + fp32 atomics)
    dw = run_conv(pkg, 1, 1, 2, x, w, None, dy, k, 1, pad)
    dw_ref = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=1, padding=pad)
    assert rel_l2(dw, dw_ref) < 1e-4, rel_l2(dw, dw_ref)


def _nhwc(t, dt, dev):
    return t.permute(0, 2, 3, 1).contiguous().to(dev, dt)


S2_SHAPES = [  # B, Cin, Cout, H, W  (H, W = spatial size of the layer INPUT)
    (2, 64, 64, 16, 16),
    (3, 32, 32, 8, 24),
    (2, 128, 128, 12, 12),
    (5, 256, 256, 6, 6),
    (1, 64, 32, 10, 6),
]


@pytest.mark.parametrize("shape", S2_SHAPES)
def test_tc_down_conv_2x2_s2(pkg, shape):
    """Conv2d(C,C,2,stride 2) (unet.py:93): strided-gather forward, scatter data gradient, weight gradient."""
    B, Cin, Cout, H, W = shape
    L, dev = pkg._capi.lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, 2, 2, generator=g) / (Cin * 4) ** 0.5).bfloat16().float()
    b = torch.randn(Cout, generator=g)
    dy = torch.randn(B, Cout, H // 2, W // 2, generator=g).bfloat16().float()
    xg, wg, bg, dyg = _nhwc(x, torch.bfloat16, dev), w.to(dev), b.to(dev), _nhwc(dy, torch.bfloat16, dev)
    y = torch.empty(B, H // 2, W // 2, Cout, device=dev, dtype=torch.bfloat16)
    dx = torch.empty(B, H, W, Cin, device=dev, dtype=torch.bfloat16)
    dw = torch.empty_like(wg)
    for mode, out in ((0, y), (1, dx), (2, None)):
        rc = L.fu_test_conv(1, 1, mode, B, H, W, Cin, Cout, 2, 2, 0, 0, _p(xg), _p(wg), _p(bg), _p(out), _p(dyg), _p(dw), None, None)
        assert rc == 0, pkg._capi.last_error(None)
    torch.cuda.synchronize()
    assert rel_l2(y.float().cpu().permute(0, 3, 1, 2), F.conv2d(x, w, b, stride=2)) < 6e-3
    assert rel_l2(dx.float().cpu().permute(0, 3, 1, 2), torch.nn.grad.conv2d_input(x.shape, w, dy, stride=2)) < 6e-3
    assert rel_l2(dw.cpu(), torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=2)) < 1e-4


@pytest.mark.parametrize("shape", S2_SHAPES)
def test_tc_up_conv_transpose_2x2_s2(pkg, shape):
    """ConvTranspose2d(Cin,Cout,2,stride 2) (unet.py:240): scatter forward, gather data gradient, weight gradient."""
    B, Cin, Cout, H, W = shape
    L, dev = pkg._capi.lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(sum(shape) + 1)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float().requires_grad_(True)
    w = (torch.randn(Cin, Cout, 2, 2, generator=g) / Cin ** 0.5).bfloat16().float().requires_grad_(True)
    b = torch.randn(Cout, generator=g)
    dy = torch.randn(B, Cout, 2 * H, 2 * W, generator=g).bfloat16().float()
    y_ref = F.conv_transpose2d(x, w, b, stride=2)
    y_ref.backward(dy)
    xg, wg, bg, dyg = _nhwc(x.detach(), torch.bfloat16, dev), w.detach().to(dev), b.to(dev), _nhwc(dy, torch.bfloat16, dev)
    y = torch.empty(B, 2 * H, 2 * W, Cout, device=dev, dtype=torch.bfloat16)
    dx = torch.empty(B, H, W, Cin, device=dev, dtype=torch.bfloat16)
    dw = torch.empty_like(wg)
    for mode, out in ((0, y), (1, dx), (2, None)):
        rc = L.fu_test_conv(1, 1, mode, B, H, W, Cin, Cout, 2, -2, 0, 0, _p(xg), _p(wg), _p(bg), _p(out), _p(dyg), _p(dw), None, None)
        assert rc == 0, pkg._capi.last_error(None)
    torch.cuda.synchronize()
    assert rel_l2(y.float().cpu().permute(0, 3, 1, 2), y_ref.detach()) < 6e-3
    assert rel_l2(dx.float().cpu().permute(0, 3, 1, 2), x.grad) < 6e-3
    assert rel_l2(dw.cpu(), w.grad) < 1e-4


HALO_SHAPES = [  # B, Cin, Cout, H, W  -- 3x3 convs wide enough for the halo kernel (W >= 24)
    (2, 32, 32, 40, 56),      # thin layer: 64-byte rows, weights resident in shared memory
    (1, 64, 64, 48, 48),      # resident weights, 128-byte rows
    (2, 128, 128, 24, 40),    # streamed weights, two pixel tiles per weight tile, 2 K chunks
    (1, 32, 64, 33, 29),      # ragged in both directions
    (3, 64, 32, 26, 70),
    (1, 256, 256, 24, 24),    # two N tiles
    (5, 64, 64, 24, 24),      # odd number of pixel tiles (last pair half empty)
    (2, 32, 64, 20, 96),      # weight gradient: 48-pixel K tiles, nine taps per CTA (Cin = 32)
    (1, 128, 64, 12, 192),    # weight gradient: 64-pixel K tiles, 128-channel B tiles
    (2, 32, 32, 3, 64),       # weight gradient with the filter rows stacked along M (Cout <= 32): fewer image rows than shifts + 1
    (1, 64, 32, 7, 130),      # same, 192 accumulator columns, three ragged row segments
]


# FU_TC_STACK=2: every 32-column layer takes box widths 16 / 32 and with them the kw-stacked MMAs (N = 96, partial sums
# shifted together by warp shuffles in the epilogue); 0: never
@pytest.mark.parametrize("env", [{}, {"FU_TC_HALO1": "0"}, {"FU_TC_PAIR": "0"}, {"FU_TC_RESIDENT": "0"}, {"FU_TC_STACK": "2"},
                                 {"FU_TC_STACK": "2", "FU_TC_EPI_SETS": "1"}, {"FU_TC_STACK": "0"}])
@pytest.mark.parametrize("shape", HALO_SHAPES)
def test_tc_halo_conv_fwd_dgrad(pkg, shape, env):
    """Second-generation 3x3 kernel (one halo load per chunk, taps = row-shifted descriptor views, paired
    pixel tiles, resident weights) and each of its fallback modes, against torch-CPU fp32."""
    import os
    B, Cin, Cout, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) + 7)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).bfloat16().float()
    b = torch.randn(Cout, generator=g)
    dy = torch.randn(B, Cout, H, W, generator=g).bfloat16().float()
    env = dict(env, FU_TC_V2_MINW="24")      # the engine only uses this kernel for W >= 48; test it from 24
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        y, stats = run_conv(pkg, 1, 1, 0, x, w, b, None, 3, 1, 1, relu=1, want_stats=True)
        dx = run_conv(pkg, 1, 1, 1, x, w, None, dy, 3, 1, 1)
        dw = run_conv(pkg, 1, 1, 2, x, w, None, dy, 3, 1, 1)       # halo weight-gradient kernel when W >= 32
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    y_ref = torch.relu(F.conv2d(x, w, b, padding=1))
    assert rel_l2(y, y_ref.bfloat16().float()) < 2e-3, rel_l2(y, y_ref)
    assert rel_l2(stats[:Cout], y.double().sum(dim=(0, 2, 3))) < 1e-4
    assert rel_l2(stats[Cout:], (y.double() ** 2).sum(dim=(0, 2, 3))) < 1e-4
    assert rel_l2(dx, torch.nn.grad.conv2d_input(x.shape, w, dy, stride=1, padding=1)) < 6e-3
    assert rel_l2(dw, torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=1, padding=1)) < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# parity mode on the tensor cores (FU_PRECISION_FP32_TC, fu_test_conv impl = 2): fp32 NHWC tensors, operands read
# from split-bf16 twins, three MMA passes per contraction.  Inputs are NOT bf16-representable here; the result
# must match torch-CPU fp32 to ~1e-5 (the dropped lo*lo products are ~2^-16 relative).
# ---------------------------------------------------------------------------------------------------------------
SPLIT_TOL = 3e-5


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_split_tc_conv_fwd_dgrad_wgrad(pkg, shape):
    B, Cin, Cout, H, W, k = shape
    pad = k // 2
    g = torch.Generator().manual_seed(sum(shape) + 11)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    dy = torch.randn(B, Cout, H, W, generator=g)
    y_ref = torch.relu(F.conv2d(x, w, b, padding=pad))
    y, stats = run_conv(pkg, 0, 2, 0, x, w, b, None, k, 1, pad, relu=1, want_stats=True)
    assert rel_l2(y, y_ref) < SPLIT_TOL, rel_l2(y, y_ref)
    assert rel_l2(stats[:Cout], y.double().sum(dim=(0, 2, 3))) < 1e-5
    assert rel_l2(stats[Cout:], (y.double() ** 2).sum(dim=(0, 2, 3))) < 1e-5
    dx = run_conv(pkg, 0, 2, 1, x, w, None, dy, k, 1, pad)
    assert rel_l2(dx, torch.nn.grad.conv2d_input(x.shape, w, dy, stride=1, padding=pad)) < SPLIT_TOL
    dw = run_conv(pkg, 0, 2, 2, x, w, None, dy, k, 1, pad)
    assert rel_l2(dw, torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=1, padding=pad)) < SPLIT_TOL


@pytest.mark.parametrize("env", [{}, {"FU_TC_HALO1": "0"}, {"FU_TC_PAIR": "0"}, {"FU_TC_RESIDENT": "0"}])
@pytest.mark.parametrize("shape", HALO_SHAPES)
def test_split_tc_halo_conv(pkg, shape, env):
    import os
    B, Cin, Cout, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) + 13)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    dy = torch.randn(B, Cout, H, W, generator=g)
    env = dict(env, FU_TC_V2_MINW="24")
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        y, stats = run_conv(pkg, 0, 2, 0, x, w, b, None, 3, 1, 1, relu=1, want_stats=True)
        dx = run_conv(pkg, 0, 2, 1, x, w, None, dy, 3, 1, 1)
        dw = run_conv(pkg, 0, 2, 2, x, w, None, dy, 3, 1, 1)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    y_ref = torch.relu(F.conv2d(x, w, b, padding=1))
    assert rel_l2(y, y_ref) < SPLIT_TOL, rel_l2(y, y_ref)
    assert rel_l2(stats[:Cout], y.double().sum(dim=(0, 2, 3))) < 1e-5
    assert rel_l2(stats[Cout:], (y.double() ** 2).sum(dim=(0, 2, 3))) < 1e-5
    assert rel_l2(dx, torch.nn.grad.conv2d_input(x.shape, w, dy, stride=1, padding=1)) < SPLIT_TOL
    assert rel_l2(dw, torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=1, padding=1)) < SPLIT_TOL


@pytest.mark.parametrize("shape", S2_SHAPES)
def test_split_tc_down_and_up_2x2_s2(pkg, shape):
    """Conv2d(C,C,2,2) and ConvTranspose2d(.,.,2,2) (unet.py:93,240) through the split path."""
    B, Cin, Cout, H, W = shape
    L, dev = pkg._capi.lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(sum(shape) + 17)
    f32 = torch.float32
    # ---- down ----
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 2, 2, generator=g) / (Cin * 4) ** 0.5
    b = torch.randn(Cout, generator=g)
    dy = torch.randn(B, Cout, H // 2, W // 2, generator=g)
    xg, wg, bg, dyg = _nhwc(x, f32, dev), w.to(dev), b.to(dev), _nhwc(dy, f32, dev)
    y = torch.empty(B, H // 2, W // 2, Cout, device=dev)
    dx = torch.empty(B, H, W, Cin, device=dev)
    dw = torch.empty_like(wg)
    for mode, out in ((0, y), (1, dx), (2, None)):
        rc = L.fu_test_conv(0, 2, mode, B, H, W, Cin, Cout, 2, 2, 0, 0, _p(xg), _p(wg), _p(bg), _p(out), _p(dyg), _p(dw), None, None)
        assert rc == 0, pkg._capi.last_error(None)
    torch.cuda.synchronize()
    assert rel_l2(y.cpu().permute(0, 3, 1, 2), F.conv2d(x, w, b, stride=2)) < SPLIT_TOL
    assert rel_l2(dx.cpu().permute(0, 3, 1, 2), torch.nn.grad.conv2d_input(x.shape, w, dy, stride=2)) < SPLIT_TOL
    assert rel_l2(dw.cpu(), torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=2)) < SPLIT_TOL
    # ---- up ----
    x = torch.randn(B, Cin, H, W, generator=g).requires_grad_(True)
    w = (torch.randn(Cin, Cout, 2, 2, generator=g) / Cin ** 0.5).requires_grad_(True)
    dy = torch.randn(B, Cout, 2 * H, 2 * W, generator=g)
    y_ref = F.conv_transpose2d(x, w, b, stride=2)
    y_ref.backward(dy)
    xg, wg, dyg = _nhwc(x.detach(), f32, dev), w.detach().to(dev), _nhwc(dy, f32, dev)
    y = torch.empty(B, 2 * H, 2 * W, Cout, device=dev)
    dx = torch.empty(B, H, W, Cin, device=dev)
    dw = torch.empty_like(wg)
    for mode, out in ((0, y), (1, dx), (2, None)):
        rc = L.fu_test_conv(0, 2, mode, B, H, W, Cin, Cout, 2, -2, 0, 0, _p(xg), _p(wg), _p(bg), _p(out), _p(dyg), _p(dw), None, None)
        assert rc == 0, pkg._capi.last_error(None)
    torch.cuda.synchronize()
    assert rel_l2(y.cpu().permute(0, 3, 1, 2), y_ref.detach()) < SPLIT_TOL
    assert rel_l2(dx.cpu().permute(0, 3, 1, 2), x.grad) < SPLIT_TOL
    assert rel_l2(dw.cpu(), w.grad) < SPLIT_TOL
