"""CUDA-event timing of the sample-preparation / post-processing kernels at BASELINE sizes (180^2 tiles padded to
192^2, 7 classes, 14 landmarks), L2 flushed between iterations.  Prints one JSON line per op with the algorithmic
bytes (compulsory reads + writes) and the achieved GB/s against MEASURED_PEAKS.json's HBM figure.
usage: python tools/prepost_time.py [B] [n_nets]"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
pp = pkg.prepost


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_nets = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        peak = float(mp.get("hbm_gbs") or 0) or None
    except Exception:
        mp = None
    h, dim, C, L = 180, 192, 7, 14
    g = torch.Generator().manual_seed(0)
    tiles = (torch.rand(B, h, h, generator=g) * 60000).to(dev)
    lands = (torch.rand(B, 2, L, generator=g) * (h - 1)).to(dev)
    segs = [torch.softmax(torch.randn(B, C, dim, dim, device=dev), 1) for _ in range(n_nets)]
    heats = [torch.randn(B, L, dim, dim, device=dev) for _ in range(n_nets)]
    tgt = pp.heatmap_targets(lands, (h, h))
    lab = torch.randint(0, 7, (B, h, h), device=dev, dtype=torch.uint8)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ops = {
        "prep_tiles": (lambda: pp.prep_tiles(tiles, pad_img_dim=dim), 4 * B * (h * h + dim * dim)),
        "heatmap_targets": (lambda: pp.heatmap_targets(lands, (h, h)), 4 * B * L * h * h),
        "ensemble_combine": (lambda: pp.ensemble_combine(segs, heats, (h, h)),
                             4 * n_nets * B * (C + L) * h * h + B * h * h * (1 + 4 * L)),
        "extract_landmarks": (lambda: pp.extract_landmarks(tgt, lab, [1] * L), B * h * h * (4 * L + 1)),
    }
    for name, (fn, nbytes) in ops.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        gbs = nbytes / ms / 1e6
        print(json.dumps({"op": name, "B": B, "n_nets": n_nets, "ms": round(ms, 4), "algorithmic_bytes": nbytes,
                          "GB/s": round(gbs, 1), "hbm_peak_GB/s": peak, "frac": round(gbs / peak, 3) if peak else None,
                          "images_per_s": round(B / ms * 1e3)}))


if __name__ == "__main__":
    main()
