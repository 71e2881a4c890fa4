"""Launch every sample-prep / post-processing kernel once at B=32 (for `ncu --set full` captures)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pp = importlib.import_module("deepfluorolabeling-ipcai2020_b200").prepost
dev = torch.device("cuda:0")
B, h, dim, C, L, n = 32, 180, 192, 7, 14, 3
g = torch.Generator().manual_seed(0)
tiles = (torch.rand(B, h, h, generator=g) * 60000).to(dev)
lands = (torch.rand(B, 2, L, generator=g) * (h - 1)).to(dev)
segs = [torch.softmax(torch.randn(B, C, dim, dim, device=dev), 1) for _ in range(n)]
heats = [torch.randn(B, L, dim, dim, device=dev) for _ in range(n)]
lab = torch.randint(0, 7, (B, h, h), device=dev, dtype=torch.uint8)
torch.cuda.synchronize()
pp.prep_tiles(tiles, pad_img_dim=dim)
tgt = pp.heatmap_targets(lands, (h, h))
pp.ensemble_combine(segs, heats, (h, h))
pp.extract_landmarks(tgt, lab, [1] * L)
torch.cuda.synchronize()
print("done")
