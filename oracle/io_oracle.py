"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's sample preparation and inference
post-processing (the callers either side of the U-Net).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the
product path (deepfluorolabeling-ipcai2020_b200/prepost.py -> libfluorounet.so) never does.

Pinned: tests/golden/make_golden_io.py runs the REAL reference in the authoring container
(dataset.RandomDataAugDataSet.__getitem__, util.seg_dataset_ensemble and the est_lands_csv.py script
itself, unmodified, with an in-memory stand-in for the absent h5py package) and stores inputs and
outputs in tests/golden/io.npz; tests/test_io_oracle_golden.py checks every function here against them.

Plain numpy / torch-CPU in fp32, following the reference expression by expression:
  prep_tiles          dataset.py:287-293
  heatmap_targets     dataset.py:295-325
  ensemble_combine    util.py:331-370
  gaussian_template   util.py:36-48
  ncc_2d              ncc.py:12-38
  extract_landmarks   est_lands_csv.py:87-134
"""
import math

import numpy as np
import torch


def calc_pad_amount(padded_img_dim, cur_img_dim):
    """dataset.py:26-40: border width that brings cur_img_dim to (at least) padded_img_dim."""
    assert padded_img_dim > cur_img_dim
    pad = (padded_img_dim - cur_img_dim) / 2
    return int(pad) + 1 if pad != int(pad) else int(pad)


def prep_tiles(tiles, pad, normalize=True):
    """dataset.py:287-293.  tiles (B,h,w) fp32 -> (B,1,h+2pad,w+2pad)."""
    out = []
    for t in tiles:
        p = torch.as_tensor(t, dtype=torch.float32)[None]
        if pad > 0:
            p = torch.from_numpy(np.pad(p.numpy(), ((0, 0), (pad, pad), (pad, pad)), "reflect"))
        if normalize:
            p = (p - p.mean()) / p.std()
        out.append(p)
    return torch.stack(out)


def heatmap_targets(lands, H, W, sigma=2.5):
    """dataset.py:295-325.  lands (B,2,L) (x = column, y = row; inf = outside the view) -> (B,L,H,W)."""
    lands = torch.as_tensor(lands, dtype=torch.float32)
    B, _, L = lands.shape
    h = torch.zeros(B, L, H, W)
    sig = torch.full([L], sigma)
    Y, X = torch.meshgrid(torch.arange(0, H), torch.arange(0, W), indexing="ij")
    Y, X = Y.float(), X.float()
    for b in range(B):
        for l in range(L):
            s = sig[l]
            mu_x, mu_y = lands[b, 0, l], lands[b, 1, l]
            if not math.isinf(mu_x) and not math.isinf(mu_y):
                h[b, l] = torch.exp(((X - mu_x).pow(2) + (Y - mu_y).pow(2)) / (s * s * -2)) / (2 * math.pi * s * s)
    return h


def _crop(img, shape):
    """util.py:92-114."""
    r0 = int((img.shape[-2] - shape[-2]) / 2)
    c0 = int((img.shape[-1] - shape[-1]) / 2)
    return img[..., r0:r0 + shape[-2], c0:c0 + shape[-1]]


def ensemble_combine(segs, heats, out_shape):
    """util.py:331-370 for one batch: segs / heats = lists (one entry per network) of (B,C,H,W) / (B,L,H,W);
    the reference runs batch size 1, so the heat min/max is taken per (network, image).
    Returns labels u8 (B,h,w) and the averaged normalised heat-maps (B,L,h,w) (None without heats)."""
    n = len(segs)
    B = segs[0].shape[0]
    labels, avg_heats = [], []
    for b in range(B):
        avg_m, avg_h = None, None
        for k in range(n):
            m = _crop(torch.as_tensor(segs[k][b:b + 1], dtype=torch.float32), out_shape).clone()
            avg_m = m if avg_m is None else avg_m + m
            if heats is not None:
                hm = _crop(torch.as_tensor(heats[k][b:b + 1], dtype=torch.float32), out_shape)
                lo, hi = hm.min().item(), hm.max().item()
                hm = (hm - lo) / (hi - lo)
                avg_h = hm if avg_h is None else avg_h + hm
        avg_m = avg_m / n
        labels.append(torch.max(avg_m, dim=1)[1].to(torch.uint8))
        if heats is not None:
            avg_heats.append(avg_h / n)
    return torch.cat(labels), (torch.cat(avg_heats) if heats is not None else None)


def gaussian_template(num_rows, num_cols, sigma):
    """util.py:36-48 with the default (centre) peak."""
    pr, pc = num_rows // 2, num_cols // 2
    Y, X = torch.meshgrid(torch.arange(0, num_rows), torch.arange(0, num_cols), indexing="ij")
    Y, X = Y.float(), X.float()
    return torch.exp(((X - pc).pow(2) + (Y - pr).pow(2)) / (sigma * sigma * -2)) / (2 * math.pi * sigma * sigma)


def ncc_2d(X, Y):
    """ncc.py:12-38."""
    N = X.shape[-1] * X.shape[-2]
    Xz = X - X.mean(dim=(-2, -1), keepdim=True)
    Yz = Y - Y.mean(dim=(-2, -1), keepdim=True)
    Xs = torch.sqrt((Xz * Xz).sum(dim=(-2, -1)) / (N - 1))
    Ys = torch.sqrt((Yz * Yz).sum(dim=(-2, -1)) / (N - 1))
    return (Xz * Yz).sum(dim=(-2, -1)) / ((N * (Xs * Ys)) + 1.0e-8)


def extract_landmarks(heats, segs=None, seg_labels=None, tmpl_dim=25, sigma=2.5, min_ncc=0.9):
    """est_lands_csv.py:87-134.  heats (P,L,h,w); segs (P,h,w) or None; seg_labels: per-landmark anatomy label
    (None / negative = unmasked).  Returns (P,L,2) int64 rows/cols with -1,-1 for "not found", and the NCC scores
    (NaN where the masked arg-max found no pixel)."""
    heats = torch.as_tensor(heats, dtype=torch.float32)
    P, L = heats.shape[:2]
    half = tmpl_dim // 2
    tmpl = gaussian_template(tmpl_dim, tmpl_dim, sigma)
    out = torch.full((P, L, 2), -1, dtype=torch.int64)
    scores = torch.full((P, L), float("nan"))
    for i in range(P):
        for l in range(L):
            cur = heats[i, l]
            pad = torch.from_numpy(np.pad(cur.numpy(), ((half, half), (half, half)), "reflect"))
            label = None if seg_labels is None else seg_labels[l]
            if segs is None or label is None or label < 0:
                idx = np.unravel_index(torch.argmax(cur).item(), cur.shape)
            else:
                tmp = cur.clone()
                tmp[torch.as_tensor(segs[i]) != label] = -math.inf
                idx = np.unravel_index(torch.argmax(tmp).item(), cur.shape)
                if tmp[idx[0], idx[1]] == -math.inf:
                    continue
            roi = pad[idx[0]:idx[0] + tmpl_dim, idx[1]:idx[1] + tmpl_dim]
            s = ncc_2d(tmpl, roi)
            scores[i, l] = s
            if not (s < min_ncc):
                out[i, l, 0], out[i, l, 1] = int(idx[0]), int(idx[1])
    return out, scores
