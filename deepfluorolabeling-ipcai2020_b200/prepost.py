"""Device versions of the reference's host code either side of the U-Net:

    prep_tiles          dataset.py:287-293   reflect pad + z-score of the raw tiles
    heatmap_targets     dataset.py:295-325   Gaussian heat-map targets from landmark coordinates
    ensemble_combine    util.py:331-370      ensemble averaging / arg-max labels / normalised heat-maps
    seg_dataset_ensemble util.py:293-377     the same loop over a dataset, networks batched on the device
    extract_landmarks   est_lands_csv.py:87-134  landmark pixel from a heat-map (+ segmentation)

Each is a thin wrapper over the C ABI (include/fluoro_unet.h) on CUDA tensors.  There is no CPU
fallback: CPU tensors are refused."""
import ctypes as C
import time

import torch

from . import _capi
from .util import center_crop  # noqa: F401  (re-exported for callers that crop by hand)


def calc_pad_amount(padded_img_dim, cur_img_dim):
    """dataset.py:26-40."""
    assert padded_img_dim > cur_img_dim
    pad = (padded_img_dim - cur_img_dim) / 2
    return int(pad) + 1 if pad != int(pad) else int(pad)


def _dev(name, t, dtype, ndim):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (this package has no CPU path)")
    if t.dtype != dtype or t.dim() not in (ndim if isinstance(ndim, tuple) else (ndim,)):
        raise TypeError(f"{name} must be a {ndim}-D {dtype} tensor, got {tuple(t.shape)} {t.dtype}")
    return t.contiguous()


def _check(rc, who):
    if rc != 0:
        raise RuntimeError(f"{who} failed ({rc}): {_capi.last_error(None)}")


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def prep_tiles(tiles, pad_img_dim=0, normalize=True):
    """(B,h,w) or (B,1,h,w) raw fp32 tiles -> (B,1,H,H) network input: reflect-padded to ``pad_img_dim`` (the
    ``--unet-img-dim`` of train.py; 0 = no padding) and z-scored per tile -- dataset.py:287-293."""
    if isinstance(tiles, torch.Tensor) and tiles.dim() == 4:
        if tiles.shape[1] != 1:
            raise ValueError("prep_tiles: tiles have one channel (dataset.py:48)")
        tiles = tiles[:, 0]
    tiles = _dev("tiles", tiles, torch.float32, 3)
    B, h, w = tiles.shape
    pad = 0
    if pad_img_dim > 0:
        if h != w:
            raise ValueError("prep_tiles: only square tiles can be padded (dataset.py:83-84)")
        pad = calc_pad_amount(pad_img_dim, w)
    out = torch.empty(B, 1, h + 2 * pad, w + 2 * pad, device=tiles.device, dtype=torch.float32)
    sums = torch.empty(2 * B, device=tiles.device, dtype=torch.float64) if normalize else None
    with torch.cuda.device(tiles.device):
        _check(_capi.lib().fu_prep_tiles(tiles.data_ptr(), B, h, w, pad, int(bool(normalize)),
                                         sums.data_ptr() if normalize else None, out.data_ptr(), _stream(tiles)),
               "fu_prep_tiles")
    return out


def heatmap_targets(lands, shape, sigma=2.5):
    """lands (B,2,L) fp32 (row 0 = column x, row 1 = row y; +-inf = outside the view) -> (B,L,H,W) Gaussian
    heat-map targets -- dataset.py:295-325."""
    lands = _dev("lands", lands, torch.float32, 3)
    if lands.shape[1] != 2:
        raise ValueError("heatmap_targets: lands must be (B,2,L) (dataset.py:59)")
    B, _, L = lands.shape
    H, W = int(shape[-2]), int(shape[-1])
    if L > 65535:
        raise ValueError("heatmap_targets: at most 65535 landmarks per sample (one launch covers 65535 planes)")
    if L > 0 and B * L > 65535:     # one launch covers at most 65535 planes: split the batch
        step = max(1, 65535 // L)
        return torch.cat([heatmap_targets(lands[i:i + step], shape, sigma) for i in range(0, B, step)])
    out = torch.empty(B, L, H, W, device=lands.device, dtype=torch.float32)
    with torch.cuda.device(lands.device):
        _check(_capi.lib().fu_heatmap_targets(lands.data_ptr(), B, L, H, W, float(sigma), out.data_ptr(), _stream(lands)),
               "fu_heatmap_targets")
    return out


def ensemble_combine(segs, heats, out_shape):
    """segs / heats: lists with one (B,C,H,W) / (B,L,H,W) fp32 CUDA tensor per network (heats None for seg-only
    networks).  Returns (labels u8 (B,h,w), averaged normalised heat-maps (B,L,h,w) or None) for the centre-crop
    window ``out_shape`` -- util.py:331-370, with the heat min/max taken per (network, image) as the reference's
    batch-size-1 loop does."""
    n = len(segs)
    if n < 1:
        raise ValueError("ensemble_combine: no networks")
    segs = [_dev("segs[%d]" % k, s, torch.float32, 4) for k, s in enumerate(segs)]
    B, NC, H, W = segs[0].shape
    h, w = int(out_shape[-2]), int(out_shape[-1])
    r0, c0 = int((H - h) / 2), int((W - w) / 2)
    dev = segs[0].device
    NL = 0
    if heats is not None:
        if len(heats) != n:
            raise ValueError("ensemble_combine: one heat-map tensor per network")
        heats = [_dev("heats[%d]" % k, t, torch.float32, 4) for k, t in enumerate(heats)]
        NL = heats[0].shape[1]
    for k in range(n):
        if tuple(segs[k].shape) != (B, NC, H, W) or (NL and tuple(heats[k].shape) != (B, NL, H, W)) or segs[k].device != dev:
            raise ValueError("ensemble_combine: the networks' outputs disagree in shape or device")
    L = _capi.lib()
    labels = torch.empty(B, h, w, device=dev, dtype=torch.uint8)
    avg = torch.empty(B, NL, h, w, device=dev, dtype=torch.float32) if NL else None
    ws = torch.empty(int(L.fu_ensemble_workspace_words(n, B)), device=dev, dtype=torch.int32) if NL else None
    seg_p = (C.c_void_p * n)(*[s.data_ptr() for s in segs])
    heat_p = (C.c_void_p * n)(*[t.data_ptr() for t in heats]) if NL else None
    with torch.cuda.device(dev):
        _check(L.fu_ensemble_combine(seg_p, heat_p, n, B, NC, NL, H, W, r0, c0, h, w,
                                     ws.data_ptr() if NL else None, labels.data_ptr(),
                                     avg.data_ptr() if NL else None, _stream(segs[0])), "fu_ensemble_combine")
    return labels, avg


def seg_dataset_ensemble(projs, nets, orig_img_shape, num_lands=0, batch_size=32, times=None):
    """util.seg_dataset_ensemble (util.py:293-377) without the HDF5 file: runs every network of the ensemble (eval
    mode, no grad) over ``projs`` (N,1,H,W) in batches and returns (labels u8 (N,h,w), heat-maps (N,L,h,w) or None)
    as CUDA tensors -- what the reference stores as 'nn-segs' / 'nn-heats'.  ``times`` receives one wall-clock
    figure per image (batch time / batch size), as util.py:321,363-366 does per image."""
    projs = _dev("projs", projs, torch.float32, 4)
    for net in nets:
        net.eval()
    lab, hts = [], []
    with torch.no_grad():
        for i in range(0, projs.shape[0], batch_size):
            if times is not None:
                torch.cuda.synchronize(projs.device)
                t0 = time.time()
            x = projs[i:i + batch_size]
            outs = [net(x) for net in nets]
            two = num_lands > 0 or isinstance(outs[0], tuple)
            segs = [o[0] if two else o for o in outs]
            heats = [o[1] for o in outs] if num_lands > 0 else None
            l, a = ensemble_combine(segs, heats, orig_img_shape)
            lab.append(l)
            hts.append(a)
            if times is not None:
                torch.cuda.synchronize(projs.device)
                times.extend([(time.time() - t0) / x.shape[0]] * x.shape[0])
    return torch.cat(lab), (torch.cat(hts) if num_lands > 0 else None)


# est_lands_csv.py:54-71
SEG_LABELS_FOR_LANDS = {"FH-l": 5, "FH-r": 6, "GSN-l": 1, "GSN-r": 2, "IOF-l": 1, "IOF-r": 2, "MOF-l": 1, "MOF-r": 2,
                        "SPS-l": 1, "SPS-r": 2, "IPS-l": 1, "IPS-r": 2, "ASIS-l": 1, "ASIS-r": 2, "PSIS-l": 1,
                        "PSIS-r": 2, "PIIS-l": 1, "PIIS-r": 2}


def extract_landmarks(heats, segs=None, seg_labels=None, tmpl_dim=25, sigma=2.5, min_ncc=0.9, return_scores=False):
    """heats (P,L,h,w) fp32, segs (P,h,w) u8 or None, seg_labels: L anatomy labels (None / negative = unmasked) or
    L landmark names looked up in SEG_LABELS_FOR_LANDS.  Returns (P,L,2) int32 (row, col), -1,-1 = not found --
    est_lands_csv.py:87-134 (rule_3)."""
    heats = _dev("heats", heats, torch.float32, 4)
    P, L, h, w = heats.shape
    lab_p = None
    if segs is not None:
        segs = _dev("segs", segs, torch.uint8, 3)
        if tuple(segs.shape) != (P, h, w) or segs.device != heats.device:
            raise ValueError("extract_landmarks: segs must be (P,h,w) on the heat-maps' device")
        if seg_labels is None or len(seg_labels) != L:
            raise ValueError("extract_landmarks: one segmentation label (or landmark name) per landmark")
        ids = [SEG_LABELS_FOR_LANDS[s] if isinstance(s, str) else (-1 if s is None else int(s)) for s in seg_labels]
        lab_p = (C.c_int32 * L)(*ids)
    out = torch.empty(P, L, 2, device=heats.device, dtype=torch.int32)
    ncc = torch.empty(P, L, device=heats.device, dtype=torch.float32) if return_scores else None
    with torch.cuda.device(heats.device):
        _check(_capi.lib().fu_extract_landmarks(heats.data_ptr(), segs.data_ptr() if segs is not None else None, lab_p,
                                                P, L, h, w, int(tmpl_dim), float(sigma), float(min_ncc), out.data_ptr(),
                                                ncc.data_ptr() if return_scores else None, _stream(heats)),
               "fu_extract_landmarks")
    return (out, ncc) if return_scores else out
