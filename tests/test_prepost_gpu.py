"""Sample preparation and inference post-processing kernels (csrc/kernels_io.cuh behind fu_prep_tiles,
fu_heatmap_targets, fu_ensemble_combine, fu_extract_landmarks) against
  (1) outputs of the REAL reference stored in tests/golden/io.npz (dataset.py, util.py, est_lands_csv.py),
  (2) the oracle (oracle/io_oracle.py) on other seeded inputs,
  (3) size-independent properties at BASELINE.json's full sizes (32 x 180^2 -> 192^2, 14 landmarks).
Integer / label / index results must be bit-exact; floating-point results within the stated tolerances
(the device computes tile statistics in fp64 and uses CUDA's expf, the reference torch-CPU fp32).  GPU only."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from oracle import io_oracle as IO   # checker only

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PREP_TOL = dict(rtol=3e-6, atol=3e-6)     # z-scored values are O(1)
HEAT_TOL = dict(rtol=3e-6, atol=1e-30)    # expf: <= 2 ulp on the device, <= 1 ulp on the CPU


@pytest.fixture(scope="module")
def pp():
    assert torch.cuda.is_available()
    p = load_pkg()
    p._capi.lib()
    return p.prepost


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLDEN, "io.npz"))
    return {k: z[k] for k in z.files}


def cu(a, dtype=None):
    t = torch.as_tensor(np.asarray(a))
    return (t.to(dtype) if dtype is not None else t).to(DEV)


@pytest.mark.parametrize("tag,step,dim", [("small", 1, 32), ("odd", 1, 32), ("paper", 5, 192)])
def test_prep_and_heat_targets_match_reference_golden(pp, gold, tag, step, dim):
    tiles = gold[f"prep_{tag}_tiles"]
    out = pp.prep_tiles(cu(tiles), pad_img_dim=dim)
    want = gold[f"prep_{tag}_out"]
    np.testing.assert_allclose(out.cpu().numpy()[..., ::step, ::step], want, **PREP_TOL)
    heat = pp.heatmap_targets(cu(gold[f"prep_{tag}_lands"]), tiles.shape[-2:])
    np.testing.assert_allclose(heat.cpu().numpy()[..., ::step, ::step], gold[f"prep_{tag}_heat"], **HEAT_TOL)
    assert float(heat[0, 1].abs().max()) == 0.0            # the +inf landmark: zero plane (dataset.py:316)


def test_prep_without_normalisation_or_padding_is_exact(pp):
    g = torch.Generator().manual_seed(3)
    t = torch.randn(4, 17, 23, generator=g)
    out = pp.prep_tiles(t.to(DEV), pad_img_dim=0, normalize=False)
    assert torch.equal(out.cpu(), t[:, None])
    sq = torch.randn(2, 19, 19, generator=g)
    out = pp.prep_tiles(sq.to(DEV), pad_img_dim=30, normalize=False)     # pad 6 (odd difference rounds up)
    assert torch.equal(out.cpu(), IO.prep_tiles(sq.numpy(), 6, normalize=False))


def test_ensemble_matches_reference_golden(pp, gold):
    segs = [cu(s) for s in gold["ens_segs"]]
    heats = [cu(s) for s in gold["ens_heats"]]
    labels, avg = pp.ensemble_combine(segs, heats, gold["ens_labels"].shape[-2:])
    assert labels.dtype == torch.uint8
    np.testing.assert_array_equal(labels.cpu().numpy(), gold["ens_labels"])          # incl. exact ties: first max wins
    np.testing.assert_allclose(avg.cpu().numpy(), gold["ens_avg_heats"], rtol=1e-6, atol=1e-7)
    labels2, none = pp.ensemble_combine(segs[:2], None, gold["ens_labels"].shape[-2:])
    assert none is None
    np.testing.assert_array_equal(labels2.cpu().numpy(), gold["ens_labels_2nets"])


@pytest.mark.parametrize("tag", ["seg", "noseg"])
def test_landmarks_match_est_lands_csv_golden(pp, gold, tag):
    segs = cu(gold["land_segs"]) if tag == "seg" else None
    labels = [int(v) for v in gold["land_labels"]]
    rc, ncc = pp.extract_landmarks(cu(gold["land_heats"]), segs, labels if segs is not None else None, return_scores=True)
    assert rc.dtype == torch.int32
    np.testing.assert_array_equal(rc.cpu().numpy().astype(np.int64), gold[f"land_rc_{tag}"])
    _, want = IO.extract_landmarks(gold["land_heats"], gold["land_segs"] if tag == "seg" else None, labels)
    got, want = ncc.cpu().numpy(), want.numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got[~np.isnan(got)], want[~np.isnan(want)], rtol=0, atol=1e-4)


def test_landmark_names_resolve_to_the_reference_label_table(pp, gold):
    names = ["FH-l", "FH-r", "GSN-l", "GSN-r", "IOF-l", "IOF-r", "MOF-l", "MOF-r", "SPS-l", "SPS-r", "IPS-l", "IPS-r",
             "ASIS-l", "ASIS-r"]
    rc = pp.extract_landmarks(cu(gold["land_heats"]), cu(gold["land_segs"]), names)
    np.testing.assert_array_equal(rc.cpu().numpy().astype(np.int64), gold["land_rc_seg"])


@pytest.mark.parametrize("B,h,dim,L,seed", [(5, 37, 48, 6, 0), (2, 64, 64, 1, 1), (3, 50, 51, 9, 2)])
def test_prep_and_heat_targets_match_oracle(pp, B, h, dim, L, seed):
    g = torch.Generator().manual_seed(seed)
    tiles = torch.rand(B, h, h, generator=g) * 4000 - 700
    pad = IO.calc_pad_amount(dim, h) if dim > h else 0
    got = pp.prep_tiles(tiles.to(DEV), pad_img_dim=dim if dim > h else 0)
    np.testing.assert_allclose(got.cpu().numpy(), IO.prep_tiles(tiles.numpy(), pad).numpy(), **PREP_TOL)
    lands = torch.rand(B, 2, L, generator=g) * (h + 20) - 10
    lands[B - 1, 1, 0] = math.inf
    heat = pp.heatmap_targets(lands.to(DEV), (h, h + 3))
    np.testing.assert_allclose(heat.cpu().numpy(), IO.heatmap_targets(lands, h, h + 3).numpy(), **HEAT_TOL)


def test_heat_targets_special_coordinates(pp):
    """NaN coordinates poison the whole plane in the reference expression, +-inf gives a zero plane (dataset.py:316), finite
    coordinates far outside the view underflow to zero, peaks just outside the border still leak in."""
    lands = torch.tensor([[[float("nan"), 5.0, 1.0e30, -3.5, 20.25, float("-inf")],
                           [7.0, float("nan"), 4.0, 10.0, 33.75, 2.0]]])
    got = pp.heatmap_targets(lands.to(DEV), (30, 34)).cpu()
    want = IO.heatmap_targets(lands, 30, 34)
    assert bool(got[0, 0].isnan().all()) and bool(got[0, 1].isnan().all())
    assert float(got[0, 2].abs().max()) == 0.0 and float(got[0, 5].abs().max()) == 0.0
    assert float(got[0, 3].max()) > 0 and float(got[0, 4].max()) > 0
    np.testing.assert_allclose(got.numpy(), want.numpy(), **HEAT_TOL)


@pytest.mark.parametrize("n,B,C,L,H,h,seed", [(1, 3, 7, 14, 24, 24, 0), (4, 2, 5, 3, 33, 20, 1), (16, 1, 2, 1, 16, 9, 2),
                                                  (4, 7, 7, 14, 192, 180, 3)])   # the last one spans two L2 chunks
def test_ensemble_matches_oracle(pp, n, B, C, L, H, h, seed):
    g = torch.Generator().manual_seed(seed)
    segs = [torch.round(torch.softmax(torch.randn(B, C, H, H, generator=g), 1) * 16) / 16 for _ in range(n)]
    heats = [torch.randn(B, L, H, H, generator=g) * (k + 1) - k for k in range(n)]
    labels, avg = pp.ensemble_combine([s.to(DEV) for s in segs], [t.to(DEV) for t in heats], (h, h))
    want_l, want_a = IO.ensemble_combine(segs, heats, (h, h))
    assert torch.equal(labels.cpu(), want_l)
    np.testing.assert_allclose(avg.cpu().numpy(), want_a.numpy(), rtol=1e-6, atol=1e-7)


def test_ensemble_rejects_more_networks_than_the_kernel_holds(pp):
    s = [torch.zeros(1, 2, 8, 8, device=DEV) for _ in range(17)]
    with pytest.raises(RuntimeError, match="at most 16"):
        pp.ensemble_combine(s, None, (8, 8))


@pytest.mark.parametrize("P,L,h,w,seed,use_seg", [(2, 5, 40, 52, 1, True), (3, 3, 30, 30, 10, False), (1, 14, 64, 64, 1, True),
                                                   (2, 4, 31, 33, 1, True)])
def test_landmarks_match_oracle(pp, P, L, h, w, seed, use_seg):
    g = torch.Generator().manual_seed(seed)
    lands = torch.stack([torch.rand(P, L, generator=g) * (w - 1), torch.rand(P, L, generator=g) * (h - 1)], dim=1)
    heats = IO.heatmap_targets(lands, h, w) + torch.randn(P, L, h, w, generator=g) * 3e-4
    heats[:, L - 1] = torch.randn(P, h, w, generator=g) * 1e-3                  # noise only: rejected by the NCC test
    segs = torch.randint(0, 3, (P, h, w), generator=g).to(torch.uint8) if use_seg else None
    labels = [l % 4 for l in range(L)] if use_seg else None                       # label 3 never occurs: "not found"
    if use_seg and L > 4:
        labels[4] = -1                                                            # unmasked landmark
    rc, ncc = pp.extract_landmarks(heats.to(DEV), segs.to(DEV) if use_seg else None, labels, return_scores=True)
    want_rc, want_ncc = IO.extract_landmarks(heats, segs, labels)
    margin = (want_ncc - 0.9).abs()
    assert bool((margin[~want_ncc.isnan()] > 1e-3).all())                         # the fixture stays off the threshold
    assert torch.equal(rc.cpu().long(), want_rc)
    assert torch.equal(ncc.cpu().isnan(), want_ncc.isnan())


def test_full_size_properties_and_round_trip(pp):
    """BASELINE configs[1] sizes: 32 tiles of 180^2 padded to 192^2, 14 landmarks per tile."""
    g = torch.Generator().manual_seed(7)
    B, h, dim, L = 32, 180, 192, 14
    tiles = (torch.rand(B, h, h, generator=g) * 60000).to(DEV)
    x = pp.prep_tiles(tiles, pad_img_dim=dim)
    assert x.shape == (B, 1, dim, dim)
    m, s = x.mean(dim=(1, 2, 3)), x.std(dim=(1, 2, 3))
    assert float(m.abs().max()) < 1e-4 and float((s - 1).abs().max()) < 1e-4       # z-scored over the padded tile
    p = 6
    for k in (1, 3, 6):                                                             # numpy 'reflect' symmetry
        assert torch.equal(x[:, :, p - k, :], x[:, :, p + k, :]) and torch.equal(x[:, :, :, p - k], x[:, :, :, p + k])
        assert torch.equal(x[:, :, dim - 1 - p + k, :], x[:, :, dim - 1 - p - k, :])
        assert torch.equal(x[:, :, :, dim - 1 - p + k], x[:, :, :, dim - 1 - p - k])
    # the interior is an affine map of the raw tile with positive slope, the same for the whole tile
    core = x[:, 0, p:p + h, p:p + h]
    a = (core[:, 0, 1] - core[:, 0, 0]) / (tiles[:, 0, 1] - tiles[:, 0, 0])
    assert bool((a > 0).all())
    # heat-map targets -> landmark extraction recovers integer landmark positions exactly
    lands = torch.stack([torch.randint(15, h - 15, (B, L), generator=g), torch.randint(15, h - 15, (B, L), generator=g)], 1).float()
    lands[3, 0, 5] = math.inf
    heat = pp.heatmap_targets(lands.to(DEV), (h, h))
    peak = 1.0 / (2 * math.pi * 2.5 * 2.5)
    assert abs(float(heat.max()) - peak) < 1e-7 and float(heat.min()) >= 0.0
    sums = heat.sum(dim=(2, 3)).cpu()
    assert float(sums[3, 5]) == 0.0
    ok = torch.ones(B, L, dtype=torch.bool)
    ok[3, 5] = False
    assert float((sums[ok] - 1).abs().max()) < 1e-3                                 # a unit-mass Gaussian inside the view
    rc, ncc = pp.extract_landmarks(heat, return_scores=True)
    rc = rc.cpu().long()
    want = torch.stack([lands[:, 1], lands[:, 0]], dim=-1).long()                   # (row, col) = (y, x)
    assert torch.equal(rc[ok], want[ok])
    # ncc.py:38 divides by N * sd_x * sd_y with the UNBIASED sd: a perfect match scores (N-1)/N = 624/625
    assert float((ncc.cpu()[ok] - 624.0 / 625.0).abs().max()) < 1e-5
    assert rc[3, 5].tolist() == [-1, -1]                                            # an empty plane has NCC 0 < 0.9
    # a one-network ensemble is arg-max + min-max normalisation
    seg = torch.softmax(torch.randn(B, 7, dim, dim, generator=g), 1).to(DEV)
    hm = torch.randn(B, L, dim, dim, generator=g).to(DEV)
    labels, avg = pp.ensemble_combine([seg], [hm], (h, h))
    assert torch.equal(labels.long(), seg[:, :, p:p + h, p:p + h].argmax(dim=1))
    assert float(avg.amin()) == 0.0 and float(avg.amax()) == 1.0
    assert float(avg.amin(dim=(1, 2, 3)).max()) == 0.0 and float(avg.amax(dim=(1, 2, 3)).min()) == 1.0


def test_seg_dataset_ensemble_runs_the_engine_networks(pp):
    pkg = load_pkg()
    kw = dict(n_classes=7, depth=3, wf=3, batch_norm=True, padding=True, max_pool=False, num_lands=14, do_res=True,
              block_depth=2)
    nets = []
    for k in range(2):
        torch.manual_seed(20 + k)
        nets.append(pkg.UNet(precision="fp32", **kw).to(DEV))
    g = torch.Generator().manual_seed(5)
    raw = torch.rand(5, 20, 20, generator=g) * 1000
    projs = pp.prep_tiles(raw.to(DEV), pad_img_dim=32)
    times = []
    labels, heats = pp.seg_dataset_ensemble(projs, nets, (20, 20), num_lands=14, batch_size=2, times=times)
    assert labels.shape == (5, 20, 20) and heats.shape == (5, 14, 20, 20) and len(times) == 5
    with torch.no_grad():                                                          # same batches as the call above
        outs = [[net(projs[i:i + 2]) for i in range(0, 5, 2)] for net in nets]
    want_l, want_h = IO.ensemble_combine([torch.cat([o[0] for o in per_net]).cpu() for per_net in outs],
                                         [torch.cat([o[1] for o in per_net]).cpu() for per_net in outs], (20, 20))
    assert torch.equal(labels.cpu(), want_l)
    np.testing.assert_allclose(heats.cpu().numpy(), want_h.numpy(), rtol=1e-5, atol=1e-6)


def test_heatmap_targets_rejects_more_landmarks_than_one_launch_covers(pp):
    """65536 landmarks per sample cannot be split over launches by batch: rejected loudly (used to recurse forever)."""
    with pytest.raises(ValueError, match="65535"):
        pp.heatmap_targets(torch.zeros(1, 2, 65536, device="cuda:0"), (8, 8))
