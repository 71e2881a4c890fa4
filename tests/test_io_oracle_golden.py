"""The sample-preparation / post-processing oracle (oracle/io_oracle.py) against outputs of the REAL reference
(tests/golden/io.npz, written by tests/golden/make_golden_io.py from dataset.py, util.py and est_lands_csv.py
run unmodified).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import io_oracle as IO


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLDEN, "io.npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("tag,step", [("small", 1), ("odd", 1), ("paper", 5)])
def test_prep_tiles_and_heatmap_targets_match_dataset_getitem(gold, tag, step):
    tiles, pad = gold[f"prep_{tag}_tiles"], int(gold[f"prep_{tag}_pad"])
    out = IO.prep_tiles(tiles, pad).numpy()[..., ::step, ::step]
    assert out.shape == gold[f"prep_{tag}_out"].shape
    np.testing.assert_array_equal(out, gold[f"prep_{tag}_out"])          # same fp32 expressions: bit-exact
    heat = IO.heatmap_targets(gold[f"prep_{tag}_lands"], tiles.shape[-2], tiles.shape[-1]).numpy()[..., ::step, ::step]
    np.testing.assert_array_equal(heat, gold[f"prep_{tag}_heat"])
    lands = gold[f"prep_{tag}_lands"]
    assert np.all(heat[0, 1] == 0) and np.isinf(lands[0, 0, 1])           # out-of-view landmark -> zero plane


def test_calc_pad_amount():
    assert IO.calc_pad_amount(192, 180) == 6 and IO.calc_pad_amount(32, 21) == 6 and IO.calc_pad_amount(32, 20) == 6
    assert IO.calc_pad_amount(736, 718) == 9 and IO.calc_pad_amount(1440, 1436) == 2


def test_ensemble_combine_matches_seg_dataset_ensemble(gold):
    segs = [torch.from_numpy(s) for s in gold["ens_segs"]]
    heats = [torch.from_numpy(s) for s in gold["ens_heats"]]
    labels, avg = IO.ensemble_combine(segs, heats, gold["ens_labels"].shape[-2:])
    np.testing.assert_array_equal(labels.numpy(), gold["ens_labels"])
    np.testing.assert_array_equal(avg.numpy(), gold["ens_avg_heats"])
    labels2, none = IO.ensemble_combine(segs[:2], None, gold["ens_labels"].shape[-2:])
    assert none is None
    np.testing.assert_array_equal(labels2.numpy(), gold["ens_labels_2nets"])
    # the fixture really exercises the tie rule: some pixels have two classes with the same averaged probability
    m = sum(IO._crop(s, gold["ens_labels"].shape[-2:]) for s in segs) / 3
    top2 = torch.topk(m, 2, dim=1)[0]
    assert int((top2[:, 0] == top2[:, 1]).sum()) > 0


def test_template_and_ncc_match_reference(gold):
    np.testing.assert_array_equal(IO.gaussian_template(25, 25, 2.5).numpy(), gold["tmpl_25"])
    got = IO.ncc_2d(torch.from_numpy(gold["ncc_a"]), torch.from_numpy(gold["ncc_b"])).numpy()
    np.testing.assert_allclose(got, gold["ncc_ab"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("tag", ["seg", "noseg"])
def test_extract_landmarks_matches_est_lands_csv(gold, tag):
    segs = gold["land_segs"] if tag == "seg" else None
    rc, scores = IO.extract_landmarks(gold["land_heats"], segs, list(gold["land_labels"]))
    np.testing.assert_array_equal(rc.numpy(), gold[f"land_rc_{tag}"])
    found = gold[f"land_rc_{tag}"][..., 0] >= 0
    assert 0 < found.sum() < found.size                                    # both outcomes are pinned
    s = scores.numpy()
    assert np.all(np.isnan(s) | (np.abs(s - 0.9) > 1e-3))                  # no score sits on the threshold


# ---- size-independent properties of the restatement itself (the same ones the GPU tests demand of the kernels) ----

def test_oracle_properties_prep_and_round_trip():
    g = torch.Generator().manual_seed(21)
    tiles = torch.rand(3, 26, 26, generator=g) * 5000
    x = IO.prep_tiles(tiles.numpy(), IO.calc_pad_amount(32, 26))
    assert x.shape == (3, 1, 32, 32)
    assert float(x.mean(dim=(1, 2, 3)).abs().max()) < 1e-5 and float((x.std(dim=(1, 2, 3)) - 1).abs().max()) < 1e-5
    p = 3
    assert torch.equal(x[:, :, p - 2, :], x[:, :, p + 2, :]) and torch.equal(x[:, :, :, 31 - p + 1], x[:, :, :, 31 - p - 1])
    lands = torch.stack([torch.randint(13, 35, (3, 4), generator=g), torch.randint(13, 35, (3, 4), generator=g)], 1).float()
    heat = IO.heatmap_targets(lands, 48, 48)
    assert float((heat.sum(dim=(2, 3)) - 1).abs().max()) < 1e-3            # unit-mass Gaussians inside the view
    rc, ncc = IO.extract_landmarks(heat)
    assert torch.equal(rc, torch.stack([lands[:, 1], lands[:, 0]], dim=-1).long())
    assert float((ncc - 624.0 / 625.0).abs().max()) < 1e-5                 # ncc.py:38: unbiased sd under an N-fold sum


def test_oracle_properties_ensemble():
    g = torch.Generator().manual_seed(22)
    seg = torch.softmax(torch.randn(2, 5, 12, 12, generator=g), 1)
    heat = torch.randn(2, 3, 12, 12, generator=g)
    l1, a1 = IO.ensemble_combine([seg], [heat], (10, 10))
    l3, a3 = IO.ensemble_combine([seg, seg, seg], [heat, heat * 7 - 2, heat * 0.01], (10, 10))
    assert torch.equal(l1, l3)                                             # identical members: the same labels
    np.testing.assert_allclose(a3.numpy(), a1.numpy(), rtol=0, atol=2e-6)  # min-max normalisation removes scale and offset
    assert torch.equal(l1.long(), IO._crop(seg, (10, 10)).argmax(dim=1))
    assert float(a1.amin()) == 0.0 and float(a1.amax()) == 1.0
