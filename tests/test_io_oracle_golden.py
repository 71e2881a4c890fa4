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
