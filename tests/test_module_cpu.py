"""Host-side checks that need no GPU: the module mirrors the reference's interface,
the C-ABI library loads and exports every declared symbol, and invalid
configurations are rejected loudly."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, golden_state, load_golden, load_pkg
from oracle import unet_oracle as O


@pytest.fixture(scope="module")
def pkg():
    p = load_pkg()
    p.build_library()
    return p


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "fluoro_unet.h")).read()
    declared = set(re.findall(r"\b(fu_[a-z_0-9]+)\s*\(", hdr))
    L = pkg._capi.lib()
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(L, sym), sym
    assert declared == set(pkg._capi.EXPORTS)
    assert b"sm_100a" in L.fu_build_info()


def test_library_has_no_torch_dependency(pkg):
    import subprocess
    out = subprocess.run(["ldd", pkg._capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


@pytest.mark.parametrize("name", ["dual_conv_down_train", "dual_maxpool_train", "seg_only_plain_train",
                                  "seg_only_bn_nores_train", "lands1_nosoftmax_train"])
def test_state_dict_schema_matches_reference(pkg, name):
    meta, rec = load_golden(name)
    net = pkg.UNet(**meta["kwargs"])
    ref = golden_state(rec)
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
        assert sd[k].dtype == ref[k].dtype, k
    net.load_state_dict(ref)          # train.py:316 / test_ensemble.py:100
    assert [n for n, _, _ in O.param_schema(O.UNetConfig(**meta["kwargs"]))] == list(sd.keys())


def test_default_init_is_rng_identical_to_reference(pkg):
    meta, _ = load_golden("paper_eval_192")
    torch.manual_seed(0)
    net = pkg.UNet(**meta["kwargs"])
    assert sum(p.numel() for p in net.parameters()) == 38100089
    for k, v in net.state_dict().items():
        s, sa = meta["param_sums"][k]
        assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(float(v.double().abs().sum()) - sa) <= 1e-9 * max(1.0, sa), k


@pytest.mark.parametrize("kw", [dict(padding=False), dict(padding=True, pad_mode="circular"),
                                dict(padding=True, up_mode="upsample"), dict(padding=True, lands_block_depth=1),
                                dict(padding=True, precision="fp8"), dict(padding=True, precision="bf16", wf=2)])
def test_unsupported_configs_are_rejected(pkg, kw):
    with pytest.raises(ValueError):
        pkg.UNet(**kw)


def test_cpu_input_raises_instead_of_falling_back(pkg):
    net = pkg.UNet(padding=True, depth=2, wf=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 1, 8, 8))


def test_engine_create_without_gpu_fails_loudly(pkg):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = pkg._capi.lib()
    cfg = pkg._capi.FuConfig(in_channels=1, n_classes=2, depth=2, wf=2, padding=1, pad_mode_zeros=1, batch_norm=1,
                             up_mode_upconv=1, max_pool=0, num_lands=0, do_res=1, block_depth=2,
                             lands_block_depth=0, lands_num_1x1=2, do_soft_max=1, precision=0)
    h = ctypes.c_void_p()
    rc = L.fu_engine_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == -3 and h.value is None
    assert b"no CPU fallback" in L.fu_last_error(None)
    cfg.padding = 0
    assert L.fu_engine_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1


def test_center_crop_matches_reference_semantics(pkg):
    x = torch.arange(2 * 3 * 10 * 12, dtype=torch.float32).reshape(2, 3, 10, 12)
    y = pkg.center_crop(x, (2, 3, 6, 7))
    assert y.shape == (2, 3, 6, 7) and torch.equal(y, x[:, :, 2:8, 2:9])
    assert pkg.center_crop(x, x.shape) is x
    assert torch.equal(pkg.center_crop(x[0], (3, 4, 4)), x[0][:, 3:7, 4:8])


def test_losses_match_reference_formulas(pkg):
    """dice.py / ncc.py mirrors against the oracle's restatement (itself pinned on the reference
    formulas) on CPU tensors."""
    g = torch.Generator().manual_seed(0)
    seg = torch.softmax(torch.randn(3, 7, 20, 20, generator=g), dim=1)
    heat = torch.randn(3, 14, 20, 20, generator=g)
    tgt_seg = torch.nn.functional.one_hot(torch.randint(0, 7, (3, 20, 20), generator=g), 7).permute(0, 3, 1, 2).float()
    tgt_heat = torch.rand(3, 14, 20, 20, generator=g)
    a = pkg.DiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)((seg, heat), (tgt_seg, tgt_heat))
    b = O.dice_and_heatmap_loss(seg, heat, tgt_seg, tgt_heat, skip_bg=False, heatmap_wgt=0.5)
    assert abs(float(a) - float(b)) < 1e-6
    assert abs(float(pkg.DiceLoss2D(skip_bg=True)(seg, tgt_seg)) - float(O.dice_loss(seg, tgt_seg, True))) < 1e-6


def test_fused_loss_and_graphed_step_refuse_cpu_tensors(pkg):
    """The device-only helpers fail loudly instead of falling back to a CPU path."""
    seg = torch.rand(1, 7, 8, 8)
    with pytest.raises(RuntimeError):
        pkg.FusedDiceLoss2D()(seg, seg)
    with pytest.raises(RuntimeError):
        pkg.FusedDiceAndHeatMapLoss2D()((seg, seg), (seg, seg))
    with pytest.raises(ValueError):
        pkg.GraphedStep(lambda x: x, (seg,))


def test_fused_loss_descriptor_matches_header(pkg):
    """ctypes mirror of fu_loss_desc: 4 x (pointer + 3 int64 strides) + 6 int32 + 2 float, no hidden padding."""
    import ctypes as C
    assert C.sizeof(pkg._capi.FuLossDesc) == 4 * (8 + 24) + 6 * 4 + 2 * 4


def test_prepost_helpers_refuse_cpu_tensors_and_validate_arguments(pkg):
    """prepost.py has no CPU path; the C ABI rejects bad shapes before touching the device."""
    import ctypes as C
    pp = pkg.prepost
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pp.prep_tiles(torch.zeros(2, 8, 8), pad_img_dim=12)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pp.heatmap_targets(torch.zeros(2, 2, 3), (8, 8))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pp.ensemble_combine([torch.zeros(1, 2, 8, 8)], None, (8, 8))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pp.extract_landmarks(torch.zeros(1, 2, 30, 30))
    assert pp.calc_pad_amount(192, 180) == 6 and pp.calc_pad_amount(32, 21) == 6       # dataset.py:26-40
    assert pp.SEG_LABELS_FOR_LANDS["FH-l"] == 5 and pp.SEG_LABELS_FOR_LANDS["ASIS-r"] == 2 and len(pp.SEG_LABELS_FOR_LANDS) == 18
    L = pkg._capi.lib()
    one = C.c_void_p(8)   # any non-null address: validation fails before it is used
    assert L.fu_prep_tiles(one, 1, 8, 8, 8, 0, None, one, None) == -2 and "pad < tile" in pkg._capi.last_error(None)
    assert L.fu_prep_tiles(one, 1, 8, 8, 2, 1, None, one, None) == -6                   # normalise without workspace
    assert L.fu_heatmap_targets(one, 1, 14, 8, 8, 0.0, one, None) == -6                 # sigma must be positive
    assert L.fu_ensemble_workspace_words(3, 32) == 192
    ptrs = (C.c_void_p * 17)(*([8] * 17))
    assert L.fu_ensemble_combine(ptrs, None, 17, 1, 7, 0, 8, 8, 0, 0, 8, 8, None, one, None, None) == -6
    assert "at most 16" in pkg._capi.last_error(None)
    assert L.fu_ensemble_combine(ptrs, None, 2, 1, 7, 0, 8, 8, 1, 0, 8, 8, None, one, None, None) == -6   # window outside
    assert L.fu_extract_landmarks(one, None, None, 1, 14, 30, 30, 24, 2.5, 0.9, one, None, None) == -6    # even template
    assert L.fu_extract_landmarks(one, None, None, 1, 14, 12, 30, 25, 2.5, 0.9, one, None, None) == -2    # 12 >= h
    assert L.fu_extract_landmarks(one, None, None, 1, 65, 30, 30, 25, 2.5, 0.9, one, None, None) == -6    # > 64 landmarks
