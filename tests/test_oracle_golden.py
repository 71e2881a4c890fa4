"""The oracle (oracle/unet_oracle.py) against the fixtures generated from the real
reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from conftest import SMALL_CASES, golden_state, load_golden, rel_l2
from oracle import unet_oracle as O

TOL = 2e-5   # fp32 CPU vs fp32 CPU, different summation order only


@pytest.mark.parametrize("name", SMALL_CASES)
def test_forward_matches_reference(name):
    meta, rec = load_golden(name)
    cfg = O.UNetConfig(**meta["kwargs"])
    sd = golden_state(rec)
    out = O.forward(sd, cfg, rec["x"], training=meta["training"])
    assert rel_l2(out["logits"], rec["logits"]) < TOL
    assert rel_l2(out["seg"], rec["seg"]) < TOL
    if "heat" in rec:
        assert rel_l2(out["heat"], rec["heat"]) < TOL
    else:
        assert out["heat"] is None
    if meta["training"]:
        after = golden_state(rec, "state_after/")
        for k, v in after.items():
            if "num_batches" in k:
                assert int(out["new_stats"][k]) == int(v)
            else:
                assert rel_l2(out["new_stats"][k], v) < TOL, k


@pytest.mark.parametrize("name", SMALL_CASES)
def test_backward_matches_reference_autograd(name):
    meta, rec = load_golden(name)
    cfg = O.UNetConfig(**meta["kwargs"])
    sd = golden_state(rec)
    out = O.forward(sd, cfg, rec["x"], training=meta["training"], want_tape=True)
    grads = O.backward(sd, cfg, out["tape"], rec["d_seg"], rec.get("d_heat"))
    ref = golden_state(rec, "grad/")
    assert set(grads) == set(ref), set(grads) ^ set(ref)
    for k in meta["none_grads"]:
        assert k not in grads
    for k, g in ref.items():
        assert grads[k].shape == g.shape, k
        assert rel_l2(grads[k], g) < 2e-4, (k, rel_l2(grads[k], g))


@pytest.mark.parametrize("name", SMALL_CASES)
def test_schema_matches_reference_state_dict(name):
    meta, rec = load_golden(name)
    cfg = O.UNetConfig(**meta["kwargs"])
    sd = golden_state(rec)
    schema = O.param_schema(cfg)
    assert [n for n, _, _ in schema] == list(sd.keys())
    for n, shape, _ in schema:
        assert tuple(sd[n].shape) == tuple(shape), n


def test_paper_config_eval_matches_reference():
    """BASELINE.json config 1.  Weights are rebuilt from the seed through the oracle's own
    schema using torch's default initialisers in reference order (see test_module_cpu for the
    engine-side module); here we only need forward parity on identical weights, so rebuild
    via torch modules in the reference's construction order."""
    from conftest import load_pkg
    meta, rec = load_golden("paper_eval_192")
    pkg = load_pkg()
    torch.manual_seed(0)
    net = pkg.UNet(**meta["kwargs"])
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    for k, (s, sa) in meta["param_sums"].items():
        assert abs(float(sd[k].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(float(sd[k].double().abs().sum()) - sa) <= 1e-9 * max(1.0, sa), k
    cfg = O.UNetConfig(**meta["kwargs"])
    g = torch.Generator().manual_seed(meta["warm"]["seed"])
    for _ in range(meta["warm"]["n_iter"]):
        out = O.forward(sd, cfg, torch.randn(*meta["warm"]["shape"], generator=g), training=True)
        sd.update(out["new_stats"])
    for k, v in golden_state(rec, "bn_after/").items():
        assert rel_l2(sd[k], v) < 1e-4, k
    out = O.forward(sd, cfg, rec["x"], training=False)
    assert rel_l2(out["logits"][:, :, ::4, ::4], rec["logits_s4"]) < 1e-4
    assert rel_l2(out["seg"][:, :, ::4, ::4], rec["seg_s4"]) < 1e-4
    assert rel_l2(out["heat"][:, :, ::4, ::4], rec["heat_s4"]) < 1e-4
    assert abs(float(out["heat"].double().sum()) - float(rec["heat_moments"][0])) < 1e-3 * abs(float(rec["heat_moments"][0])) + 1e-2


def test_losses_match_reference_dice_and_ncc():
    import numpy as np
    import os
    from conftest import GOLDEN, load_pkg
    z = np.load(os.path.join(GOLDEN, "losses.npz"))
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    pkg = load_pkg()
    for impl in ("oracle", "product"):
        seg = t["seg"].clone().requires_grad_(True)
        heat = t["heat"].clone().requires_grad_(True)
        if impl == "oracle":
            l = O.dice_and_heatmap_loss(seg, heat, t["tgt_seg"], t["tgt_heat"], skip_bg=False, heatmap_wgt=0.5)
            l_bg = O.dice_loss(t["seg"], t["tgt_seg"], skip_bg=True)
            l_d = O.dice_loss(t["seg"], t["tgt_seg"], skip_bg=False)
        else:
            l = pkg.DiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=0.5)((seg, heat), (t["tgt_seg"], t["tgt_heat"]))
            l_bg = pkg.DiceLoss2D(skip_bg=True)(t["seg"], t["tgt_seg"])
            l_d = pkg.DiceLoss2D(skip_bg=False)(t["seg"], t["tgt_seg"])
        l.backward()
        assert abs(float(l) - float(t["l_dual"])) < 1e-6
        assert abs(float(l_bg) - float(t["l_dice_bg"])) < 1e-6
        assert abs(float(l_d) - float(t["l_dice"])) < 1e-6
        assert rel_l2(seg.grad, t["d_seg"]) < 1e-5
        assert rel_l2(heat.grad, t["d_heat"]) < 1e-5


def test_oracle_training_step_matches_reference_at_the_benchmarked_size():
    """BASELINE.json configs[1]: the oracle's explicit forward + hand-derived backward on 32 tiles at 192x192 against
    the real reference's training step (tests/golden/make_golden_large.py): outputs, loss and every parameter
    gradient.  Pins the checker at the size the bench runs."""
    import importlib.util
    import os
    from conftest import GOLDEN, load_pkg
    spec = importlib.util.spec_from_file_location("make_golden_large", os.path.join(GOLDEN, "make_golden_large.py"))
    mgl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgl)
    meta, rec = load_golden("large_paper_b32_192")
    case = meta["case"]
    torch.manual_seed(meta["init_seed"])
    net = load_pkg().UNet(**case["kwargs"])           # parameter container: same init stream as the reference
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.UNetConfig(**case["kwargs"])
    x, mask, heat_t = mgl.make_inputs(case, seed=meta["input_seed"])
    out = O.forward(sd, cfg, x, training=True, want_tape=True)
    st = case["out_stride"]
    assert rel_l2(out["seg"][:, :, ::st, ::st], rec["seg_s"]) < 1e-4
    assert rel_l2(out["heat"][:, :, ::st, ::st], rec["heat_s"]) < 1e-4
    loss, d_seg, d_heat = O.loss_and_output_grads(out, cfg, mask, heat_t, heatmap_wgt=case["heatmap_wgt"])
    assert abs(float(loss) - float(rec["loss"])) < 1e-5
    grads = O.backward(sd, cfg, out["tape"], d_seg, d_heat)
    for n, stride in meta["grad_strides"].items():
        # two fp32 CPU computations in different summation orders, judged against the reference's fp64 run: the oracle
        # may be as far from that truth as the fp32 reference is (its `floor`), within a small factor
        smp = grads[n].flatten()[::stride]
        e = rel_l2(smp, rec["grad_sample64/" + n])
        floor = rel_l2(rec["grad_sample/" + n], rec["grad_sample64/" + n])
        assert e < max(1e-3 if grads[n].dim() > 1 else 1e-2, 6 * floor), (n, e, floor)
    for k, v in rec.items():
        if k.startswith("state_after/"):
            assert rel_l2(out["new_stats"][k[len("state_after/"):]], v) < 1e-4, k
