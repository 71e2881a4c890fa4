"""Parity at BASELINE.json's real sizes against outputs of the REAL reference (tests/golden/make_golden_large.py):
configs[1] B=32 @192^2 training step, configs[2] B=8 @736^2 seg-only, configs[4] B=2 @1440^2 dual head with
heatmap_wgt 1.0.  One training step each (train.py:405-422): outputs, loss, BN running statistics and EVERY parameter
gradient (norm + a strided sample of each tensor).

Bounds
  * parity modes (fp32 on the CUDA cores, parity_tc on the tensor cores): north_star's 1e-3 on the outputs (measured
    ~1e-5); gradients per tensor against the reference's own fp64 run: 3e-2, or eight times the distance of the fp32
    REFERENCE to that truth, whichever is larger, and cosine > 0.999.  (A ReLU whose pre-activation sits within
    rounding of zero may legitimately flip; at B=32 @192^2 that alone puts two correct fp32 implementations up to
    9e-3 apart on single tensors, and the noise grows with the square root of the forward perturbation.)
  * bf16 throughput mode: single-pass bf16 operands cannot meet 1e-3; what is asserted per gradient
    tensor is structural: the direction (cosine over the sample) and the norm must agree with the reference, so a
    wrong tap, a dropped bias gradient or a wrong split-K partial in ONE layer fails, while rounding noise passes.
    The per-tensor relative errors are written to gpurun_out/parity_report.jsonl.
"""
import importlib.util
import json
import os

import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden, load_pkg, rel_l2

pytestmark = pytest.mark.gpu
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")

_spec = importlib.util.spec_from_file_location("make_golden_large", os.path.join(GOLDEN, "make_golden_large.py"))
_mgl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mgl)          # only make_inputs / grad_sample are used; nothing here touches /root/reference


@pytest.fixture(scope="module")
def pkg():
    assert torch.cuda.is_available()
    return load_pkg()


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))


CASES = [("paper_b32_192", "fp32"), ("paper_b32_192", "parity_tc"), ("paper_b32_192", "bf16"),
         ("seg_b8_736", "parity_tc"), ("seg_b8_736", "bf16"),
         ("dual_b2_1440", "parity_tc"), ("dual_b2_1440", "bf16")]


@pytest.mark.parametrize("name,precision", CASES)
def test_training_step_matches_the_reference_at_full_size(pkg, name, precision):
    meta, rec = load_golden("large_" + name)
    case = meta["case"]
    dev = torch.device("cuda:0")
    torch.manual_seed(meta["init_seed"])
    net = pkg.UNet(precision=precision, **case["kwargs"])
    for k, v in net.state_dict().items():          # same init as the reference run
        s = meta["param_sums"][k]
        assert abs(float(v.double().sum()) - s[0]) <= 1e-6 * max(1.0, abs(s[1])), k
    x, mask, heat_t = _mgl.make_inputs(case, seed=meta["input_seed"])
    assert abs(float(x.double().sum()) - float(rec["x_sums"][0])) < 1e-6 * float(rec["x_sums"][1])
    assert float(mask.double().sum()) == float(rec["mask_sum"])
    net.to(dev).train()
    if case["heatmap_wgt"] is None:
        seg, heat = net(x.to(dev)), None
        loss = pkg.FusedDiceLoss2D(skip_bg=False)(seg, mask.to(dev))
    else:
        seg, heat = net(x.to(dev))
        loss = pkg.FusedDiceAndHeatMapLoss2D(skip_bg=False, heatmap_wgt=case["heatmap_wgt"])((seg, heat), (mask.to(dev), heat_t.to(dev)))
    loss.backward()
    torch.cuda.synchronize()
    st = case["out_stride"]
    out = {"loss": abs(float(loss.detach()) - float(rec["loss"])),
           "seg": rel_l2(seg.detach().cpu()[:, :, ::st, ::st], rec["seg_s"])}
    if heat is not None:
        out["heat"] = rel_l2(heat.detach().cpu()[:, :, ::st, ::st], rec["heat_s"])
    sd = net.state_dict()
    out["bn_running"] = max(rel_l2(sd[k[len("state_after/"):]].cpu(), v) for k, v in rec.items() if k.startswith("state_after/"))
    per, norms = {}, {}
    for n, p in net.named_parameters():
        if n in meta["none_grads"]:
            assert p.grad is None, n
            continue
        assert p.grad is not None, n
        smp = p.grad.detach().cpu().flatten()[::meta["grad_strides"][n]]
        ref, ref64 = rec["grad_sample/" + n], rec["grad_sample64/" + n]
        nrm = float(p.grad.double().norm()) / (float(rec["grad_norm/" + n][0]) + 1e-300)
        # error against the fp64 truth; floor = the fp32 reference's own distance to it
        per[n] = (rel_l2(smp, ref64), _cos(smp, ref64), nrm, p.dim(), rel_l2(ref, ref64))
        norms[n] = (float(ref64.double().norm()), float(smp.double().norm()), float(ref.double().norm()))
    # gradients that vanish identically (dual_b2_1440: with heatmap_wgt = 1 the loss is the NCC, which is invariant to
    # the per-channel constants the last block's biases add, so their gradient is pure rounding noise in every
    # implementation) are held to an absolute bound instead of a relative one: ten times the fp32 reference's own noise
    med = sorted(v[0] for v in norms.values())[len(norms) // 2]
    for n in [k for k, v in norms.items() if v[0] < 1e-6 * med]:
        # (bf16: the exact cancellation of ~4M rounded terms leaves ~1e-2 of a typical gradient norm)
        assert norms[n][1] < (5e-2 * med if precision == "bf16" else 10 * norms[n][2] + 1e-4 * med), (n, norms[n], med)
        del per[n]
    worst = sorted(per.items(), key=lambda kv: -kv[1][0])[:5]
    cnt = net.engine_counters()
    with open(REPORT, "a") as f:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        f.write(json.dumps({"test": "large_golden", "case": name, "precision": precision, "out": out,
                            "worst_grads(err_vs_fp64,cos,norm_ratio,ndim,fp32_reference_floor)": [(k, [round(x, 6) for x in v]) for k, v in worst],
                            "n_tensors_worse_than_4x_floor": sum(1 for v in per.values() if v[0] > 4 * v[4] + 1e-4),
                            "median_grad_err": sorted(v[0] for v in per.values())[len(per) // 2],
                            "min_cos": min(v[1] for v in per.values()),
                            "tc_kernel_launches": cnt["tc_kernel_launches"]}) + "\n")
    if precision == "bf16":
        assert out["seg"] < 3e-2 and out.get("heat", 0.0) < 3e-2 and out["loss"] < 2e-2, out
        assert out["bn_running"] < 3e-2, out
        for n, (err, cos, nrm, dim, floor) in per.items():
            assert cos > (0.9 if dim > 1 else 0.7), (n, err, cos, nrm)
            assert 0.7 < nrm < 1.4, (n, err, cos, nrm)
    else:
        assert out["seg"] < 1e-3 and out.get("heat", 0.0) < 1e-3 and out["loss"] < 1e-5, out       # north_star: 1e-3
        assert out["bn_running"] < 1e-4, out
        # Gradient noise at these sizes is dominated by ReLU masks that flip within rounding of zero: a forward
        # perturbation eps flips a fraction ~eps of the masks and moves the gradient by ~sqrt(eps) (measured: fp32
        # reference vs its fp64 run 1e-6 -> 4e-3 median; parity_tc 1e-5 -> 1.2e-2; bf16 1e-2 -> 0.3).  A parity mode may
        # therefore sit a small factor above the fp32 reference's own distance to the truth, never more.
        for n, (err, cos, nrm, dim, floor) in per.items():
            assert err < max(3e-2, 8 * floor), (n, err, cos, nrm, floor)
            assert cos > 0.999, (n, err, cos, nrm, floor)
    assert (cnt["tc_kernel_launches"] > 0) == (precision != "fp32")
