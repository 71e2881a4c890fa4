import importlib
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "deepfluorolabeling-ipcai2020_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def load_pkg():
    return importlib.import_module(PKG)


def load_golden(name):
    """Returns (meta, dict of torch tensors)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    rec = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return meta, rec


def golden_state(rec, prefix="state/"):
    return {k[len(prefix):]: v for k, v in rec.items() if k.startswith(prefix)}


SMALL_CASES = ["dual_conv_down_train", "dual_conv_down_eval", "dual_maxpool_train", "seg_only_plain_train",
               "seg_only_bn_nores_train", "lands1_nosoftmax_train", "deep4_wf3_train"]


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()
