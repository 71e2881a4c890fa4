"""Generate tests/golden/io.npz from the REAL reference (run once in the authoring container):

    python tests/golden/make_golden_io.py

Runs, unmodified and on seeded synthetic data,
  * dataset.RandomDataAugDataSet.__getitem__ (dataset.py:91-328, augmentation probability 0) -> padded,
    z-scored tiles and Gaussian heat-map targets,
  * util.seg_dataset_ensemble (util.py:293-377) with stand-in networks that return stored outputs,
  * the est_lands_csv.py script itself (as a subprocess), with and without --use-seg,
with tests/golden/_h5stub/h5py.py standing in for the h5py package this image does not have.
Nothing at test/bench time reads /root/reference; the tests read only io.npz."""
import csv
import math
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference/train_test_code"
HERE = os.path.dirname(os.path.abspath(__file__))
STUB = os.path.join(HERE, "_h5stub")

LAND_NAMES = ["FH-l", "FH-r", "GSN-l", "GSN-r", "IOF-l", "IOF-r", "MOF-l", "MOF-r", "SPS-l", "SPS-r",
              "IPS-l", "IPS-r", "ASIS-l", "ASIS-r"]
# est_lands_csv.py:54-71
LAND_LABELS = [5, 6, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2]


def gen_prep(out):
    import dataset  # the reference module
    g = torch.Generator().manual_seed(11)
    for tag, (n, dim, pad_dim, L) in {"small": (3, 20, 32, 5), "odd": (2, 21, 32, 3), "paper": (1, 180, 192, 14)}.items():
        projs = torch.rand(n, 1, dim, dim, generator=g) * 3000.0 + 500.0 * torch.randn(n, 1, 1, 1, generator=g)
        segs = torch.zeros(n, 7, dim, dim)
        lands = torch.rand(n, 2, L, generator=g) * (dim + 10) - 5.0   # some outside the tile but finite
        lands[0, :, 1] = math.inf                                       # dataset.py: out-of-view landmark
        if L > 2:
            lands[n - 1, 0, 2] = -math.inf
        ds = dataset.RandomDataAugDataSet(projs.clone(), segs, lands.clone(), proj_pad_dim=pad_dim)
        ds.prob_of_aug = 0
        ps, hs = [], []
        for i in range(n):
            p, s, cl, h = ds[i]
            ps.append(p)
            hs.append(h[:, 0])
        out[f"prep_{tag}_tiles"] = projs[:, 0].numpy()
        out[f"prep_{tag}_pad"] = np.int64(ds.extra_pad)
        out[f"prep_{tag}_lands"] = lands.numpy()
        P, Hm = torch.stack(ps).numpy(), torch.stack(hs).numpy()
        if tag == "paper":   # keep the fixture small: every 5th row/column
            P, Hm = P[..., ::5, ::5], Hm[..., ::5, ::5]
        out[f"prep_{tag}_out"] = P
        out[f"prep_{tag}_heat"] = Hm


def gen_ensemble(out):
    import h5py  # the stand-in
    import util  # the reference module
    g = torch.Generator().manual_seed(12)
    n_nets, n_img, C, L, H, h = 3, 2, 7, 14, 40, 36
    segs = torch.softmax(torch.randn(n_nets, n_img, C, H, H, generator=g) * 2, dim=2)
    segs = torch.round(segs * 8) / 8          # coarse values: plenty of exact ties between classes
    heats = torch.randn(n_nets, n_img, L, H, H, generator=g) * torch.tensor([0.02, 1.0, 30.0]).view(3, 1, 1, 1, 1) + 0.3

    class FakeNet(torch.nn.Module):
        def __init__(self, k):
            super().__init__()
            self.k = k

        def forward(self, projs):
            i = int(projs.flatten()[0].item())
            return segs[self.k, i:i + 1].clone(), heats[self.k, i:i + 1].clone()

    class DS(torch.utils.data.Dataset):
        rob_orig_img_shape = (h, h)

        def __len__(self):
            return n_img

        def __getitem__(self, i):
            return (torch.full((1, H, H), float(i)),)

    with tempfile.TemporaryDirectory() as td:
        f = h5py.File(os.path.join(td, "o.pkl"), "w")
        times = []
        util.seg_dataset_ensemble(DS(), [FakeNet(k) for k in range(n_nets)], f, dev=None, num_lands=L, times=times)
        out["ens_segs"] = segs.numpy()
        out["ens_heats"] = heats.numpy()
        out["ens_labels"] = f.store["nn-segs"].copy()
        out["ens_avg_heats"] = f.store["nn-heats"].copy()
        # seg-only ensemble (num_lands = 0, nets return a single tensor)
        class SegNet(FakeNet):
            def forward(self, projs):
                return super().forward(projs)[0]
        f2 = h5py.File(os.path.join(td, "o2.pkl"), "w")
        util.seg_dataset_ensemble(DS(), [SegNet(k) for k in range(2)], f2, dev=None, num_lands=0)
        out["ens_labels_2nets"] = f2.store["nn-segs"].copy()


def gen_landmarks(out):
    g = torch.Generator().manual_seed(13)
    P, L, h = 3, 14, 48
    Y, X = torch.meshgrid(torch.arange(h).float(), torch.arange(h).float(), indexing="ij")
    heats = torch.randn(P, L, h, h, generator=g) * 2e-4
    segs = torch.randint(0, 7, (P, h, h), generator=g).to(torch.uint8)
    for p in range(P):
        for l in range(L):
            cy, cx = [float(v) for v in torch.rand(2, generator=g) * (h - 1)]
            if l % 5 == 1:
                cy = float(l % 3)                    # peaks at the border: the reflect-padded window matters
            sig = 2.5 if l % 4 != 3 else 6.0         # wide blobs fail the NCC >= 0.9 test
            amp = 1.0 if l % 7 != 6 else 0.0         # pure-noise planes fail it too
            heats[p, l] += amp * torch.exp(-((X - cx) ** 2 + (Y - cy) ** 2) / (2 * sig * sig)) / (2 * math.pi * sig * sig)
            if l % 3 == 0:                           # make the landmark's anatomy label present around the peak
                r, c = int(round(cy)), int(round(cx))
                segs[p, max(r - 2, 0):r + 3, max(c - 2, 0):c + 3] = LAND_LABELS[l]
    segs[2][segs[2] == 5] = 0                        # label 5 absent in projection 2: FH-l cannot be found there
    store = {"nn-heats": heats.numpy(), "nn-segs": segs.numpy(), "land-names/num-lands": np.int64(L)}
    for l, nm in enumerate(LAND_NAMES):
        store[f"land-names/land-{l:02d}"] = nm
    res = {}
    with tempfile.TemporaryDirectory() as td:
        fp = os.path.join(td, "heats.pkl")
        with open(fp, "wb") as f:
            pickle.dump(store, f)
        env = dict(os.environ, PYTHONPATH=STUB + os.pathsep + os.environ.get("PYTHONPATH", ""))
        for tag, extra in (("seg", ["--use-seg", "nn-segs"]), ("noseg", [])):
            oc = os.path.join(td, tag + ".csv")
            subprocess.run([sys.executable, os.path.join(REF, "est_lands_csv.py"), fp, "nn-heats", "--out", oc,
                            "--pat", "1"] + extra, check=True, env=env, cwd=td, stdout=subprocess.DEVNULL)
            rc = np.zeros((P, L, 2), dtype=np.int64)
            with open(oc) as f:
                for row in csv.DictReader(f):
                    rc[int(row["proj"]), int(row["land"])] = (int(row["row"]), int(row["col"]))
            res[tag] = rc
    out["land_heats"] = heats.numpy()
    out["land_segs"] = segs.numpy()
    out["land_labels"] = np.asarray(LAND_LABELS, dtype=np.int64)
    out["land_rc_seg"] = res["seg"]
    out["land_rc_noseg"] = res["noseg"]
    # pins of the two helpers the script leans on
    import ncc
    import util
    t = util.get_gaussian_2d_heatmap(25, 25, 2.5)
    out["tmpl_25"] = t.numpy()
    a, b = torch.rand(4, 9, 11, generator=g), torch.rand(4, 9, 11, generator=g)
    out["ncc_a"], out["ncc_b"], out["ncc_ab"] = a.numpy(), b.numpy(), ncc.ncc_2d(a, b).numpy()


def main():
    sys.path.insert(0, STUB)
    sys.path.insert(0, REF)
    out = {}
    gen_prep(out)
    gen_ensemble(out)
    gen_landmarks(out)
    path = os.path.join(HERE, "io.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})
    print("landmarks found (seg / noseg):", int((out["land_rc_seg"][..., 0] >= 0).sum()), int((out["land_rc_noseg"][..., 0] >= 0).sum()),
          "of", out["land_rc_seg"].shape[0] * out["land_rc_seg"].shape[1])


if __name__ == "__main__":
    main()
