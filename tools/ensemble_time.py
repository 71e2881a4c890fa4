"""Inference path of test_ensemble.py / est_lands_csv.py on the device, timed end to end:
pinned raw 180^2 tiles -> H2D -> prep_tiles -> n_nets paper networks (eval, no grad, bf16 engine) -> ensemble_combine
-> extract_landmarks -> D2H of the u1 label maps, normalised heat-maps' landmarks (row, col).  CUDA events around the
whole loop and around each stage; one JSON line.  The reference times the same span per image with time.time()
(util.py:321,363-366; est_lands_csv.py:94,131-133); its CPU forward is timed here on the host cores through the
oracle port (1 image, eval) for scale.
usage: python tools/ensemble_time.py [n_images] [n_nets] [batch]"""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
pp = pkg.prepost
PAPER = dict(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True, max_pool=False, num_lands=14, do_res=True,
             block_depth=2)
NAMES = ["FH-l", "FH-r", "GSN-l", "GSN-r", "IOF-l", "IOF-r", "MOF-l", "MOF-r", "SPS-l", "SPS-r", "IPS-l", "IPS-r",
         "ASIS-l", "ASIS-r"]


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    N = int(argv[0]) if len(argv) > 0 else 256
    n_nets = int(argv[1]) if len(argv) > 1 else 3
    B = int(argv[2]) if len(argv) > 2 else 32
    dev = torch.device("cuda:0")
    nets = []
    for k in range(n_nets):
        torch.manual_seed(k)
        nets.append(pkg.UNet(precision="bf16", **PAPER).to(dev).eval())
    g = torch.Generator().manual_seed(0)
    raw = (torch.rand(N, 180, 180, generator=g) * 4000.0).pin_memory()
    out_labels = torch.empty(N, 180, 180, dtype=torch.uint8).pin_memory()
    out_rc = torch.empty(N, 14, 2, dtype=torch.int32).pin_memory()
    stages = ["h2d+prep", "forward", "combine", "landmarks+d2h"]
    # one CUDA graph holds the eval forwards of all networks (pkg.GraphedForward); "--eager" launches them from Python
    graphed = None if "--eager" in sys.argv else pkg.GraphedForward(nets, torch.zeros(B, 1, 192, 192, device=dev))
    ev = None

    def run(record):
        nonlocal ev
        ev = []
        with torch.no_grad():
            for i in range(0, N, B):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record else None
                if record:
                    e[0].record()
                x = pp.prep_tiles(raw[i:i + B].to(dev, non_blocking=True), pad_img_dim=192)
                if record:
                    e[1].record()
                outs = graphed(x) if (graphed is not None and x.shape[0] == B) else [net(x) for net in nets]
                if record:
                    e[2].record()
                labels, heats = pp.ensemble_combine([o[0] for o in outs], [o[1] for o in outs], (180, 180))
                if record:
                    e[3].record()
                rc = pp.extract_landmarks(heats, labels, NAMES)
                out_labels[i:i + B].copy_(labels, non_blocking=True)
                out_rc[i:i + B].copy_(rc, non_blocking=True)
                if record:
                    e[4].record()
                    ev.append(e)

    run(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(False)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    run(True)
    torch.cuda.synchronize()
    split = {s: round(sum(e[k].elapsed_time(e[k + 1]) for e in ev) / N, 5) for k, s in enumerate(stages)}
    found = int((out_rc[..., 0] >= 0).sum())
    # the reference's CPU forward for scale (oracle port, eval, 1 image, all host threads)
    cpu_ms = None
    try:
        from oracle import unet_oracle as O
        cfg = O.UNetConfig(**PAPER)
        sd = {k: v.detach().cpu().clone() for k, v in nets[0].state_dict().items()}
        x1 = torch.randn(1, 1, 192, 192)
        with torch.no_grad():
            O.forward(sd, cfg, x1, training=False)
            t0 = time.perf_counter()
            for _ in range(3):
                O.forward(sd, cfg, x1, training=False)
            cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    except Exception as ex:  # the timing above stands on its own
        print("cpu forward not timed:", ex, file=sys.stderr)
    print(json.dumps({"op": "ensemble inference, tiles -> labels + landmark pixels", "n_images": N, "n_nets": n_nets, "batch": B,
                      "precision": "bf16", "forward_launch": "eager" if graphed is None else "one CUDA graph for all networks", "ms_per_image": round(total_ms / N, 4), "images_per_s": round(N / total_ms * 1e3, 1),
                      "ms_per_image_by_stage": split, "landmarks_reported": found,
                      "h2d_bytes_per_image": 180 * 180 * 4, "d2h_bytes_per_image": 180 * 180 + 14 * 2 * 4,
                      "cpu_forward_ms_per_image_per_net": cpu_ms and round(cpu_ms, 1), "cpu_threads": torch.get_num_threads()}))


if __name__ == "__main__":
    main()
