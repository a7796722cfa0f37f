#!/usr/bin/env python
"""Development check: class maps / logits of the fused conv1+conv2 kernel against the separate kernels (PCSEG_FUSE12=0)
on n A4 pages; prints where they differ."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, out):
    import torch
    from page_segmentation_b200 import synth
    from page_segmentation_b200.lib.network import Network
    from oracle import pipeline as opipe
    imgs = []
    for s in range(min(n, 4)):
        page = synth.make_page(100 + s)
        imgs.append(opipe.prepare_images(page, page, 6, 18)[0])
    imgs = np.stack([imgs[i % len(imgs)] for i in range(n)])
    W = synth.make_weights("fcn_skip", 3, seed=0)
    net = Network("Predict", n_classes=3, weights=W, precision=os.environ.get("PCSEG_PRECISION", "fp16"))
    c = net._context()
    h, w = imgs.shape[1:]
    d_img = torch.from_numpy(imgs).cuda()
    d_labels = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_logits = torch.empty((n, h, w, 3), dtype=torch.float32, device="cuda")
    for rep in range(3):
        c.forward(d_img, None, n, h, w, d_labels, d_logits, None)
        torch.cuda.synchronize()
        np.save(f"{out}_{rep}.npy", d_logits.cpu().numpy())


if __name__ == "__main__":
    if len(sys.argv) > 2:
        run(int(sys.argv[1]), sys.argv[2])
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    for mode in ("1", "0"):
        env = dict(os.environ, PCSEG_FUSE12=mode)
        subprocess.run([sys.executable, __file__, str(n), f"/tmp/f12_{mode}"], env=env, check=True)
    ref = np.load("/tmp/f12_0_0.npy")
    for rep in range(3):
        a = np.load(f"/tmp/f12_1_{rep}.npy")
        d = np.abs(a - ref).max(-1)
        print(f"rep {rep}: max |d logit| {d.max():.4g}, pixels with |d| > 1e-2: {(d > 1e-2).sum()}, argmax differs: {(a.argmax(-1) != ref.argmax(-1)).sum()}")
        bad = d > 1e-2
        for pg in range(n):
            if bad[pg].any():
                ys, xs = np.nonzero(bad[pg])
                print(f"   page {pg}: {bad[pg].sum()} px, rows {ys.min()}..{ys.max()}, cols {xs.min()}..{xs.max()}; rows hist (per 64): "
                      f"{np.bincount(ys // 64, minlength=19).tolist()}; cols hist (per 124): {np.bincount(xs // 124, minlength=7).tolist()}")
