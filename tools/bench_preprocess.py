#!/usr/bin/env python
"""Device time of the preprocess stage alone (prepare_images, dataset.py:114-150) at BASELINE configs[1]'s shape:
64 binarised A4 pages -> 1169x827, CUDA events on the launch stream, inputs (557 MB) larger than L2.  Checks the first
pages against the oracle before timing.

    python tools/bench_preprocess.py [pages] > profiles/rNN_preprocess.jsonl
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from page_segmentation_b200 import synth  # noqa: E402


def main():
    import torch
    from page_segmentation_b200 import runtime
    from oracle import pipeline as opipe
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ctx = runtime.get_context(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6550.0)
    H, W = synth.A4_H, synth.A4_W
    Hs, Ws = synth.scaled_shape(H, W, 1 / 3)
    pages = np.stack([synth.make_page(s) for s in range(8)])
    d_pages = torch.from_numpy(np.concatenate([pages] * (n // 8))).cuda()
    d_image = torch.empty((n, Hs, Ws), dtype=torch.uint8, device="cuda")
    d_binary = torch.empty((n, Hs, Ws), dtype=torch.uint8, device="cuda")
    ctx.preprocess(d_pages, d_pages, n, H, W, Hs, Ws, d_image, d_binary, None)
    torch.cuda.synchronize()
    img, binr = d_image.cpu().numpy(), d_binary.cpu().numpy()
    bad = 0
    for s in (0, 5):
        o_img, o_bin = opipe.prepare_images(pages[s], pages[s], 6, 18)
        bad += int((img[s] != o_img).sum()) + int((binr[s] != o_bin).sum())
        bad += int((img[n - 8 + s] != o_img).sum())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    for _ in range(3):
        ctx.preprocess(d_pages, d_pages, n, H, W, Hs, Ws, d_image, d_binary, None)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ctx.preprocess(d_pages, d_pages, n, H, W, Hs, Ws, d_image, d_binary, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_per_page = H * W + 2 * Hs * Ws
    gbs = bytes_per_page * n / ms / 1e6
    print(json.dumps({"stage": "preprocess", "pages": n, "ms": round(ms, 4), "mismatching_bytes_vs_oracle": bad,
                      "algorithmic_MB_per_page": round(bytes_per_page / 1e6, 2), "achieved_GBs": round(gbs, 1),
                      "frac_of_measured_hbm": round(gbs / peak, 4),
                      "tiles_per_block": os.environ.get("PCSEG_RESAMPLE_TILES", "default")}), flush=True)


if __name__ == "__main__":
    main()
