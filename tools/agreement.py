#!/usr/bin/env python
"""Class-map agreement and logit error of the device forward on full A4 pages against the fp32 / fp64 CPU oracle
(oracle/network.py) -- the numbers DESIGN.md section 4 quotes.

    python tools/agreement.py [seeds...] > profiles/rNN_agreement_full_a4.jsonl
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
    from oracle import network as onet
    from oracle import pipeline as opipe
    from page_segmentation_b200.lib.network import Network
    seeds = [int(a) for a in sys.argv[1:]] or [0, 1]
    W = synth.make_weights("fcn_skip", 3, seed=0)
    for s in seeds:
        page = synth.make_page(s)
        img, _ = opipe.prepare_images(page, page, 6, 18)
        l32 = onet.Forward("fcn_skip", W, 3).logits(img)[0]
        l64 = onet.Forward("fcn_skip", W, 3, dtype=torch.float64).logits(img)[0]
        ref = l64.argmax(-1)
        srt = np.sort(l64, -1)
        margin = srt[..., -1] - srt[..., -2]
        for precision in ("bf16", "fp16"):
            net = Network("Predict", n_classes=3, weights=W, precision=precision)
            c = net._context()
            h, w = img.shape
            d_img = torch.from_numpy(img[None]).cuda()
            d_labels = torch.empty((1, h, w), dtype=torch.uint8, device="cuda")
            d_logits = torch.empty((1, h, w, 3), dtype=torch.float32, device="cuda")
            c.forward(d_img, None, 1, h, w, d_labels, d_logits, None)
            torch.cuda.synchronize()
            lg, pred = d_logits[0].cpu().numpy(), d_labels[0].cpu().numpy()
            bad = pred != ref
            print(json.dumps({"page_seed": s, "precision": precision, "pixels": int(pred.size),
                              "max_abs_logit_err_vs_fp64": float(np.abs(lg - l64).max()),
                              "mean_abs_logit_err_vs_fp64": float(np.abs(lg - l64).mean()),
                              "max_abs_logit_err_vs_fp32": float(np.abs(lg - l32).max()),
                              "argmax_agreement_vs_fp64": float((~bad).mean()),
                              "max_fp64_margin_at_disagreement": float(margin[bad].max()) if bad.any() else 0.0,
                              "median_fp64_margin": float(np.median(margin)),
                              "fp32_oracle_agreement_vs_fp64": float((l32.argmax(-1) == ref).mean()),
                              "conv1_debug": os.environ.get("PCSEG_C1_DEBUG", "0")}), flush=True)


if __name__ == "__main__":
    main()
