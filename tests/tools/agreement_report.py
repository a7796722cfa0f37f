"""Full-size parity report (run on the GPU box): device logits / argmax vs the fp32 and fp64 CPU oracle on
synthetic A4 pages, for both operand precisions.  Writes one JSON line per (precision, page)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from oracle import network as onet
from oracle import pipeline as opipe
from page_segmentation_b200 import synth
from page_segmentation_b200.lib.dataset import SingleData
from page_segmentation_b200.lib.network import Network

W = synth.make_weights("fcn_skip", 3, seed=0)
f32 = onet.Forward("fcn_skip", W, 3)
f64 = onet.Forward("fcn_skip", W, 3, dtype=torch.float64)
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    page = synth.make_page(seed)
    img, b = opipe.prepare_images(page, page, 6, 18)
    l32, _ = f32.logits(img)
    l64, _ = f64.logits(img)
    s = np.sort(l64, -1)
    margin = s[..., -1] - s[..., -2]
    for prec in ("bf16", "fp16"):
        net = Network("Predict", n_classes=3, weights=W, precision=prec)
        logit, prob, pred = net.predict_single_data(SingleData(image=img))
        bad = pred != l64.argmax(-1)
        print(json.dumps({"page_seed": seed, "precision": prec, "pixels": int(pred.size),
                          "max_abs_logit_err_vs_fp64": float(np.abs(logit - l64).max()),
                          "mean_abs_logit_err_vs_fp64": float(np.abs(logit - l64).mean()),
                          "max_abs_logit_err_vs_fp32": float(np.abs(logit - l32).max()),
                          "argmax_agreement_vs_fp64": float(1 - bad.mean()),
                          "max_fp64_margin_at_disagreement": float(margin[bad].max()) if bad.any() else 0.0,
                          "median_fp64_margin": float(np.median(margin)),
                          "fp32_oracle_agreement_vs_fp64": float((l32.argmax(-1) == l64.argmax(-1)).mean())}))
