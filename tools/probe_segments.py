#!/usr/bin/env python
"""Device time of pcs_class_components (segment extraction, BASELINE configs[3]) on the class maps the bench produces
(random-init weights: noise-like) and on blocky maps (what a trained model yields); run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from page_segmentation_b200 import synth  # noqa: E402


def main():
    import torch
    from page_segmentation_b200.runtime import PageBatchEngine
    n = int(os.environ.get("PROBE_PAGES", "64"))
    lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], np.uint8)
    eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, seed=0), 3, lut=lut)
    pages = np.stack([synth.make_page(s) for s in range(8)])
    d_pages = torch.from_numpy(np.concatenate([pages] * (n // 8))).cuda()
    Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 1 / 3)
    b = eng.run_device(d_pages, 1 / 3, cc_majority=True, masks=False)
    noisy = b["labels"].clone()
    rng = np.random.default_rng(0)
    coarse = rng.integers(0, 3, (n, Hs // 24 + 1, Ws // 24 + 1)).astype(np.uint8)
    blocky = torch.from_numpy(np.ascontiguousarray(np.kron(coarse, np.ones((1, 24, 24), np.uint8))[:, :Hs, :Ws])).cuda()
    maxc = 4096
    d_stats = torch.empty((n, 3, maxc, 5), dtype=torch.int32, device="cuda")
    d_ncomp = torch.empty((n, 3), dtype=torch.int32, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, pred in (("noise-like (random-init network + cc_majority)", noisy), ("blocky 24x24", blocky)):
        reps = 1 if os.environ.get("PROBE_QUICK") else 5
        eng.ctx.class_components(pred, n, Hs, Ws, 3, d_stats, maxc, d_ncomp)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            eng.ctx.class_components(pred, n, Hs, Ws, 3, d_stats, maxc, d_ncomp)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"class_map": name, "pages": n, "ms": e0.elapsed_time(e1) / reps,
                          "components_per_page_per_class": d_ncomp.float().mean(0).tolist()}), flush=True)


if __name__ == "__main__":
    main()
