#!/usr/bin/env python
"""Interleaved A/B of programmatic dependent launch in one process (same clocks, same buffers):
device-resident batches of 8 and 64 pages and the host-buffer pipeline.  Development tool."""
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
    lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], dtype=np.uint8)
    eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, seed=0), 3, lut=lut)
    n = 64
    base = np.stack([synth.make_page(s) for s in range(8)])
    h_pages = torch.empty((n, synth.A4_H, synth.A4_W), dtype=torch.uint8).pin_memory()
    for i in range(n):
        h_pages[i] = torch.from_numpy(base[i % 8])
    Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 1 / 3)
    h_out = {k: torch.empty((n, Hs, Ws) + ((3,) if k != "labels" else ()), dtype=torch.uint8).pin_memory().numpy()
             for k in ("labels", "color", "overlay", "inverted")}
    hp = h_pages.numpy()
    d_pages = h_pages.cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def t(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    res = {}
    for rnd in range(4):
        for pdl in (0, 1):
            eng.ctx.set_pdl(bool(pdl))
            res.setdefault(("dev64", pdl), []).append(t(lambda: eng.run_device(d_pages, 1 / 3), 5))
            res.setdefault(("dev8", pdl), []).append(t(lambda: eng.run_device(d_pages[:8], 1 / 3), 20))
            res.setdefault(("host64", pdl), []).append(t(lambda: eng.run_host(hp, 1 / 3, h_out), 4))
    for k in sorted(res):
        print(json.dumps({"case": k[0], "pdl": k[1], "ms": [round(v, 3) for v in res[k]], "median": round(float(np.median(res[k])), 3)}))


if __name__ == "__main__":
    main()
