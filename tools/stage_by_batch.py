"""Per-stage device time of the bench step against the number of pages per launch (library timing events); `cc` adds
cc_majority and segment extraction.  Development tool: python tools/stage_by_batch.py [cc]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
from page_segmentation_b200 import synth
import torch
from page_segmentation_b200.runtime import PageBatchEngine
lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], np.uint8)
ARCH = "unet" if "unet" in sys.argv[1:] else "fcn_skip"
eng = PageBatchEngine(ARCH, synth.make_weights(ARCH, 3, seed=0), 3, lut=lut)
base = np.stack([synth.make_page(s) for s in range(8)])
Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 1 / 3)
CC = "cc" in sys.argv[1:]          # with cc_majority + segment extraction (BASELINE configs[3])
for n in ((8,) if ARCH == "unet" else (4, 8, 16, 32, 64)):
    d = torch.from_numpy(np.concatenate([base] * max(1, n // 8))[:n]).cuda()
    d_stats = torch.empty((n, 3, 4096, 5), dtype=torch.int32, device="cuda")
    d_ncomp = torch.empty((n, 3), dtype=torch.int32, device="cuda")

    def step():
        b = eng.run_device(d, 1 / 3, masks=False, cc_majority=CC)
        if CC:
            eng.ctx.class_components(b["labels"], n, Hs, Ws, 3, d_stats, 4096, d_ncomp)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    eng.ctx.set_timing(True)
    acc = {}
    reps = 5
    for _ in range(reps):
        step()
        for k, v in eng.ctx.timings():
            acc[k] = acc.get(k, 0.0) + v / reps
    eng.ctx.set_timing(False)
    tot = sum(acc.values())
    print(json.dumps({"pages": n, "total_ms": round(tot, 4), "us_per_page": round(tot / n * 1e3, 1),
                      "stage_us_per_page": {k: round(v / n * 1e3, 1) for k, v in acc.items()}}), flush=True)
