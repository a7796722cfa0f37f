import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
from page_segmentation_b200 import synth
import torch
from page_segmentation_b200.runtime import PageBatchEngine
lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], np.uint8)
eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, seed=0), 3, lut=lut)
base = np.stack([synth.make_page(s) for s in range(8)])
for n in (4, 8, 16, 32, 64):
    d = torch.from_numpy(np.concatenate([base] * max(1, n // 8))[:n]).cuda()
    for _ in range(3):
        eng.run_device(d, 1 / 3, masks=False)
    torch.cuda.synchronize()
    eng.ctx.set_timing(True)
    acc = {}
    reps = 5
    for _ in range(reps):
        eng.run_device(d, 1 / 3, masks=False)
        for k, v in eng.ctx.timings():
            acc[k] = acc.get(k, 0.0) + v / reps
    eng.ctx.set_timing(False)
    tot = sum(acc.values())
    print(json.dumps({"pages": n, "total_ms": round(tot, 4), "us_per_page": round(tot / n * 1e3, 1),
                      "stage_us_per_page": {k: round(v / n * 1e3, 1) for k, v in acc.items()}}), flush=True)
