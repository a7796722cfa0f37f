"""SURVEY section 8(d)'s stress variant timed on the device: `line_height_px = 6` (scale 1), the network on the 3520x2496
grid of the page itself -- 8.92x the FLOPs of a normalised page (999 GFLOP per page).  Per-stage library timing events.
Development tool: python tools/stress_variant.py [pages per launch, default 8]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
from page_segmentation_b200 import synth
import torch
from page_segmentation_b200.runtime import PageBatchEngine
lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], np.uint8)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, seed=0), 3, lut=lut)
d = torch.from_numpy(np.stack([synth.make_page(s) for s in range(n)])).cuda()
for _ in range(3):
    eng.run_device(d, 1.0, masks=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    eng.run_device(d, 1.0, masks=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
eng.ctx.set_timing(True)
acc = {}
for _ in range(reps):
    eng.run_device(d, 1.0, masks=True)
    for k, v in eng.ctx.timings():
        acc[k] = acc.get(k, 0.0) + v / reps
eng.ctx.set_timing(False)
gflop_page = 111.999 * (3520 * 2496) / (1184 * 832)
print(json.dumps({"workload": f"stress variant: {n} synthetic 2480x3508 pages at scale 1 (3520x2496 grid), preprocess + fcn_skip + argmax + colour masks, one launch sequence",
                  "ms_per_step": round(ms, 3), "pages_per_s": round(n / ms * 1e3, 1), "mpixel_per_s": round(n * 8.69984 / ms * 1e3, 1),
                  "normalised_page_equivalents_per_s": round(n / ms * 1e3 * 3520 * 2496 / (1184 * 832), 1),
                  "body_tflops": round(n * gflop_page / sum(v for k, v in acc.items() if k not in ("preprocess", "masks")) , 1),
                  "stage_ms": {k: round(v, 4) for k, v in acc.items()}}), flush=True)
