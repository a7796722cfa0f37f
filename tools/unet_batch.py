"""U-Net device time against the number of pages per launch (8 / 16 / 32): no gain beyond 8 (557 / 566 / 547 pages/s on one
B200), which is why bench.py runs the 256 pages of BASELINE configs[2] in sub-batches of 8.  Development tool."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
from page_segmentation_b200 import synth
import torch
from page_segmentation_b200.runtime import PageBatchEngine
lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], np.uint8)
eng = PageBatchEngine("unet", synth.make_weights("unet", 3, seed=0), 3, lut=lut)
base = np.stack([synth.make_page(s) for s in range(8)])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (8, 16, 32):
    d = torch.from_numpy(np.concatenate([base] * (n // 8))).cuda()
    eng.run_device(d, 1 / 3)
    torch.cuda.synchronize()
    reps = 32 // n * 2
    e0.record()
    for _ in range(reps):
        eng.run_device(d, 1 / 3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"pages_per_launch": n, "ms": round(ms, 3), "pages_per_s": round(n / ms * 1e3, 1),
                      "mem_GB": round(torch.cuda.mem_get_info()[0] / 1e9, 1)}), flush=True)
    del d
