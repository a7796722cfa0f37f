import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from page_segmentation_b200 import synth
from page_segmentation_b200.runtime import PageBatchEngine
LUT = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], dtype=np.uint8)
n = 64
eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, 0), 3, lut=LUT)
h_pages = torch.empty((n, synth.A4_H, synth.A4_W), dtype=torch.uint8).pin_memory()
p = torch.from_numpy(synth.make_page(0))
for i in range(n): h_pages[i] = p
Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 1 / 3)
keep = {k: torch.empty((n, Hs, Ws) + ((3,) if k != "labels" else ()), dtype=torch.uint8).pin_memory() for k in ("labels", "color", "overlay", "inverted")}
outs = {k: v.numpy() for k, v in keep.items()}
print("pinned:", h_pages.is_pinned(), [v.is_pinned() for v in keep.values()])
eng.run_host(h_pages.numpy(), 1 / 3, outs)
os.environ["PCSEG_TRACE_HOST"] = "1"
eng.run_host(h_pages.numpy(), 1 / 3, outs)
