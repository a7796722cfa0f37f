"""Diagnostic: e2e host pipeline time vs sub-batch size (PCSEG_HOST_CHUNK) and output set."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from page_segmentation_b200 import synth
from page_segmentation_b200.runtime import PageBatchEngine
LUT = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], dtype=np.uint8)
n = 64
eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, 0), 3, lut=LUT)
base = np.stack([synth.make_page(s) for s in range(4)])
h_pages = torch.empty((n, synth.A4_H, synth.A4_W), dtype=torch.uint8).pin_memory()
for i in range(n):
    h_pages[i] = torch.from_numpy(base[i % 4])
Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 1 / 3)
outs = {k: torch.empty((n, Hs, Ws) + ((3,) if k != "labels" else ()), dtype=torch.uint8).pin_memory().numpy()
        for k in ("labels", "color", "overlay", "inverted")}
hp = h_pages.numpy()
d_pages = h_pages.cuda()
for _ in range(2):
    eng.run_device(d_pages, 1 / 3)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(3):
    eng.run_device(d_pages, 1 / 3)
torch.cuda.synchronize()
print("device-resident ms/step", (time.perf_counter() - t) / 3 * 1e3)
for chunk in (64, 32, 16, 8, 4, 2):
    os.environ["PCSEG_HOST_CHUNK"] = str(chunk)
    for name, o in (("all", outs), ("labels-only", {"labels": outs["labels"]})):
        eng.run_host(hp, 1 / 3, o)
        t = time.perf_counter()
        for _ in range(3):
            eng.run_host(hp, 1 / 3, o)
        dt = (time.perf_counter() - t) / 3
        print(f"chunk {chunk:3d} outputs {name:12s} {dt * 1e3:7.2f} ms/step  {n / dt:7.0f} pages/s")
