"""Diagnostic: PCIe copy bandwidth alone, both directions at once, and concurrent with the compute step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from page_segmentation_b200 import synth
from page_segmentation_b200.runtime import PageBatchEngine
LUT = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], dtype=np.uint8)
eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, 0), 3, lut=LUT)
n = 32
d_pages = torch.from_numpy(np.stack([synth.make_page(s % 2) for s in range(n)])).cuda()
N = 512 << 20
h_a = torch.empty(N, dtype=torch.uint8).pin_memory(); h_b = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device="cuda"); d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps

def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
def comp():
    eng.run_device(d_pages, 1 / 3)
t_h2d, t_d2h, t_comp = timed(h2d), timed(d2h), timed(comp)
print(f"h2d alone {N/t_h2d/1e9:.1f} GB/s  d2h alone {N/t_d2h/1e9:.1f} GB/s  compute {t_comp*1e3:.2f} ms")
t = timed(lambda: (h2d(), d2h()))
print(f"h2d+d2h concurrently: {t*1e3:.2f} ms (sum alone {1e3*(t_h2d+t_d2h):.2f}, max alone {1e3*max(t_h2d,t_d2h):.2f})")
t = timed(lambda: (d2h(), comp()))
print(f"d2h + compute: {t*1e3:.2f} ms (d2h {t_d2h*1e3:.2f}, compute {t_comp*1e3:.2f})")
t = timed(lambda: (h2d(), comp()))
print(f"h2d + compute: {t*1e3:.2f} ms (h2d {t_h2d*1e3:.2f}, compute {t_comp*1e3:.2f})")
t = timed(lambda: (h2d(), d2h(), comp()))
print(f"h2d + d2h + compute: {t*1e3:.2f} ms")
