#!/usr/bin/env python
"""Explicit chunk schedules (PCSEG_HOST_CHUNKS, read per call) for the compact host-buffer call on one GPU: wall time of
64 A4 pages, median of several calls.  Development tool.

    python tools/sweep_chunks.py ["4,8" "4,12" ...]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from page_segmentation_b200 import synth  # noqa: E402


def main():
    import torch
    from page_segmentation_b200.runtime import PageBatchEngine
    mode = "compact"
    if len(sys.argv) > 1 and sys.argv[1] in ("compact", "segments"):
        mode = sys.argv.pop(1)
    scheds = sys.argv[1:] or ["", "4,8", "4,12", "8,12", "4,12,12,12,12,8,4", "4,8,12,12,12,8,8", "4,16", "6,10", "4,10", "8,8,12,12,12,8,4"]
    n = 64
    lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], np.uint8)
    eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, seed=0), 3, lut=lut)
    base = np.stack([synth.make_page(s) for s in range(8)])
    h_pages = torch.empty((n, synth.A4_H, synth.A4_W), dtype=torch.uint8).pin_memory().numpy()
    for i in range(n):
        h_pages[i] = base[i % 8]
    Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 1 / 3)
    out = {"labels": torch.empty((n, Hs, Ws), dtype=torch.uint8).pin_memory().numpy(),
           "binary_bits": torch.empty((n, (Hs * Ws + 31) // 32), dtype=torch.int32).pin_memory().numpy().view(np.uint32),
           "stats": torch.empty((n, 3, 4096, 5), dtype=torch.int32).pin_memory().numpy(),
           "ncomp": torch.empty((n, 3), dtype=torch.int32).pin_memory().numpy()}
    for rnd in range(2):                                    # two rounds: the order of the schedules must not matter
        for s in scheds:
            if s:
                os.environ["PCSEG_HOST_CHUNKS"] = s
            else:
                os.environ.pop("PCSEG_HOST_CHUNKS", None)
            ts = []
            for k in range(8):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if mode == "segments":
                    eng.run_host_segments_compact(h_pages, 1 / 3, out, max_components=4096, cc_majority=True)
                else:
                    eng.run_host_compact(h_pages, 1 / 3, out, cc_majority=False)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            ts = sorted(ts[2:])
            print(json.dumps({"chunks": s or "default", "ms_median": round(ts[len(ts) // 2], 3), "ms_min": round(ts[0], 3),
                              "pages_per_s": round(n / ts[len(ts) // 2] * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
