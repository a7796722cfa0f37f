#!/usr/bin/env python
"""Per-chunk device timeline of one host-buffer call (PCSEG_TRACE_HOST=1 is read per call): compact transport by default,
`segments` for pcs_predict_pages_segments_compact.  Prints the library's trace (stderr) for the LAST of three calls."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from page_segmentation_b200 import synth  # noqa: E402


def main():
    import torch
    from page_segmentation_b200.runtime import PageBatchEngine
    mode = sys.argv[1] if len(sys.argv) > 1 else "compact"
    n = int(os.environ.get("TRACE_PAGES", "64"))
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
    for k in range(3):
        if k == 2:
            os.environ["PCSEG_TRACE_HOST"] = "1"
        if mode == "segments":
            eng.run_host_segments_compact(h_pages, 1 / 3, out, max_components=4096, cc_majority=True)
        else:
            eng.run_host_compact(h_pages, 1 / 3, out, cc_majority=False)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
