#!/usr/bin/env python
"""The drop-in flow of bench.py's e2e_api, several passes in a row with the time of each phase: a check that the flow's
throughput does not depend on what earlier passes left behind."""
import gc
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from page_segmentation_b200 import synth  # noqa: E402


def main():
    import torch
    from page_segmentation_b200 import lazy
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import DatasetLoader, SingleData
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.output import flush_outputs, output_data
    from page_segmentation_b200.lib.postprocess import find_postprocessor
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    n = 64
    base = [np.array(synth.make_page(s)) for s in range(8)]
    pages = [base[i % 8].copy() for i in range(n)]
    root = tempfile.mkdtemp(prefix="pcseg_api_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    net = Network("Predict", n_classes=3, weights=synth.make_weights("fcn_skip", 3, seed=0), precision="fp16")
    settings = PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, output=root, post_process=[find_postprocessor("cc_majority")])
    predictor = Predictor(settings, network=net)
    loader = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True)
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        entries = [SingleData(image=p, line_height_px=18, output_path=f"page{i:04d}.png") for i, p in enumerate(pages)]
        dataset = loader.load_data(entries)
        t1 = time.perf_counter()
        for pred in predictor.predict(dataset):
            output_data(root, pred.labels, pred.data, DEFAULT_COLOR_MAP)
        t2 = time.perf_counter()
        flush_outputs()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"pass {it}: {n / (t3 - t0):8.1f} pages/s  load_data {1e3 * (t1 - t0):6.1f} ms  predict+output {1e3 * (t2 - t1):6.1f} ms  flush {1e3 * (t3 - t2):6.1f} ms  "
              f"pool {lazy.pinned_pool().stats}  cuda {torch.cuda.memory_allocated() >> 20} MB reserved {torch.cuda.memory_reserved() >> 20} MB  gc {gc.get_count()}", flush=True)
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
