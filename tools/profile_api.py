#!/usr/bin/env python
"""cProfile of the per-page API flows (main thread) + wall time of the background stages (diagnostics)."""
import cProfile
import os
import pstats
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from page_segmentation_b200 import synth  # noqa: E402


def main():
    import torch
    from page_segmentation_b200 import pipeline
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
    loader = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True)
    net = Network("Predict", n_classes=3, weights=synth.make_weights("fcn_skip", 3, seed=0))
    which = sys.argv[1] if len(sys.argv) > 1 else "flow"
    with tempfile.TemporaryDirectory(dir="/dev/shm") as out:
        pred = Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, output=out,
                                         post_process=[find_postprocessor("cc_majority")]), network=net)

        def entries():
            return [SingleData(image=pages[i], line_height_px=18, output_path=f"page_{i:04d}.png") for i in range(n)]

        def flow():
            ds = loader.load_data(entries())
            for p in pred.predict(ds):
                output_data(out, p.labels, p.data, DEFAULT_COLOR_MAP)
            flush_outputs()

        def page_by_page():
            for e in entries():
                p = pred.predict_single(loader.load_images(e))
                output_data(out, p.labels, p.data, DEFAULT_COLOR_MAP)
            flush_outputs()

        fn = flow if which == "flow" else page_by_page
        fn()
        fn()
        pipeline.TRACE.clear()
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        fn()
        pr.disable()
        dt = time.perf_counter() - t0
        print(f"{which}: {dt * 1e3:.1f} ms for {n} pages")
        pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
        agg = {}
        for k, v in pipeline.TRACE:
            agg[k] = agg.get(k, 0.0) + v
        print({k: round(v * 1e3, 2) for k, v in agg.items()})


if __name__ == "__main__":
    main()
