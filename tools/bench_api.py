#!/usr/bin/env python
"""Throughput of the reference-named Python API on one GPU (what a caller of the drop-in sees):
DatasetLoader.load_data -> Predictor.predict -> output_data on synthetic A4 pages.  Wall clock, host included."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from page_segmentation_b200 import synth  # noqa: E402


def main():
    import torch
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import DatasetLoader, SingleData
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.output import output_data
    from page_segmentation_b200.lib.postprocess import find_postprocessor
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    n = int(os.environ.get("PCSEG_API_PAGES", "32"))
    pages = [synth.make_page(s) for s in range(8)]
    entries = [SingleData(image=pages[i % 8], line_height_px=18, image_path=f"/in/page_{i:04d}.png") for i in range(n)]
    loader = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True)
    net = Network("Predict", n_classes=3, weights=synth.make_weights("fcn_skip", 3, seed=0))

    def clock(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t

    loader.load_data(entries[:2])                                           # warm-up
    ds, t_load = clock(lambda: loader.load_data(entries))
    for label, post in (("predict", []), ("predict + cc_majority", [find_postprocessor("cc_majority")])):
        pred = Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, post_process=post), network=net)
        list(pred.predict(type(ds)(ds.data[:2], ds.color_map)))
        def consume():
            # what the reference's front ends do: use a Prediction, drop it, take the next one
            k = 0
            for p in pred.predict(ds):
                k += int(p.labels[0, 0]) + 1
            return k
        preds = None                                    # kept results of the previous round would hold page-locked blocks
        consume()
        _, t = clock(consume)
        print(json.dumps({"stage": f"Predictor.{label}, each result consumed and dropped", "pages": n, "s": round(t, 4),
                          "pages_per_s": round(n / t, 1)}), flush=True)
        preds, t = clock(lambda: list(pred.predict(ds)))
        print(json.dumps({"stage": f"Predictor.{label}, all results kept", "pages": n, "s": round(t, 4),
                          "pages_per_s": round(n / t, 1)}), flush=True)
    print(json.dumps({"stage": "DatasetLoader.load_data (in-memory pages)", "pages": n, "s": round(t_load, 4),
                      "pages_per_s": round(n / t_load, 1)}), flush=True)
    with tempfile.TemporaryDirectory() as out:
        for sub in ("color", "overlay", "inverted"):
            os.makedirs(os.path.join(out, sub))
        output_data(out, preds[0].labels, preds[0].data, DEFAULT_COLOR_MAP)
        _, t = clock(lambda: [output_data(out, p.labels, p.data, DEFAULT_COLOR_MAP) for p in preds])
        print(json.dumps({"stage": "output_data (three PNG files per page, device encoder)", "pages": n, "s": round(t, 4),
                          "pages_per_s": round(n / t, 1)}), flush=True)
        del preds, ds

        def flow():
            # one page at a time through the whole per-page API, nothing kept
            for i in range(n):
                e = SingleData(image=pages[i % 8], line_height_px=18, image_path=f"/in/page_{i:04d}.png")
                p = pred.predict_single(loader.load_images(e))
                output_data(out, p.labels, p.data, DEFAULT_COLOR_MAP)
        flow()
        _, t = clock(flow)
        print(json.dumps({"stage": "load_images -> predict_single (+ cc_majority) -> output_data, page by page, nothing kept",
                          "pages": n, "s": round(t, 4), "pages_per_s": round(n / t, 1)}), flush=True)


if __name__ == "__main__":
    main()
