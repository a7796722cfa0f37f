#!/usr/bin/env python
"""Throughput of the reference-named Python API on one GPU (what a caller of the drop-in sees):
DatasetLoader.load_data -> Predictor.predict -> output_data on synthetic A4 pages.  Wall clock, host included;
every line ends with all device work finished and (where files are written) all files on disk."""
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
    from page_segmentation_b200.lib.output import flush_outputs, output_data
    from page_segmentation_b200.lib.postprocess import find_postprocessor
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    n = int(os.environ.get("PCSEG_API_PAGES", "64"))
    pages = [np.array(synth.make_page(s)) for s in range(8)]
    pages = [pages[i % 8].copy() for i in range(n)]                         # n separate pageable arrays, as a caller holds them

    def entries():
        return [SingleData(image=pages[i], line_height_px=18, output_path=f"page_{i:04d}.png") for i in range(n)]

    loader = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True)
    net = Network("Predict", n_classes=3, weights=synth.make_weights("fcn_skip", 3, seed=0))

    def clock(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t

    def line(stage, t, **kw):
        print(json.dumps({"stage": stage, "pages": n, "s": round(t, 4), "pages_per_s": round(n / t, 1), **kw}), flush=True)

    def staged(ds):
        # wait for the background stager: every page's image is on the device
        from page_segmentation_b200.lazy import peek
        for d in ds.data:
            peek(d, "image").device_tensor()
        return ds

    staged(loader.load_data(entries()[:8]))                                 # warm-up
    ds, t = clock(lambda: staged(loader.load_data(entries())))
    line("DatasetLoader.load_data (in-memory pages) until every page is staged on the device", t)
    ds, t = clock(lambda: (lambda d: [np.asarray(x.image) for x in d.data] and d)(loader.load_data(entries())))
    line("DatasetLoader.load_data, every image read on the host afterwards", t)
    for label, post in (("predict", []), ("predict + cc_majority", [find_postprocessor("cc_majority")])):
        pred = Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, post_process=post), network=net)
        ds = staged(loader.load_data(entries()))
        list(pred.predict(type(ds)(ds.data[:8], ds.color_map)))

        def on_device():
            for p in pred.predict(ds):
                pass
        on_device()
        _, t = clock(on_device)
        line(f"Predictor.{label}, results left on the device", t)

        def consume():
            # a front end that looks at every class map on the host, then drops the Prediction
            k = 0
            for p in pred.predict(ds):
                k += int(p.labels[0, 0]) + 1
            return k
        consume()
        _, t = clock(consume)
        line(f"Predictor.{label}, each class map read on the host and dropped", t)
        _, t = clock(lambda: [np.asarray(p.labels) for p in pred.predict(ds)])
        line(f"Predictor.{label}, all class maps read and kept", t)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as out:
        pred = Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, output=out,
                                         post_process=[find_postprocessor("cc_majority")]), network=net)
        preds = list(pred.predict(ds))

        def write_all():
            for p in preds:
                output_data(out, p.labels, p.data, DEFAULT_COLOR_MAP)
            flush_outputs()
        write_all()
        from page_segmentation_b200.runtime import get_context
        c = get_context()
        c.set_timing(True)
        output_data(out, preds[0].labels, preds[0].data, DEFAULT_COLOR_MAP)
        print(json.dumps({"output_pages_device_ms_per_page": dict(c.timings())}), flush=True)
        c.set_timing(False)
        _, t = clock(write_all)
        size = sum(os.path.getsize(os.path.join(out, c, f)) for c in ("color", "overlay", "inverted") for f in os.listdir(os.path.join(out, c)))
        line("output_data (three PNG files per page, device encoder) + flush_outputs", t, mb_per_page=round(size / n / 1e6, 2))
        del preds, ds

        def flow():
            ds = loader.load_data(entries())
            for p in pred.predict(ds):
                output_data(out, p.labels, p.data, DEFAULT_COLOR_MAP)
            flush_outputs()
        flow()
        for _ in range(2):
            _, t = clock(flow)
            line("load_data -> predict (+ cc_majority) -> output_data -> flush_outputs (the drop-in flow)", t)

        def page_by_page():
            for e in entries():
                p = pred.predict_single(loader.load_images(e))
                output_data(out, p.labels, p.data, DEFAULT_COLOR_MAP)
            flush_outputs()
        page_by_page()
        _, t = clock(page_by_page)
        line("load_images -> predict_single (+ cc_majority) -> output_data, page by page, + flush_outputs", t)
    from page_segmentation_b200.lazy import pinned_pool
    print(json.dumps({"pinned_pool": pinned_pool().stats}), flush=True)


if __name__ == "__main__":
    main()
