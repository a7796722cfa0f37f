#!/usr/bin/env python
"""Device time of the stages around the network (not on bench.py's step): cc_majority, bounding boxes,
compute_char_height, nearest resize to the original shape, region extraction.  CUDA events on the launch stream,
inputs larger than L2 (batches of A4 pages), algorithmic bytes per page as in SURVEY.md section 8(d).

    python tools/bench_stages.py > profiles/rNN_stage_bench.jsonl
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from page_segmentation_b200 import synth  # noqa: E402


def main():
    import torch
    from page_segmentation_b200 import runtime
    ctx = runtime.get_context(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6550.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    quick = os.environ.get("PCSEG_STAGE_QUICK") is not None      # one warm-up, one repetition: for ncu launch lists

    def timed(name, fn, pages, bytes_per_page, note, reps=5):
        if quick:
            reps = 1
        for _ in range(1 if quick else 3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = bytes_per_page * pages / ms / 1e6
        print(json.dumps({"stage": name, "pages": pages, "ms": round(ms, 4), "pages_per_s": round(pages / ms * 1e3, 1),
                          "algorithmic_MB_per_page": round(bytes_per_page / 1e6, 2), "achieved_GBs": round(gbs, 1),
                          "frac_of_measured_hbm": round(gbs / peak, 4), "note": note}), flush=True)

    n = 32
    H, W = synth.A4_H, synth.A4_W
    Hs, Ws = synth.scaled_shape(H, W, 1 / 3)
    pages = np.stack([synth.make_page(s) for s in range(8)])
    d_pages = torch.from_numpy(np.concatenate([pages] * (n // 8))).cuda()
    # scaled binary + a blocky class map, as the network stage leaves them
    d_image = torch.empty((n, Hs, Ws), dtype=torch.uint8, device="cuda")
    d_binary = torch.empty((n, Hs, Ws), dtype=torch.uint8, device="cuda")
    ctx.preprocess(d_pages, d_pages, n, H, W, Hs, Ws, d_image, d_binary, None)
    rng = np.random.default_rng(0)
    coarse = rng.integers(0, 3, (n, Hs // 24 + 1, Ws // 24 + 1)).astype(np.uint8)
    pred = np.kron(coarse, np.ones((1, 24, 24), np.uint8))[:, :Hs, :Ws]
    noise = rng.random(pred.shape) < 0.1
    pred[noise] = rng.integers(0, 3, int(noise.sum()))
    d_pred0 = torch.from_numpy(np.ascontiguousarray(pred)).cuda()
    d_pred = d_pred0.clone()
    px = Hs * Ws

    def vote():
        d_pred.copy_(d_pred0)
        ctx.cc_majority(d_pred, d_binary, n, Hs, Ws, 3)
    timed("cc_majority", vote, n, 7 * px, "binary + pred in, labels i32 + pred out (single ideal pass); includes the reset copy of pred")

    d_boxes = torch.empty_like(d_pred0)
    timed("bounding_boxes", lambda: ctx.bounding_boxes(d_pred0, n, Hs, Ws, 3, d_boxes), n, 2 * px, "pred in, pred out")

    d_h = torch.empty((n,), dtype=torch.int32, device="cuda")
    timed("compute_char_height", lambda: ctx.char_height(d_pages, n, H, W, False, d_h), n, H * W,
          "full-resolution grey page read once (Otsu histogram + 8-connected components + box filter + median)")

    d_up = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
    timed("resize_nearest_to_original", lambda: ctx.resize_nearest(d_pred0, n, Hs, Ws, d_up, H, W), n, px + H * W,
          "scale_to_original_shape: class map in, full-resolution class map out")

    lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], np.uint8)
    d_c = [torch.empty((n, Hs, Ws, 3), dtype=torch.uint8, device="cuda") for _ in range(3)]
    timed("generate_output_masks", lambda: ctx.masks(d_pred0, d_binary, n, Hs, Ws, lut, *d_c), n, 2 * px + 9 * px,
          "labels + binary in, three colour images out")

    size = ctx.png_bytes(Hs, Ws, 3)
    stride = (size + 255) // 256 * 256
    d_png = torch.empty((3 * n, stride), dtype=torch.uint8, device="cuda")
    d_all = torch.cat(d_c)                                    # (3n, Hs, Ws, 3): the three masks of every page
    timed("png_encode (three masks per page)", lambda: ctx.png_encode(d_all, 3 * n, Hs, Ws, 3, d_png, stride), n, 2 * 9 * px,
          "three colour masks in, three complete PNG files out (level 1: Sub/Up filter choice, fixed-Huffman run-length deflate, Adler-32, CRC-32)")

    # region extraction works page by page on the full-resolution `inverted` image (26 MB each)
    inv = [torch.from_numpy(synth.make_inverted_image(s, H, W, 40)).cuda() for s in range(4)]
    d_text = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    d_region = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    col = np.array([255, 0, 0], np.uint8)

    def regions():
        for im in inv:
            ctx.text_regions(im, H, W, col, 40, 13, 36, d_text, d_region)
    timed("text_regions (get_text_contours pixel work)", regions, len(inv), 3 * H * W + 2 * H * W,
          "RGB page in, canvas + region_text out; 16 launches per page (inRange, 7 x 2 morphology passes, unpack)")

    Ho, Wo = 300, int(W * (300 / H))
    d_masks = torch.empty((2, Ho, Wo), dtype=torch.uint8, device="cuda")
    d_sat = torch.empty((2, Ho + 1, Wo + 1), dtype=torch.int32, device="cuda")
    cols = np.array([[0, 255, 0], [255, 0, 0]], np.uint8)

    def segs():
        for im in inv:
            ctx.segment_masks(im, H, W, Ho, Wo, cols, d_masks)
            ctx.integral_image(d_masks, 2, Ho, Wo, d_sat)
    timed("find_segments pixel work (300-row working size)", segs, len(inv), 9 * Ho * Wo * 3 + 2 * Ho * Wo * 5,
          "gathers 9 source pixels per output pixel; launch-latency bound (3 launches of a 64k-pixel grid)")


if __name__ == "__main__":
    main()
