"""Generates tests/golden/ref_*.npz by running the REFERENCE's own functions (see reference_import.py) on
seeded synthetic inputs:   python tests/golden/make_reference_golden.py

  ref_postprocess.npz  postprocess.vote_connected_component_class, the cv2 error of add_bounding_boxes,
                       find_postprocessor's key normalisation, output.generate_output_masks,
                       image_ops.compute_char_height (through a PNG file, as the reference reads it)
  ref_regions.npz      xycut.do_xy_cut, pc_segmentation.find_segments / get_text_contours

Inputs are stored next to the outputs, so the tests never need the reference tree.
"""
import os
import sys
import tempfile

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import reference_import as ref  # noqa: E402
from page_segmentation_b200 import synth  # noqa: E402
from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP  # noqa: E402


def rects(segs):
    return np.array([[s.x_start, s.y_start, s.x_end, s.y_end] for s in segs], dtype=np.int64).reshape(-1, 4)


def pack_contours(contours):
    pts = [np.asarray(c.contour, dtype=np.int32).reshape(-1, 2) for c in contours]
    offs = np.cumsum([0] + [len(p) for p in pts]).astype(np.int64)
    return (np.concatenate(pts) if pts else np.zeros((0, 2), np.int32)), offs


def class_map(rng, h, w, n_classes):
    """Blocky random class map with speckle, like an argmax of a noisy network."""
    coarse = rng.integers(0, n_classes, (h // 16 + 1, w // 16 + 1))
    pred = np.kron(coarse, np.ones((16, 16), dtype=np.int64))[:h, :w]
    flip = rng.random((h, w)) < 0.15
    pred[flip] = rng.integers(0, n_classes, int(flip.sum()))
    return pred.astype(np.int64)


def main():
    pp = ref.load("postprocess")
    out = ref.load("output")
    ds = ref.load("dataset")
    iops = ref.load("image_ops")
    xy = ref.load("xycut")
    pcs = ref.load("pc_segmentation")
    rng = np.random.default_rng(20260)

    # ---------------- postprocess / output / image_ops ----------------
    store = {}
    cases = [(0, 96, 128, 18, 3), (1, 257, 191, 12, 3), (2, 400, 300, 18, 5)]
    store["n_vote"] = np.array(len(cases))
    for i, (seed, h, w, lh, ncls) in enumerate(cases):
        page = synth.make_page(seed, h, w, lh)
        binary = (page == 0).astype(np.uint8)                         # 1 = ink, as dataset.py:146 leaves it
        pred = class_map(rng, h, w, ncls)
        data = ds.SingleData(binary=binary)
        voted = pp.vote_connected_component_class(pred.copy(), data)
        store[f"vote{i}_binary"], store[f"vote{i}_pred"], store[f"vote{i}_voted"] = binary, pred.astype(np.uint8), voted.astype(np.uint8)
        masks = out.generate_output_masks(data, voted, DEFAULT_COLOR_MAP) if ncls == 3 else None
        if masks is not None:
            store[f"vote{i}_color"], store[f"vote{i}_overlay"] = masks.color, masks.overlay
            store[f"vote{i}_inverted"], store[f"vote{i}_fg"] = masks.inverted_overlay, masks.fg_color_mask
    try:
        pp.add_bounding_boxes(class_map(rng, 40, 40, 3), ds.SingleData(binary=np.zeros((40, 40), np.uint8)))
        store["bbox_error"] = np.array("")
    except Exception as e:                                              # cv2 rejects the bool array (postprocess.py:33)
        store["bbox_error"] = np.array(type(e).__name__)
    keys = ["cc_majority", "CC-Vote", "vote_connected_components", "votecomponents", "Bounding_Boxes", "bbox"]
    store["pp_keys"] = np.array(keys)
    store["pp_names"] = np.array([pp.find_postprocessor(k).__name__ for k in keys])

    ch_cases = [(3, 900, 700, 24, False), (4, 900, 700, 30, True), (5, 640, 480, 18, False), (6, 200, 200, 4, False)]
    store["n_char"] = np.array(len(ch_cases))
    with tempfile.TemporaryDirectory() as tmp:
        for i, (seed, h, w, lh, inverse) in enumerate(ch_cases):
            page = synth.make_grey_page(seed, h, w, lh) if i % 2 == 0 else synth.make_page(seed, h, w, lh)
            if inverse:
                page = 255 - page
            path = os.path.join(tmp, f"p{i}.png")
            cv2.imwrite(path, page)
            got = iops.compute_char_height(path, inverse)
            store[f"char{i}_page"], store[f"char{i}_inverse"] = page, np.array(inverse)
            store[f"char{i}_height"] = np.array(-1 if got is None else int(got))
    # evaluation metrics (image_ops.fgpa / fgoverlap_per_class) on the vote cases: the voted map against the raw one
    for i, (seed, h, w, lh, ncls) in enumerate(cases):
        pred, voted, binary = store[f"vote{i}_pred"].astype(np.int64), store[f"vote{i}_voted"].astype(np.int64), store[f"vote{i}_binary"].astype(np.int64)
        store[f"eval{i}_fgpa"] = np.array(iops.fgpa(voted, pred, binary))
        ov, tp, fp, fn = iops.fgoverlap_per_class(voted, pred, binary, ncls)
        store[f"eval{i}_overlap"], store[f"eval{i}_tp"], store[f"eval{i}_fp"], store[f"eval{i}_fn"] = np.array(ov), np.array(tp), np.array(fp), np.array(fn)
        store[f"eval{i}_ncls"] = np.array(ncls)
    np.savez_compressed(os.path.join(HERE, "ref_postprocess.npz"), **store)

    # ---------------- region extraction ----------------
    store = {}
    xy_cases = []
    for seed, h, w, ch in [(0, 300, 214, 8), (1, 300, 214, 8), (2, 240, 320, 6), (5, 180, 127, 5)]:
        inv = synth.make_inverted_image(seed, h, w, ch)
        for colour in ((255, 0, 0), (0, 255, 0)):
            mask = np.all(inv == np.array(colour, np.uint8), axis=-1)
            mask = cv2.dilate(mask.astype(np.uint8), np.ones((3, 3), np.uint8)).astype(bool)
            for params in [(ch, ch, 2 * ch, ch), (max(1, ch // 2), ch, ch, 2 * ch), (2, 2, 3, 3)]:
                xy_cases.append((mask, params))
    # degenerate masks: empty, full, one row of ink
    xy_cases.append((np.zeros((40, 30), bool), (3, 3, 5, 5)))
    xy_cases.append((np.ones((40, 30), bool), (3, 3, 5, 5)))
    line = np.zeros((50, 60), bool)
    line[20:24, 5:50] = True
    xy_cases.append((line, (2, 2, 4, 4)))
    store["n_xy"] = np.array(len(xy_cases))
    for i, (mask, params) in enumerate(xy_cases):
        store[f"xy{i}_mask"] = np.packbits(mask, axis=1)
        store[f"xy{i}_shape"] = np.array(mask.shape)
        store[f"xy{i}_params"] = np.array(params)
        store[f"xy{i}_rects"] = rects(xy.do_xy_cut(mask, *params))

    seg_cases = [(0, 700, 500, 18, 300), (1, 700, 500, 14, 250), (2, 1000, 707, 24, 300), (3, 512, 384, 12, 511), (4, 333, 517, 9, 200)]
    store["n_seg"] = np.array(len(seg_cases))
    for i, (seed, h, w, ch, rh) in enumerate(seg_cases):
        inv = synth.make_inverted_image(seed, h, w, ch)
        store[f"seg{i}_image"] = inv
        store[f"seg{i}_args"] = np.array([h, ch, rh])
        t, im = pcs.find_segments(h, inv, ch, rh, DEFAULT_COLOR_MAP)
        store[f"seg{i}_text"], store[f"seg{i}_pictures"] = rects(t), rects(im)
        _, only = pcs.find_segments(h, inv, ch, rh, DEFAULT_COLOR_MAP, only_images=True)
        assert np.array_equal(rects(only), rects(im))
        pts, offs = pack_contours(pcs.get_text_contours(inv, ch, DEFAULT_COLOR_MAP))
        store[f"seg{i}_contour_points"], store[f"seg{i}_contour_offsets"] = pts, offs
    np.savez_compressed(os.path.join(HERE, "ref_regions.npz"), **store)
    for f in ("ref_postprocess.npz", "ref_regions.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
