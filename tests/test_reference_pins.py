"""Pins against the REFERENCE's own code (CPU tests).

tests/golden/ref_*.npz were produced by executing the reference's functions themselves
(tests/golden/make_reference_golden.py + reference_import.py; the modules below import with numpy + cv2 alone).
Here the oracle restatements and the host-side logic of the product are held to those vectors; where the reference
tree is mounted (the build container) the same comparisons also run live on fresh random cases.
"""
import os
import sys

import cv2
import numpy as np
import pytest

from oracle import image_ops as oio
from oracle import pipeline as opipe
from oracle import regions as oreg
from page_segmentation_b200 import synth
from page_segmentation_b200.lib import xycut as pxy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import reference_import as ref  # noqa: E402

LUT = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}
POST = np.load(os.path.join(HERE, "golden", "ref_postprocess.npz"))
REG = np.load(os.path.join(HERE, "golden", "ref_regions.npz"))
live = pytest.mark.skipif(not ref.available(), reason="the reference tree is only mounted in the build container")


def _rect_array(rects):
    return np.array([[r.x_start, r.y_start, r.x_end, r.y_end] for r in rects], dtype=np.int64).reshape(-1, 4)


def _unpack_mask(i):
    h, w = REG[f"xy{i}_shape"]
    return np.unpackbits(REG[f"xy{i}_mask"], axis=1)[:, :w].astype(bool)


# ------------------------------------------------------------------ postprocess / output / image_ops
@pytest.mark.parametrize("i", range(int(POST["n_vote"])))
def test_oracle_vote_and_masks_match_reference_vectors(i):
    binary, pred = POST[f"vote{i}_binary"], POST[f"vote{i}_pred"].astype(np.int64)
    voted = opipe.vote_connected_component_class(pred.copy(), binary)
    assert np.array_equal(voted, POST[f"vote{i}_voted"])
    if f"vote{i}_color" in POST:
        color, overlay, inverted, fg = opipe.generate_output_masks(binary, voted, LUT)
        for got, key in ((color, "color"), (overlay, "overlay"), (inverted, "inverted"), (fg, "fg")):
            assert np.array_equal(got, POST[f"vote{i}_{key}"]), key


def test_reference_add_bounding_boxes_is_broken_and_registry_names():
    # postprocess.py:33 hands cv2 a bool array; the reference raises there, the product implements the intent
    assert str(POST["bbox_error"]) == "error"
    from page_segmentation_b200.lib.postprocess import find_postprocessor
    for key, name in zip(POST["pp_keys"], POST["pp_names"]):
        assert find_postprocessor(str(key)).__name__ == str(name)


@pytest.mark.parametrize("i", range(int(POST["n_char"])))
def test_oracle_char_height_matches_reference_vectors(i):
    got = oio.compute_char_height_array(POST[f"char{i}_page"], bool(POST[f"char{i}_inverse"]))
    assert (-1 if got is None else int(got)) == int(POST[f"char{i}_height"])


# ------------------------------------------------------------------ xycut
@pytest.mark.parametrize("i", range(int(REG["n_xy"])))
def test_xy_cut_matches_reference_vectors(i):
    mask, params = _unpack_mask(i), [int(v) for v in REG[f"xy{i}_params"]]
    exp = REG[f"xy{i}_rects"]
    assert np.array_equal(np.array(oreg.do_xy_cut(mask, *params), dtype=np.int64).reshape(-1, 4), exp)
    # the product's host recursion, fed with a CPU-made summed-area table (the device makes it in production)
    got = pxy.xy_cut_from_integral(oreg.integral_image(mask), *params)
    assert np.array_equal(_rect_array(got), exp)


@pytest.mark.parametrize("i", range(int(REG["n_seg"])))
def test_oracle_find_segments_and_contours_match_reference_vectors(i):
    image = REG[f"seg{i}_image"]
    h, ch, rh = (int(v) for v in REG[f"seg{i}_args"])
    text, pictures = oreg.find_segments(h, image, ch, rh, LUT[2], LUT[1])
    assert np.array_equal(np.array(text, dtype=np.int64).reshape(-1, 4), REG[f"seg{i}_text"])
    assert np.array_equal(np.array(pictures, dtype=np.int64).reshape(-1, 4), REG[f"seg{i}_pictures"])
    contours = oreg.get_text_contours(image, ch, LUT[1])
    offs = REG[f"seg{i}_contour_offsets"]
    assert len(contours) == len(offs) - 1
    for c, a, b in zip(contours, offs, offs[1:]):
        assert np.array_equal(np.asarray(c).reshape(-1, 2), REG[f"seg{i}_contour_points"][a:b])


def test_region_types_behave_like_the_reference():
    r = pxy.RectSegment(3, 5, 11, 17)
    assert r.scale(2.5) == pxy.RectSegment(7, 12, 27, 42)              # int() truncation, xycut.py:43-49
    assert r.as_xy() == [(5, 3), (17, 11)]
    assert r.polygon_coords() == [(3, 5), (11, 5), (11, 17), (3, 17)]
    img = np.arange(20 * 20).reshape(20, 20)
    assert np.array_equal(r.of(img), img[5:17, 3:11])
    c = pxy.CVContour(np.array([[[1, 2]], [[3, 4]], [[5, 7]]]))
    assert c.contour.shape == (3, 2) and np.array_equal(c.scale(1.5).contour, [[1, 3], [4, 6], [7, 10]])
    assert len(pxy.Segment1D(4, 9)) == 5
    assert pxy.single_color(np.array([[[1, 2, 3], [1, 2, 4]]]), np.array([1, 2, 3])).tolist() == [[True, False]]


# ------------------------------------------------------------------ OpenCV semantics the device kernels restate
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 13, 16, 18, 33, 40])
def test_rect_morphology_restatement_matches_cv2(k):
    rng = np.random.default_rng(k)
    img = ((rng.random((57, 83)) < 0.12) * 255).astype(np.uint8)
    img[0, :5] = 255
    img[-1, -3:] = 255
    kern = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
    assert np.array_equal(oreg.rect_morph(img, k, erode=False), cv2.dilate(img, kern))
    assert np.array_equal(oreg.rect_morph(img, k, erode=True), cv2.erode(img, kern))
    closed = oreg.rect_morph(oreg.rect_morph(img, k, False), k, True)
    assert np.array_equal(closed, cv2.morphologyEx(img, cv2.MORPH_CLOSE, kern))
    opened = oreg.rect_morph(oreg.rect_morph(img, k, True), k, False)
    assert np.array_equal(opened, cv2.morphologyEx(img, cv2.MORPH_OPEN, kern))


@pytest.mark.parametrize("src,dst", [((700, 500), (300, 214)), ((3508, 2480), (300, 212)), ((333, 517), (200, 310)),
                                     ((100, 100), (250, 130)), ((512, 384), (511, 383))])
def test_nearest_resize_index_restatement_matches_cv2(src, dst):
    rng = np.random.default_rng(src[0])
    img = rng.integers(0, 255, src + (3,), dtype=np.uint8)
    got = img[oreg.resize_nearest_index(dst[0], src[0])][:, oreg.resize_nearest_index(dst[1], src[1])]
    assert np.array_equal(got, cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_NEAREST))


def test_integral_image_restatement():
    rng = np.random.default_rng(1)
    m = rng.random((37, 53)) < 0.3
    sat = oreg.integral_image(m)
    assert sat.shape == (38, 54) and sat[0].sum() == 0 and sat[:, 0].sum() == 0
    for (r0, r1, c0, c1) in [(0, 37, 0, 53), (5, 20, 7, 8), (36, 37, 0, 53), (10, 10, 3, 9)]:
        assert sat[r1, c1] - sat[r0, c1] - sat[r1, c0] + sat[r0, c0] == m[r0:r1, c0:c1].sum()


# ------------------------------------------------------------------ live against the mounted reference
@live
def test_live_reference_xy_cut_random_masks():
    xy = ref.load("xycut")
    rng = np.random.default_rng(77)
    for case in range(60):
        h, w = int(rng.integers(20, 160)), int(rng.integers(20, 160))
        mask = np.zeros((h, w), bool)
        for _ in range(int(rng.integers(1, 8))):                       # a few random blocks of speckled ink
            y0, x0 = int(rng.integers(0, h - 4)), int(rng.integers(0, w - 4))
            y1, x1 = int(rng.integers(y0 + 2, h + 1)), int(rng.integers(x0 + 2, w + 1))
            mask[y0:y1, x0:x1] |= rng.random((y1 - y0, x1 - x0)) < rng.uniform(0.2, 0.9)
        params = [int(v) for v in rng.integers(1, 12, 4)]
        exp = _rect_array(xy.do_xy_cut(mask, *params))
        assert np.array_equal(np.array(oreg.do_xy_cut(mask, *params), dtype=np.int64).reshape(-1, 4), exp), (case, params)
        got = pxy.xy_cut_from_integral(oreg.integral_image(mask), *params)
        assert np.array_equal(_rect_array(got), exp), (case, params)


@live
def test_live_reference_region_extraction_and_vote():
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    pcs, pp, ds = ref.load("pc_segmentation"), ref.load("postprocess"), ref.load("dataset")
    for seed, h, w, ch, rh in [(11, 640, 460, 16, 300), (12, 450, 620, 10, 220)]:
        image = synth.make_inverted_image(seed, h, w, ch)
        t, im = pcs.find_segments(h, image, ch, rh, DEFAULT_COLOR_MAP)
        ot, oi = oreg.find_segments(h, image, ch, rh, LUT[2], LUT[1])
        assert np.array_equal(_rect_array(t), np.array(ot, dtype=np.int64).reshape(-1, 4))
        assert np.array_equal(_rect_array(im), np.array(oi, dtype=np.int64).reshape(-1, 4))
        exp = pcs.get_text_contours(image, ch, DEFAULT_COLOR_MAP)
        got = oreg.get_text_contours(image, ch, LUT[1])
        assert len(exp) == len(got) and all(np.array_equal(a.contour, b) for a, b in zip(exp, got))
    rng = np.random.default_rng(5)
    page = synth.make_page(21, 300, 260, 14)
    binary = (page == 0).astype(np.uint8)
    pred = rng.integers(0, 4, binary.shape).astype(np.int64)
    exp = pp.vote_connected_component_class(pred.copy(), ds.SingleData(binary=binary))
    assert np.array_equal(opipe.vote_connected_component_class(pred.copy(), binary), exp)


@live
def test_live_reference_api_surface():
    """Field names / order / defaults of the value types, the Architecture enum and the two pure helpers."""
    import dataclasses
    import importlib

    def fields(c):
        if dataclasses.is_dataclass(c):
            return [(f.name, None if f.default is dataclasses.MISSING else f.default) for f in dataclasses.fields(c)]
        return [(f, c._field_defaults.get(f)) for f in c._fields]

    for mod, names in [("dataset", ["SingleData", "Dataset"]), ("predictor_data", ["PredictSettings", "Prediction"]),
                       ("output", ["Masks"]), ("xycut", ["RectSegment", "CVContour", "Segment1D", "Gap"])]:
        theirs, ours = ref.load(mod), importlib.import_module("page_segmentation_b200.lib." + mod)
        for n in names:
            assert fields(getattr(theirs, n)) == fields(getattr(ours, n)), n
    ra = ref.load("architecture")
    from page_segmentation_b200.lib import architecture as ma
    assert [(a.name, a.value) for a in ra.Architecture] == [(a.name, a.value) for a in ma.Architecture]
    x = np.arange(12, dtype=np.uint8).reshape(3, 4)
    assert np.array_equal(ra.default_preprocess(x), ma.default_preprocess(x))
    ru = ref.load("util")
    from page_segmentation_b200.lib import util as mu
    assert np.array_equal(ru.gray_to_rgb(x), mu.gray_to_rgb(x))
    assert np.array_equal(ru.image_to_batch(x), mu.image_to_batch(x))
    rp = ref.load("postprocess")
    from page_segmentation_b200.lib import postprocess as mp
    assert sorted(rp.POSTPROCESSORS) == sorted(mp.POSTPROCESSORS) and rp.postprocess_help() == mp.postprocess_help()


@live
def test_live_reference_list_dataset(tmp_path):
    import json
    from page_segmentation_b200.lib import dataset as md
    rd = ref.load("dataset")
    root = tmp_path / "ds"
    for sub in ("binary_images", "images", "masks", "normalizations", "together"):
        (root / sub).mkdir(parents=True)
    names = ["p003", "p001", "p010", "p002"]
    for i, n in enumerate(names):
        (root / "binary_images" / f"{n}.bin.png").write_bytes(b"x")
        (root / "images" / f"{n}.png").write_bytes(b"x")
        (root / "masks" / f"{n}.mask.png").write_bytes(b"x")
        (root / "normalizations" / f"{n}.norm").write_text(json.dumps({"char_height": 17 + i}))
        (root / "together" / f"{n}.png").write_bytes(b"x")
        (root / "together" / f"{n}_GT.png").write_bytes(b"x")
    (root / "images" / "extra.png").write_bytes(b"x")                       # page without mask / binary
    key = lambda rec: rec["image_path"]                                      # noqa: E731
    assert md.list_dataset(str(root), line_height_px=9, images_dir_="together", masks_dir_="together",
                           masks_postfix="_GT.png") == \
        rd.list_dataset(str(root), line_height_px=9, images_dir_="together", masks_dir_="together", masks_postfix="_GT.png")
    for kw in (dict(verify_filenames=True), dict(verify_filenames=True, line_height_px=12)):
        assert sorted(md.list_dataset(str(root), **kw), key=key) == sorted(rd.list_dataset(str(root), **kw), key=key)
    for kw in (dict(), dict(line_height_px=5)):                             # 5 images vs 4 masks
        with pytest.raises(Exception) as e1:
            md.list_dataset(str(root), **kw)
        with pytest.raises(Exception) as e2:
            rd.list_dataset(str(root), **kw)
        assert type(e1.value) is type(e2.value) and str(e1.value) == str(e2.value)
    (root / "images" / "extra.png").unlink()
    assert md.list_dataset(str(root)) == rd.list_dataset(str(root))
    with pytest.raises(Exception, match="Dataset dir does not exist"):
        md.list_dataset(str(root / "nope"))
    with pytest.raises(Exception, match="Norm dir does not exist"):
        md.list_dataset(str(root), normalizations_dir="absent")


@pytest.mark.parametrize("i", range(int(POST["n_vote"])))
def test_oracle_eval_metrics_match_reference_vectors(i):
    pred, voted, binary = (POST[f"vote{i}_{k}"].astype(np.int64) for k in ("pred", "voted", "binary"))
    assert oio.fgpa(voted, pred, binary) == float(POST[f"eval{i}_fgpa"])
    ov, tp, fp, fn = oio.fgoverlap_per_class(voted, pred, binary, int(POST[f"eval{i}_ncls"]))
    assert np.array_equal(np.array(ov), POST[f"eval{i}_overlap"], equal_nan=True)
    assert tp == POST[f"eval{i}_tp"].tolist() and fp == POST[f"eval{i}_fp"].tolist() and fn == POST[f"eval{i}_fn"].tolist()


@live
def test_live_reference_eval_metrics():
    iops = ref.load("image_ops")
    rng = np.random.default_rng(9)
    for n_classes, shape in ((3, (50, 70)), (5, (33, 20)), (2, (8, 8))):
        pred, mask = rng.integers(0, n_classes + 2, shape), rng.integers(0, n_classes, shape)       # pred also holds out-of-range classes
        binary = (rng.random(shape) < 0.4).astype(np.int64)
        assert oio.fgpa(pred, mask, binary) == iops.fgpa(pred, mask, binary)
        exp, got = iops.fgoverlap_per_class(pred, mask, binary, n_classes), oio.fgoverlap_per_class(pred, mask, binary, n_classes)
        assert np.array_equal(np.array(exp[0]), np.array(got[0]), equal_nan=True) and list(exp[1:]) == list(got[1:])
    with np.errstate(invalid="ignore"):
        assert np.isnan(iops.fgpa(pred, mask, np.zeros(shape, np.int64))) and np.isnan(oio.fgpa(pred, mask, np.zeros(shape, np.int64)))
