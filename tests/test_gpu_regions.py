"""Parity of the device region-extraction path (pcs_segment_masks, pcs_dilate3x3, pcs_integral_image,
pcs_text_regions and the lib/pc_segmentation.py + lib/xycut.py mirrors on top of them) with
  * the vectors the REFERENCE's own code produced (tests/golden/ref_regions.npz, ref_postprocess.npz), and
  * the CPU oracle (oracle/regions.py, live cv2) on further seeded cases up to full A4 size.
Integer / index work throughout: the bar is bit-exact."""
import os

import cv2
import numpy as np
import pytest

from oracle import regions as oreg
from page_segmentation_b200 import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REG = np.load(os.path.join(HERE, "golden", "ref_regions.npz"))
POST = np.load(os.path.join(HERE, "golden", "ref_postprocess.npz"))
TEXT, PICTURE = (255, 0, 0), (0, 255, 0)


def _rects(segs):
    return np.array([[s.x_start, s.y_start, s.x_end, s.y_end] for s in segs], dtype=np.int64).reshape(-1, 4)


def _unpack_mask(i):
    h, w = REG[f"xy{i}_shape"]
    return np.unpackbits(REG[f"xy{i}_mask"], axis=1)[:, :w].astype(bool)


# ------------------------------------------------------------------ against the reference-made vectors
@pytest.mark.parametrize("i", range(int(REG["n_xy"])))
def test_do_xy_cut_reproduces_reference(ctx, i):
    from page_segmentation_b200.lib.xycut import do_xy_cut
    got = do_xy_cut(_unpack_mask(i), *[int(v) for v in REG[f"xy{i}_params"]])
    assert np.array_equal(_rects(got), REG[f"xy{i}_rects"])


@pytest.mark.parametrize("i", range(int(REG["n_seg"])))
def test_find_segments_and_text_contours_reproduce_reference(ctx, i):
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.pc_segmentation import find_segments, get_text_contours
    image = REG[f"seg{i}_image"]
    h, ch, rh = (int(v) for v in REG[f"seg{i}_args"])
    text, pictures = find_segments(h, image, ch, rh, DEFAULT_COLOR_MAP)
    assert np.array_equal(_rects(text), REG[f"seg{i}_text"])
    assert np.array_equal(_rects(pictures), REG[f"seg{i}_pictures"])
    t2, p2 = find_segments(h, image, ch, rh, DEFAULT_COLOR_MAP, only_images=True)
    assert t2 == [] and np.array_equal(_rects(p2), REG[f"seg{i}_pictures"])
    contours = get_text_contours(image, ch, DEFAULT_COLOR_MAP)
    offs = REG[f"seg{i}_contour_offsets"]
    assert len(contours) == len(offs) - 1
    for c, a, b in zip(contours, offs, offs[1:]):
        assert np.array_equal(np.asarray(c.contour).reshape(-1, 2), REG[f"seg{i}_contour_points"][a:b])


@pytest.mark.parametrize("i", range(int(POST["n_vote"])))
def test_vote_and_masks_reproduce_reference(ctx, i):
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import SingleData
    from page_segmentation_b200.lib.output import generate_output_masks
    from page_segmentation_b200.lib.postprocess import vote_connected_component_class
    binary, pred = POST[f"vote{i}_binary"], POST[f"vote{i}_pred"].astype(np.int64)
    data = SingleData(binary=binary)
    voted = vote_connected_component_class(pred.copy(), data)
    assert voted.dtype == np.int64 and np.array_equal(voted, POST[f"vote{i}_voted"])
    if f"vote{i}_color" in POST:
        m = generate_output_masks(data, voted, DEFAULT_COLOR_MAP)
        assert np.array_equal(m.color, POST[f"vote{i}_color"])
        assert np.array_equal(m.overlay, POST[f"vote{i}_overlay"])
        assert np.array_equal(m.inverted_overlay, POST[f"vote{i}_inverted"])
        assert np.array_equal(m.fg_color_mask, POST[f"vote{i}_fg"])


@pytest.mark.parametrize("i", range(int(POST["n_char"])))
def test_char_height_reproduces_reference(ctx, i):
    from page_segmentation_b200.lib.image_ops import compute_char_height_array
    got = compute_char_height_array(POST[f"char{i}_page"], bool(POST[f"char{i}_inverse"]))
    assert (-1 if got is None else int(got)) == int(POST[f"char{i}_height"])


# ------------------------------------------------------------------ kernels against the oracle / cv2
def _device_segment_masks(ctx, image, Ho, Wo, colours):
    import torch
    d_rgb = torch.from_numpy(np.ascontiguousarray(image)).cuda()
    d_masks = torch.empty((len(colours), Ho, Wo), dtype=torch.uint8, device="cuda")
    ctx.segment_masks(d_rgb, image.shape[0], image.shape[1], Ho, Wo, np.array(colours, np.uint8), d_masks)
    return d_masks.cpu().numpy()


@pytest.mark.parametrize("seed,shape,out", [(0, (700, 500), (300, 214)), (1, (3508, 2480), (300, 212)), (2, (333, 517), (200, 310)),
                                            (3, (120, 90), (260, 195)), (4, (512, 384), (511, 383)), (5, (65, 33), (7, 3))])
def test_segment_masks_match_cv2(ctx, seed, shape, out):
    image = synth.make_inverted_image(seed, shape[0], shape[1], max(4, shape[0] // 40))
    got = _device_segment_masks(ctx, image, out[0], out[1], [PICTURE, TEXT, (0, 0, 0)])
    exp = oreg.segment_masks(image, out[0], out[1], [PICTURE, TEXT, (0, 0, 0)])
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("shape,c", [((64, 80), 1), ((233, 97), 3), ((50, 31), 4)])
def test_dilate_matches_cv2(ctx, shape, c):
    from page_segmentation_b200.lib.pc_segmentation import dilate
    rng = np.random.default_rng(shape[0])
    img = (rng.integers(0, 256, shape + ((c,) if c > 1 else ()), dtype=np.uint8) * (rng.random(shape + ((c,) if c > 1 else ())) < 0.2)).astype(np.uint8)
    assert np.array_equal(dilate(img), cv2.dilate(img, np.ones((3, 3), np.uint8), iterations=1))


@pytest.mark.parametrize("n,h,w", [(1, 1, 1), (2, 37, 53), (1, 300, 212), (3, 129, 1000), (1, 3508, 2480)])
def test_integral_image_matches_numpy(ctx, n, h, w):
    from page_segmentation_b200.lib.xycut import integral_image
    rng = np.random.default_rng(h * 7 + w)
    masks = (rng.random((n, h, w)) < 0.4) * rng.integers(1, 255, (n, h, w))
    got = integral_image(masks.astype(np.uint8))
    for i in range(n):
        assert np.array_equal(got[i], oreg.integral_image(masks[i]))


@pytest.mark.parametrize("seed,shape,ch", [(0, (700, 500), 18), (1, (333, 517), 9), (2, (64, 31), 3), (3, (100, 257), 40),
                                           (4, (480, 640), 33), (5, (96, 96), 64)])
def test_text_region_masks_match_cv2(ctx, seed, shape, ch):
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.pc_segmentation import text_region_masks
    image = synth.make_inverted_image(seed, shape[0], shape[1], min(ch, 24))
    canvas, region = text_region_masks(image, ch, DEFAULT_COLOR_MAP)
    exp_canvas, exp_region = oreg.text_region_masks(image, ch, TEXT)
    assert np.array_equal(canvas, exp_canvas)
    assert np.array_equal(region, exp_region)


def test_text_regions_reject_empty_structuring_element(ctx):
    from page_segmentation_b200._native import PcsError
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.pc_segmentation import text_region_masks
    with pytest.raises(PcsError):                       # int(2 / 3) == 0: cv2 raises on the empty element as well
        text_region_masks(synth.make_inverted_image(0, 64, 64, 6), 2, DEFAULT_COLOR_MAP)
    with pytest.raises(cv2.error):
        oreg.text_region_masks(synth.make_inverted_image(0, 64, 64, 6), 2, TEXT)


def test_full_a4_region_extraction_matches_oracle(ctx):
    """The `inverted` image of a whole A4 page at 300 dpi (26 MB), char height 40 px."""
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.pc_segmentation import find_segments, get_text_contours
    image = synth.make_inverted_image(9, synth.A4_H, synth.A4_W, 40)
    text, pictures = find_segments(synth.A4_H, image, 40, 300, DEFAULT_COLOR_MAP)
    otext, opictures = oreg.find_segments(synth.A4_H, image, 40, 300, PICTURE, TEXT)
    assert len(otext) >= 2 and len(opictures) >= 1
    assert np.array_equal(_rects(text), np.array(otext, dtype=np.int64).reshape(-1, 4))
    assert np.array_equal(_rects(pictures), np.array(opictures, dtype=np.int64).reshape(-1, 4))
    got = get_text_contours(image, 40, DEFAULT_COLOR_MAP)
    exp = oreg.get_text_contours(image, 40, TEXT)
    assert len(exp) >= 2 and len(got) == len(exp)
    assert all(np.array_equal(a.contour, b) for a, b in zip(got, exp))
