"""Parity of pcs_char_height with the CPU oracle restating compute_char_height (image_ops.py:58-82).
Integer result: bit-exact bar."""
import os

import numpy as np
import pytest

from oracle import image_ops as oio
from page_segmentation_b200 import synth

pytestmark = pytest.mark.gpu


def _text_page(seed, h, w, glyph_h):
    """grey page with glyph-like boxes of about glyph_h pixels plus noise, so that Otsu has work to do"""
    rng = np.random.default_rng(seed)
    page = rng.normal(215, 8, (h, w))
    y = 20
    while y + glyph_h + 10 < h:
        x = 15
        while x + 40 < w:
            gh = int(glyph_h + rng.integers(-3, 4))
            gw = int(max(6, gh * rng.uniform(0.45, 1.4)))
            page[y:y + gh, x:x + gw] = rng.normal(40, 10, (gh, gw))
            if rng.random() < 0.2:                     # a dot touching the glyph only through a corner
                page[y + gh:y + gh + 3, x + gw:x + gw + 3] = 35
            x += gw + int(rng.integers(4, 14))
        y += glyph_h + int(rng.integers(8, 20))
    return np.clip(page, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("seed,shape,gh", [(0, (400, 600), 18), (1, (700, 500), 30), (2, (333, 517), 12), (3, (256, 256), 45)])
def test_char_height_matches_oracle(ctx, seed, shape, gh, inverse):
    from page_segmentation_b200.lib.image_ops import compute_char_height_array
    page = _text_page(seed, *shape, gh)
    if inverse:
        page = 255 - page
    got = compute_char_height_array(page, inverse)
    exp = oio.compute_char_height_array(page, inverse)
    assert (got is None) == (exp is None)
    if exp is not None:
        assert int(got) == int(exp)
        assert abs(int(got) - gh) <= 4


# Two half glyphs that touch only through one corner, placed so that the corner sits on a border of the 256 x 32
# labelling tiles (csrc/ccl.cu): merged (8-connectivity) the glyph is 24 rows high, unmerged there are two of 12.
_CORNERS = {
    "tile corner, down-right": ((20, 249), (32, 256)),          # (31,255) - (32,256): row 32 starts a tile row
    "tile corner, down-left": ((20, 256), (32, 249)),           # (31,256) - (32,255)
    "vertical tile border, down-right": ((40, 249), (52, 256)),  # (51,255) - (52,256): inside a tile row
    "vertical tile border, down-left": ((40, 256), (52, 249)),   # (51,256) - (52,255): last pixel of a tile's last segment
    "horizontal tile border, down-right": ((52, 100), (64, 107)),  # (63,106) - (64,107): inside a tile column
    "horizontal tile border, down-left": ((52, 107), (64, 100)),
    "segment border inside a tile, down-right": ((70, 57), (82, 64)),  # (81,63) - (82,64)
    "segment border inside a tile, down-left": ((70, 64), (82, 57)),
}


@pytest.mark.parametrize("case", sorted(_CORNERS))
def test_char_height_diagonal_contacts_on_tile_borders(ctx, case):
    from page_segmentation_b200.lib.image_ops import compute_char_height_array
    (ya, xa), (yb, xb) = _CORNERS[case]
    page = np.full((128, 600), 255, np.uint8)
    page[ya:ya + 12, xa:xa + 7] = 0
    page[yb:yb + 12, xb:xb + 7] = 0
    got, exp = compute_char_height_array(page, False), oio.compute_char_height_array(page, False)
    assert exp == 24 and got == exp
    page[yb:yb + 12, xb:xb + 7] = 255                       # the same halves one column apart: two letters of 12 rows
    xb += 1 if xb > xa else -1
    page[yb:yb + 12, xb:xb + 7] = 0
    got, exp = compute_char_height_array(page, False), oio.compute_char_height_array(page, False)
    assert exp == 12 and got == exp


def test_char_height_binarised_page_and_none(ctx):
    from page_segmentation_b200.lib.image_ops import compute_char_height_array
    page = synth.make_page(3, 900, 700, 18)
    assert compute_char_height_array(page, False) == oio.compute_char_height_array(page, False)
    blank = np.full((200, 300), 255, np.uint8)
    blank[0, 0] = 0                                      # Otsu needs two levels; a single dot is no letter
    assert compute_char_height_array(blank, False) is None and oio.compute_char_height_array(blank, False) is None


def test_char_height_file_api(ctx, tmp_path):
    import cv2
    from page_segmentation_b200.lib.image_ops import compute_char_height
    page = _text_page(5, 300, 420, 20)
    f = os.path.join(tmp_path, "page.png")
    cv2.imwrite(f, page)
    assert int(compute_char_height(f, False)) == int(oio.compute_char_height_array(page, False))
    with pytest.raises(Exception, match="File does not exist"):
        compute_char_height(os.path.join(tmp_path, "missing.png"), False)


def test_char_height_batch(ctx):
    import torch
    pages = np.stack([_text_page(s, 320, 480, 14 + 6 * s) for s in range(3)])
    d = torch.from_numpy(pages).cuda()
    out = torch.empty((3,), dtype=torch.int32, device="cuda")
    ctx.char_height(d, 3, 320, 480, False, out)
    for i in range(3):
        assert int(out[i].cpu()) == int(oio.compute_char_height_array(pages[i], False))


def test_eval_metrics_reproduce_reference(ctx):
    """fgpa / fgoverlap_per_class (image_ops.py:8-55) against vectors produced by the reference's own functions."""
    from page_segmentation_b200.lib.image_ops import fgoverlap_per_class, fgpa
    post = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_postprocess.npz"))
    for i in range(int(post["n_vote"])):
        pred, voted, binary = (post[f"vote{i}_{k}"] for k in ("pred", "voted", "binary"))
        assert fgpa(voted, pred, binary) == float(post[f"eval{i}_fgpa"])
        ov, tp, fp, fn = fgoverlap_per_class(voted, pred, binary, int(post[f"eval{i}_ncls"]))
        assert np.array_equal(np.array(ov), post[f"eval{i}_overlap"], equal_nan=True)
        assert tp == post[f"eval{i}_tp"].tolist() and fp == post[f"eval{i}_fp"].tolist() and fn == post[f"eval{i}_fn"].tolist()
    rng = np.random.default_rng(4)
    pred, mask = rng.integers(0, 6, (1169, 827)), rng.integers(0, 4, (1169, 827))        # classes above n_classes in pred
    binary = (rng.random(pred.shape) < 0.3).astype(np.uint8)
    assert fgpa(pred, mask, binary) == oio.fgpa(pred, mask, binary)
    got, exp = fgoverlap_per_class(pred, mask, binary, 4), oio.fgoverlap_per_class(pred, mask, binary, 4)
    assert np.array_equal(np.array(got[0]), np.array(exp[0]), equal_nan=True) and list(got[1:]) == list(exp[1:])
    assert np.isnan(fgpa(pred, mask, np.zeros_like(binary)))
