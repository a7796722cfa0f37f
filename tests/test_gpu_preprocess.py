"""Parity of the device preprocess (pcs_preprocess / pcs_resize_nearest) with the
CPU oracle restating dataset.py:114-150 and util.py:21-29.  Bit-exact bar."""
import numpy as np
import pytest

from oracle import pipeline as opipe
from oracle import resize as osk
from page_segmentation_b200 import synth

pytestmark = pytest.mark.gpu


def _run(page_grey, page_bin, target_lh, lh):
    from page_segmentation_b200.lib.dataset import prepare_images
    return prepare_images(page_grey, page_bin, target_lh, lh, keep_orig_bin=True)


@pytest.mark.parametrize("shape,lh", [((3508, 2480), 18), ((700, 500), 18), ((333, 517), 11), ((64, 64), 6),
                                      ((97, 131), 5), ((1200, 900), 24)])
def test_binarised_page_bit_exact(ctx, shape, lh):
    page = synth.make_page(3, shape[0], shape[1], lh)
    img, b, ob = _run(page, page, 6, lh)
    eimg, eb, eob = opipe.prepare_images(page, page, 6, lh, keep_orig_bin=True)
    assert img.shape == eimg.shape and img.dtype == np.uint8
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(ob, eob)
    np.testing.assert_array_equal(img, eimg)


def test_upscale_reflect_borders(ctx):
    # scale > 1 exercises the reflect border handling of the 4x4 cubic taps
    page = synth.make_page(5, 120, 90, 4)
    img, b, ob = _run(page, page, 6, 4)
    eimg, eb, eob = opipe.prepare_images(page, page, 6, 4, keep_orig_bin=True)
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(img, eimg)


def test_separate_binary_input(ctx):
    grey = synth.make_page(1, 400, 300, 18)
    binary = synth.make_page(2, 400, 300, 18)
    img, b, ob = _run(grey, binary, 6, 18)
    eimg, eb, eob = opipe.prepare_images(grey, binary, 6, 18, keep_orig_bin=True)
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(ob, eob)
    np.testing.assert_array_equal(img, eimg)


def test_binary_01_input(ctx):
    # `binary / 255 if max > 1 else binary` (dataset.py:135): a {0,1} page is used as is
    grey = synth.make_page(1, 300, 200, 18)
    b01 = (synth.make_page(2, 300, 200, 18) > 0).astype(np.uint8)
    img, b, ob = _run(grey, b01, 6, 18)
    eimg, eb, eob = opipe.prepare_images(grey, b01, 6, 18, keep_orig_bin=True)
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(ob, eob)


@pytest.mark.parametrize("shape,lh", [((600, 420), 18), ((301, 203), 9)])
def test_grey_page_antialiased(ctx, shape, lh):
    """>2 grey levels => anti_aliasing=True (dataset.py:127).  The Gaussian
    weights go through exp(); tolerance: <= 1 grey level on <= 1e-4 of pixels."""
    grey = synth.make_grey_page(4, shape[0], shape[1], lh)
    assert len(np.unique(grey)) > 2
    img, b, ob = _run(grey, grey, 6, lh)
    eimg, eb, eob = opipe.prepare_images(grey, grey, 6, lh, keep_orig_bin=True)
    np.testing.assert_array_equal(b, eb)
    diff = np.abs(img.astype(int) - eimg.astype(int))
    assert diff.max() <= 1
    assert (diff > 0).mean() <= 1e-4


def test_batched_pages_match_single(ctx):
    import torch
    pages = np.stack([synth.make_page(s, 480, 360, 18) for s in range(3)])
    Hs, Ws = synth.scaled_shape(480, 360, 6 / 18)
    d = torch.from_numpy(pages).cuda()
    d_img = torch.empty((3, Hs, Ws), dtype=torch.uint8, device="cuda")
    d_bin = torch.empty_like(d_img)
    ctx.preprocess(d, d, 3, 480, 360, Hs, Ws, d_img, d_bin, None)
    for i in range(3):
        eimg, eb = opipe.prepare_images(pages[i], pages[i], 6, 18)
        np.testing.assert_array_equal(d_img[i].cpu().numpy(), eimg)
        np.testing.assert_array_equal(d_bin[i].cpu().numpy(), eb)


# ---- two-level fast path (bitmap + 16-pattern LUT): pages of a multiple of 32 bytes, scale factor <= 4 ----
@pytest.mark.parametrize("shape,lh,target", [((128, 96), 4, 6), ((256, 160), 6, 6), ((640, 480), 20, 6), ((352, 512), 23, 6),
                                             ((640, 480), 30, 6)])
def test_fast_path_shapes_and_scales(ctx, shape, lh, target):
    """up-scale (reflect borders inside the tile), scale 1, ~1/3.3, ~1/3.8 and 1/5 (beyond the fast path's
    span limit -> general kernel) on shapes that are eligible for the fast path"""
    page = synth.make_page(11, shape[0], shape[1], max(lh, 6))
    img, b, ob = _run(page, page, target, lh)
    eimg, eb, eob = opipe.prepare_images(page, page, target, lh, keep_orig_bin=True)
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(ob, eob)
    np.testing.assert_array_equal(img, eimg)


@pytest.mark.parametrize("levels,first_is_ink", [((0, 255), False), ((0, 255), True), ((10, 200), True), ((1, 2), False)])
def test_fast_path_levels(ctx, levels, first_is_ink):
    """the bitmap is relative to page[0]: either level may come first; levels need not be 0/255"""
    rng = np.random.default_rng(5)
    mask = rng.random((320, 256)) < 0.3
    mask[0, 0] = first_is_ink
    page = np.where(mask, levels[0], levels[1]).astype(np.uint8)
    img, b, ob = _run(page, page, 6, 15)
    eimg, eb, eob = opipe.prepare_images(page, page, 6, 15, keep_orig_bin=True)
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(ob, eob)
    np.testing.assert_array_equal(img, eimg)


def test_fast_path_mixed_batch(ctx):
    """one call, pages of different kinds: binarised, blank (one level), three levels (anti-aliased general
    path), grey; each page must match its own single-page oracle result"""
    import torch
    H, W, lh = 480, 352, 18
    two = synth.make_page(2, H, W, lh)
    blank = np.full((H, W), 255, np.uint8)
    three = two.copy(); three[100:140, 50:90] = 128
    grey = synth.make_grey_page(4, H, W, lh)
    pages = np.stack([two, blank, three, grey, two[::-1].copy()])
    n = len(pages)
    Hs, Ws = synth.scaled_shape(H, W, 6 / lh)
    d = torch.from_numpy(pages).cuda()
    d_img = torch.empty((n, Hs, Ws), dtype=torch.uint8, device="cuda")
    d_bin = torch.empty_like(d_img)
    ctx.preprocess(d, d, n, H, W, Hs, Ws, d_img, d_bin, None)
    for i in range(n):
        eimg, eb = opipe.prepare_images(pages[i], pages[i], 6, lh)
        np.testing.assert_array_equal(d_bin[i].cpu().numpy(), eb)
        diff = np.abs(d_img[i].cpu().numpy().astype(int) - eimg.astype(int))
        if len(np.unique(pages[i])) <= 2:
            assert diff.max() == 0, i
        else:
            assert diff.max() <= 1 and (diff > 0).mean() <= 1e-4, i


def test_fast_path_separate_binary(ctx):
    grey = synth.make_page(1, 384, 320, 18)
    binary = synth.make_page(2, 384, 320, 18)
    img, b, ob = _run(grey, binary, 6, 18)
    eimg, eb, eob = opipe.prepare_images(grey, binary, 6, 18, keep_orig_bin=True)
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(ob, eob)
    np.testing.assert_array_equal(img, eimg)


@pytest.mark.parametrize("kind,shape,lh,max_width", [("bin", (600, 480), 18, 100), ("bin", (333, 517), 11, 200),
                                                     ("grey", (480, 360), 12, 90), ("bin", (640, 480), 18, 1000)])
def test_max_width_second_pass(ctx, kind, shape, lh, max_width):
    """dataset.py:139-143: a second rescale when the first result is wider than max_width (the last case is
    not: no second pass).  The second pass anti-aliases the fp64 image with Gaussian weights that go through
    exp(): tolerance <= 1 grey level on <= 1e-3 of the pixels; the binary is exact."""
    from page_segmentation_b200.lib.dataset import prepare_images
    page = synth.make_page(7, *shape, lh) if kind == "bin" else synth.make_grey_page(7, *shape, lh)
    img, b, ob = prepare_images(page, page, 6, lh, max_width=max_width, keep_orig_bin=True)
    eimg, eb, eob = opipe.prepare_images(page, page, 6, lh, max_width=max_width, keep_orig_bin=True)
    assert img.shape == eimg.shape and b.shape == eb.shape
    assert img.shape[1] <= max(max_width, 1) or max_width >= round(shape[1] * 6 / lh)
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(ob, eob)
    diff = np.abs(img.astype(int) - eimg.astype(int))
    assert diff.max() <= 1, diff.max()
    assert (diff > 0).mean() <= 1e-3, (diff > 0).mean()


@pytest.mark.parametrize("src,dst", [((50, 40), (150, 121)), ((389, 275), (1169, 827)), ((120, 90), (40, 30))])
def test_preserving_resize(ctx, src, dst):
    from page_segmentation_b200.lib.util import preserving_resize
    rng = np.random.default_rng(0)
    a = rng.integers(0, 4, size=src).astype(np.int64)
    got = preserving_resize(a, dst)
    exp = osk.resize(a, dst, order=0)
    assert got.dtype == np.float64
    np.testing.assert_array_equal(got, exp)


def _max_width_on(c, page, lh, max_width):
    """pcs_preprocess_max_width on context `c` -> (image, binary) numpy."""
    import torch
    H, W = page.shape
    H1, W1 = synth.scaled_shape(H, W, 6 / lh)
    H2, W2 = synth.scaled_shape(H1, W1, max_width / W1)
    d = torch.from_numpy(page).cuda()
    d_img = torch.empty((H2, W2), dtype=torch.uint8, device="cuda")
    d_bin = torch.empty((H2, W2), dtype=torch.uint8, device="cuda")
    c.preprocess_max_width(d, d, 1, H, W, H1, W1, H2, W2, d_img, d_bin, None)
    c.synchronize()
    return d_img.cpu().numpy(), d_bin.cpu().numpy()


def _assert_max_width(got, page, lh, max_width):
    eimg, eb = opipe.prepare_images(page, page, 6, lh, max_width=max_width)
    np.testing.assert_array_equal(got[1], eb)
    diff = np.abs(got[0].astype(int) - eimg.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() <= 1e-3, (diff.max(), (diff > 0).mean())


def test_max_width_is_the_first_call_on_a_fresh_context():
    """the max_width pass carves its fp64 planes out of a second scratch buffer and then runs a whole
    first pass, which grows the first scratch buffer; growing it must not free the second one under its feet.  A fresh
    pcs_ctx (not the session's, whose scratch has long been grown) makes the first pass grow scratch inside the call."""
    import torch
    from page_segmentation_b200 import _native
    assert torch.cuda.is_available()
    c = _native.Context(0)
    try:
        c.use_torch_stream()
        page = synth.make_page(7, 600, 480, 18)
        _assert_max_width(_max_width_on(c, page, 18, 100), page, 18, 100)
    finally:
        c.close()


def test_scratch_growth_between_two_max_width_calls():
    """... and a scratch-growing call (connected components of a much larger page) between two max_width calls must
    leave the second scratch buffer alive: the second call re-uses it."""
    import torch
    from page_segmentation_b200 import _native
    c = _native.Context(0)
    try:
        c.use_torch_stream()
        page = synth.make_page(8, 333, 517, 11)
        first = _max_width_on(c, page, 11, 200)
        _assert_max_width(first, page, 11, 200)
        big = torch.from_numpy((synth.make_page(9, 2000, 1500, 18) < 128).astype(np.uint8)).cuda()
        labels = torch.empty(big.shape, dtype=torch.int32, device="cuda")
        ncomp = torch.zeros((1,), dtype=torch.int32, device="cuda")
        c.ccl(big, 1, big.shape[0], big.shape[1], labels, None, 0, ncomp)          # grows ctx->scratch
        c.synchronize()
        second = _max_width_on(c, page, 11, 200)
        np.testing.assert_array_equal(first[0], second[0])
        np.testing.assert_array_equal(first[1], second[1])
    finally:
        c.close()


@pytest.mark.parametrize("case", ["one_stray_pixel_last", "one_stray_pixel_first_row", "sum_neutral_pair", "levels_around_mean",
                                  "one_level", "one_pixel_differs"])
def test_two_level_verdict_is_exact(ctx, case):
    """The fast path decides "at most two grey levels" from three sums per page (scan_pack_kernel: count, sum and sum of
    squares of byte ^ page[0]).  Pages built to fool anything weaker than the exact test -- a single pixel of a third
    level, third levels that leave the sum unchanged -- must take the anti-aliased general path like the oracle
    (`len(np.unique(image)) > 2`, dataset.py:127); pages with one or two levels must stay bit-exact."""
    H, W, lh = 480, 352, 18
    page = synth.make_page(21, H, W, lh)                     # levels {0, 255}, page[0, 0] == 255
    exact = True
    if case == "one_stray_pixel_last":
        page[-1, -1] = 254
        exact = False
    elif case == "one_stray_pixel_first_row":
        page[0, 5] = 1
        exact = False
    elif case == "sum_neutral_pair":                         # two ink pixels 0 -> 1 and 0 -> 255 ^ ... keep the count, move the sums
        ys, xs = np.nonzero(page == 0)
        page[ys[0], xs[0]] = 1
        page[ys[1], xs[1]] = 2
        exact = False
    elif case == "levels_around_mean":                       # x = byte ^ 255 in {127, 128, 129} in equal numbers: mean 128
        page[:] = 255
        page[10:40, 10:40] = 255 ^ 128
        page[10:20, 10:40] = 255 ^ 127
        page[30:40, 10:40] = 255 ^ 129
        exact = False
    elif case == "one_level":
        page[:] = 77
    elif case == "one_pixel_differs":
        page[:] = 255
        page[H // 2, W // 2] = 0
    img, b, ob = _run(page, page, 6, lh)
    eimg, eb, eob = opipe.prepare_images(page, page, 6, lh, keep_orig_bin=True)
    np.testing.assert_array_equal(b, eb)
    np.testing.assert_array_equal(ob, eob)
    if exact:
        np.testing.assert_array_equal(img, eimg)
    else:
        assert len(np.unique(page)) > 2
        diff = np.abs(img.astype(int) - eimg.astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() <= 1e-4
        # and the anti-aliased result differs from what the two-level kernel would have produced
        two = np.where(page == page[0, 0], page[0, 0], page[page != page[0, 0]][0]).astype(np.uint8)
        timg = opipe.prepare_images(two, two, 6, lh)[0]
        assert (timg != eimg).any()
