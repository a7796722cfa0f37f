"""Known-answer tests pinning the CPU oracle (the reference ships no tests or golden
vectors - SURVEY.md section 8(c) - so the pins are hand-computed here)."""
import cv2
import numpy as np
import pytest
import torch

from oracle import network as onet
from oracle import pipeline as opipe
from oracle import resize as osk
from page_segmentation_b200 import synth


# ---------------- TF/Keras operator semantics (model.py) ----------------
def test_conv_same_is_cross_correlation_no_flip():
    k = np.arange(25, dtype=np.float32).reshape(5, 5, 1, 1)
    x = torch.zeros(1, 1, 9, 9)
    x[0, 0, 4, 4] = 1.0
    y = onet._conv_same(x, torch.from_numpy(k), torch.zeros(1), 5)[0, 0].numpy()
    # y[h,w] = sum x[h+i-2, w+j-2] K[i,j]: a delta at (4,4) paints K flipped around it
    np.testing.assert_array_equal(y[2:7, 2:7], k[::-1, ::-1, 0, 0])
    # and a single output pixel at (4,4) of an arbitrary image is sum(x * K) unflipped
    img = torch.arange(81, dtype=torch.float32).reshape(1, 1, 9, 9)
    out = onet._conv_same(img, torch.from_numpy(k), torch.zeros(1), 5)[0, 0, 4, 4].item()
    assert out == float((img[0, 0, 2:7, 2:7].numpy() * k[:, :, 0, 0]).sum())


def test_conv_same_even_kernel_pads_after():
    # Conv2D(2, 'same'): y[h,w] = sum_{i,j in {0,1}} x[h+i, w+j] K[i,j]  (TF pads bottom/right)
    k = np.array([[1, 2], [3, 4]], dtype=np.float32).reshape(2, 2, 1, 1)
    x = torch.arange(9, dtype=torch.float32).reshape(1, 1, 3, 3)
    y = onet._conv_same(x, torch.from_numpy(k), torch.zeros(1), 2)[0, 0].numpy()
    assert y.shape == (3, 3)
    assert y[0, 0] == 0 * 1 + 1 * 2 + 3 * 3 + 4 * 4
    assert y[2, 2] == 8 * 1                                    # only the top-left tap is inside


def test_conv_transpose_s1_is_gradient_of_conv():
    # Conv2DTranspose(5,'same',s=1), kernel (kh,kw,Cout,Cin): y[h,w,o] = sum x[h-i+2, w-j+2, c] K[i,j,o,c]
    rng = np.random.default_rng(0)
    k = rng.standard_normal((5, 5, 2, 3)).astype(np.float32)      # Cout=2, Cin=3
    x = torch.zeros(1, 3, 7, 7)
    x[0, 1, 3, 3] = 1.0
    y = onet._deconv_same(x, torch.from_numpy(k), torch.zeros(2), 5, 1)[0].numpy()
    # a delta in input channel 1 paints K[:, :, o, 1] UNflipped, centred
    for o in range(2):
        np.testing.assert_allclose(y[o, 1:6, 1:6], k[:, :, o, 1], rtol=0, atol=0)


def test_conv_transpose_2x2_stride2_blocks():
    # y[2h+i, 2w+j, o] = b[o] + sum_c x[h,w,c] K[i,j,o,c]
    k = np.arange(2 * 2 * 1 * 2, dtype=np.float32).reshape(2, 2, 1, 2)
    x = torch.tensor([[[[1.0, 2.0]], [[10.0, 20.0]]]])          # (1, Cin=2, 1, 2)
    y = onet._deconv_same(x, torch.from_numpy(k), torch.tensor([0.5]), 2, 2)[0, 0].numpy()
    assert y.shape == (2, 4)
    for w in range(2):
        for i in range(2):
            for j in range(2):
                exp = 0.5 + x[0, 0, 0, w].item() * k[i, j, 0, 0] + x[0, 1, 0, w].item() * k[i, j, 0, 1]
                assert y[i, 2 * w + j] == exp


@pytest.mark.parametrize("d,exp", [(31, 1), (32, 0), (33, 31), (1169, 15), (827, 5), (64, 0)])
def test_calculate_padding(d, exp):
    assert onet.calculate_padding(d, d)[0] == exp
    assert synth.padded_shape(d, 7)[0] == d + exp


def test_pad_is_bottom_right_and_crop_top_left():
    W = synth.make_weights("fcn", 2, seed=1)
    img = np.zeros((33, 40), np.uint8)
    img[32, 39] = 255                                          # last real pixel
    f = onet.Forward("fcn", W, 2)
    logit, kept = f.logits(img, keep=["conv1"])
    assert logit.shape == (33, 40, 2)
    assert kept["conv1"].shape == (64, 64, 20)                 # padded to multiples of 32


def test_concat_order_decoder_then_encoder():
    """fcn_skip: logits see [deconv5 (20), conv2 (30)] in that order (model.py:85)."""
    W = [(np.zeros_like(k), np.zeros_like(b)) for k, b in synth.make_weights("fcn_skip", 2, seed=0)]
    W[0][0][2, 2, 0, 0] = 1.0                                  # conv1 ch0 = x
    W[1][0][2, 2, 0, 7] = 1.0                                  # conv2 ch7 = conv1 ch0
    W[-1][0][0, 0, 20 + 7, 1] = 1.0                            # logits class 1 reads concat channel 20+7
    img = (np.arange(32 * 32) % 251).astype(np.uint8).reshape(32, 32)
    logit, _ = onet.Forward("fcn_skip", W, 2).logits(img)
    np.testing.assert_allclose(logit[..., 1], img.astype(np.float32) / 255.0, atol=1e-6)
    assert np.all(logit[..., 0] == 0)


def test_argmax_ties_take_lowest_class():
    logit = np.zeros((2, 2, 3), np.float32)
    logit[0, 0] = [1, 1, 0]
    logit[0, 1] = [0, 2, 2]
    _, pred = opipe.softmax_argmax(logit)
    assert pred.dtype == np.int64 and pred[0, 0] == 0 and pred[0, 1] == 1 and pred[1, 1] == 0


def test_parameter_counts_match_survey():
    def count(arch):
        return sum(k.size + b.size for k, b in synth.make_weights(arch, 3, 0))
    assert count("fcn_skip") == 673013
    assert count("fcn") == 602523


# ---------------- skimage resize restatement (dataset.py:114-128) ----------------
def test_rescale_output_shape_half_to_even():
    assert osk.rescale_output_shape((5, 7), 0.5) == (2, 4)      # 2.5 -> 2, 3.5 -> 4
    assert osk.rescale_output_shape((3508, 2480), 6 / 18) == (1169, 827)


def test_nearest_downscale_by_3_picks_the_centre_pixel():
    a = np.arange(9 * 12, dtype=np.float64).reshape(9, 12)
    out = osk.resize(a, (3, 4), order=0)
    np.testing.assert_array_equal(out, a[1::3, 1::3])


def test_nearest_round_half_away_from_zero():
    # 4 -> 2: r = 2*y + 0.5 -> round(0.5) = 1, round(2.5) = 3 (C round(), not banker's)
    a = np.arange(4, dtype=np.float64).reshape(4, 1)
    np.testing.assert_array_equal(osk.resize(a, (2, 1), order=0)[:, 0], [1, 3])


def test_bicubic_reproduces_constants_pixels_and_ramps():
    c = np.full((20, 30), 77.0)
    np.testing.assert_allclose(osk.resize(c, (7, 11), order=3), 77.0, rtol=0, atol=1e-12)
    a = np.random.default_rng(0).integers(0, 255, (12, 12)).astype(np.float64)
    # integer sampling positions (scale 1/3 -> coordinates 3x+1) return the pixel itself
    np.testing.assert_allclose(osk.resize(a, (4, 4), order=3), a[1::3, 1::3], rtol=0, atol=1e-9)
    # the Catmull-Rom cubic reproduces a linear ramp away from the borders
    ramp = np.tile(np.arange(40, dtype=np.float64), (8, 1))
    out = osk.resize(ramp, (8, 16), order=3)
    xs = 2.5 * np.arange(16) + 0.75
    np.testing.assert_allclose(out[0, 1:-1], xs[1:-1], rtol=0, atol=1e-9)


def test_bicubic_clips_to_input_range():
    a = np.zeros((8, 8))
    a[:, 4:] = 255.0
    out = osk.resize(a, (8, 13), order=3)
    assert out.min() >= 0.0 and out.max() <= 255.0             # overshoot removed by clip=True


def test_reflect_is_mirror_without_edge_repeat():
    idx = osk._reflect(np.array([-2, -1, 0, 4, 5, 6]), 5)
    np.testing.assert_array_equal(idx, [2, 1, 0, 4, 3, 2])


def test_antialias_branch_uses_scipy_gaussian():
    from scipy import ndimage as ndi
    a = np.random.default_rng(1).integers(0, 255, (30, 30)).astype(np.float64)
    got = osk.resize(a, (10, 10), order=3, anti_aliasing=True)
    blurred = ndi.gaussian_filter(a, (1.0, 1.0), mode="mirror")
    np.testing.assert_allclose(got, osk.resize(blurred, (10, 10), order=3), rtol=0, atol=1e-12)


def test_prepare_images_semantics():
    page = np.full((36, 36), 255, np.uint8)
    page[9:27, 9:27] = 0                                       # one ink square
    img, b, ob = opipe.prepare_images(page, page, 6, 18, keep_orig_bin=True)
    assert img.shape == (12, 12) and img.dtype == np.uint8 and b.dtype == np.uint8
    assert set(np.unique(b)) == {0, 1} and b[6, 6] == 1 and b[0, 0] == 0     # 1 = ink
    assert img[6, 6] == 255 and img[0, 0] == 0                               # inverted grey
    np.testing.assert_array_equal(ob, (page == 0).astype(np.uint8))


# ---------------- epilogue / post-processing (output.py, postprocess.py) ----------------
def test_generate_output_masks_toy():
    pred = np.array([[0, 1], [2, 1]])
    binary = np.array([[1, 0], [1, 1]], np.uint8)
    lut = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}
    color, overlay, inverted, fg = opipe.generate_output_masks(binary, pred, lut)
    assert color[1, 0].tolist() == [0, 255, 0]
    assert overlay[0, 1].tolist() == [255, 0, 0] and overlay[0, 0].tolist() == [0, 0, 0]     # ink blanked
    assert inverted[0, 1].tolist() == [0, 0, 0] and inverted[1, 1].tolist() == [255, 0, 0]   # paper blanked
    np.testing.assert_array_equal(fg, inverted)                # identity noted in SURVEY 3.1a-7


CC_EXAMPLE = np.array([[1, 1, 0, 0, 0, 1, 0, 0],
                       [0, 1, 0, 1, 0, 1, 0, 1],
                       [0, 0, 0, 1, 1, 1, 0, 1],
                       [1, 0, 0, 0, 0, 0, 0, 0],
                       [1, 1, 0, 1, 0, 0, 1, 1],
                       [0, 0, 0, 1, 0, 0, 1, 0]], np.uint8)


def test_cv2_label_order_is_raster_order_of_first_pixel():
    n, labels, stats, _ = cv2.connectedComponentsWithStats(CC_EXAMPLE, connectivity=4)
    en, elabels, estats = opipe.connected_components_4(CC_EXAMPLE)
    assert n == en == 7
    np.testing.assert_array_equal(labels, elabels)
    np.testing.assert_array_equal(stats, estats)
    assert labels[0, 0] == 1 and labels[0, 5] == 2 and labels[1, 7] == 3 and labels[3, 0] == 4
    assert stats[2].tolist() == [3, 0, 3, 3, 6]                # left, top, width, height, area


def test_cc_restatement_matches_cv2_on_random_images():
    rng = np.random.default_rng(3)
    for _ in range(5):
        a = (rng.random((40, 57)) < 0.45).astype(np.uint8)
        n, labels, stats, _ = cv2.connectedComponentsWithStats(a, connectivity=4)
        en, elabels, estats = opipe.connected_components_4(a)
        assert n == en
        np.testing.assert_array_equal(labels, elabels)
        np.testing.assert_array_equal(stats, estats)


def test_vote_connected_component_class_toy():
    binary = np.array([[1, 1, 0, 1],
                       [1, 0, 0, 1],
                       [0, 0, 0, 1]], np.uint8)
    pred = np.array([[2, 1, 0, 1],
                     [1, 0, 2, 2],
                     [0, 0, 0, 2]], np.int64)
    out = opipe.vote_connected_component_class(pred.copy(), binary)
    # component A = {(0,0),(0,1),(1,0)} votes {2,1,1} -> 1 ; component B = column 3 votes {1,2,2} -> 2
    assert out[0, 0] == 1 and out[0, 1] == 1 and out[1, 0] == 1
    assert out[0, 3] == 2 and out[1, 3] == 2 and out[2, 3] == 2
    assert out[1, 2] == 2 and out[2, 0] == 0                   # background pixels untouched
    tie = opipe.vote_connected_component_class(np.array([[2, 1]]), np.array([[1, 1]], np.uint8))
    assert tie.tolist() == [[1, 1]]                            # tie -> lowest class


def test_add_bounding_boxes_toy():
    pred = np.zeros((5, 6), np.int64)
    pred[0, 0] = 1
    pred[2, 2] = 1                                             # two separate class-1 components
    pred[1, 4] = 2
    pred[3, 5] = 2
    pred[2, 5] = 2                                             # one L-shaped ... (1,4) is separate from (2,5),(3,5)
    out = opipe.add_bounding_boxes(pred)
    assert out[0, 0] == 1 and out[2, 2] == 1 and out[1, 1] == 0
    assert out[1, 4] == 2 and out[2, 5] == 2 and out[3, 5] == 2 and out[2, 4] == 0


def test_bf16_twin_rounding_points():
    W = synth.make_weights("fcn_skip", 3, seed=0)
    page = synth.make_page(1, 96, 96, 18)
    img, _ = opipe.prepare_images(page, page, 6, 18)
    twin = onet.Forward("fcn_skip", W, 3, bf16=True)
    _, kept = twin.logits(img, keep=["conv1", "conv2", "deconv4"])
    for v in kept.values():                                    # every stored activation is bf16-representable
        t = torch.from_numpy(v)
        assert torch.equal(t, t.to(torch.bfloat16).to(torch.float32))
    lt, _ = twin.logits(img)
    l32, _ = onet.Forward("fcn_skip", W, 3).logits(img)
    assert 0 < np.abs(lt - l32).max() < 8e-3


# ---------------------------------------------------------------------------
# compute_char_height (image_ops.py:58-82)
# ---------------------------------------------------------------------------
def _glyph_page(h=300, w=400, sizes=((20, 12), (22, 14), (30, 20), (8, 8), (70, 30), (16, 40)), paper=230, ink=30):
    page = np.full((h, w), paper, np.uint8)
    x = 10
    for gh, gw in sizes:
        page[50:50 + gh, x:x + gw] = ink
        x += gw + 15
    return page


def test_char_height_positional_4_means_8_connectivity():
    """the reference's `cv2.connectedComponentsWithStats(img, 4)` ignores the 4 (it fills the `labels` slot):
    cv2's default 8-connectivity applies -- two diagonal squares are ONE component"""
    import cv2
    img = np.zeros((40, 40), np.uint8)
    img[5:10, 5:10] = 255
    img[10:15, 10:15] = 255           # touches the first square only through a corner
    n_ref = cv2.connectedComponentsWithStats(img, 4)[0]
    assert n_ref == cv2.connectedComponentsWithStats(img, connectivity=8)[0] == 2
    assert cv2.connectedComponentsWithStats(img, connectivity=4)[0] == 3


def test_char_height_known_answer():
    from oracle import image_ops as oio
    page = _glyph_page()
    # letter-like boxes: (20x12), (22x14), (30x20); (8x8) too small, (70x30) too tall, (16x40): w/h = 2.5
    assert oio.compute_char_height_array(page, inverse=False) == 22       # sorted [20, 22, 30][3 // 2]
    # inverse=True analyses the paper instead: one huge component, no letters
    assert oio.compute_char_height_array(page, inverse=True) is None


@pytest.mark.parametrize("down_right", [True, False])
def test_char_height_half_glyphs_touching_through_a_corner(down_right):
    """known answer behind tests/test_gpu_char_height.py::test_char_height_diagonal_contacts_on_tile_borders: two 12 x 7
    halves touching through one corner are ONE letter of 24 rows (8-connectivity); one column apart they are two of 12"""
    from oracle import image_ops as oio
    page = np.full((128, 600), 255, np.uint8)
    xa, xb = (249, 256) if down_right else (256, 249)
    page[20:32, xa:xa + 7] = 0
    page[32:44, xb:xb + 7] = 0
    assert oio.compute_char_height_array(page, inverse=False) == 24
    page[32:44, xb:xb + 7] = 255
    xb += 1 if down_right else -1
    page[32:44, xb:xb + 7] = 0
    assert oio.compute_char_height_array(page, inverse=False) == 12


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_otsu_restatement_matches_cv2(seed):
    import cv2
    from oracle import image_ops as oio
    rng = np.random.default_rng(seed)
    img = np.clip(rng.normal(90, 30, (120, 160)) + (rng.random((120, 160)) < 0.3) * 110, 0, 255).astype(np.uint8)
    assert oio.otsu_threshold(img) == int(cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[0])
