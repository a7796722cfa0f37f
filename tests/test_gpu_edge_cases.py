"""Edge cases of the device path: degenerate page sizes, ragged datasets, empty foregrounds and the error
behaviour of the C ABI (status codes + messages instead of exceptions or crashes)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import network as onet
from oracle import pipeline as opipe
from page_segmentation_b200 import synth

pytestmark = pytest.mark.gpu
LUT = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}


def _twin_fp16(arch, W, n_classes, img):
    saved = onet._bf16
    onet._bf16 = lambda t: t.to(torch.float16).to(t.dtype)
    try:
        return onet.Forward(arch, W, n_classes, bf16=True, fused_head=True, conv1_rounded=True).logits(img)[0]
    finally:
        onet._bf16 = saved


@pytest.mark.parametrize("hw", [(1, 1), (1, 40), (37, 1), (5, 7), (31, 33), (32, 33), (129, 125)])
def test_degenerate_page_sizes(ctx, hw):
    """Pages smaller than one 32-pixel padding unit, one pixel wide / high, and one past a strip boundary."""
    from page_segmentation_b200.lib.dataset import SingleData
    from page_segmentation_b200.lib.network import Network
    rng = np.random.default_rng(hw[0] * 100 + hw[1])
    img = rng.integers(0, 256, hw, dtype=np.uint8)
    W = synth.make_weights("fcn_skip", 3, seed=3)
    logit, prob, pred = Network("Predict", n_classes=3, weights=W, precision="fp16").predict_single_data(SingleData(image=img))
    assert logit.shape == hw + (3,) and prob.shape == hw + (3,) and pred.shape == hw and pred.dtype == np.int64
    exp = _twin_fp16("fcn_skip", W, 3, img)
    assert np.abs(logit - exp).max() <= 3e-4
    np.testing.assert_allclose(prob.sum(-1), 1.0, atol=1e-5)
    l64 = onet.Forward("fcn_skip", W, 3, dtype=torch.float64).logits(img)[0]
    bad = pred != l64.argmax(-1)
    margin = np.sort(l64, -1)[..., -1] - np.sort(l64, -1)[..., -2]
    assert (margin[bad] <= 2e-3).all()                      # only near-ties may differ


def test_ragged_dataset_through_predictor(ctx):
    """Pages of different sizes and line heights in one dataset: every page equals its stand-alone result."""
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import DatasetLoader, SingleData
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    specs = [(0, 300, 240, 18), (1, 517, 333, 11), (2, 96, 700, 6), (3, 301, 203, 23)]
    loader = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True)
    entries = [SingleData(image=synth.make_page(s, h, w, lh), line_height_px=lh) for s, h, w, lh in specs]
    ds = loader.load_data(entries)
    W = synth.make_weights("fcn_skip", 3, seed=2)
    net = Network("Predict", n_classes=3, weights=W, precision="fp16")
    preds = list(Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP), network=net).predict(ds))
    assert len(preds) == len(specs)
    for (s, h, w, lh), p in zip(specs, preds):
        page = synth.make_page(s, h, w, lh)
        eimg, eb = opipe.prepare_images(page, page, 6, lh)
        np.testing.assert_array_equal(p.data.image, eimg)
        np.testing.assert_array_equal(p.data.binary, eb)
        assert p.labels.shape == eimg.shape
        alone = Network("Predict", n_classes=3, weights=W, precision="fp16").predict_single_data(SingleData(image=eimg))[2]
        np.testing.assert_array_equal(p.labels, alone)


def test_empty_and_full_foreground(ctx):
    """No ink at all / all ink: cc_majority, bounding boxes, masks and compute_char_height stay well defined."""
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import SingleData
    from page_segmentation_b200.lib.image_ops import compute_char_height_array
    from page_segmentation_b200.lib.output import generate_output_masks
    from page_segmentation_b200.lib.postprocess import add_bounding_boxes, vote_connected_component_class
    from page_segmentation_b200.runtime import connected_components_with_stats
    rng = np.random.default_rng(0)
    pred = rng.integers(0, 3, (45, 67)).astype(np.int64)
    for fill in (0, 1):
        binary = np.full(pred.shape, fill, np.uint8)
        data = SingleData(binary=binary)
        exp = opipe.vote_connected_component_class(pred.copy(), binary)
        np.testing.assert_array_equal(vote_connected_component_class(pred.copy(), data), exp)
        m = generate_output_masks(data, pred, DEFAULT_COLOR_MAP)
        c, o, i, f = opipe.generate_output_masks(binary, pred, LUT)
        np.testing.assert_array_equal(m.overlay, o)
        np.testing.assert_array_equal(m.inverted_overlay, i)
        n, labels, stats = connected_components_with_stats(binary)
        assert n == 1 + fill and labels.max() == fill
    flat = np.zeros((30, 30), np.int64)
    np.testing.assert_array_equal(add_bounding_boxes(flat, SingleData(binary=flat.astype(np.uint8))), opipe.add_bounding_boxes(flat))
    assert compute_char_height_array(np.full((64, 64), 200, np.uint8), False) is None       # one grey level: no letters


def test_c_abi_reports_errors_instead_of_crashing(ctx):
    """Every entry point returns a negative pcs_status with a message; nothing throws across the ABI."""
    from page_segmentation_b200 import _native
    lib = _native.load()
    fresh = _native.Context(ctx.device)                      # a context without a model
    try:
        d = torch.zeros((8, 8), dtype=torch.uint8, device="cuda")
        rc = lib.pcs_forward(fresh.h, d.data_ptr(), None, 1, 8, 8, d.data_ptr(), None, None, None, None, None, None)
        assert rc == -3 and b"model" in lib.pcs_last_error(fresh.h).lower()                  # PCS_ERR_STATE
        assert lib.pcs_preprocess(fresh.h, None, None, 1, 8, 8, 3, 3, None, None, None) == -1         # PCS_ERR_ARG
        assert lib.pcs_preprocess(fresh.h, d.data_ptr(), d.data_ptr(), 1, 8, 8, 0, 3, d.data_ptr(), d.data_ptr(), None) == -1
        assert lib.pcs_ccl(fresh.h, d.data_ptr(), 1, 0, 8, d.data_ptr(), None, 0, None) == -1
        assert lib.pcs_cc_majority(fresh.h, d.data_ptr(), d.data_ptr(), 1, 8, 8, 0) == -1
        assert lib.pcs_masks(fresh.h, d.data_ptr(), d.data_ptr(), 1, 8, 8, None, 3, d.data_ptr(), None, None) < 0
        assert lib.pcs_text_regions(fresh.h, d.data_ptr(), 8, 8, np.zeros(3, np.uint8).ctypes.data, 0, 1, 1, None, None) == -1
        assert lib.pcs_segment_masks(fresh.h, d.data_ptr(), 8, 8, 4, 4, np.zeros(27, np.uint8).ctypes.data, 9, d.data_ptr()) == -1
        assert lib.pcs_integral_image(fresh.h, None, 1, 8, 8, None) == -1
        assert len(lib.pcs_last_error(fresh.h)) > 0
        bad = (_native.LayerWeights * 1)()
        assert lib.pcs_model_load(fresh.h, 0, 3, 0, bad, 1) < 0                               # wrong layer count / null tensors
        assert lib.pcs_model_load(fresh.h, 7, 3, 0, bad, 1) < 0                               # unknown architecture
        assert lib.pcs_forward(None, None, None, 0, 0, 0, None, None, None, None, None, None, None) == -1
    finally:
        del fresh
    # the Python layer turns the status into an exception with the library's message
    with pytest.raises(_native.PcsError):
        ctx.ccl(torch.zeros((4, 4), dtype=torch.uint8, device="cuda"), 1, 0, 4, torch.zeros((4, 4), dtype=torch.int32, device="cuda"), None, 0, None)


def test_fp16_saturation_is_counted_and_reported(ctx):
    """fp16 stores saturate at +-65504 (umma_ptx.cuh); the first forward after a model load scans the stored activations
    and the host side warns and points at bf16.  A sane model counts zero, and bf16 (no saturation bound in reach) is
    never scanned."""
    import warnings
    from page_segmentation_b200.lib.dataset import SingleData
    from page_segmentation_b200.lib.network import Network
    page = synth.make_page(4, 96, 128, 6)
    data = SingleData(image=255 - page, binary=(page == 0).astype(np.uint8))
    W = synth.make_weights("fcn_skip", 3, seed=2)
    with warnings.catch_warnings():
        warnings.simplefilter("error")                                   # a sane model must not warn
        Network("Predict", n_classes=3, weights=W, precision="fp16").predict_single_data(data)
    assert ctx.saturation_count() == 0
    hot = [(k.copy(), b.copy()) for k, b in W]
    hot[2] = (hot[2][0] * 3e5, hot[2][1])                                # conv3: activations far beyond 65504
    with pytest.warns(RuntimeWarning, match="saturated"):
        Network("Predict", n_classes=3, weights=hot, precision="fp16").predict_single_data(data)
    assert ctx.saturation_count() > 0
    n_first = ctx.saturation_count()
    net = Network("Predict", n_classes=3, weights=hot, precision="fp16")
    with pytest.warns(RuntimeWarning):
        net.predict_single_data(data)
    with warnings.catch_warnings():
        warnings.simplefilter("error")                                   # reported once per model load, scanned once
        net.predict_single_data(data)
    assert ctx.saturation_count() == n_first
    ctx.set_saturation_check(2)
    try:
        net.predict_single_data(data)
        assert ctx.saturation_count() == 2 * n_first                      # mode 2: every forward
    finally:
        ctx.set_saturation_check(1)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        Network("Predict", n_classes=3, weights=hot, precision="bf16").predict_single_data(data)
    assert ctx.saturation_count() == 0
