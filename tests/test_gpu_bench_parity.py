"""Parity AT THE BENCH CONFIGURATION (BASELINE configs[1]): the batched forward over many A4 pages, where the CTA
ranges of the marching kernels (conv_fold.cu: `rows_per_cta`) cut across page and strip boundaries, and the
host-buffer batch call (pcs_predict_pages_host) with its chunk schedule -- both against the CPU oracle that restates
network.py:248-260 / model.py:45-92 / dataset.py:131-150 / output.py:44-60, never against another device run.

Tolerances (the ones of test_gpu_network.py):
  * device logits vs the fp32 oracle: max |d| <= 1e-3 (fp16 operands) / 8e-3 (bf16 operands);
  * device class map vs the fp64 oracle: >= 99.9 % (fp16) / 99.7 % (bf16) of the pixels, and every disagreeing
    pixel is a near-tie (fp64 top-2 margin <= 2 x the largest logit error of that page).
"""
import numpy as np
import pytest
import torch

from oracle import network as onet
from oracle import pipeline as opipe
from page_segmentation_b200 import synth

pytestmark = pytest.mark.gpu

TOL = {"bf16": dict(f32_max=8e-3, agree=0.997), "fp16": dict(f32_max=1e-3, agree=0.999)}
LUT = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}
A4_SEEDS = tuple(range(100, 116))          # 16 distinct pages
_cache = {}


def _a4_inputs():
    """16 distinct synthetic A4 pages, preprocessed by the oracle (so this file tests the network, not the resampler)."""
    if "a4" not in _cache:
        imgs, bins = [], []
        for s in A4_SEEDS:
            page = synth.make_page(s)
            img, b = opipe.prepare_images(page, page, 6, 18)
            imgs.append(img)
            bins.append(b)
        _cache["a4"] = (np.stack(imgs), np.stack(bins))
    return _cache["a4"]


def _oracle_logits(arch, W, img, key):
    """fp32 and fp64 oracle logits of one page, computed once per (page, weights)."""
    k = ("logits", arch, key)
    if k not in _cache:
        l32 = onet.Forward(arch, W, 3).logits(img)[0]
        l64 = onet.Forward(arch, W, 3, dtype=torch.float64).logits(img)[0]
        _cache[k] = (l32, l64)
    return _cache[k]


def _check_page(logit, pred, l32, l64, tol, what):
    err32 = np.abs(logit - l32).max()
    assert err32 <= tol["f32_max"], (what, err32)
    ref = l64.argmax(-1)
    agree = (pred == ref).mean()
    assert agree >= tol["agree"], (what, agree)
    bad = pred != ref
    if bad.any():
        s = np.sort(l64, -1)
        margin = s[..., -1] - s[..., -2]
        assert margin[bad].max() <= 2 * np.abs(logit - l64).max(), (what, margin[bad].max())
    np.testing.assert_array_equal(pred, logit.argmax(-1))


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_batched_a4_forward_vs_oracle(ctx, precision):
    """16 A4 pages (16 distinct seeds) in ONE pcs_forward: at this size conv2's ranges are 896 rows long against 1184-row
    strips, so most CTAs start and end inside a strip and many cross a strip or page boundary.  First, middle and last
    page against the oracle; every page's class map must equal the argmax of its own logits."""
    from page_segmentation_b200.lib.network import Network
    imgs, _ = _a4_inputs()
    n, h, w = imgs.shape
    W = synth.make_weights("fcn_skip", 3, seed=0)
    net = Network("Predict", n_classes=3, weights=W, precision=precision)
    c = net._context()
    d_img = torch.from_numpy(imgs).cuda()
    d_labels = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_logits = torch.empty((n, h, w, 3), dtype=torch.float32, device="cuda")
    c.forward(d_img, None, n, h, w, d_labels, d_logits, None)
    torch.cuda.synchronize()
    for i in (0, n // 2 - 1, n // 2, n - 1):
        l32, l64 = _oracle_logits("fcn_skip", W, imgs[i], ("a4", i, 0))
        _check_page(d_logits[i].cpu().numpy(), d_labels[i].cpu().numpy(), l32, l64, TOL[precision], f"page {i}")
    # every page: labels are the first-max argmax of the logits the same launch produced
    ref = d_logits.argmax(-1).to(torch.uint8)
    assert bool((ref == d_labels).all())
    # and the labels-only schedule (what bench.py runs: no logits output) gives the same class maps
    d_labels2 = torch.empty_like(d_labels)
    c.forward(d_img, None, n, h, w, d_labels2)
    assert bool((d_labels2 == d_labels).all())


@pytest.mark.parametrize("arch", ["fcn_skip", "fcn"])
@pytest.mark.parametrize("n,h,w", [(3, 72, 130), (5, 150, 480), (7, 200, 300), (2, 330, 250)])
def test_cta_ranges_crossing_pages_and_strips(ctx, arch, n, h, w):
    """Shapes whose `rows_per_cta` does NOT divide the layer height, so that CTA ranges of the marching kernels begin
    and end in the middle of strips and run across strip and page boundaries at several layers:
      (3,72,130):  padded 96x160; conv5/conv6/deconv3 level h = 24 with 16-row ranges;
      (5,150,480): padded 160x480, 4 strips: conv2 ranges of 24 rows against h = 160;
      (7,200,300): padded 224x320, 3 strips: conv2 ranges of 32 rows, conv3 ranges of 16 rows against h = 112;
      (2,330,250): padded 352x256, 3 strips of which the last has 8 pixels.
    Every page against the fp32 / fp64 oracle."""
    from page_segmentation_b200.lib.network import Network
    imgs = np.stack([opipe.prepare_images(p, p, 6, 18)[0] for p in
                     (synth.make_page(200 + s, h * 3, w * 3, 18) for s in range(n))])
    assert imgs.shape == (n, h, w)
    W = synth.make_weights(arch, 3, seed=4)
    from page_segmentation_b200.lib.architecture import Architecture
    net = Network("Predict", n_classes=3, model_constructor=Architecture(arch), weights=W, precision="fp16")
    c = net._context()
    d_img = torch.from_numpy(imgs).cuda()
    d_labels = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_logits = torch.empty((n, h, w, 3), dtype=torch.float32, device="cuda")
    c.forward(d_img, None, n, h, w, d_labels, d_logits, None)
    torch.cuda.synchronize()
    for i in range(n):
        l32, l64 = _oracle_logits(arch, W, imgs[i], ("small", n, h, w, i))
        _check_page(d_logits[i].cpu().numpy(), d_labels[i].cpu().numpy(), l32, l64, TOL["fp16"], f"page {i}")


@pytest.mark.parametrize("cc", [False, True])
def test_host_batch_call_a4_vs_oracle(ctx, cc):
    """pcs_predict_pages_host on 16 A4 pages with the default chunk schedule (2, 4, 8, 2): preprocess outputs bit-exact
    against the oracle, class maps against the fp64 oracle (and, with cc_majority, against the oracle's vote over the
    device's raw class map), colour masks bit-exact against the oracle's generate_output_masks."""
    from page_segmentation_b200.runtime import PageBatchEngine
    n = 16
    pages = np.stack([synth.make_page(s) for s in A4_SEEDS[:n]])
    W = synth.make_weights("fcn_skip", 3, seed=0)
    lut = np.array([LUT[i] for i in range(3)], dtype=np.uint8)
    eng = PageBatchEngine("fcn_skip", W, 3, precision="fp16", lut=lut)
    Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 6 / 18)
    h_pages = torch.from_numpy(pages).pin_memory().numpy()
    out = {k: torch.empty((n, Hs, Ws) + ((3,) if k in ("color", "overlay", "inverted") else ()), dtype=torch.uint8).pin_memory().numpy()
           for k in ("image", "binary", "labels", "color", "overlay", "inverted")}
    eng.ctx.predict_pages_host(h_pages, h_pages, n, synth.A4_H, synth.A4_W, Hs, Ws, cc, lut,
                               out["image"], out["binary"], out["labels"], out["color"], out["overlay"], out["inverted"])
    imgs, bins = _a4_inputs()
    np.testing.assert_array_equal(out["image"], imgs[:n])
    np.testing.assert_array_equal(out["binary"], bins[:n])
    raw = None
    if cc:      # the raw class maps of the same pages (labels-only call), voted by the ORACLE
        raw = {k: torch.empty((n, Hs, Ws), dtype=torch.uint8).pin_memory().numpy() for k in ("labels",)}
        eng.ctx.predict_pages_host(h_pages, h_pages, n, synth.A4_H, synth.A4_W, Hs, Ws, False, None, None, None, raw["labels"])
    for i in (0, 1, 2, 7, 13, 15):          # first chunk, second chunk, a steady chunk, the tail chunk
        l32, l64 = _oracle_logits("fcn_skip", W, imgs[i], ("a4", i, 0))
        ref = l64.argmax(-1)
        if cc:
            assert (raw["labels"][i] == ref).mean() >= TOL["fp16"]["agree"], i
            exp = opipe.vote_connected_component_class(raw["labels"][i].astype(np.int64), bins[i])
            np.testing.assert_array_equal(out["labels"][i], exp)
        else:
            assert (out["labels"][i] == ref).mean() >= TOL["fp16"]["agree"], i
        col, ov, inv, _ = opipe.generate_output_masks(bins[i], out["labels"][i].astype(np.int64), LUT)
        np.testing.assert_array_equal(out["color"][i], col)
        np.testing.assert_array_equal(out["overlay"][i], ov)
        np.testing.assert_array_equal(out["inverted"][i], inv)


def test_two_owners_of_one_context_do_not_share_weights(ctx):
    """The per-device context holds ONE model: a Network, then a PageBatchEngine with other weights, then the
    Network again must each run their own weights."""
    from page_segmentation_b200.lib.dataset import SingleData
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.runtime import PageBatchEngine
    page = synth.make_page(5, 300, 240, 18)
    img, _ = opipe.prepare_images(page, page, 6, 18)
    Wa, Wb = synth.make_weights("fcn_skip", 3, seed=11), synth.make_weights("fcn_skip", 3, seed=12)
    net = Network("Predict", n_classes=3, weights=Wa, precision="fp16")
    first = net.predict_single_data(SingleData(image=img))[0]
    eng = PageBatchEngine("fcn_skip", Wb, 3, precision="fp16")
    d = {k: v.clone() for k, v in eng.run_device(torch.from_numpy(page[None]).cuda(), 6 / 18, masks=False).items()}
    lb = onet.Forward("fcn_skip", Wb, 3).logits(img)[0]
    assert (d["labels"][0].cpu().numpy() == lb.argmax(-1)).mean() >= 0.999
    again = net.predict_single_data(SingleData(image=img))[0]
    np.testing.assert_array_equal(first, again)
    la = onet.Forward("fcn_skip", Wa, 3).logits(img)[0]
    assert np.abs(again - la).max() <= 1e-3
    d2 = eng.run_device(torch.from_numpy(page[None]).cuda(), 6 / 18, masks=False)      # and the engine reloads its own
    assert bool((d2["labels"] == d["labels"]).all())


UNET_TOL = {"fp16": dict(f32_max=8e-3, agree=0.999), "bf16": dict(f32_max=6e-2, agree=0.998)}


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_unet_a4_vs_oracle(ctx, precision):
    """U-Net (model.py:151-203) at the size BASELINE configs[2] names: two A4 pages in one pcs_forward, the first against
    the fp32 / fp64 oracle.  Stated tolerances: logits max |d| <= 8e-3 (fp16 operands; 5.1e-3 measured) / 6e-2 (bf16 operands; 4.0e-2 measured) on logits
    within +-3; class map >= 99.9 % of the fp64 oracle's with fp16 operands (99.978 % measured), >= 99.8 % with bf16 operands (99.878 % measured; He-normal weights give a median top-2
    margin of 0.5), every disagreeing pixel a near-tie (fp64 margin <= 2 x the page's largest logit error)."""
    from page_segmentation_b200.lib.architecture import Architecture
    from page_segmentation_b200.lib.network import Network
    imgs, _ = _a4_inputs()
    imgs = imgs[:2]
    n, h, w = imgs.shape
    W = synth.make_weights("unet", 3, seed=0)
    net = Network("Predict", n_classes=3, model_constructor=Architecture("unet"), weights=W, precision=precision)
    c = net._context()
    d_img = torch.from_numpy(imgs).cuda()
    d_labels = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_logits = torch.empty((n, h, w, 3), dtype=torch.float32, device="cuda")
    c.forward(d_img, None, n, h, w, d_labels, d_logits, None)
    torch.cuda.synchronize()
    l32, l64 = _oracle_logits("unet", W, imgs[0], ("a4", 0, 0))
    logit, pred = d_logits[0].cpu().numpy(), d_labels[0].cpu().numpy()
    print(f"unet {precision}: max |logit - fp32| = {np.abs(logit - l32).max():.3e}, agreement with fp64 = {(pred == l64.argmax(-1)).mean():.6f}")
    _check_page(logit, pred, l32, l64, UNET_TOL[precision], f"unet {precision}")
    assert bool((d_logits.argmax(-1).to(torch.uint8) == d_labels).all())
    # the same page alone gives the same class map (batch independence at A4)
    d_one = torch.empty((1, h, w), dtype=torch.uint8, device="cuda")
    c.forward(d_img[1:2].contiguous(), None, 1, h, w, d_one)
    assert bool((d_one[0] == d_labels[1]).all())


def test_stress_variant_full_resolution_page(ctx):
    """SURVEY section 8(d)'s stress variant: `line_height_px = 6`, i.e. scale 1 -- the network runs on the 3508x2480 page
    itself (padded grid 3520x2496: 21 strips, 8.9x the pixels of a normalised page, planes of > 2^27 elements).  The largest
    single-page size the path is specified for: device preprocess bit-exact against the oracle (prepare_images,
    dataset.py:131-150, at scale 1), device logits within the fp16-operand tolerance of the fp32 oracle
    (network.py:248-260), class map = first-max argmax of the device logits, and every pixel whose class differs from the
    fp32 oracle's a near-tie (fp32 top-2 margin <= 2 x the page's largest logit error).  (No fp64 run at this size: it costs
    a minute of CPU; the fp64 near-tie analysis is done at A4 above.)"""
    from page_segmentation_b200.lib.network import Network
    page = synth.make_page(7)
    H, W_ = page.shape
    img, binary = opipe.prepare_images(page, page, 6, 6)
    assert img.shape == (H, W_)
    W = synth.make_weights("fcn_skip", 3, seed=0)
    net = Network("Predict", n_classes=3, weights=W, precision="fp16")
    c = net._context()
    d_page = torch.from_numpy(page[None]).cuda()
    d_img = torch.empty((1, H, W_), dtype=torch.uint8, device="cuda")
    d_bin = torch.empty((1, H, W_), dtype=torch.uint8, device="cuda")
    c.preprocess(d_page, d_page, 1, H, W_, H, W_, d_img, d_bin)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_img[0].cpu().numpy(), img)
    np.testing.assert_array_equal(d_bin[0].cpu().numpy(), binary)
    d_labels = torch.empty((1, H, W_), dtype=torch.uint8, device="cuda")
    d_logits = torch.empty((1, H, W_, 3), dtype=torch.float32, device="cuda")
    c.forward(d_img, None, 1, H, W_, d_labels, d_logits, None)
    torch.cuda.synchronize()
    logit, pred = d_logits[0].cpu().numpy(), d_labels[0].cpu().numpy()
    l32 = onet.Forward("fcn_skip", W, 3).logits(img)[0]
    err = np.abs(logit - l32).max()
    print(f"stress variant: max |logit - fp32| = {err:.3e}, agreement with the fp32 oracle = {(pred == l32.argmax(-1)).mean():.6f}")
    assert err <= TOL["fp16"]["f32_max"], err
    np.testing.assert_array_equal(pred, logit.argmax(-1))
    ref = l32.argmax(-1)
    assert (pred == ref).mean() >= TOL["fp16"]["agree"]
    bad = pred != ref
    if bad.any():
        s = np.sort(l32, -1)
        assert (s[..., -1] - s[..., -2])[bad].max() <= 2 * err
    # labels-only schedule (the bench's) on the same page
    d_labels2 = torch.empty_like(d_labels)
    c.forward(d_img, None, 1, H, W_, d_labels2)
    assert bool((d_labels2 == d_labels).all())
