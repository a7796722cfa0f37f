"""Parity of the device network (pcs_forward) with the CPU oracle restating
model.py / network.py:248-260.

Tolerances (stated, see DESIGN.md "precision"):
  * device vs numerics twin (same bf16|fp16 rounding points, fp32 accumulate):
    logits max |d| <= 2e-3 (bf16) / 3e-4 (fp16), mean |d| <= 6e-5 / 1e-5 -- only
    accumulation order and rare 1-ulp operand flips differ;
  * device vs fp32 oracle: logits max |d| <= 8e-3 (bf16) / 1e-3 (fp16);
  * argmax vs the fp64 oracle: agreement >= 99.9 % for fp16 operands, >= 99.7 % for
    bf16 operands, and every disagreeing pixel is a near-tie: fp64 top-2 margin
    <= 2 * max logit error.
"""
import numpy as np
import pytest
import torch

from oracle import network as onet
from oracle import pipeline as opipe
from page_segmentation_b200 import synth

pytestmark = pytest.mark.gpu

TOL = {"bf16": dict(twin_max=2e-3, twin_mean=6e-5, f32_max=8e-3, agree=0.997),
       "fp16": dict(twin_max=3e-4, twin_mean=1e-5, f32_max=1e-3, agree=0.999)}


def _device_predict(arch, weights, n_classes, image, precision, engine, keep=False):
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.architecture import Architecture
    from page_segmentation_b200.lib.dataset import SingleData
    net = Network("Predict", n_classes=n_classes, model_constructor=Architecture(arch), weights=weights,
                  precision=precision)
    net.engine = engine
    # keep=True: also store the activations the fused kernels never write (conv2 of fcn_skip)
    net._context().set_keep_activations(keep)
    try:
        return net, net.predict_single_data(SingleData(image=image))
    finally:
        net._context().set_keep_activations(False)


def _small_input(seed, h, w):
    page = synth.make_page(seed, h * 3, w * 3, 18)
    img, b = opipe.prepare_images(page, page, 6, 18)
    return img, b


@pytest.mark.parametrize("engine", ["direct", "umma"])
@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("arch,hw", [("fcn_skip", (150, 203)), ("fcn", (97, 64)), ("fcn_skip", (32, 32))])
def test_fcn_logits_and_argmax(ctx, arch, hw, precision, engine):
    img, _ = _small_input(1, *hw)
    W = synth.make_weights(arch, 3, seed=2)
    net, (logit, prob, pred) = _device_predict(arch, W, 3, img, precision, engine)
    assert logit.shape == img.shape + (3,) and logit.dtype == np.float32
    assert prob.shape == logit.shape and prob.dtype == np.float32
    assert pred.shape == img.shape and pred.dtype == np.int64
    tol = TOL[precision]

    fused = engine == "umma"          # the tensor engine composes deconv5 with the logits (fp32-grade weights)
    twin = onet.Forward(arch, W, 3, bf16=True, fused_head=fused)
    if precision == "fp16":
        onet._bf16, saved = (lambda t: t.to(torch.float16).to(t.dtype)), onet._bf16
        try:
            twin = onet.Forward(arch, W, 3, bf16=True, fused_head=fused, conv1_rounded=fused)   # fp16 tensor engine: one operand
            lt, _ = twin.logits(img)
        finally:
            onet._bf16 = saved
    else:
        lt, _ = twin.logits(img)
    d = np.abs(logit - lt)
    assert d.max() <= tol["twin_max"], d.max()
    assert d.mean() <= tol["twin_mean"], d.mean()

    l32, _ = onet.Forward(arch, W, 3).logits(img)
    assert np.abs(logit - l32).max() <= tol["f32_max"]

    l64, _ = onet.Forward(arch, W, 3, dtype=torch.float64).logits(img)
    ref = l64.argmax(-1)
    agree = (pred == ref).mean()
    assert agree >= tol["agree"], agree
    s = np.sort(l64, -1)
    margin = s[..., -1] - s[..., -2]
    bad = pred != ref
    if bad.any():
        assert margin[bad].max() <= 2 * np.abs(logit - l64).max()

    # softmax / argmax are exact functions of the returned logits
    np.testing.assert_array_equal(pred, logit.argmax(-1))
    eprob, _ = opipe.softmax_argmax(logit)
    np.testing.assert_allclose(prob, eprob, rtol=0, atol=2e-6)


@pytest.mark.parametrize("engine", ["direct", "umma"])
def test_fcn_skip_layerwise_vs_twin(ctx, engine):
    """Every stored activation against the twin: localises layout / weight-transform errors."""
    img, _ = _small_input(3, 96, 128)
    W = synth.make_weights("fcn_skip", 3, seed=5)
    net, (logit, _, _) = _device_predict("fcn_skip", W, 3, img, "bf16", engine, keep=True)
    c = net._context()
    names = ["conv1", "conv2", "conv3", "conv5", "conv6", "conv7", "deconv1", "deconv2", "deconv3", "deconv4"]
    twin = onet.Forward("fcn_skip", W, 3, bf16=True)
    _, kept = twin.logits(img, keep=names)
    for nme in names:
        got = c.debug_activation(nme)[0]
        exp = kept[nme]
        assert got.shape == exp.shape, (nme, got.shape, exp.shape)
        d = np.abs(got - exp)
        scale = max(1e-3, np.abs(exp).max())
        assert d.max() <= 2 ** -7 * scale + 1e-6, (nme, d.max(), scale)      # <= ~1 bf16 ulp of the range
        assert d.mean() <= 4e-4 * scale + 1e-7, (nme, d.mean())


@pytest.mark.parametrize("n_classes,engine", [(2, "umma"), (4, "umma"), (5, "umma"), (2, "direct"), (5, "direct"), (8, "direct")])
def test_n_classes(ctx, n_classes, engine):
    """<= 4 classes run the fused tensor-core head, more fall back to the CUDA-core head kernel."""
    img, _ = _small_input(7, 64, 96)
    W = synth.make_weights("fcn_skip", n_classes, seed=3)
    _, (logit, prob, pred) = _device_predict("fcn_skip", W, n_classes, img, "fp16", engine)
    lt, _ = onet.Forward("fcn_skip", W, n_classes).logits(img)
    assert np.abs(logit - lt).max() <= 1e-3
    np.testing.assert_array_equal(pred, logit.argmax(-1))
    np.testing.assert_allclose(prob.sum(-1), 1.0, atol=1e-5)


@pytest.mark.parametrize("engine,hw", [("direct", (32, 64)), ("umma", (32, 64)), ("umma", (70, 150))])
def test_unet_small(ctx, engine, hw):
    """U-Net (model.py:151-203); on the tensor engine the four UpSampling2D + Conv2D(2x2) blocks run as 2x2
    convolutions on the low-resolution grid with pre-summed weights (rounded once, so a little closer to the
    fp32 graph than the twin, which rounds every tap)"""
    img, _ = _small_input(2, *hw)
    W = synth.make_weights("unet", 3, seed=1)
    _, (logit, prob, pred) = _device_predict("unet", W, 3, img, "fp16", engine)
    onet._bf16, saved = (lambda t: t.to(torch.float16).to(t.dtype)), onet._bf16
    try:
        lt, _ = onet.Forward("unet", W, 3, bf16=True, conv1_rounded=engine == "umma").logits(img)
    finally:
        onet._bf16 = saved
    scale = np.abs(lt).max()
    assert np.abs(logit - lt).max() <= 2e-3 * max(1.0, scale)
    l32, _ = onet.Forward("unet", W, 3).logits(img)
    assert (pred == l32.argmax(-1)).mean() >= 0.995


def test_batch_equals_single(ctx):
    imgs = [_small_input(s, 64, 96)[0] for s in range(3)]
    W = synth.make_weights("fcn_skip", 3, seed=2)
    from page_segmentation_b200.lib.network import Network
    net = Network("Predict", n_classes=3, weights=W, precision="bf16")
    d = torch.from_numpy(np.stack(imgs)).cuda()
    out = torch.empty_like(d)
    net.predict_labels_device(d, out)
    from page_segmentation_b200.lib.dataset import SingleData
    for i, im in enumerate(imgs):
        _, _, pred = net.predict_single_data(SingleData(image=im))
        np.testing.assert_array_equal(out[i].cpu().numpy(), pred)


@pytest.mark.parametrize("tap", [(2, 2), (0, 0), (4, 4), (1, 3), (3, 0)])
@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_umma_one_hot_tap_is_a_shifted_copy(ctx, tap, precision):
    """Sharp layout check of the tcgen05 path: with one-hot conv2 weights at a single tap,
    conv2 must be a bit-exact shifted copy of conv1 (1.0 * x accumulates exactly)."""
    ty, tx = tap
    img, _ = _small_input(4, 70, 300)
    W = [(np.zeros_like(k), np.zeros_like(b)) for k, b in synth.make_weights("fcn_skip", 3, seed=0)]
    for c in range(20):
        W[0][0][2, 2, 0, c] = (c + 1) / 32.0
        W[1][0][ty, tx, c, c] = 1.0
    net, _ = _device_predict("fcn_skip", W, 3, img, precision, "umma", keep=True)
    c = net._context()
    conv1 = c.debug_activation("conv1")[0]
    conv2 = c.debug_activation("conv2")[0]
    pool2 = c.debug_activation("pool2")[0]
    H, Wd, _ = conv1.shape
    exp = np.zeros((H, Wd, 30), np.float32)
    ys, xs = np.arange(H)[:, None] + ty - 2, np.arange(Wd)[None, :] + tx - 2
    ok = (ys >= 0) & (ys < H) & (xs >= 0) & (xs < Wd)
    src = conv1[np.clip(ys, 0, H - 1), np.clip(xs, 0, Wd - 1), :]
    exp[..., :20] = np.where(ok[..., None], src, 0.0)
    assert conv1.max() > 0
    # conv1 itself in closed form (guards against a stale model on the device)
    x32 = (np.pad(img, ((0, H - img.shape[0]), (0, Wd - img.shape[1]))).astype(np.float64) / 255.0).astype(np.float32)
    c1 = torch.from_numpy(x32[..., None] * (np.arange(1, 21, dtype=np.float32) / np.float32(32.0)))
    c1 = c1.to(torch.bfloat16 if precision == "bf16" else torch.float16).to(torch.float32).numpy()
    np.testing.assert_array_equal(conv1, c1)
    bad = (conv2 != exp).any(-1)
    if bad.any():       # tile map of the mismatches (8 rows x 124 px tiles) for diagnosis
        lines = []
        for rb in range(H // 8):
            lines.append(" ".join(f"{int(bad[rb * 8:(rb + 1) * 8, st * 124:(st + 1) * 124].sum()):4d}"
                                  for st in range((Wd + 123) // 124)))
        yy, xx = np.nonzero(bad)
        print("mismatch tile map:\n" + "\n".join(lines))
    np.testing.assert_array_equal(conv2, exp)
    np.testing.assert_array_equal(pool2, exp.reshape(H // 2, 2, Wd // 2, 2, 30).max(axis=(1, 3)))


@pytest.mark.parametrize("tap", [(2, 0), (2, 1), (2, 2), (2, 3), (2, 4), (0, 4), (4, 0), (1, 3)])
@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_pixel_pair_handoff_one_hot_tap(ctx, tap, precision):
    """The default tensor path hands conv1's channels 16..19 to conv2 as pixel-pair units (w + 1 units per row, two taps per
    K half, 7 MMAs per input row): with one-hot conv2 weights pool2 must be the bit-exact 2x2 maximum of a shifted copy of
    conv1 (closed form), for every horizontal tap, across strip borders and at both page edges."""
    ty, tx = tap
    img, _ = _small_input(4, 70, 300)
    W = [(np.zeros_like(k), np.zeros_like(b)) for k, b in synth.make_weights("fcn_skip", 3, seed=0)]
    for c in range(20):
        W[0][0][2, 2, 0, c] = (c + 1) / 32.0
        W[1][0][ty, tx, c, c] = 1.0
    net, _ = _device_predict("fcn_skip", W, 3, img, precision, "umma", keep=False)
    pool2 = net._context().debug_activation("pool2")[0]
    H, Wd = pool2.shape[0] * 2, pool2.shape[1] * 2
    x32 = (np.pad(img, ((0, H - img.shape[0]), (0, Wd - img.shape[1]))).astype(np.float64) / 255.0).astype(np.float32)
    c1 = torch.from_numpy(x32[..., None] * (np.arange(1, 21, dtype=np.float32) / np.float32(32.0)))
    c1 = c1.to(torch.bfloat16 if precision == "bf16" else torch.float16).to(torch.float32).numpy()
    assert c1.max() > 0
    exp = np.zeros((H, Wd, 30), np.float32)
    ys, xs = np.arange(H)[:, None] + ty - 2, np.arange(Wd)[None, :] + tx - 2
    ok = (ys >= 0) & (ys < H) & (xs >= 0) & (xs < Wd)
    exp[..., :20] = np.where(ok[..., None], c1[np.clip(ys, 0, H - 1), np.clip(xs, 0, Wd - 1), :], 0.0)
    np.testing.assert_array_equal(pool2, exp.reshape(H // 2, 2, Wd // 2, 2, 30).max(axis=(1, 3)))
