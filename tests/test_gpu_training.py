"""Parity of the device training step (csrc/train.cu, csrc/train_tc.cu + lib/trainer.py) with the CPU oracle
(oracle/train.py: torch autograd on the restated graph, Keras Adam with per-variable clipnorm restated in numpy).

Tolerances of the fp32 engine (fp32 on both sides, different summation orders):
  loss                 |d| <= 2e-5
  every gradient       max |d| <= 2e-4 * max |g_oracle| + 1e-7   per variable
  parameters after 3 Adam steps   max |d| <= 4e-5  (lr 1e-3: every coordinate moves ~1e-3 per step at the start of Adam,
                                  whatever the size of its gradient, so m / sqrt(v) amplifies the relative gradient
                                  error of near-zero gradients; 4e-5 is 1.3 % of the distance travelled)

Tolerances of the tensor engine (bf16 activations and activation gradients, 8 significand bits, fp32 accumulation;
the error of a gradient tensor is dominated by the rounding of its operands, so it is measured in the L2 norm):
  loss                 |d| <= 5e-3
  every gradient       ||g - g_oracle||_2 <= 0.25 * ||g_oracle||_2   per variable on pages of a few thousand pixels
                       (fcn without skip connections: conv1's gradient crosses all 24 roundings of the chain and is
                       summed over 40 x 72 pixels only; measured 0.18), <= 0.08 at 256 x 384, <= 0.05 on an A4 page
The rounding noise of a gradient tensor averages out over the pixels it is summed over: measured (tools/check_train_tc.py)
0.3 % .. 7.6 % per variable at 64 x 96 (conv7 sees 8 x 12 pixels), 0.3 % .. 4.1 % at 256 x 384, 0.2 % .. 2.8 % at
1169 x 827; weight and bias gradients of a layer carry the same error, i.e. it comes from the bf16 activation-gradient
chain, not from the weight-gradient kernel.
"""
import numpy as np
import pytest

from oracle import train as otr
from page_segmentation_b200 import synth

pytestmark = pytest.mark.gpu


def _page(seed, h, w):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    img[: h // 2, : w // 3] = 255                                   # a flat region: ReLU / max-pool ties
    lab = rng.integers(0, 3, (h, w)).astype(np.uint8)
    return img, lab


@pytest.mark.parametrize("arch,hw", [("fcn_skip", (40, 50)), ("fcn_skip", (64, 96)), ("fcn", (33, 71))])
def test_loss_and_gradients_match_autograd(ctx, arch, hw):
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    W = synth.make_weights(arch, 3, seed=5)
    img, lab = _page(hw[0], *hw)
    eng = FcnTrainStep(arch, W, 3, l_rate=1e-3, engine="fp32")
    loss = eng.forward_backward(img, lab)
    exp_loss, exp_grads, _ = otr.loss_and_grads(arch, W, img, lab, 3)
    assert abs(loss - exp_loss) <= 2e-5
    for (name, *_r), (gk, gb), (ek, eb) in zip(eng.table, eng.gradients(), exp_grads):
        for got, exp, what in ((gk, ek, "kernel"), (gb, eb, "bias")):
            assert got.shape == exp.shape, (name, what)
            assert np.abs(got - exp).max() <= 2e-4 * np.abs(exp).max() + 1e-7, (name, what, np.abs(got - exp).max(), np.abs(exp).max())


def test_adam_with_clipnorm_matches_keras_restatement(ctx):
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    arch = "fcn_skip"
    W = synth.make_weights(arch, 3, seed=6)
    eng = FcnTrainStep(arch, W, 3, l_rate=1e-3, clipnorm=0.05, engine="fp32")      # small enough that several variables are clipped
    params = [a.astype(np.float64) for pair in W for a in pair]
    m = [np.zeros_like(p) for p in params]
    v = [np.zeros_like(p) for p in params]
    losses = []
    for step in range(3):
        img, lab = _page(10 + step, 48, 64)
        cur = [(params[2 * i].astype(np.float32), params[2 * i + 1].astype(np.float32)) for i in range(len(W))]
        exp_loss, exp_grads, _ = otr.loss_and_grads(arch, cur, img, lab, 3)
        flat = [g for pair in exp_grads for g in pair]
        assert any(np.sqrt((g.astype(np.float64) ** 2).sum()) > 0.05 for g in flat)
        otr.adam_clipnorm_step(params, flat, m, v, step + 1, lr=1e-3, clipnorm=0.05)
        losses.append(eng.step(img, lab))
        assert abs(losses[-1] - exp_loss) <= 5e-5
    for (gk, gb), i in zip(eng.weights(), range(len(W))):
        assert np.abs(gk - params[2 * i]).max() <= 4e-5
        assert np.abs(gb - params[2 * i + 1]).max() <= 4e-5
        assert np.abs(gk - W[i][0]).max() > 2e-3                 # and the parameters did move


def _rel(g, e):
    return float(np.linalg.norm((g - e).ravel()) / max(np.linalg.norm(e.ravel()), 1e-30))


@pytest.mark.parametrize("arch,hw", [("fcn_skip", (40, 50)), ("fcn_skip", (64, 96)), ("fcn", (33, 71)), ("fcn_skip", (150, 260))])
def test_tensor_engine_loss_and_gradients_match_autograd(ctx, arch, hw):
    """(150, 260) -> 160 x 288 padded: three 112-pixel strips and several row bands per weight-gradient launch."""
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    W = synth.make_weights(arch, 3, seed=5)
    img, lab = _page(hw[0], *hw)
    eng = FcnTrainStep(arch, W, 3, l_rate=1e-3, engine="tensor")
    loss = eng.forward_backward(img, lab)
    exp_loss, exp_grads, _ = otr.loss_and_grads(arch, W, img, lab, 3)
    assert abs(loss - exp_loss) <= 5e-3
    for (name, *_r), (gk, gb), (ek, eb) in zip(eng.table, eng.gradients(), exp_grads):
        for got, exp, what in ((gk, ek, "kernel"), (gb, eb, "bias")):
            assert got.shape == exp.shape and np.isfinite(got).all(), (name, what)
            assert _rel(got, exp) <= 0.25, (name, what, _rel(got, exp))


def test_tensor_engine_matches_the_fp32_engine_on_an_a4_page(ctx):
    """Full-size page (1169 x 827 -> 1184 x 832): every tile shape of the bench configuration; the fp32 CUDA-core engine
    (held to the oracle above) is the reference here, the CPU oracle would take a minute."""
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    W = synth.make_weights("fcn_skip", 3, seed=7)
    page = synth.make_page(3, 1169, 827, 18)
    img = 255 - page
    lab = ((page == 0).astype(np.uint8) * (1 + (np.arange(827)[None, :] > 400))).astype(np.uint8)
    ref = FcnTrainStep("fcn_skip", W, 3, engine="fp32")
    eng = FcnTrainStep("fcn_skip", W, 3, engine="tensor")
    l0, l1 = ref.forward_backward(img, lab), eng.forward_backward(img, lab)
    assert abs(l0 - l1) <= 5e-3
    for (name, *_r), (gk, gb), (ek, eb) in zip(eng.table, eng.gradients(), ref.gradients()):
        assert _rel(gk, ek) <= 0.05 and _rel(gb, eb) <= 0.05, (name, _rel(gk, ek), _rel(gb, eb))


def _plane_major_bf16(a_hwc, planes, torch, dev):
    """(H, W, C) float array of bf16-representable values -> [planes][H][W][8] bfloat16 device tensor, zero padded"""
    h, w, c = a_hwc.shape
    full = np.zeros((h, w, planes * 8), dtype=np.float32)
    full[:, :, :c] = a_hwc
    t = torch.from_numpy(full.reshape(h, w, planes, 8).transpose(2, 0, 1, 3).copy()).to(dev)
    return t.to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("h,w,ci,co,k", [(40, 50, 20, 30, 5), (37, 250, 1, 20, 5), (64, 130, 60, 60, 5), (20, 30, 120, 40, 5),
                                         (33, 40, 80, 80, 5), (150, 260, 40, 40, 5), (40, 140, 80, 250, 1), (31, 129, 104, 128, 1)])
def test_weight_gradient_kernel_is_exact_on_small_integers(ctx, h, w, ci, co, k):
    """pcs_train_tc_wgrad (tcgen05, MN-major operands straight from the plane-major layout) against a numpy restatement of
    dw[co][ci][ky][kx] = sum x[r + ky - p][c + kx - p][ci] * dy[r][c][co] on small-integer tensors: every product and every
    partial sum is exact in bf16 / fp32, so the result is bit-exact whatever the order of the atomics.  Shapes cover one
    to five groups of vertical taps (1 .. 15 input planes), several strips and row bands, N from 32 to 256, and k = 1."""
    import torch
    rng = np.random.default_rng(h * 7 + w + ci + co)
    x = rng.integers(-3, 4, (h, w, ci)).astype(np.float32)
    dy = rng.integers(-2, 3, (h, w, co)).astype(np.float32)
    dy[rng.random((h, w)) < 0.3] = 0                                   # masked pixels, as a ReLU leaves them
    p = (k - 1) // 2
    xp = np.pad(x, ((p, p), (p, p), (0, 0))).astype(np.float64)
    exp = np.zeros((co, ci, k, k))
    for ky in range(k):
        for kx in range(k):
            exp[:, :, ky, kx] = np.einsum("rco,rci->oi", dy.astype(np.float64), xp[ky:ky + h, kx:kx + w])
    dev = f"cuda:{ctx.device}"
    xplanes, dplanes = (ci + 7) // 8, (co + 7) // 8
    d_x, d_dy = _plane_major_bf16(x, xplanes, torch, dev), _plane_major_bf16(dy, dplanes, torch, dev)
    d_dw = torch.zeros((co, ci, k, k), dtype=torch.float32, device=dev)
    ctx.use_torch_stream()
    ctx.train_call("tc_wgrad", d_x, xplanes, d_dy, dplanes, h, w, k, ci, co, d_dw)
    got = d_dw.cpu().numpy().astype(np.float64)
    assert np.array_equal(got, exp), (np.abs(got - exp).max(), np.abs(exp).max())
    ctx.train_call("tc_wgrad", d_x, xplanes, d_dy, dplanes, h, w, k, ci, co, d_dw)     # accumulates
    assert np.array_equal(d_dw.cpu().numpy().astype(np.float64), 2 * exp)


@pytest.mark.parametrize("arch,tol", [("fcn", 0.25), ("fcn_skip", 0.08)])
def test_tensor_engine_matches_the_fp32_engine_at_medium_size(ctx, arch, tol):
    """fcn has no skip connections: the gradients of its first layers cross every rounding of the chain and are nearly
    cancelling sums at random initialisation (measured 0.17 for conv1, against 0.003 with skips)."""
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    W = synth.make_weights(arch, 3, seed=9)
    img, lab = _page(77, 256, 384)
    ref = FcnTrainStep(arch, W, 3, engine="fp32")
    eng = FcnTrainStep(arch, W, 3, engine="tensor")
    l0, l1 = ref.forward_backward(img, lab), eng.forward_backward(img, lab)
    assert abs(l0 - l1) <= 5e-3
    for (name, *_r), (gk, gb), (ek, eb) in zip(eng.table, eng.gradients(), ref.gradients()):
        assert _rel(gk, ek) <= tol and _rel(gb, eb) <= tol, (name, _rel(gk, ek), _rel(gb, eb))


def test_tensor_engine_training_follows_the_fp32_trajectory(ctx):
    """Five Adam steps on the same pages: the losses of the two engines stay together and the parameters move alike (Adam
    moves every coordinate by about lr per step whatever the size of its gradient, so coordinates whose gradient is below
    the bf16 noise go either way: the updates are compared by their direction, not coordinate by coordinate)."""
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    W = synth.make_weights("fcn_skip", 3, seed=8)
    a = FcnTrainStep("fcn_skip", W, 3, l_rate=1e-3, engine="fp32")
    b = FcnTrainStep("fcn_skip", W, 3, l_rate=1e-3, engine="tensor")
    for step in range(5):
        img, lab = _page(20 + step, 96, 128)
        la, lb = a.step(img, lab), b.step(img, lab)
        assert abs(la - lb) <= 2e-2, (step, la, lb)
    for (name, *_r), (ka, ba), (kb, bb), (k0, _b0) in zip(a.table, a.weights(), b.weights(), W):
        ua, ub = (ka - k0).ravel().astype(np.float64), (kb - k0).ravel().astype(np.float64)
        cos = float(ua @ ub / (np.linalg.norm(ua) * np.linalg.norm(ub)))
        assert cos >= 0.9, (name, cos)


def test_tensor_engine_rejects_what_it_does_not_cover(ctx):
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    W = synth.make_weights("fcn_skip", 6, seed=1)
    with pytest.raises(ValueError):
        FcnTrainStep("fcn_skip", W, 6, engine="tensor")
    assert FcnTrainStep("fcn_skip", W, 6).engine == "fp32"               # the default falls back by class count
    eng = FcnTrainStep("fcn_skip", synth.make_weights("fcn_skip", 3, seed=1), 3)
    assert eng.engine == "tensor"
    with pytest.raises(ValueError):
        eng.forward_backward(np.zeros((40, 50), np.uint8), np.full((40, 50), 3, np.uint8))      # label outside 0..2


def test_trainer_reduces_the_loss_and_saves_a_loadable_model(ctx, tmp_path):
    """Overfits two small pages whose labels are a function of the image; the written .h5 predicts through Network."""
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import Dataset, SingleData
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.trainer import Trainer, TrainSettings
    pages = []
    for s in range(2):
        page = synth.make_page(s, 192, 160, 6)
        img = 255 - page                                            # `data.image` is the inverted grey page
        pages.append(SingleData(image=img, binary=(page == 0).astype(np.uint8), mask=(page == 0).astype(np.uint8)))
    settings = TrainSettings(n_epoch=12, n_classes=3, l_rate=1e-3, train_data=Dataset(pages, DEFAULT_COLOR_MAP), validation_data=None,
                             display=0, output_dir=str(tmp_path), threads=1)
    tr = Trainer(settings)
    tr.train()
    assert len(tr.losses) == 24 and np.isfinite(tr.losses).all()
    assert np.mean(tr.losses[-4:]) < 0.6 * np.mean(tr.losses[:2])
    net = Network("Predict", n_classes=3, model=str(tmp_path / "model.h5"))
    _, _, pred = net.predict_single_data(pages[0])
    assert (pred == pages[0].mask).mean() > 0.8
