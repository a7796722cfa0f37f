"""Parity of the device training step (csrc/train.cu + lib/trainer.py) with the CPU oracle (oracle/train.py: torch
autograd on the restated graph, Keras Adam with per-variable clipnorm restated in numpy).

Tolerances (fp32 on both sides, different summation orders):
  loss                 |d| <= 2e-5
  every gradient       max |d| <= 2e-4 * max |g_oracle| + 1e-7   per variable
  parameters after 3 Adam steps   max |d| <= 4e-5  (lr 1e-3: every coordinate moves ~1e-3 per step at the start of Adam,
                                  whatever the size of its gradient, so m / sqrt(v) amplifies the relative gradient
                                  error of near-zero gradients; 4e-5 is 1.3 % of the distance travelled)
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
    eng = FcnTrainStep(arch, W, 3, l_rate=1e-3)
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
    eng = FcnTrainStep(arch, W, 3, l_rate=1e-3, clipnorm=0.05)      # small enough that several variables are clipped
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
