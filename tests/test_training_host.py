"""CPU tests of the training host logic: weight layout conversions, the oracle's optimizer restatement, the
data-parallel gradient averaging contract (gloo, world size 2) and the TrainSettings surface."""
import os

import numpy as np
import pytest
import torch

from oracle import train as otr
from page_segmentation_b200 import synth
from page_segmentation_b200.lib import trainer as tr


def test_weight_layout_roundtrip_and_correlation_forms():
    rng = np.random.default_rng(0)
    for kind, shape in (("conv", (5, 5, 20, 30)), ("deconv", (5, 5, 40, 120)), ("deconv_s2", (2, 2, 60, 80)), ("logits", (1, 1, 50, 3))):
        k = rng.normal(size=shape).astype(np.float32)
        assert np.array_equal(tr.from_internal(kind, tr.to_internal(kind, k), shape[0]), k)
    # the stride-1 transposed convolution equals a correlation with the internal (flipped) weights
    k = rng.normal(size=(5, 5, 7, 4)).astype(np.float32)                     # (kh, kw, C_out, C_in)
    x = torch.from_numpy(rng.normal(size=(1, 4, 9, 11)).astype(np.float32))
    ref = torch.nn.functional.conv_transpose2d(x, torch.from_numpy(k).permute(3, 2, 0, 1).contiguous(), padding=2)
    got = torch.nn.functional.conv2d(x, torch.from_numpy(tr.to_internal("deconv", k)), padding=2)
    assert torch.allclose(ref, got, atol=1e-5)
    # the input gradient of a correlation layer is the correlation of dy with bwd_data_weights(w)
    w = torch.from_numpy(rng.normal(size=(6, 4, 5, 5)).astype(np.float32))
    xx = x.clone().requires_grad_(True)
    y = torch.nn.functional.conv2d(xx, w, padding=2)
    dy = torch.from_numpy(rng.normal(size=tuple(y.shape)).astype(np.float32))
    y.backward(dy)
    got = torch.nn.functional.conv2d(dy, tr.bwd_data_weights("conv", w), padding=2)
    assert torch.allclose(xx.grad, got, atol=1e-4)


def test_oracle_adam_clipnorm_known_answer():
    p, g = [np.array([1.0, -2.0])], [np.array([3.0, 4.0])]                   # norm 5 -> clipped to [0.6, 0.8]
    m, v = [np.zeros(2)], [np.zeros(2)]
    otr.adam_clipnorm_step(p, g, m, v, 1, lr=0.1, clipnorm=1.0)
    assert np.allclose(m[0], [0.06, 0.08]) and np.allclose(v[0], [0.00036, 0.00064])
    # first step of Adam moves every coordinate by lr (up to eps): lr_t * m / sqrt(v) = lr
    assert np.allclose(p[0], [0.9, -2.1], atol=1e-5)


def test_oracle_gradients_match_finite_differences():
    arch = "fcn_skip"
    W = [(k.astype(np.float64), b.astype(np.float64)) for k, b in synth.make_weights(arch, 3, seed=1)]
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (16, 24), dtype=np.uint8)
    lab = rng.integers(0, 3, (16, 24)).astype(np.uint8)
    _, grads, _ = otr.loss_and_grads(arch, W, img, lab, 3, dtype=torch.float64)
    for layer, idx in ((0, (2, 3, 0, 5)), (7, (1, 4, 3, 2)), (10, (1, 0, 2, 7)), (12, (0, 0, 11, 1))):
        Wp = [(k.copy(), b.copy()) for k, b in W]
        Wm = [(k.copy(), b.copy()) for k, b in W]
        Wp[layer][0][idx] += 1e-5
        Wm[layer][0][idx] -= 1e-5
        lp = otr.loss_and_grads(arch, Wp, img, lab, 3, dtype=torch.float64)[0]
        lm = otr.loss_and_grads(arch, Wm, img, lab, 3, dtype=torch.float64)[0]
        assert abs((lp - lm) / 2e-5 - grads[layer][0][idx]) <= 1e-6 + 1e-4 * abs(grads[layer][0][idx])


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Stub:                                           # the two members allreduce_gradients touches
            grads = torch.full((5,), float(rank + 1))
        scale = tr.FcnTrainStep.allreduce_gradients(Stub)
        out.put((rank, scale, Stub.grads.tolist()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_average_gloo():
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctxm.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, scale, grads in res:
        assert scale == 0.5 and grads == [3.0] * 5          # SUM over ranks; the 1/world scale goes into the update


def test_train_settings_surface():
    names = tr.TrainSettings._fields
    for f in ("n_epoch", "n_classes", "l_rate", "train_data", "validation_data", "output_dir", "architecture", "optimizer_norm_clipping",
              "optimizer_norm_clip_value", "foreground_masks", "load", "model_name", "model_suffix", "save_weights_only"):
        assert f in names
    d = tr.TrainSettings._field_defaults
    assert d["optimizer_norm_clipping"] is True and d["optimizer_norm_clip_value"] == 1.0 and d["model_suffix"] == ".h5"
