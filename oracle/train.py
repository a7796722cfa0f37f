"""CPU restatement of the reference's training step (test infrastructure, see oracle/__init__).

Follows ocr4all_pixel_classifier/lib/network.py:151-161 (one page per step: `image_to_batch(preprocess(i))`, mask as
sparse labels), lib/metrics.py:8-9 (`loss` = mean sparse categorical cross entropy from logits), lib/model.py:45-92 /
:206-234 (the graphs, through oracle.network's TF-semantics primitives) and the optimizer compiled at
lib/network.py:91-103: tf.keras.optimizers.Adam(lr, clipnorm) of TensorFlow <= 2.5, i.e. every variable's gradient
is clipped by its own norm (tf.clip_by_norm) before  m, v, lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t),
p -= lr_t * m / (sqrt(v) + eps)  with eps = 1e-7 [recalled: Keras OptimizerV2 Adam, non-amsgrad; unpinned].
Gradients come from torch autograd on the restated graph (float64 available for tolerance derivation).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import network as onet


def loss_and_grads(arch: str, weights: Sequence[Tuple[np.ndarray, np.ndarray]], image_u8: np.ndarray, labels: np.ndarray,
                   n_classes: int, dtype=torch.float32):
    """-> (loss, [(dKernel, dBias) in Keras layout per layer], logits HWC)."""
    from page_segmentation_b200.synth import layer_table
    assert arch in ("fcn_skip", "fcn")
    table = layer_table(arch, n_classes)
    params = {}
    for (name, kind, k, ci, co, act), (w, b) in zip(table, weights):
        wt = torch.tensor(np.ascontiguousarray(w), dtype=dtype, requires_grad=True)
        bt = torch.tensor(np.ascontiguousarray(b), dtype=dtype, requires_grad=True)
        params[name] = (wt, bt, kind, k, act)

    def L(name, x):
        w, b, kind, k, act = params[name]
        if kind in ("conv", "logits"):
            y = onet._conv_same(x, w, b, k)
        elif kind == "deconv":
            y = onet._deconv_same(x, w, b, k, 1)
        else:
            y = onet._deconv_same(x, w, b, k, 2)
        return onet._act(y, act)

    h, w_ = image_u8.shape
    px, py = onet.calculate_padding(h, w_)
    x = torch.from_numpy((image_u8.astype(np.float64) / 255.0).astype(np.float32)).to(dtype)[None, None]
    x = F.pad(x, (0, py, 0, px))
    skip = arch == "fcn_skip"
    conv1 = L("conv1", x)
    conv2 = L("conv2", conv1)
    conv3 = L("conv3", F.max_pool2d(conv2, 2, 2))
    conv4 = L("conv4", conv3)
    conv5 = L("conv5", F.max_pool2d(conv4, 2, 2))
    conv6 = L("conv6", conv5)
    conv7 = L("conv7", F.max_pool2d(conv6, 2, 2))
    d1 = L("deconv1", conv7)
    d2 = L("deconv2", d1)
    if skip:
        d2 = torch.cat([d2, conv6], 1)
    d3 = L("deconv3", d2)
    if skip:
        d3 = torch.cat([d3, conv5], 1)
    d4 = L("deconv4", d3)
    if skip:
        d4 = torch.cat([d4, conv3], 1)
    d5 = L("deconv5", d4)
    if skip:
        d5 = torch.cat([d5, conv2], 1)
    logits = L("logits", d5[:, :, :h, :w_])
    target = torch.from_numpy(labels.astype(np.int64))[None]
    loss = F.cross_entropy(logits, target, reduction="mean")           # metrics.py:8-9
    loss.backward()
    grads = [(params[n][0].grad.numpy().copy(), params[n][1].grad.numpy().copy()) for (n, *_r) in table]
    return float(loss.detach()), grads, logits[0].permute(1, 2, 0).detach().numpy()


def adam_clipnorm_step(params: List[np.ndarray], grads: List[np.ndarray], m: List[np.ndarray], v: List[np.ndarray], t: int,
                       lr: float, clipnorm: float = 1.0, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-7):
    """One Keras-Adam update in place (float64 arithmetic); t counts from 1.  Every array is one variable."""
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for p, g, mi, vi in zip(params, grads, m, v):
        g = g.astype(np.float64)
        if clipnorm and clipnorm > 0:
            norm = np.sqrt((g * g).sum())
            g = g * (clipnorm / max(norm, clipnorm))                   # tf.clip_by_norm
        mi[...] = b1 * mi + (1 - b1) * g
        vi[...] = b2 * vi + (1 - b2) * g * g
        p[...] = p - lr_t * mi / (np.sqrt(vi) + eps)
