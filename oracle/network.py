"""CPU restatement of the reference Keras graphs (test infrastructure, see
oracle/__init__): torch-CPU conv primitives with TF/Keras 2.x semantics.

Follows ocr4all_pixel_classifier/lib/model.py:10-42 (calculate_padding / pad /
crop), :45-92 (model_fcn_skip), :206-234 (model_fcn), :151-203 (unet),
lib/architecture.py:67-68 (default_preprocess x/255.0), lib/util.py:12-18
(image_to_batch) and lib/network.py:248-260 (predict_single_data).

TF semantics restated:
  Conv2D(k,'same',s=1): cross-correlation, kernel (kh,kw,Cin,Cout), pad k//2
      (k=2: pad 0 before, 1 after);
  Conv2DTranspose(5,'same',s=1): kernel (kh,kw,Cout,Cin), gradient-of-conv
      form == F.conv_transpose2d(x, K.permute(3,2,0,1), padding=2);
  Conv2DTranspose(2,'same',s=2): y[2h+i,2w+j,o] = b[o] + sum_c x[h,w,c] K[i,j,o,c];
  MaxPooling2D(2,2); UpSampling2D(2) nearest; Dropout = identity at inference;
  bias added before the activation.

`bf16=True` gives the numerics twin of the device path: weights of every layer
but conv1 and logits are rounded to bf16, every stored activation is rounded to
bf16, accumulation stays fp32 (see DESIGN.md "precision").  `conv1_rounded=True`
rounds the weights of the first layer as well: the tensor engine with fp16 operands
uses them as ONE operand (conv1_umma.cu; with bf16 operands they are split hi + lo).  `fused_head=True`
additionally mirrors the tensor-engine head: deconv5 keeps its fp32 weights (it is
composed with the logits layer on the host and split into two bf16 operand halves)
and the conv2 skip enters the logits unrounded (taken from conv2's fp32 accumulators).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def calculate_padding(h: int, w: int, f: int = 32) -> Tuple[int, int]:
    """model.py:10-17 (px pads H = shape[1], py pads W = shape[2])."""
    return (f - h % f) % f, (f - w % f) % f


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype)


def _conv_same(x, w_keras, b, k):
    """x NCHW; Keras Conv2D kernel (kh,kw,Cin,Cout)."""
    w = w_keras.permute(3, 2, 0, 1).contiguous()
    if k % 2 == 1:
        return F.conv2d(x, w, b, padding=k // 2)
    # even kernel: TF SAME pads 0 before, 1 after
    x = F.pad(x, (0, k - 1, 0, k - 1))
    return F.conv2d(x, w, b)


def _deconv_same(x, w_keras, b, k, stride):
    """Keras Conv2DTranspose kernel (kh,kw,Cout,Cin) -> torch (Cin,Cout,kh,kw)."""
    w = w_keras.permute(3, 2, 0, 1).contiguous()
    if stride == 1:
        return F.conv_transpose2d(x, w, b, padding=k // 2)
    assert stride == k == 2
    return F.conv_transpose2d(x, w, b, stride=2)


def _act(x, act):
    return F.relu(x) if act == "relu" else x


class Forward:
    """Runs one of the three reference graphs on a single page."""

    def __init__(self, arch: str, weights: Sequence[Tuple[np.ndarray, np.ndarray]], n_classes: int,
                 dtype=torch.float32, bf16: bool = False, fused_head: bool = False, conv1_rounded: bool = False):
        from page_segmentation_b200.synth import layer_table
        self.arch = arch
        self.table = layer_table(arch, n_classes)
        assert len(weights) == len(self.table)
        self.dtype = dtype
        self.bf16 = bf16
        self.fused_head = fused_head and bf16
        self.params: Dict[str, Tuple[torch.Tensor, torch.Tensor, tuple]] = {}
        for (name, kind, k, ci, co, act), (w, b) in zip(self.table, weights):
            wt = torch.from_numpy(np.ascontiguousarray(w)).to(dtype)
            bt = torch.from_numpy(np.ascontiguousarray(b)).to(dtype)
            if bf16 and name not in ("logits",) + (() if conv1_rounded else ("conv1", "conv1a")) and \
                    not (self.fused_head and name == "deconv5"):
                wt = _bf16(wt)
            self.params[name] = (wt, bt, (kind, k, ci, co, act))

    def _layer(self, name, x, store=True):
        w, b, (kind, k, ci, co, act) = self.params[name]
        if kind in ("conv", "logits"):
            y = _conv_same(x, w, b, k)
        elif kind == "deconv":
            y = _deconv_same(x, w, b, k, 1)
        else:
            y = _deconv_same(x, w, b, k, 2)
        y = _act(y, act)
        if self.bf16 and store:
            y = _bf16(y)
        return y

    @torch.no_grad()
    def logits(self, image_u8: np.ndarray, keep: Optional[List[str]] = None):
        """image_u8: (H,W) uint8 `data.image`; returns logits (H,W,n) float and
        a dict of kept intermediate activations (NHWC numpy)."""
        h, w = image_u8.shape
        px, py = calculate_padding(h, w)
        # architecture.py:67-68 x/255.0 (float64 in numpy) -> float32 at the TF boundary
        x = (image_u8.astype(np.float64) / 255.0).astype(np.float32)
        x = torch.from_numpy(x).to(self.dtype)[None, None]
        x = F.pad(x, (0, py, 0, px))                          # model.py:20-26, bottom/right zeros
        kept: Dict[str, np.ndarray] = {}

        def K(name, t):
            if keep is not None and name in keep:
                kept[name] = t[0].permute(1, 2, 0).to(torch.float32).numpy().copy()
            return t

        L = self._layer
        if self.arch in ("fcn_skip", "fcn"):
            skip = self.arch == "fcn_skip"
            conv1 = K("conv1", L("conv1", x))
            conv2_f32 = L("conv2", conv1, store=False)
            conv2 = K("conv2", _bf16(conv2_f32) if self.bf16 else conv2_f32)
            pool2 = F.max_pool2d(conv2, 2, 2)
            conv3 = K("conv3", L("conv3", pool2))
            conv4 = K("conv4", L("conv4", conv3))
            pool4 = F.max_pool2d(conv4, 2, 2)
            conv5 = K("conv5", L("conv5", pool4))
            conv6 = K("conv6", L("conv6", conv5))
            pool6 = F.max_pool2d(conv6, 2, 2)
            conv7 = K("conv7", L("conv7", pool6))
            d1 = K("deconv1", L("deconv1", conv7))
            d2 = K("deconv2", L("deconv2", d1))
            if skip:
                d2 = torch.cat([d2, conv6], 1)
            d3 = K("deconv3", L("deconv3", d2))
            if skip:
                d3 = torch.cat([d3, conv5], 1)
            d4 = K("deconv4", L("deconv4", d3))
            if skip:
                d4 = torch.cat([d4, conv3], 1)
            # device path keeps deconv5 in fp32 registers and feeds the logits directly
            d5 = K("deconv5", L("deconv5", d4, store=False))
            if skip:
                d5 = torch.cat([d5, conv2_f32 if self.fused_head else conv2], 1)
            last = d5
        elif self.arch == "unet":
            def up(name, t):
                return L(name, F.interpolate(t, scale_factor=2, mode="nearest"))
            c1 = K("conv1b", L("conv1b", L("conv1a", x)))
            c2 = K("conv2b", L("conv2b", L("conv2a", F.max_pool2d(c1, 2, 2))))
            c3 = K("conv3b", L("conv3b", L("conv3a", F.max_pool2d(c2, 2, 2))))
            c4 = K("conv4b", L("conv4b", L("conv4a", F.max_pool2d(c3, 2, 2))))
            c5 = K("conv5b", L("conv5b", L("conv5a", F.max_pool2d(c4, 2, 2))))
            u6 = K("up6", up("up6", c5))
            c6 = K("conv6b", L("conv6b", L("conv6a", torch.cat([c4, u6], 1))))
            u7 = up("up7", c6)
            c7 = K("conv7b", L("conv7b", L("conv7a", torch.cat([c3, u7], 1))))
            u8 = up("up8", c7)
            c8 = K("conv8b", L("conv8b", L("conv8a", torch.cat([c2, u8], 1))))
            u9 = up("up9", c8)
            c9 = K("conv9b", L("conv9b", L("conv9a", torch.cat([c1, u9], 1))))
            last = c9
        else:
            raise KeyError(self.arch)
        last = last[:, :, :h, :w]                              # model.py:29-42 crop
        logit = self._layer("logits", last, store=False)
        return logit[0].permute(1, 2, 0).to(torch.float32).numpy(), kept

    def predict(self, image_u8: np.ndarray):
        """network.py:248-260 -> (logit f32 HWC, prob f32 HWC, pred int64 HW)."""
        from .pipeline import softmax_argmax
        logit, _ = self.logits(image_u8)
        prob, pred = softmax_argmax(logit)
        return logit, prob, pred
