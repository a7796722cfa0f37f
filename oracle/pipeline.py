# Test infrastructure: restates functions of ocr4all_pixel_classifier (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation,
# (c) its authors, licensed Apache-2.0 OR GPL-3.0-or-later) on the CPU as the parity oracle; never imported by the product path.
"""CPU restatement of the page-level stages either side of the network
(test infrastructure, see oracle/__init__).

Follows, line by line:
  prepare_images                 ocr4all_pixel_classifier/lib/dataset.py:131-150
  softmax / argmax               lib/network.py:258-259
  generate_output_masks          lib/output.py:44-60
  scale_to_original_shape        lib/output.py:63-79 (+ lib/util.py:21-29)
  vote_connected_component_class lib/postprocess.py:9-26
  add_bounding_boxes             lib/postprocess.py:29-42 (intent; the reference
                                 passes a bool array to cv2, which cv2 rejects)
  ColorMap.to_rgb_array          ocr4all-pylib 0.2.6 (absent): out[pred==label]=colour,
                                 unknown labels -> (0,0,0)   [SURVEY appendix D]
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import cv2
import numpy as np

from . import resize as sk


def prepare_images(image: np.ndarray, binary: np.ndarray, target_line_height: int, line_height_px: int,
                   max_width: Optional[int] = None, keep_orig_bin: bool = False):
    """dataset.py:131-150."""
    scale = target_line_height / line_height_px
    orig_bin = binary / 255 if np.max(binary) > 1 else binary
    bin_ = 1.0 - sk.rescale(orig_bin, scale, order=0, anti_aliasing=False)
    img = 1.0 - sk.resize(image, bin_.shape, order=3, anti_aliasing=len(np.unique(image)) > 2) / 255
    if max_width is not None:
        n_scale = max_width / bin_.shape[1]
        if n_scale < 1.0:
            bin_ = sk.rescale(bin_, n_scale, order=0, anti_aliasing=False)
            img = sk.resize(img, bin_.shape, order=3, anti_aliasing=len(np.unique(img)) > 2)
    img = np.ascontiguousarray((img * 255).astype(np.uint8))
    bin_ = np.ascontiguousarray(bin_.astype(np.uint8))
    if keep_orig_bin:
        return img, bin_, (1 - orig_bin).astype(np.uint8)
    return img, bin_


def softmax_argmax(logit: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """network.py:258-259: scipy.special.softmax(logit, -1), np.argmax(logit, -1)."""
    from scipy.special import softmax
    return softmax(logit, -1), np.argmax(logit, -1)


def to_rgb_array(pred: np.ndarray, lut: Dict[int, Tuple[int, int, int]]) -> np.ndarray:
    out = np.zeros(pred.shape + (3,), dtype=np.uint8)
    for label, colour in lut.items():
        out[pred == label] = colour
    return out


def generate_output_masks(binary: np.ndarray, pred: np.ndarray, lut: Dict[int, Tuple[int, int, int]]):
    """output.py:44-60; returns (color, overlay, inverted_overlay, fg_color_mask)."""
    color_mask = to_rgb_array(pred, lut)
    foreground = np.stack([(1 - binary)] * 3, axis=-1)
    binary3d = np.stack([binary] * 3, axis=-1)
    overlay_mask = color_mask.copy()
    overlay_mask[foreground == 0] = 0
    inverted_overlay_mask = color_mask.copy()
    inverted_overlay_mask[binary3d == 0] = 0
    fg_color_mask = color_mask.copy()
    fg_color_mask[foreground != 0] = 0
    return color_mask, overlay_mask, inverted_overlay_mask, fg_color_mask


def preserving_resize(image: np.ndarray, target_shape) -> np.ndarray:
    """util.py:21-29."""
    return sk.resize(image, target_shape, order=0, anti_aliasing=False)


def scale_to_original_shape(image, binary, orig_binary, original_shape, pred):
    """output.py:63-79; returns (image, binary, pred) at original_shape."""
    resized_image = preserving_resize(image, original_shape)
    pred = preserving_resize(pred, original_shape).astype("int64")
    if binary.shape != tuple(original_shape):
        if orig_binary is not None:
            resized_binary = orig_binary
        else:
            resized_binary = preserving_resize(binary, original_shape).astype("bool")
    else:
        resized_binary = binary
    return resized_image, resized_binary, pred


def vote_connected_component_class(pred: np.ndarray, binary: np.ndarray) -> np.ndarray:
    """postprocess.py:9-26 (mutates and returns pred, like the reference)."""
    num_labels, labels, stats, _ = cv2.connectedComponentsWithStats(binary, connectivity=4)
    for i in range(1, num_labels):
        left = stats[i, cv2.CC_STAT_LEFT]
        top = stats[i, cv2.CC_STAT_TOP]
        w = stats[i, cv2.CC_STAT_WIDTH]
        h = stats[i, cv2.CC_STAT_HEIGHT]
        pred_slice = pred[top:top + h, left:left + w]
        mask = (labels[top:top + h, left:left + w] == i)
        prebin = np.reshape((pred_slice + 1) * mask, pred_slice.size)
        bins = np.bincount(prebin)
        maxclass = np.argmax(bins[1:])
        pred[top:top + h, left:left + w] = pred_slice - mask * pred_slice + mask * maxclass
    return pred


def add_bounding_boxes(pred: np.ndarray) -> np.ndarray:
    """postprocess.py:29-42 with `(pred == c).astype(uint8)` (evident intent)."""
    classes = np.unique(pred)
    newpred = np.zeros_like(pred)
    for c in classes:
        num_labels, labels, stats, _ = cv2.connectedComponentsWithStats((pred == c).astype(np.uint8),
                                                                        connectivity=4)
        for i in range(1, num_labels):
            left = stats[i, cv2.CC_STAT_LEFT]
            top = stats[i, cv2.CC_STAT_TOP]
            w = stats[i, cv2.CC_STAT_WIDTH]
            h = stats[i, cv2.CC_STAT_HEIGHT]
            newpred[top:top + h, left:left + w] = c
    return newpred


def connected_components_4(binary: np.ndarray):
    """Independent restatement of cv2.connectedComponentsWithStats(connectivity=4)
    semantics used to pin cv2 itself on small cases: labels are numbered in
    raster order of each component's first pixel; stats rows are
    [left, top, width, height, area]; row 0 is the background."""
    h, w = binary.shape
    labels = np.zeros((h, w), dtype=np.int32)
    stats = [[0, 0, 0, 0, 0]]
    nxt = 1
    fg = binary != 0
    for y in range(h):
        for x in range(w):
            if fg[y, x] and labels[y, x] == 0:
                stack = [(y, x)]
                labels[y, x] = nxt
                x0 = x1 = x
                y0 = y1 = y
                area = 0
                while stack:
                    cy, cx = stack.pop()
                    area += 1
                    x0, x1 = min(x0, cx), max(x1, cx)
                    y0, y1 = min(y0, cy), max(y1, cy)
                    for ny, nx in ((cy - 1, cx), (cy + 1, cx), (cy, cx - 1), (cy, cx + 1)):
                        if 0 <= ny < h and 0 <= nx < w and fg[ny, nx] and labels[ny, nx] == 0:
                            labels[ny, nx] = nxt
                            stack.append((ny, nx))
                stats.append([x0, y0, x1 - x0 + 1, y1 - y0 + 1, area])
                nxt += 1
    bg = ~fg
    if bg.any():
        ys, xs = np.nonzero(bg)
        stats[0] = [int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1),
                    int(ys.max() - ys.min() + 1), int(bg.sum())]
    return nxt, labels, np.asarray(stats, dtype=np.int32)
