# Test infrastructure: restates functions of ocr4all_pixel_classifier (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation,
# (c) its authors, licensed Apache-2.0 OR GPL-3.0-or-later) on the CPU as the parity oracle; never imported by the product path.
"""CPU restatement of `compute_char_height` (test infrastructure, see oracle/__init__).

Follows ocr4all_pixel_classifier/lib/image_ops.py:58-82 with cv2 (installed here, 4.13) as the live
implementation of its three library calls: Otsu threshold, subtraction from 255, connected components with
stats.  The reference passes `4` POSITIONALLY to cv2.connectedComponentsWithStats, where it fills the `labels`
output slot and is ignored, so cv2's default 8-connectivity applies [probed here]; the restatement makes that
explicit with `connectivity=8` and tests/test_oracle_known_answers.py pins the equivalence.
"""
from __future__ import annotations

import cv2
import numpy as np


def otsu_threshold(img: np.ndarray) -> int:
    """cv2 getThreshVal_Otsu_8u restated in numpy float64 (same operation order): pins what the device computes."""
    h = np.bincount(img.ravel(), minlength=256).astype(np.float64)
    scale = 1.0 / float(img.size)
    mu = 0.0
    for i in range(256):
        mu += i * h[i]
    mu *= scale
    mu1 = q1 = max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = h[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma, max_val = sigma, i
    return max_val


def compute_char_height_array(img: np.ndarray, inverse: bool):
    """image_ops.py:62-82 on an in-memory grey page."""
    _, bw = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    if not inverse:
        bw = cv2.subtract(255, bw)
    _, _, stats, _ = cv2.connectedComponentsWithStats(bw, connectivity=8)
    w, h = stats[1:, cv2.CC_STAT_WIDTH], stats[1:, cv2.CC_STAT_HEIGHT]
    ok = (0.5 < w / np.maximum(h, 1)) & (w / np.maximum(h, 1) < 2) & (10 < h) & (h < 60) & (5 < w) & (w < 50)
    heights = np.sort(h[ok])
    if len(heights) == 0:
        return None
    return heights[int(len(heights) / 2)]


def fgpa(pred: np.ndarray, mask: np.ndarray, bin: np.ndarray) -> float:
    """image_ops.py:8-19."""
    fg = np.count_nonzero(bin)
    with np.errstate(invalid="ignore"):
        return np.float64(fg - np.count_nonzero((pred * bin) != (mask * bin))) / np.float64(fg)      # nan without foreground


def fgoverlap_per_class(pred: np.ndarray, mask: np.ndarray, bin: np.ndarray, n_classes: int):
    """image_ops.py:22-55 -> (overlaps, tps, fps, fns), n_classes + 1 entries each."""
    pfg, mfg = (pred.astype(np.int64) + 1) * bin - 1, (mask.astype(np.int64) + 1) * bin - 1
    out = ([], [], [], [])
    for i in range(n_classes + 1):
        a, e = pfg == i, mfg == i
        tp, fp, fn = int((a & e).sum()), int((a & ~e).sum()), int((e & ~a).sum())
        vals = (np.nan, 0, 0, 0) if tp + fp + fn == 0 else (tp / (tp + fp + fn), tp, fp, fn)
        for lst, v in zip(out, vals):
            lst.append(v)
    return out
