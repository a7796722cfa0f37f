# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/image_ops.py on the B200 path: `compute_char_height` (:58-82, the producer of
the `line_height_px` normalisation input; pcs_char_height) and the evaluation metrics `fgpa` (:8-19) and
`fgoverlap_per_class` (:22-55) as one masked counting pass (pcs_eval_counts)."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from typing import List, Tuple

from .. import runtime


def compute_char_height_array(img: np.ndarray, inverse: bool, device: Optional[int] = None):
    """`compute_char_height` for a grey page already in memory (uint8, H x W); returns the height (numpy int32, as
    indexing cv2's stats gives) or None when no component looks like a letter."""
    torch = runtime._torch()
    ctx = runtime.get_context(device)
    img = np.asarray(img)
    if img.ndim != 2 or img.dtype != np.uint8:
        raise ValueError("compute_char_height expects a 2-D uint8 grey image (cv2.IMREAD_GRAYSCALE)")
    d_img = runtime.to_device_u8(img, ctx.device)
    d_out = torch.empty((1,), dtype=torch.int32, device=d_img.device)
    ctx.char_height(d_img, 1, img.shape[0], img.shape[1], inverse, d_out)
    h = int(d_out.cpu()[0])
    return None if h < 0 else np.int32(h)


def compute_char_height(file_name: str, inverse: bool):
    """image_ops.py:58-82: same signature, same error for a missing file, same None for "no letters"."""
    if not os.path.exists(file_name):
        raise Exception(f"File does not exist at {file_name}")
    import cv2                      # file decoding stays on the host, as in the reference (cv2.imread)
    img = cv2.imread(file_name, cv2.IMREAD_GRAYSCALE)
    return compute_char_height_array(img, inverse)


def _eval_counts(pred: np.ndarray, mask: np.ndarray, bin: np.ndarray, n_classes: int, device: Optional[int] = None):
    torch = runtime._torch()
    ctx = runtime.get_context(device)
    pred, mask, bin = np.asarray(pred), np.asarray(mask), np.asarray(bin)
    if pred.shape != mask.shape or pred.shape != bin.shape:
        raise ValueError("pred, mask and bin must have one shape")
    if bin.size and (bin.min() < 0 or bin.max() > 1):
        raise ValueError("bin must be a {0, 1} foreground map (1 is foreground), as dataset.py:146 leaves it")
    d = [runtime.to_device_u8(a, ctx.device) for a in (pred, mask, bin)]
    nb = n_classes + 2
    d_out = torch.zeros((2 + nb * nb,), dtype=torch.int64, device=d[0].device)
    ctx.eval_counts(d[0], d[1], d[2], pred.size, n_classes, d_out)
    out = d_out.cpu().numpy()
    return int(out[0]), int(out[1]), out[2:].reshape(nb, nb)


def fgpa(pred: np.ndarray, mask: np.ndarray, bin: np.ndarray):
    """image_ops.py:8-19: foreground pixel accuracy.  A page without foreground gives nan, as the reference does with the
    numpy scalars np.count_nonzero returns today (with numpy 1.x ints it raised ZeroDivisionError)."""
    fg_count, neq, _ = _eval_counts(pred, mask, bin, 1)
    return np.float64(fg_count - neq) / np.float64(fg_count) if fg_count else np.float64("nan")


def fgoverlap_per_class(pred: np.ndarray, mask: np.ndarray, bin: np.ndarray, n_classes: int) \
        -> Tuple[List[float], List[int], List[int], List[int]]:
    """image_ops.py:22-55: per class 0..n_classes (n_classes + 1 entries) the foreground overlap tp / (tp + fp + fn) and
    tp, fp, fn; nan and zeros for a class that occurs in neither map."""
    _, _, conf = _eval_counts(pred, mask, bin, n_classes)
    overlaps, tps, fps, fns = [], [], [], []
    for i in range(n_classes + 1):
        tp = int(conf[i, i])
        fp = int(conf[i, :].sum()) - tp          # predicted i, expected something else
        fn = int(conf[:, i].sum()) - tp          # expected i, predicted something else
        if tp + fp + fn == 0:
            overlaps.append(np.nan); tps.append(0); fps.append(0); fns.append(0)
        else:
            overlaps.append(tp / (tp + fp + fn)); tps.append(tp); fps.append(fp); fns.append(fn)
    return overlaps, tps, fps, fns
