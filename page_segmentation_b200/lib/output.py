"""Mirror of ocr4all_pixel_classifier/lib/output.py: Masks (:12-17),
output_data (:20-41), generate_output_masks (:44-60),
scale_to_original_shape (:63-79)."""
from __future__ import annotations

import os
from dataclasses import dataclass, replace
from typing import Optional

import numpy as np

from .colors import ColorMap
from .dataset import SingleData


@dataclass
class Masks:
    color: np.ndarray
    overlay: np.ndarray
    inverted_overlay: np.ndarray
    fg_color_mask: Optional[np.ndarray] = None


def generate_output_masks(data: SingleData, pred: np.ndarray, color_map: ColorMap) -> Masks:
    """output.py:44-60 through the device epilogue (pcs_masks)."""
    import torch
    from ..runtime import get_context, to_device_u8
    ctx = get_context()
    pred = np.asarray(pred)
    h, w = pred.shape
    lut = color_map.lut()
    if pred.size and (pred.min() < 0 or pred.max() > 255):
        raise ValueError("labels outside 0..255")
    binary = np.asarray(data.binary)
    d_pred = to_device_u8(pred, ctx.device)
    d_bin = to_device_u8(binary, ctx.device)
    outs = [torch.empty((h, w, 3), dtype=torch.uint8, device=d_pred.device) for _ in range(3)]
    ctx.masks(d_pred, d_bin, 1, h, w, lut, outs[0], outs[1], outs[2])
    color, overlay, inverted = (o.cpu().numpy() for o in outs)
    # fg_color_mask[foreground != 0] = 0 is arithmetically the inverted overlay (output.py:50-53)
    return Masks(color=color, overlay=overlay, inverted_overlay=inverted, fg_color_mask=inverted.copy())


def output_data(output_dir, pred, data: SingleData, color_map):
    """output.py:20-41 (cv2.imwrite in place of skimage.io.imsave)."""
    import cv2
    if len(pred.shape) == 3:
        assert (pred.shape[0] == 1)
        pred = pred[0]
    if data.output_path:
        filename = data.output_path
        dir = os.path.dirname(filename)
        if os.path.isabs(dir):
            os.makedirs(dir, exist_ok=True)
        elif dir:
            for category in ["color", "overlay", "inverted"]:
                os.makedirs(os.path.join(output_dir, category, dir), exist_ok=True)
    else:
        filename = os.path.basename(data.image_path)
    masks = generate_output_masks(data, pred, color_map)
    for category, img in (("color", masks.color), ("overlay", masks.overlay), ("inverted", masks.inverted_overlay)):
        path = os.path.join(output_dir, category, filename)
        if not cv2.imwrite(path, np.ascontiguousarray(img[..., ::-1])):
            raise IOError(f"could not write {path}")


def scale_to_original_shape(data: SingleData, pred):
    """output.py:63-79."""
    from .util import preserving_resize
    resized_image = preserving_resize(data.image, data.original_shape)
    pred = preserving_resize(pred, data.original_shape).astype('int64')
    if data.binary.shape != tuple(data.original_shape):
        if data.orig_binary is not None:
            resized_binary = data.orig_binary
        else:
            resized_binary = preserving_resize(data.binary, data.original_shape).astype('bool')
    else:
        resized_binary = data.binary
    data = replace(data, binary=resized_binary, image=resized_image)
    return data, pred
