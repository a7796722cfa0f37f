# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
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


def _class_map_and_binary_on_device(data: SingleData, pred):
    """(ctx, class map, data.binary) as contiguous (H, W) uint8 device tensors.  `pred` / `data.binary` that are still on
    the device (lazy.DeviceArray, what Predictor.predict and DatasetLoader hand out) are used where they are."""
    from ..lazy import DeviceArray, device_tensor_of, peek
    from ..runtime import get_context
    ctx = get_context()
    h, w = pred.shape
    from .postprocess import _pred_to_device
    if isinstance(pred, DeviceArray) and pred.on_device:
        d_pred = pred.device_tensor()
    else:
        d_pred, _ = _pred_to_device(np.asarray(pred), ctx.device)   # int64 from np.argmax: checked and narrowed on the device
    binary = peek(data, "binary")
    if not isinstance(binary, DeviceArray):
        binary = np.asarray(binary)
        if binary.dtype != np.uint8 and binary.dtype != np.bool_:
            binary = (binary != 0)
    d_bin = device_tensor_of(binary, ctx.device)
    if tuple(d_bin.shape) != (h, w):
        raise ValueError(f"data.binary {tuple(d_bin.shape)} and the class map {(h, w)} differ in shape")
    return ctx, d_pred.contiguous(), d_bin.contiguous()


def _masks_on_device(data: SingleData, pred, color_map: ColorMap):
    """The three masks as one (3, H, W, 3) uint8 device tensor: color, overlay, inverted."""
    import torch
    ctx, d_pred, d_bin = _class_map_and_binary_on_device(data, pred)
    h, w = d_pred.shape
    outs = torch.empty((3, h, w, 3), dtype=torch.uint8, device=d_pred.device)
    ctx.masks(d_pred, d_bin, 1, h, w, color_map.lut(), outs[0], outs[1], outs[2])
    return ctx, outs


def encode_png(images, level: int = 1) -> list:
    """PNG files (bytes) of a stack of equally sized uint8 images (n, H, W) or (n, H, W, 3|4), built on the device
    (pcs_png_encode).  level 1 (default): Sub filter + fixed-Huffman deflate with run-length matches - class-colour
    masks come out 30-100x smaller than raw; level 0: stored blocks, raw size + 0.2 %.  Lossless like any PNG.
    `images` may be a numpy array or a CUDA tensor."""
    import torch
    from ..runtime import get_context
    ctx = get_context()
    d_img = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images)).to(f"cuda:{ctx.device}")
    if d_img.dtype != torch.uint8 or d_img.dim() not in (3, 4):
        raise ValueError("encode_png expects uint8 images (n, H, W) or (n, H, W, C)")
    n, h, w = d_img.shape[:3]
    c = 1 if d_img.dim() == 3 else d_img.shape[3]
    if level and ctx.png_bytes(h, w, c, level) == 0:
        level = 0                                        # scanlines too long for the run tables: stored blocks
    size = ctx.png_bytes(h, w, c, level)
    if size == 0:
        raise ValueError(f"images of shape {h} x {w} x {c} cannot be written as PNG by the device encoder")
    stride = (size + 255) // 256 * 256
    d_out = torch.empty((n, stride), dtype=torch.uint8, device=d_img.device)
    d_sizes = torch.empty((n,), dtype=torch.int64, device=d_img.device)
    ctx.png_encode(d_img.contiguous(), n, h, w, c, d_out, stride, d_sizes, level)
    sizes = d_sizes.cpu().tolist()
    if level:                                            # only the bytes of the files travel, not the worst-case buffers
        return [d_out[i, :sizes[i]].cpu().numpy().tobytes() for i in range(n)]
    from ..runtime import to_host
    files = to_host(d_out)
    return [files[i, :sizes[i]].tobytes() for i in range(n)]


def generate_output_masks(data: SingleData, pred: np.ndarray, color_map: ColorMap) -> Masks:
    """output.py:44-60 through the device epilogue (pcs_masks)."""
    _, outs = _masks_on_device(data, pred, color_map)
    from ..runtime import results_to_host
    # fg_color_mask[foreground != 0] = 0 is arithmetically the inverted overlay (output.py:50-53)
    (color, overlay, inverted), fg_color = results_to_host(outs, outs[2], site="masks")
    return Masks(color=color, overlay=overlay, inverted_overlay=inverted, fg_color_mask=fg_color)


def output_data(output_dir, pred, data: SingleData, color_map):
    """output.py:20-41.  `.png` targets (the frontend's default): ONE library call per page (pcs_output_pages) builds the
    three masks and their PNG files on the device; the library's worker threads write them, so the call returns before
    the files exist -- `flush_outputs()` waits for them (it also runs at interpreter exit; PCSEG_OUTPUT_ASYNC=0 makes
    every call wait).  Other extensions go through cv2.imwrite in place of skimage.io.imsave."""
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
    categories = ("color", "overlay", "inverted")
    if filename.lower().endswith(".png"):
        from .. import pipeline
        ctx, d_pred, d_bin = _class_map_and_binary_on_device(data, pred)
        h, w = d_pred.shape
        ctx.output_pages(d_pred, d_bin, 1, h, w, color_map.lut(), [os.path.join(output_dir, c, filename) for c in categories])
        if not pipeline.OUTPUT_ASYNC:
            ctx.output_flush()
        return
    _, d_masks = _masks_on_device(data, pred, color_map)
    import cv2
    for category, img in zip(categories, d_masks.cpu().numpy()):
        path = os.path.join(output_dir, category, filename)
        if not cv2.imwrite(path, np.ascontiguousarray(img[..., ::-1])):
            raise IOError(f"could not write {path}")


def flush_outputs():
    """Waits until every PNG file handed to output_data is on disk (they are written by a background thread; also runs
    at interpreter exit).  Not part of the reference API, whose output_data writes synchronously."""
    from ..pipeline import flush_outputs as f
    f()


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
