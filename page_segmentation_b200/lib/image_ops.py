"""Mirror of ocr4all_pixel_classifier/lib/image_ops.py:58-82 (`compute_char_height`), the producer of the
`line_height_px` normalisation input, on the B200 path (pcs_char_height)."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

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
