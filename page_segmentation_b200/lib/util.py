# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/util.py."""
import numpy as np


def gray_to_rgb(img):
    """util.py:4-9."""
    if len(img.shape) != 3 or img.shape[2] != 3:
        img = img[..., np.newaxis]
        return np.concatenate(3 * (img,), axis=-1)
    return img


def image_to_batch(img):
    """util.py:12-18."""
    if len(img.shape) == 2:
        return np.expand_dims(np.expand_dims(img, axis=0), axis=-1)
    assert img.shape != 3
    return np.expand_dims(img, axis=0)


def preserving_resize(image: np.ndarray, target_shape) -> np.ndarray:
    """util.py:21-29: order-0 resize without anti-aliasing or range change.

    Runs the device kernel behind pcs_resize_nearest.  Like skimage the result
    is float64 (callers in output.py cast it back); integer inputs in 0..255 and
    bool are supported, which covers image / pred / binary planes."""
    from ..runtime import resize_nearest_plane
    return resize_nearest_plane(np.asarray(image), tuple(int(v) for v in target_shape)).astype(np.float64)
