# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/pc_segmentation.py: region extraction from the `inverted`
colour image (`find_segments` :24-60, `dilate` :63-67, `get_text_contours` :70-116).

Pixel work runs on the device (pcs_segment_masks, pcs_integral_image, pcs_text_regions); what is left on
the host is what the survey leaves there: the XY-cut recursion over projection profiles (lib/xycut.py) and
OpenCV's contour tracing / polygon filling, which are sequential border-following algorithms.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from .colors import ColorMap
from .xycut import CVContour, RectSegment, xy_cut_from_integral

ColorMapping = Dict[str, np.ndarray]


def seg(left_upper: Tuple[int, int], right_lower: Tuple[int, int]) -> RectSegment:
    return RectSegment(left_upper[0], left_upper[1], right_lower[0], right_lower[1])


DEFAULT_COLOR_MAPPING = {
    "image": np.array([0, 255, 0]),
    "text": np.array([0, 0, 255]),
}


def _rgb_u8(image: np.ndarray) -> np.ndarray:
    image = np.asarray(image)
    if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
        raise ValueError("expected an (H, W, 3) uint8 colour image")
    return np.ascontiguousarray(image)


def find_segments(orig_height: int, image: np.ndarray, char_height: int, resize_height: int,
                  color_map: ColorMap, only_images=False, device: Optional[int] = None) \
        -> Tuple[List[RectSegment], List[RectSegment]]:
    """pc_segmentation.py:24-60 -> (segments_text, segments_image)."""
    from .. import runtime
    torch = runtime._torch()
    ctx = runtime.get_context(device)
    image = _rgb_u8(image)
    # working size and thresholds exactly as the reference derives them (:27-42)
    scale_percent = resize_height / image.shape[0]
    height = resize_height
    width = int(image.shape[1] * scale_percent)
    absolute_resize_factor = height / orig_height
    px_threshold_line = int(char_height * absolute_resize_factor)
    px_threshold_column = int(char_height * absolute_resize_factor)
    split_size_horizontal = int(char_height * 2 * absolute_resize_factor)
    split_size_vertical = int(char_height * absolute_resize_factor)

    labels = ["image"] if only_images else ["image", "text"]
    colours = np.array([color_map.color_for_label(name) for name in labels], dtype=np.uint8)
    d_rgb = torch.from_numpy(image).to(f"cuda:{ctx.device}")
    d_masks = torch.empty((len(labels), height, width), dtype=torch.uint8, device=d_rgb.device)
    d_sat = torch.empty((len(labels), height + 1, width + 1), dtype=torch.int32, device=d_rgb.device)
    ctx.segment_masks(d_rgb, image.shape[0], image.shape[1], height, width, colours, d_masks)
    ctx.integral_image(d_masks, len(labels), height, width, d_sat)
    sat = d_sat.cpu().numpy()

    def cut(i):
        found = xy_cut_from_integral(sat[i], px_threshold_line, px_threshold_column, split_size_horizontal,
                                     split_size_vertical)
        return [s.scale(1.0 / absolute_resize_factor) for s in found]

    segments_image = cut(0)
    segments_text = [] if only_images else cut(1)
    return segments_text, segments_image


def dilate(bin_image: np.ndarray, device: Optional[int] = None):
    """pc_segmentation.py:63-67: 3x3 rectangular dilation of a uint8 image (any channel count)."""
    from .. import runtime
    torch = runtime._torch()
    ctx = runtime.get_context(device)
    a = np.ascontiguousarray(bin_image)
    if a.dtype != np.uint8 or a.ndim not in (2, 3):
        raise ValueError("dilate expects a uint8 image")
    h, w = a.shape[:2]
    c = 1 if a.ndim == 2 else a.shape[2]
    d_src = torch.from_numpy(a).to(f"cuda:{ctx.device}")
    d_dst = torch.empty_like(d_src)
    ctx.dilate3x3(d_src, h, w, c, d_dst)
    return d_dst.cpu().numpy()


def text_region_masks(image: np.ndarray, char_height: int, color_map: ColorMap, device: Optional[int] = None):
    """Device part of get_text_contours (:71-96): returns (255 - image_after_opening, region_text), uint8."""
    from .. import runtime
    torch = runtime._torch()
    ctx = runtime.get_context(device)
    image = _rgb_u8(image)
    colour = np.array(color_map.color_for_label("text"), dtype=np.uint8)
    h, w = image.shape[:2]
    d_rgb = torch.from_numpy(image).to(f"cuda:{ctx.device}")
    d_text_inv = torch.empty((h, w), dtype=torch.uint8, device=d_rgb.device)
    d_region = torch.empty((h, w), dtype=torch.uint8, device=d_rgb.device)
    ctx.text_regions(d_rgb, h, w, colour, int(char_height), int(char_height / 3), int(char_height / 1.1),
                     d_text_inv, d_region)
    return d_text_inv.cpu().numpy(), d_region.cpu().numpy()


def get_text_contours(image, char_height: int, color_map: ColorMap, device: Optional[int] = None):
    """pc_segmentation.py:70-116 -> list of CVContour, in the reference's order."""
    import cv2
    canvas, region_text = text_region_masks(image, char_height, color_map, device)
    contours, _ = cv2.findContours(region_text, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    # the reference's draw colour is `color.tolist().reverse()`, i.e. None (:100), which OpenCV takes as an
    # all-zero scalar: every traced region, holes included, is painted black onto the canvas
    for contour in contours:
        cv2.drawContours(canvas, [contour], 0, 0, cv2.FILLED)
    canvas = cv2.copyMakeBorder(canvas, 1, 1, 1, 1, cv2.BORDER_CONSTANT, value=(255, 255, 255))
    contours, _ = cv2.findContours(canvas, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    # the first contour is the page frame; the rest are returned back to front (:114-116)
    return list(map(CVContour, contours[1:][::-1]))
