# Test infrastructure: restates functions of ocr4all_pixel_classifier (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation,
# (c) its authors, licensed Apache-2.0 OR GPL-3.0-or-later) on the CPU as the parity oracle; never imported by the product path.
"""CPU restatement of the region-extraction stage (test infrastructure, see oracle/__init__).

Follows ocr4all_pixel_classifier/lib/pc_segmentation.py (`find_segments` :24-60, `dilate` :63-67,
`get_text_contours` :70-116) and lib/xycut.py (`do_xy_cut` :95-109, `recursive_cut` :127-161, `_get_gaps`
:112-117, `_get_segments` :164-173, `_relative_seg` :120-124, `RectSegment.scale` :43-49), with cv2
(installed here, 4.13; the reference pins 4.5.5.62) as the live implementation of its OpenCV calls.

PINNED: these two reference modules import with numpy + cv2 alone, so tests/golden/make_reference_golden.py
runs the REFERENCE's own functions in the build container (ocr4all.colors.ColorMap, which is absent, is
replaced by the stand-in of SURVEY appendix D) and commits their outputs as tests/golden/ref_regions.npz;
tests/test_reference_pins.py holds this restatement to those vectors and, where /root/reference is
mounted, to the live reference on fresh random cases.

Rectangles are plain tuples (x_start, y_start, x_end, y_end) in the reference's own (mixed-up) convention.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import cv2
import numpy as np

Rect = Tuple[int, int, int, int]


# ---------------------------------------------------------------------------
# xycut.py
# ---------------------------------------------------------------------------
def _false_runs(flags: np.ndarray) -> List[Tuple[int, int]]:
    """(start, length) of every maximal run of False (xycut.py:112-117)."""
    runs, start = [], None
    for i, f in enumerate(flags.tolist()):
        if not f and start is None:
            start = i
        elif f and start is not None:
            runs.append((start, i - start))
            start = None
    if start is not None:
        runs.append((start, len(flags) - start))
    return runs


def _intervals(runs: Sequence[Tuple[int, int]], length: int, px_threshold: int, split_size: int) -> List[Tuple[int, int]]:
    """xycut.py:164-173: stretches between the runs of at least `split_size`, kept when longer than the threshold."""
    marks = [(0, 0)] + [r for r in runs if r[1] >= split_size] + [(length, 0)]
    out = []
    for (s0, l0), (s1, _l1) in zip(marks, marks[1:]):
        if s1 - (s0 + l0) > px_threshold:
            out.append((s0 + l0, s1))
    return out


def _rect(shape, start, end, pos) -> Rect:
    """xycut.py:120-124 (x from pos[1], y from pos[0], extent shape[1] - whatever the axis)."""
    return (pos[1] + start, pos[0], pos[1] + end, pos[0] + shape[1])


def recursive_cut(image: np.ndarray, threshold, split_size, axis=0, position=(0, 0), end_recurse=False) -> List[Rect]:
    """xycut.py:127-161."""
    enough = np.count_nonzero(image, axis=axis) >= threshold[axis]
    runs = _false_runs(enough)
    if not runs:
        return [_rect(image.shape, 0, image.shape[axis], position)]
    parts = _intervals(runs, image.shape[axis], threshold[axis], split_size[axis])
    if end_recurse:
        return [_rect(image.shape, a, b, position) for a, b in parts]
    out: List[Rect] = []
    for a, b in parts:
        if b - a > threshold[axis]:
            if axis == 1:
                sub, pos = image[a:b, :], (position[0], position[1] + a)
            else:
                sub, pos = image[:, a:b], (position[0] + a, position[1])
            if 0 in sub.shape:
                return out
            out += recursive_cut(sub, threshold, split_size, 1 - axis, pos, len(parts) == 1)
    return out


def do_xy_cut(binary_image, px_threshold_line, px_threshold_column, split_size_horizontal, split_size_vertical) -> List[Rect]:
    """xycut.py:95-109."""
    return recursive_cut(np.asarray(binary_image), (px_threshold_line, px_threshold_column),
                         (split_size_horizontal, split_size_vertical), 0)


def scale_rect(r: Rect, factor: float) -> Rect:
    """RectSegment.scale, xycut.py:43-49 (truncation)."""
    return tuple(int(v * factor) for v in r)


def integral_image(mask: np.ndarray) -> np.ndarray:
    """sat[(H+1) x (W+1)] int32 of (mask != 0): what the device hands to the host recursion."""
    sat = np.zeros((mask.shape[0] + 1, mask.shape[1] + 1), dtype=np.int32)
    sat[1:, 1:] = np.cumsum(np.cumsum(mask != 0, axis=0, dtype=np.int64), axis=1).astype(np.int32)
    return sat


# ---------------------------------------------------------------------------
# pc_segmentation.py
# ---------------------------------------------------------------------------
def resize_nearest_index(n_dst: int, n_src: int) -> np.ndarray:
    """cv::resize INTER_NEAREST source index table (resizeNN): min(floor(i * (1 / (n_dst / n_src))), n_src - 1),
    restated so that cv2.resize itself is pinned by tests/test_oracle_known_answers.py."""
    inv = 1.0 / (float(n_dst) / float(n_src))
    return np.minimum(np.floor(np.arange(n_dst, dtype=np.float64) * inv).astype(np.int64), n_src - 1)


def segment_masks(image: np.ndarray, height: int, width: int, colours: Sequence[Sequence[int]]) -> np.ndarray:
    """pc_segmentation.py:31-32 + filter_label (:48, :56): [m][height][width] uint8 {0,1}."""
    small = cv2.resize(image, (width, height), interpolation=cv2.INTER_NEAREST)
    small = cv2.dilate(small, np.ones((3, 3), np.uint8), iterations=1)
    return np.stack([np.all(small == np.asarray(c, dtype=np.uint8), axis=-1) for c in colours]).astype(np.uint8)


def find_segments(orig_height: int, image: np.ndarray, char_height: int, resize_height: int, colour_image, colour_text,
                  only_images: bool = False):
    """pc_segmentation.py:24-60 -> (rects_text, rects_image)."""
    scale_percent = resize_height / image.shape[0]
    height = resize_height
    width = int(image.shape[1] * scale_percent)
    factor = height / orig_height
    thr_line = thr_col = int(char_height * factor)
    split_h, split_v = int(char_height * 2 * factor), int(char_height * factor)
    masks = segment_masks(image, height, width, [colour_image, colour_text])

    def cut(m):
        return [scale_rect(r, 1.0 / factor) for r in do_xy_cut(m.astype(bool), thr_line, thr_col, split_h, split_v)]

    return ([] if only_images else cut(masks[1])), cut(masks[0])


def text_region_masks(image: np.ndarray, char_height: int, colour) -> Tuple[np.ndarray, np.ndarray]:
    """pc_segmentation.py:71-96 -> (255 - image after the opening, region_text)."""
    colour = np.array(colour)
    img = cv2.inRange(image, colour, colour)
    k = cv2.getStructuringElement(cv2.MORPH_RECT, (int(char_height), int(char_height)))
    img = cv2.morphologyEx(img, cv2.MORPH_CLOSE, k)
    k = cv2.getStructuringElement(cv2.MORPH_RECT, (int(char_height / 3), int(char_height / 3)))
    img = cv2.morphologyEx(img, cv2.MORPH_OPEN, k)
    k = cv2.getStructuringElement(cv2.MORPH_RECT, (int(char_height / 1.1), int(char_height / 1.1)))
    region_chars = cv2.dilate(img, k, iterations=1)
    region_text = cv2.morphologyEx(region_chars, cv2.MORPH_CLOSE, k)
    return 255 - img, region_text


def get_text_contours(image: np.ndarray, char_height: int, colour) -> List[np.ndarray]:
    """pc_segmentation.py:70-116 -> squeezed contour arrays in the reference's order."""
    canvas, region_text = text_region_masks(image, char_height, colour)
    contours, _ = cv2.findContours(region_text, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    for c in contours:
        cv2.drawContours(canvas, [c], 0, None, cv2.FILLED)       # :100 `list.reverse()` is None: OpenCV paints zeros
    canvas = cv2.copyMakeBorder(canvas, 1, 1, 1, 1, cv2.BORDER_CONSTANT, value=(255, 255, 255))
    contours, _ = cv2.findContours(canvas, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    return [np.squeeze(c) for c in contours[1:][::-1]]


# numpy restatement of OpenCV's rectangular erode / dilate (anchor k // 2, borders ignored), to pin the window
# orientation for even-sized elements that the device kernels implement
def rect_morph(img: np.ndarray, k: int, erode: bool) -> np.ndarray:
    h, w = img.shape
    a = k // 2
    fill = 255 if erode else 0
    pad = np.full((h + k, w + k), fill, dtype=np.uint8)
    pad[a:a + h, a:a + w] = img
    out = np.full((h, w), fill, dtype=np.uint8)
    for i in range(k):
        for j in range(k):
            win = pad[i:i + h, j:j + w]
            out = np.minimum(out, win) if erode else np.maximum(out, win)
    return out
