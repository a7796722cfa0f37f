# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/xycut.py: region types (:14-67) and the recursive XY-cut
(`do_xy_cut` :95-109, `recursive_cut` :127-161, `_get_gaps` :112-117, `_get_segments` :164-173).

Every `np.count_nonzero(sub_image, axis)` of the recursion (:135) is a difference of two rows / columns of
one summed-area table, which the device builds once per mask (pcs_integral_image); the recursion itself is
data-dependent control flow over a few hundred profile entries and stays on the host, touching no pixel.

The reference's coordinate bookkeeping is reproduced as it is, not as it was probably meant: the segment
list of either axis is closed with `image.shape[axis]` (:140) although the profile runs along the other
axis, `_relative_seg` (:120-124) always adds the interval to `pos[1]` and takes the extent from `shape[1]`,
and a cut along axis 0 advances `pos[0]` (:154).  Consumers get exactly the rectangles the reference yields.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, TypeVar, Union

import numpy as np

RGBColor = Tuple[int, int, int]


class Region(ABC):
    @abstractmethod
    def polygon_coords(self) -> Union[List[Tuple[int, int]], np.ndarray]:
        ...

    @abstractmethod
    def scale(self, factor: float) -> "Region":
        ...


@dataclass
class CVContour(Region):
    contour: np.ndarray

    def __post_init__(self):
        self.contour = np.squeeze(self.contour)

    def polygon_coords(self):
        return np.squeeze(self.contour)

    def scale(self, factor: float) -> "CVContour":
        return CVContour((self.contour * factor).astype("int32"))


@dataclass
class RectSegment(Region):
    x_start: int
    y_start: int
    x_end: int
    y_end: int

    def of(self, image: np.ndarray):
        return image[self.y_start:self.y_end, self.x_start:self.x_end]

    def scale(self, factor: float) -> "RectSegment":
        return RectSegment(*(int(v * factor) for v in (self.x_start, self.y_start, self.x_end, self.y_end)))

    def as_xy(self) -> List[Tuple[int, int]]:
        return [(self.y_start, self.x_start), (self.y_end, self.x_end)]

    def polygon_coords(self):
        # clockwise from the upper left corner
        return [(self.x_start, self.y_start), (self.x_end, self.y_start),
                (self.x_end, self.y_end), (self.x_start, self.y_end)]


AnyRegion = TypeVar("AnyRegion", Region, RectSegment, CVContour)


@dataclass
class Segment1D:
    start: int
    end: int

    def __len__(self):
        return self.end - self.start


@dataclass
class Gap:
    start: int
    length: int


def single_color(image: np.ndarray, color: Union[int, np.ndarray]):
    mask = image == color
    return mask.all(axis=-1) if image.ndim > 2 else mask


# ---------------------------------------------------------------------------
# recursion over a summed-area table
# ---------------------------------------------------------------------------
class _Profiles:
    """Projection profiles of sub-rectangles from sat[(H+1) x (W+1)] (int32, first row/column zero)."""

    def __init__(self, sat: np.ndarray):
        self.sat = sat
        self.shape = (sat.shape[0] - 1, sat.shape[1] - 1)

    def counts(self, r0: int, r1: int, c0: int, c1: int, axis: int) -> np.ndarray:
        s = self.sat
        if axis == 0:           # per column, counted down the rows
            col = s[r1, c0:c1 + 1] - s[r0, c0:c1 + 1]
            return col[1:] - col[:-1]
        row = s[r0:r1 + 1, c1] - s[r0:r1 + 1, c0]
        return row[1:] - row[:-1]


def _get_gaps(indication: np.ndarray) -> List[Gap]:
    """Maximal runs of False (xycut.py:112-117)."""
    edge = np.diff(np.concatenate(([1], indication.astype(np.int8), [1])))
    starts, ends = np.flatnonzero(edge == -1), np.flatnonzero(edge == 1)
    return [Gap(int(a), int(b - a)) for a, b in zip(starts, ends)]


def _get_segments(gaps: Sequence[Gap], length: int, px_threshold, split_size) -> List[Segment1D]:
    """Intervals between the gaps that are wide enough to split at (xycut.py:164-173)."""
    out, cursor = [], 0
    for g in [g for g in gaps if g.length >= split_size] + [Gap(length, 0)]:
        if g.start - cursor > px_threshold:
            out.append(Segment1D(cursor, g.start))
        cursor = g.start + g.length
    return out


def _relative_seg(shape, start, end, pos) -> RectSegment:
    return RectSegment(x_start=pos[1] + start, x_end=pos[1] + end, y_start=pos[0], y_end=pos[0] + shape[1])


def _cut(prof: _Profiles, rect, threshold, split_size, axis, position, end_recurse) -> List[RectSegment]:
    r0, r1, c0, c1 = rect
    shape = (r1 - r0, c1 - c0)
    enough = prof.counts(r0, r1, c0, c1, axis) >= threshold[axis]
    gaps = _get_gaps(enough)
    if not gaps:
        return [_relative_seg(shape, 0, shape[axis], position)]
    pieces = _get_segments(gaps, shape[axis], threshold[axis], split_size[axis])
    if end_recurse:
        return [_relative_seg(shape, s.start, s.end, position) for s in pieces]
    found: List[RectSegment] = []
    for s in pieces:
        if len(s) <= threshold[axis]:
            continue
        if axis == 1:           # numpy slicing clamps to the extent of the parent
            sub = (min(r0 + s.start, r1), min(r0 + s.end, r1), c0, c1)
            pos = (position[0], position[1] + s.start)
        else:
            sub = (r0, r1, min(c0 + s.start, c1), min(c0 + s.end, c1))
            pos = (position[0] + s.start, position[1])
        if sub[1] - sub[0] == 0 or sub[3] - sub[2] == 0:
            return found        # xycut.py:156: the reference stops here, dropping the remaining pieces
        found += _cut(prof, sub, threshold, split_size, 1 - axis, pos, len(pieces) == 1)
    return found


def xy_cut_from_integral(sat: np.ndarray, px_threshold_line: int, px_threshold_column: int,
                         split_size_horizontal: int, split_size_vertical: int) -> List[RectSegment]:
    """The recursion of `do_xy_cut` on a summed-area table already fetched from the device."""
    prof = _Profiles(np.asarray(sat))
    h, w = prof.shape
    return _cut(prof, (0, h, 0, w), (px_threshold_line, px_threshold_column),
                (split_size_horizontal, split_size_vertical), 0, (0, 0), False)


def integral_image(masks: np.ndarray, device: Optional[int] = None) -> np.ndarray:
    """Summed-area tables of (mask != 0) for a stack of masks [n][H][W], computed on the device."""
    from .. import runtime
    torch = runtime._torch()
    ctx = runtime.get_context(device)
    n, h, w = masks.shape
    d_mask = runtime.to_device_u8(np.asarray(masks) != 0, ctx.device)
    d_sat = torch.empty((n, h + 1, w + 1), dtype=torch.int32, device=d_mask.device)
    ctx.integral_image(d_mask, n, h, w, d_sat)
    return d_sat.cpu().numpy()


def do_xy_cut(binary_image: np.ndarray, px_threshold_line: int, px_threshold_column: int,
              split_size_horizontal: int, split_size_vertical: int) -> List[RectSegment]:
    """xycut.py:95-109: rectangular regions of a boolean image (True / non-zero = foreground)."""
    binary_image = np.asarray(binary_image)
    if binary_image.ndim != 2:
        raise ValueError("do_xy_cut expects a 2-D boolean image")
    sat = integral_image(binary_image[None])[0]
    return xy_cut_from_integral(sat, px_threshold_line, px_threshold_column, split_size_horizontal, split_size_vertical)
