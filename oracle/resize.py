"""Restatement of skimage.transform.resize / rescale (scikit-image 0.17.2) for
the two call patterns of the reference (test infrastructure, see oracle/__init__).

Reference call sites: ocr4all_pixel_classifier/lib/dataset.py:114-119
(`rescale(order=0, anti_aliasing=False, preserve_range=True)`), :122-128
(`resize(order=3, anti_aliasing=len(np.unique(img)) > 2, preserve_range=True)`),
lib/util.py:21-29 (`resize(order=0, anti_aliasing=False, preserve_range=True)`).

skimage 0.17.2 is not installed; this follows its published algorithm
(unpinned):
  * rescale: output_shape = np.round(scale * shape) (half-to-even), then resize;
  * resize (2-D): image -> float64; factors = in/out; optional
    scipy.ndimage.gaussian_filter(sigma=(factors-1)/2, mode='mirror'); affine
    warp output->input  r = f_r*y + (0.5*f_r - 0.5),  c = f_c*x + (0.5*f_c - 0.5)
    through `_warp_fast`, mode='reflect' (mirror without edge repeat);
  * order 0: pixel at C round() (half away from zero) of (r, c);
  * order 3: 4x4 cubic convolution (Catmull-Rom form of `cubic_interpolation`),
    columns first, then rows;
  * clip=True: result clamped to [image.min(), image.max()] for order > 0.
The affine estimate() of skimage adds ~1e-16 relative noise to the matrix; that
noise is not modelled (the ideal matrix is used).
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi


def rescale_output_shape(shape, scale: float):
    return tuple(int(v) for v in np.round(scale * np.asarray(shape, dtype=np.float64)))


def _reflect(coord: np.ndarray, dim: int) -> np.ndarray:
    """`coord_map(dim, coord, 'R')` of skimage/_shared/interpolation.pxd."""
    if dim == 1:
        return np.zeros_like(coord)
    cmax = dim - 1
    c = coord.astype(np.int64)
    out = c.copy()
    neg = c < 0
    if neg.any():
        a = -c[neg]
        out[neg] = np.where((a // cmax) % 2 != 0, cmax - (a % cmax), a % cmax)
    big = c > cmax
    if big.any():
        a = c[big]
        out[big] = np.where((a // cmax) % 2 != 0, cmax - (a % cmax), a % cmax)
    return out


def _c_round(v: np.ndarray) -> np.ndarray:
    """C round(): half away from zero."""
    return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int64)


def _coords(n_out: int, n_in: int) -> np.ndarray:
    f = np.float64(n_in) / np.float64(n_out)
    t = 0.5 * f - 0.5
    return f * np.arange(n_out, dtype=np.float64) + t


def _cubic(x: np.ndarray, f0, f1, f2, f3):
    """`cubic_interpolation(x, f)` of skimage/_shared/interpolation.pxd."""
    return f1 + 0.5 * x * (f2 - f0 + x * (2.0 * f0 - 5.0 * f1 + 4.0 * f2 - f3
                                          + x * (3.0 * (f1 - f2) + f3 - f0)))


def resize(image: np.ndarray, output_shape, order: int = 0, anti_aliasing: bool = False) -> np.ndarray:
    """skimage.transform.resize(image, output_shape, order=order, mode='reflect',
    cval=0, clip=True, preserve_range=True, anti_aliasing=anti_aliasing) for 2-D
    input; returns float64."""
    assert image.ndim == 2
    out_h, out_w = int(output_shape[0]), int(output_shape[1])
    in_h, in_w = image.shape
    img = image.astype(np.float64)
    if anti_aliasing:
        factors = np.array([in_h / out_h, in_w / out_w], dtype=np.float64)
        sigma = np.maximum(0, (factors - 1) / 2)
        img = ndi.gaussian_filter(img, sigma, cval=0, mode="mirror")
    r = _coords(out_h, in_h)
    c = _coords(out_w, in_w)
    if order == 0:
        ri = _reflect(_c_round(r), in_h)
        ci = _reflect(_c_round(c), in_w)
        return img[ri[:, None], ci[None, :]]
    if order != 3:
        raise NotImplementedError("the reference only uses order 0 and 3")
    r0 = np.floor(r).astype(np.int64)
    c0 = np.floor(c).astype(np.int64)
    xr = r - r0
    xc = c - c0
    rows = [_reflect(r0 - 1 + k, in_h) for k in range(4)]
    cols = [_reflect(c0 - 1 + k, in_w) for k in range(4)]
    fr = []
    for pr in range(4):
        line = img[rows[pr], :]                              # (out_h, in_w)
        fc = [line[:, cols[pc]] for pc in range(4)]          # each (out_h, out_w)
        fr.append(_cubic(xc[None, :], fc[0], fc[1], fc[2], fc[3]))
    out = _cubic(xr[:, None], fr[0], fr[1], fr[2], fr[3])
    return np.clip(out, img.min(), img.max())


def rescale(image: np.ndarray, scale: float, order: int = 0, anti_aliasing: bool = False) -> np.ndarray:
    """skimage.transform.rescale(..., multichannel=False, preserve_range=True)."""
    return resize(image, rescale_output_shape(image.shape, scale), order=order,
                  anti_aliasing=anti_aliasing)
