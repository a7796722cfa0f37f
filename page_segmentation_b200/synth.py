"""Seeded synthetic pages and random-init weights.

There is no network access for datasets or checkpoints, so benchmarks and parity
tests run on synthetic A4-300dpi pages (2480x3508, paper=255, ink=0) and on
seeded random-init weights of the reference architectures.  Both the device path
and the CPU oracle consume exactly these arrays.

Weight shapes follow what Keras stores for the reference graphs
(reference: ocr4all_pixel_classifier/lib/model.py:45-92 fcn_skip, :206-234 fcn,
:151-203 unet): Conv2D kernels are (kh, kw, C_in, C_out), Conv2DTranspose
kernels are (kh, kw, C_out, C_in).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

A4_H, A4_W = 3508, 2480          # rows, cols of an A4 page at 300 dpi
A4_MPX = A4_H * A4_W / 1e6       # 8.69984 Mpx = one "normalised page"

# (layer name, kind, k, C_in, C_out, activation) in Keras creation order.
# kind: "conv" = Conv2D same s1; "deconv" = Conv2DTranspose same s1;
#       "deconv_s2" = Conv2DTranspose 2x2 stride 2; "logits" = 1x1 Conv2D.
FCN_SKIP_LAYERS = [
    ("conv1", "conv", 5, 1, 20, "relu"),
    ("conv2", "conv", 5, 20, 30, "linear"),
    ("conv3", "conv", 5, 30, 40, "relu"),
    ("conv4", "conv", 5, 40, 40, "linear"),
    ("conv5", "conv", 5, 40, 60, "relu"),
    ("conv6", "conv", 5, 60, 60, "linear"),
    ("conv7", "conv", 5, 60, 80, "relu"),
    ("deconv1", "deconv", 5, 80, 80, "relu"),
    ("deconv2", "deconv_s2", 2, 80, 60, "relu"),
    ("deconv3", "deconv", 5, 120, 40, "relu"),
    ("deconv4", "deconv_s2", 2, 100, 30, "relu"),
    ("deconv5", "deconv_s2", 2, 70, 20, "linear"),
    ("logits", "logits", 1, 50, None, "linear"),
]

FCN_LAYERS = [
    ("conv1", "conv", 5, 1, 20, "relu"),
    ("conv2", "conv", 5, 20, 30, "linear"),
    ("conv3", "conv", 5, 30, 40, "relu"),
    ("conv4", "conv", 5, 40, 40, "linear"),
    ("conv5", "conv", 5, 40, 60, "relu"),
    ("conv6", "conv", 5, 60, 60, "linear"),
    ("conv7", "conv", 5, 60, 80, "relu"),
    ("deconv1", "deconv", 5, 80, 80, "relu"),
    ("deconv2", "deconv_s2", 2, 80, 60, "relu"),
    ("deconv3", "deconv", 5, 60, 40, "relu"),
    ("deconv4", "deconv_s2", 2, 40, 30, "relu"),
    ("deconv5", "deconv_s2", 2, 30, 20, "linear"),
    ("logits", "logits", 1, 20, None, "linear"),
]

# U-Net: conv3x3 relu pairs, "up" = UpSampling2D(2) -> Conv2D(2x2 same, relu).
UNET_LAYERS = [
    ("conv1a", "conv", 3, 1, 64, "relu"), ("conv1b", "conv", 3, 64, 64, "relu"),
    ("conv2a", "conv", 3, 64, 128, "relu"), ("conv2b", "conv", 3, 128, 128, "relu"),
    ("conv3a", "conv", 3, 128, 256, "relu"), ("conv3b", "conv", 3, 256, 256, "relu"),
    ("conv4a", "conv", 3, 256, 512, "relu"), ("conv4b", "conv", 3, 512, 512, "relu"),
    ("conv5a", "conv", 3, 512, 1024, "relu"), ("conv5b", "conv", 3, 1024, 1024, "relu"),
    ("up6", "conv", 2, 1024, 512, "relu"),
    ("conv6a", "conv", 3, 1024, 512, "relu"), ("conv6b", "conv", 3, 512, 512, "relu"),
    ("up7", "conv", 2, 512, 256, "relu"),
    ("conv7a", "conv", 3, 512, 256, "relu"), ("conv7b", "conv", 3, 256, 256, "relu"),
    ("up8", "conv", 2, 256, 128, "relu"),
    ("conv8a", "conv", 3, 256, 128, "relu"), ("conv8b", "conv", 3, 128, 128, "relu"),
    ("up9", "conv", 2, 128, 64, "relu"),
    ("conv9a", "conv", 3, 128, 64, "relu"), ("conv9b", "conv", 3, 64, 64, "relu"),
    ("logits", "logits", 1, 64, None, "linear"),
]

ARCH_LAYERS = {"fcn_skip": FCN_SKIP_LAYERS, "fcn": FCN_LAYERS, "unet": UNET_LAYERS}

DEFAULT_LUT = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}


def layer_table(arch: str, n_classes: int):
    """Layer list with the logits C_out filled in."""
    return [(n, k, ks, ci, (n_classes if co is None else co), act)
            for (n, k, ks, ci, co, act) in ARCH_LAYERS[arch]]


def keras_kernel_shape(kind: str, k: int, c_in: int, c_out: int) -> Tuple[int, int, int, int]:
    if kind in ("deconv", "deconv_s2"):
        return (k, k, c_out, c_in)      # Conv2DTranspose stores (kh, kw, out, in)
    return (k, k, c_in, c_out)


def make_weights(arch: str = "fcn_skip", n_classes: int = 3, seed: int = 0,
                 bias_range: float = 0.1) -> List[Tuple[np.ndarray, np.ndarray]]:
    """Seeded random-init weights in Keras storage order and shapes (float32).

    FCN layers: Glorot-uniform, limit = sqrt(6 / (k*k*(C_in + C_out))) (Keras
    default initialiser).  U-Net convs: He-normal (truncated at 2 sigma,
    sigma = sqrt(2/fan_in)/0.87962566) as `kernel_initializer='he_normal'`
    (model.py:156-196), logits Glorot.  Biases are U(-bias_range, bias_range)
    instead of Keras' zeros so that class margins are not degenerate.
    """
    rng = np.random.default_rng(seed)
    out = []
    for (name, kind, k, c_in, c_out, _act) in layer_table(arch, n_classes):
        shape = keras_kernel_shape(kind, k, c_in, c_out)
        if arch == "unet" and kind != "logits":
            fan_in = k * k * c_in
            sigma = np.sqrt(2.0 / fan_in) / 0.87962566103423978
            w = rng.standard_normal(shape)
            bad = np.abs(w) > 2.0
            while bad.any():                       # truncated normal by resampling
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            w = (w * sigma).astype(np.float32)
        else:
            limit = np.sqrt(6.0 / (k * k * (c_in + c_out)))
            w = rng.uniform(-limit, limit, size=shape).astype(np.float32)
        b = rng.uniform(-bias_range, bias_range, size=(c_out,)).astype(np.float32)
        out.append((w, b))
    return out


def make_page(seed: int = 0, height: int = A4_H, width: int = A4_W,
              line_height_px: int = 18) -> np.ndarray:
    """One synthetic binarised page: uint8 (H, W), paper = 255, ink = 0.

    ~45 text "lines" of glyph-like rectangles/blobs, 1-2 noisy "image" blocks
    and 0.05 % salt-and-pepper specks.  The same array
    serves as grey image and as binary image (mirrors dataset.py:169-172).
    """
    rng = np.random.default_rng(seed)
    page = np.full((height, width), 255, dtype=np.uint8)
    margin = min(150, height // 8, width // 8)
    lh = max(2, int(line_height_px))

    # "image" rectangles with 50 % salt noise
    n_img = int(rng.integers(1, 3))
    img_boxes = []
    for _ in range(n_img):
        bh = int(rng.integers(min(300, max(2, height // 6)), min(900, max(3, height // 3)) + 1))
        bw = int(rng.integers(min(300, max(2, width // 6)), min(900, max(3, width // 2)) + 1))
        y0 = int(rng.integers(margin, max(margin + 1, height - margin - bh)))
        x0 = int(rng.integers(margin, max(margin + 1, width - margin - bw)))
        noise = rng.random((bh, bw)) < 0.5
        page[y0:y0 + bh, x0:x0 + bw] = np.where(noise, 0, 255).astype(np.uint8)[
            :page[y0:y0 + bh, x0:x0 + bw].shape[0], :page[y0:y0 + bh, x0:x0 + bw].shape[1]]
        img_boxes.append((y0, x0, bh, bw))

    # text lines
    y = margin
    while y + lh < height - margin:
        x = margin + int(rng.integers(0, 40))
        word_left = int(rng.integers(3, 9))
        while x + 16 < width - margin:
            gw = int(rng.integers(6, 17))
            gh = int(rng.integers(max(2, lh * 2 // 3), lh + 1))
            gy = y + (lh - gh)
            skip = any(by - lh <= gy <= by + bh and bx - 16 <= x <= bx + bw for (by, bx, bh, bw) in img_boxes)
            if not skip:
                glyph = rng.random((gh, gw)) < 0.72
                # carve a blob-ish glyph: solid frame-ish strokes
                glyph[:, :2] = True
                glyph[:2, :] |= rng.random(gw) < 0.8
                page[gy:gy + gh, x:x + gw] = np.where(glyph, 0, page[gy:gy + gh, x:x + gw])
            x += gw + int(rng.integers(3, 7))
            word_left -= 1
            if word_left == 0:
                x += int(rng.integers(20, 31))
                word_left = int(rng.integers(3, 9))
        y += int(rng.integers(60, 71)) if lh == 18 else int(lh * rng.uniform(3.3, 3.9))

    # salt-and-pepper specks, 0.05 %
    n_specks = int(0.0005 * height * width)
    ys = rng.integers(0, height, n_specks)
    xs = rng.integers(0, width, n_specks)
    page[ys, xs] = np.where(rng.random(n_specks) < 0.5, 0, 255).astype(np.uint8)
    return page


def make_grey_page(seed: int = 0, height: int = A4_H, width: int = A4_W,
                   line_height_px: int = 18) -> np.ndarray:
    """A grey-level variant (more than two levels => the reference's
    anti-aliasing branch, dataset.py:127) built from the binarised page."""
    rng = np.random.default_rng(seed + 7919)
    page = make_page(seed, height, width, line_height_px).astype(np.int32)
    shade = rng.integers(0, 48, size=page.shape)
    grey = np.where(page > 0, 255 - shade // 2, shade)
    return grey.astype(np.uint8)


def scaled_shape(h: int, w: int, scale: float) -> Tuple[int, int]:
    """np.round(scale * shape) (half-to-even) as skimage.transform.rescale does."""
    return int(np.round(scale * h)), int(np.round(scale * w))


def padded_shape(h: int, w: int, f: int = 32) -> Tuple[int, int]:
    """model.py:10-17 calculate_padding applied to (H, W)."""
    return h + (f - h % f) % f, w + (f - w % f) % f


def make_inverted_image(seed: int = 0, height: int = 700, width: int = 500, char_height: int = 18,
                        text_colour=(255, 0, 0), image_colour=(0, 255, 0)) -> np.ndarray:
    """A synthetic `inverted` output image (output.py:50-51: class colour where ink, black elsewhere) as the
    region-extraction stage reads it: text lines of glyph blobs in `text_colour` arranged in one or two columns,
    one or two noisy picture blocks in `image_colour`, and a sprinkle of misclassified specks of either colour."""
    rng = np.random.default_rng(seed + 104729)
    img = np.zeros((height, width, 3), dtype=np.uint8)
    ch = max(3, int(char_height))
    margin = max(4, min(height, width) // 12)
    two_cols = bool(rng.integers(0, 2)) and width > 20 * ch
    gutter = 4 * ch
    cols = [(margin, width - margin)] if not two_cols else [(margin, width // 2 - gutter // 2), (width // 2 + gutter // 2, width - margin)]
    # picture blocks at the bottom of a column
    pic_top = {}
    for ci, (x0, x1) in enumerate(cols):
        if rng.random() < 0.7:
            bh = int(rng.integers(height // 6, height // 3))
            top = height - margin - bh
            block = rng.random((bh, x1 - x0)) < 0.5
            img[top:top + bh, x0:x1][block] = image_colour
            pic_top[ci] = top
    for ci, (x0, x1) in enumerate(cols):
        y = margin
        bottom = pic_top.get(ci, height - margin) - 2 * ch
        para_left = int(rng.integers(4, 9))
        while y + ch < bottom:
            x = x0 + int(rng.integers(0, ch))
            while x + ch < x1:
                gw = int(rng.integers(max(2, ch // 3), max(3, ch * 9 // 10)))
                gh = int(rng.integers(max(2, ch * 2 // 3), ch + 1))
                glyph = rng.random((gh, gw)) < 0.7
                glyph[:, :max(1, gw // 4)] = True
                img[y + ch - gh:y + ch, x:x + gw][glyph[:, :max(0, min(gw, x1 - x))]] = text_colour
                x += gw + int(rng.integers(1, max(2, ch // 4)))
                if rng.random() < 0.15:
                    x += ch
            y += int(ch * rng.uniform(1.5, 1.9))
            para_left -= 1
            if para_left == 0:
                y += int(ch * rng.uniform(2.5, 4.0))
                para_left = int(rng.integers(4, 9))
    n = int(0.001 * height * width)
    ys, xs = rng.integers(0, height, n), rng.integers(0, width, n)
    pick = rng.random(n) < 0.5
    img[ys[pick], xs[pick]] = text_colour
    img[ys[~pick], xs[~pick]] = image_colour
    return img
