"""The device PNG encoder behind output_data (lib/output.py:38-41): files must be valid PNGs that decode to exactly the
input bytes (the only parity an image file has), and byte-identical to the same container assembled on the host with
zlib's crc32 / adler32."""
import struct
import zlib

import cv2
import numpy as np
import pytest

from page_segmentation_b200 import synth

pytestmark = pytest.mark.gpu


def host_png(img: np.ndarray) -> bytes:
    """The same container (stored deflate blocks of whole scanlines) assembled with Python's zlib checksums."""
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(h))
    line = 1 + w * c
    per = (65535 // line) * line
    z = b"\x78\x01"
    for off in range(0, len(raw), per):
        part = raw[off:off + per]
        z += struct.pack("<BHH", 1 if off + per >= len(raw) else 0, len(part), len(part) ^ 0xFFFF) + part
    z += struct.pack(">I", zlib.adler32(raw))

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data))
    ihdr = struct.pack(">IIBBBBB", w, h, 8, {1: 0, 3: 2, 4: 6}[c], 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", z) + chunk(b"IEND", b"")


@pytest.mark.parametrize("shape", [(1, 1, 3), (7, 5, 3), (61, 83, 3), (389, 275, 3), (1169, 827, 3), (40, 21844, 3), (33, 65)])
def test_png_files_are_valid_and_lossless(ctx, shape):
    from page_segmentation_b200.lib.output import encode_png
    rng = np.random.default_rng(shape[0])
    imgs = rng.integers(0, 256, (2,) + shape, dtype=np.uint8)
    imgs[1][: shape[0] // 2] = 255                                   # long runs of 0xff stress the Adler sums
    files = encode_png(imgs)
    assert len(files) == 2
    for img, blob in zip(imgs, files):
        assert blob == host_png(img)                                 # every byte: headers, block framing, Adler-32, CRC-32
        assert zlib.decompress(blob[41:-16]) == b"".join(b"\x00" + img[r].tobytes() for r in range(img.shape[0]))
        dec = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)
        assert dec is not None
        np.testing.assert_array_equal(dec if img.ndim == 2 else dec[..., ::-1], img)


def test_png_full_a4_masks(ctx):
    """Three full-resolution colour masks (3508 x 2480 x 3 = 26 MB each, 400 stored blocks per file)."""
    from page_segmentation_b200.lib.output import encode_png
    inv = synth.make_inverted_image(5, synth.A4_H, synth.A4_W, 40)
    imgs = np.stack([inv, 255 - inv, inv[::-1].copy()])
    for img, blob in zip(imgs, encode_png(imgs)):
        assert len(blob) == len(host_png(img)) and blob == host_png(img)


def test_png_rejects_unsupported_shapes(ctx):
    from page_segmentation_b200.lib.output import encode_png
    with pytest.raises(ValueError):
        encode_png(np.zeros((1, 4, 30000, 3), np.uint8))             # a scanline longer than one stored block
    with pytest.raises(ValueError):
        encode_png(np.zeros((1, 4, 4, 2), np.uint8))                 # grey + alpha is not offered


def test_output_data_writes_device_encoded_pngs(ctx, tmp_path):
    from oracle import pipeline as opipe
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import SingleData
    from page_segmentation_b200.lib.output import output_data
    rng = np.random.default_rng(3)
    page = synth.make_page(2, 300, 240, 18)
    binary = (page == 0).astype(np.uint8)
    pred = rng.integers(0, 3, binary.shape).astype(np.int64)
    for sub in ("color", "overlay", "inverted"):
        (tmp_path / sub).mkdir()
    data = SingleData(binary=binary, image_path="/somewhere/page_0001.png")
    output_data(str(tmp_path), pred[None], data, DEFAULT_COLOR_MAP)
    exp = opipe.generate_output_masks(binary, pred, {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)})
    for sub, e in zip(("color", "overlay", "inverted"), exp[:3]):
        got = cv2.imread(str(tmp_path / sub / "page_0001.png"), cv2.IMREAD_COLOR)[..., ::-1]
        np.testing.assert_array_equal(got, e)
