"""The device PNG encoder behind output_data (lib/output.py:38-41): files must be valid PNGs that decode to exactly the
input bytes (the only parity an image file has), and byte-identical to the same container assembled on the host with
zlib's crc32 / adler32."""
import struct
import zlib

import cv2
import os

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


def check_container(blob: bytes, img: np.ndarray, sub_filter: bool):
    """Chunk structure, both checksums, the zlib stream and the decoded pixels of one PNG file."""
    assert blob[:8] == b"\x89PNG\r\n\x1a\n"
    pos, kinds, idat = 8, [], b""
    while pos < len(blob):
        (length,), kind = struct.unpack(">I", blob[pos:pos + 4]), blob[pos + 4:pos + 8]
        data = blob[pos + 8:pos + 8 + length]
        assert struct.unpack(">I", blob[pos + 8 + length:pos + 12 + length])[0] == zlib.crc32(kind + data), kind
        kinds.append(kind)
        if kind == b"IDAT":
            idat += data
        pos += 12 + length
    assert kinds == [b"IHDR", b"IDAT", b"IEND"] and pos == len(blob)
    h = img.shape[0]
    c = 1 if img.ndim == 2 else img.shape[2]
    stream = zlib.decompress(idat)                                  # inflate also verifies the Adler-32
    line = 1 + img.shape[1] * c
    assert len(stream) == h * line
    lines = np.frombuffer(stream, np.uint8).reshape(h, line)
    if sub_filter:                                                  # level 1: Sub or Up per scanline, undone here by hand
        assert set(np.unique(lines[:, 0])) <= {1, 2}
        prev = np.zeros(line - 1, np.uint8)
        for r in range(h):
            f = lines[r, 1:]
            if lines[r, 0] == 1:
                cur = np.cumsum(f.reshape(-1, c).astype(np.uint64), axis=0).astype(np.uint8).reshape(-1)
            else:
                cur = (f.astype(np.uint16) + prev).astype(np.uint8)
            assert np.array_equal(cur, img[r].reshape(-1)), r
            prev = cur
    else:
        assert stream == b"".join(b"\x00" + img[r].tobytes() for r in range(h))
    dec = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)
    assert dec is not None
    np.testing.assert_array_equal(dec if img.ndim == 2 else (dec[..., ::-1] if c == 3 else dec[..., [2, 1, 0, 3]]), img)


def mask_like(rng, shape):
    """Blocky class-colour image with some speckle: what the encoder is for."""
    h, w = shape[:2]
    lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 0]], np.uint8)
    coarse = rng.integers(0, 4, (h // 24 + 1, w // 24 + 1))
    lab = np.kron(coarse, np.ones((24, 24), np.int64))[:h, :w]
    flip = rng.random((h, w)) < 0.01
    lab[flip] = rng.integers(0, 4, int(flip.sum()))
    img = lut[lab]
    return img if len(shape) == 3 else img[..., 0]


@pytest.mark.parametrize("shape", [(1, 1, 3), (7, 5, 3), (61, 83, 3), (389, 275, 3), (1169, 827, 3), (5, 5461, 3), (33, 65), (300, 260, 4)])
def test_png_level1_is_valid_lossless_and_small(ctx, shape):
    from page_segmentation_b200.lib.output import encode_png
    rng = np.random.default_rng(shape[0] + 1)
    noise = rng.integers(0, 256, shape, dtype=np.uint8)             # worst case: no runs at all, 9-bit literals
    runs = np.zeros(shape, np.uint8)
    runs[shape[0] // 3:] = 200                                       # runs far longer than 258 bytes, literals >= 144
    if len(shape) == 3 and shape[2] == 3:
        masks = mask_like(rng, shape)
    else:
        masks = (rng.random(shape) < 0.5).astype(np.uint8) * 255
    imgs = np.stack([noise, runs, masks])
    files = encode_png(imgs, level=1)
    for img, blob in zip(imgs, files):
        check_container(blob, img, sub_filter=True)
    assert len(files[0]) <= imgs[0].size * 9 // 8 + shape[0] * 2 + 80
    if shape[0] * shape[1] > 10000:
        assert len(files[1]) < imgs[1].size // 50
        if len(shape) == 3 and shape[2] == 3:
            assert len(files[2]) < imgs[2].size // 15, (len(files[2]), imgs[2].size)


@pytest.mark.parametrize("shape", [(1, 1, 3), (7, 5, 3), (61, 83, 3), (389, 275, 3), (1169, 827, 3), (40, 21844, 3), (33, 65)])
def test_png_files_are_valid_and_lossless(ctx, shape):
    from page_segmentation_b200.lib.output import encode_png
    rng = np.random.default_rng(shape[0])
    imgs = rng.integers(0, 256, (2,) + shape, dtype=np.uint8)
    imgs[1][: shape[0] // 2] = 255                                   # long runs of 0xff stress the Adler sums
    files = encode_png(imgs, level=0)
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
    for img, blob in zip(imgs, encode_png(imgs, level=0)):
        assert len(blob) == len(host_png(img)) and blob == host_png(img)
    for img, blob in zip(imgs, encode_png(imgs, level=1)):
        check_container(blob, img, sub_filter=True)
        assert len(blob) < img.size // 4


def test_png_rejects_unsupported_shapes(ctx):
    from page_segmentation_b200.lib.output import encode_png
    with pytest.raises(ValueError):
        encode_png(np.zeros((1, 4, 30000, 3), np.uint8))             # a scanline longer than one stored block
    long_lines = np.zeros((1, 3, 6000, 3), np.uint8)                 # too long for the run tables: falls back to stored blocks
    assert encode_png(long_lines, level=1) == encode_png(long_lines, level=0)
    with pytest.raises(ValueError):
        encode_png(np.zeros((1, 4, 4, 2), np.uint8))                 # grey + alpha is not offered


def test_output_data_writes_device_encoded_pngs(ctx, tmp_path):
    from oracle import pipeline as opipe
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import SingleData
    from page_segmentation_b200.lib.output import flush_outputs, output_data
    rng = np.random.default_rng(3)
    page = synth.make_page(2, 300, 240, 18)
    binary = (page == 0).astype(np.uint8)
    pred = rng.integers(0, 3, binary.shape).astype(np.int64)
    for sub in ("color", "overlay", "inverted"):
        (tmp_path / sub).mkdir()
    data = SingleData(binary=binary, image_path="/somewhere/page_0001.png")
    output_data(str(tmp_path), pred[None], data, DEFAULT_COLOR_MAP)
    flush_outputs()                                                  # the library's worker threads write the files
    exp = opipe.generate_output_masks(binary, pred, {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)})
    for sub, e in zip(("color", "overlay", "inverted"), exp[:3]):
        got = cv2.imread(str(tmp_path / sub / "page_0001.png"), cv2.IMREAD_COLOR)[..., ::-1]
        np.testing.assert_array_equal(got, e)


def test_output_pages_batch_and_write_errors(ctx, tmp_path):
    """pcs_output_pages for several pages in one call (paths page-major), more calls in flight than the library has
    slots, and a path that cannot be written: the error surfaces at pcs_output_flush and the writer keeps working."""
    import torch
    from oracle import pipeline as opipe
    from page_segmentation_b200._native import PcsError
    rng = np.random.default_rng(5)
    n, H, W = 5, 140, 200
    lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], dtype=np.uint8)
    kinds = ("color", "overlay", "inverted")
    cases = []
    for call in range(9):                                            # 9 calls > 6 slots
        labels = rng.integers(0, 3, (n, H, W)).astype(np.uint8)
        binary = (rng.random((n, H, W)) < 0.3).astype(np.uint8)
        paths = [str(tmp_path / f"c{call}_p{p}_{k}.png") for p in range(n) for k in kinds]
        ctx.output_pages(torch.from_numpy(labels).cuda(), torch.from_numpy(binary).cuda(), n, H, W, lut, paths)
        cases.append((labels, binary, paths))
    ctx.output_flush()
    for labels, binary, paths in cases:
        for p in range(n):
            exp = opipe.generate_output_masks(binary[p], labels[p].astype(np.int64), {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)})
            for k in range(3):
                got = cv2.imread(paths[3 * p + k], cv2.IMREAD_COLOR)[..., ::-1]
                np.testing.assert_array_equal(got, exp[k])
    labels, binary, _ = cases[0]
    bad = [str(tmp_path / "no_such_dir" / f"x{k}.png") for k in range(3)]
    ctx.output_pages(torch.from_numpy(labels[:1]).cuda(), torch.from_numpy(binary[:1]).cuda(), 1, H, W, lut, bad)
    with pytest.raises(PcsError, match="cannot write"):
        ctx.output_flush()
    good = [str(tmp_path / f"again_{k}.png") for k in range(3)]
    ctx.output_pages(torch.from_numpy(labels[:1]).cuda(), torch.from_numpy(binary[:1]).cuda(), 1, H, W, lut, good)
    ctx.output_flush()
    assert all(os.path.getsize(g) > 0 for g in good)
    with pytest.raises(PcsError):
        ctx.output_pages(torch.from_numpy(labels[:1]).cuda(), torch.from_numpy(binary[:1]).cuda(), 1, H, W, lut, good[:2])


@pytest.mark.parametrize("cc", [False, True])
def test_batch_pipeline_with_png_output_matches_raw_masks(ctx, cc):
    """pcs_predict_pages_files: the PNG files of a page batch decode to exactly the raw masks of pcs_predict_pages_host."""
    import torch
    from page_segmentation_b200.runtime import PageBatchEngine
    lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], dtype=np.uint8)
    eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, seed=4), 3, lut=lut)
    n, H, W = 11, 420, 333                                          # 11 pages: ragged chunk schedule 2, 4, 1, 2, 2
    pages = np.stack([synth.make_page(s, H, W, 18) for s in range(n)])
    Hs, Ws = synth.scaled_shape(H, W, 1 / 3)
    raw = {k: np.zeros((n, Hs, Ws) + ((3,) if k != "labels" else ()), np.uint8) for k in ("labels", "color", "overlay", "inverted")}
    eng.run_host(pages, 1 / 3, raw, cc_majority=cc)
    stride = (eng.ctx.png_bytes(Hs, Ws, 3, 1) + 255) // 256 * 256
    out = {"labels": np.zeros((n, Hs, Ws), np.uint8), "png": torch.zeros((n, 3, stride), dtype=torch.uint8).pin_memory().numpy(),
           "png_sizes": np.zeros((n, 3), np.uint64)}
    eng.run_host_files(pages, 1 / 3, out, cc_majority=cc)
    np.testing.assert_array_equal(out["labels"], raw["labels"])
    for p in range(n):
        for k, kind in enumerate(("color", "overlay", "inverted")):
            blob = out["png"][p, k, :int(out["png_sizes"][p, k])].tobytes()
            check_container(blob, raw[kind][p], sub_filter=True)
    assert int(out["png_sizes"].max()) < raw["color"][0].size // 2
