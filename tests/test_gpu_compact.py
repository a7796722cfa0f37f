"""Compact transport formats of the batch pipeline: bit-packed binarised pages in
(pcs_preprocess_bits / pcs_predict_pages_packed), class map + bit-packed `data.binary` out (pcs_predict_pages_compact),
colour masks materialised on request on the device.  Everything is held to the oracle's prepare_images
(dataset.py:131-150), network (network.py:248-260) and generate_output_masks (output.py:44-60)."""
import numpy as np
import pytest
import torch

from oracle import network as onet
from oracle import pipeline as opipe
from page_segmentation_b200 import synth
from page_segmentation_b200.runtime import pack_pages, unpack_bits_host

pytestmark = pytest.mark.gpu
LUT = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}


@pytest.mark.parametrize("npix", [1, 31, 32, 33, 1000, 966763, 1024 * 37])
def test_pack_unpack_bits_roundtrip(ctx, npix):
    rng = np.random.default_rng(npix)
    n = 3
    src = (rng.random((n, npix)) < 0.3).astype(np.uint8) * rng.integers(1, 255, size=(n, npix)).astype(np.uint8)
    words = (npix + 31) // 32 + 2                                            # padded pitch: the padding must come back zero
    d_src = torch.from_numpy(src).cuda()
    d_bits = torch.full((n, words), -1, dtype=torch.int32, device="cuda")
    ctx.pack_bits(d_src, n, npix, d_bits, words)
    bits = d_bits.cpu().numpy().view(np.uint32)
    exp = np.packbits(np.pad(src != 0, ((0, 0), (0, words * 32 - npix))), axis=1, bitorder="little").view("<u4")
    np.testing.assert_array_equal(bits, exp)
    d_back = torch.empty((n, npix), dtype=torch.uint8, device="cuda")
    ctx.unpack_bits(d_bits, n, words, npix, d_back)
    np.testing.assert_array_equal(d_back.cpu().numpy(), (src != 0).astype(np.uint8))


@pytest.mark.parametrize("shape,lh,first_is_ink", [((3508, 2480), 18, False), ((700, 500), 18, False), ((333, 517), 11, True),
                                                   ((1200, 900), 24, True), ((97, 131), 5, False)])
def test_preprocess_from_packed_pages_is_bit_exact(ctx, shape, lh, first_is_ink):
    """pcs_preprocess_bits == the oracle's prepare_images of the uint8 page; also for pages whose first pixel is ink (the
    uint8 fast path keys its bit plane on page[0]; the packed entry point has fixed level semantics) and whose size is not
    a multiple of 32 pixels."""
    page = synth.make_page(3, shape[0], shape[1], lh)
    if first_is_ink:
        page = page.copy()
        page[0, :7] = 0
    bits, l0, l1 = pack_pages(page[None])
    assert (l0, l1) == (0, 255)
    H, W = shape
    Hs, Ws = synth.scaled_shape(H, W, 6 / lh)
    words = H * W // 32 + 1
    padded = np.zeros((1, words), np.uint32)
    padded[:, :bits.shape[1]] = bits
    d_bits = torch.from_numpy(padded.view(np.int32)).cuda()
    d_img = torch.empty((Hs, Ws), dtype=torch.uint8, device="cuda")
    d_bin = torch.empty((Hs, Ws), dtype=torch.uint8, device="cuda")
    ctx.preprocess_bits(d_bits, words, 1, H, W, l0, l1, Hs, Ws, d_img, d_bin)
    eimg, eb = opipe.prepare_images(page, page, 6, lh)
    np.testing.assert_array_equal(d_bin.cpu().numpy(), eb)
    np.testing.assert_array_equal(d_img.cpu().numpy(), eimg)


def test_pack_pages_layout():
    page = np.full((2, 3, 40), 255, np.uint8)
    page[0, 0, 1] = 0
    page[1, 2, 39] = 0
    bits, l0, l1 = pack_pages(page)
    assert bits.shape == (2, 4) and bits.dtype == np.uint32 and (l0, l1) == (0, 255)
    assert bits[0, 0] == 0xffffffff & ~2 and bits[1, 3] == 0x00ffffff & ~(1 << 23)
    np.testing.assert_array_equal(unpack_bits_host(bits, (3, 40)), (page == 255).astype(np.uint8))


@pytest.mark.parametrize("cc", [False, True])
def test_compact_and_packed_calls_vs_oracle(ctx, cc, monkeypatch):
    """uint8 pages in / packed pages in -> class map + packed binary out, over a chunked call: identical class maps from
    both input forms, `binary` bits == the oracle's binary, class maps against the fp64 oracle (the vote against the
    oracle's vote over the raw device map), and the masks materialised from the compact results == the oracle's."""
    from page_segmentation_b200.runtime import PageBatchEngine
    monkeypatch.setenv("PCSEG_HOST_CHUNK", "2")
    n, H, W_ = 5, 393, 300                       # 117 900 pixels per page: not a multiple of 32
    pages = np.stack([synth.make_page(40 + s, H, W_, 18) for s in range(n)])
    weights = synth.make_weights("fcn_skip", 3, seed=9)
    lut = np.array([LUT[i] for i in range(3)], dtype=np.uint8)
    eng = PageBatchEngine("fcn_skip", weights, 3, precision="fp16", lut=lut)
    Hs, Ws = synth.scaled_shape(H, W_, 6 / 18)
    bw = (Hs * Ws + 31) // 32
    a = {"labels": np.zeros((n, Hs, Ws), np.uint8), "binary_bits": np.zeros((n, bw), np.uint32)}
    b = {"labels": np.zeros((n, Hs, Ws), np.uint8), "binary_bits": np.zeros((n, bw), np.uint32)}
    eng.run_host_compact(pages, 6 / 18, a, cc_majority=cc)
    bits, l0, l1 = pack_pages(pages)
    eng.run_host_packed(bits, l0, l1, H, W_, 6 / 18, b, cc_majority=cc)
    np.testing.assert_array_equal(a["labels"], b["labels"])
    np.testing.assert_array_equal(a["binary_bits"], b["binary_bits"])
    raw = {"labels": np.zeros((n, Hs, Ws), np.uint8)}
    eng.run_host_compact(pages, 6 / 18, raw, cc_majority=False)
    binary = unpack_bits_host(a["binary_bits"], (Hs, Ws))
    masks = eng.masks_from_compact(a["labels"], a["binary_bits"])
    for i in range(n):
        eimg, eb = opipe.prepare_images(pages[i], pages[i], 6, 18)
        np.testing.assert_array_equal(binary[i], eb)
        l64 = onet.Forward("fcn_skip", weights, 3, dtype=torch.float64).logits(eimg)[0]
        assert (raw["labels"][i] == l64.argmax(-1)).mean() >= 0.999
        exp = opipe.vote_connected_component_class(raw["labels"][i].astype(np.int64), eb) if cc else raw["labels"][i]
        np.testing.assert_array_equal(a["labels"][i], exp)
        c, o, inv, _ = opipe.generate_output_masks(eb, a["labels"][i].astype(np.int64), LUT)
        np.testing.assert_array_equal(masks["color"][i], c)
        np.testing.assert_array_equal(masks["overlay"][i], o)
        np.testing.assert_array_equal(masks["inverted"][i], inv)


def _pinned(shape, dtype):
    t = torch.empty(shape, dtype={np.uint8: torch.uint8, np.uint32: torch.int32, np.int32: torch.int32}[dtype]).pin_memory().numpy()
    return t.view(dtype)


@pytest.mark.parametrize("cc", [False, True])
def test_streaming_submits_match_blocking_calls_and_oracle(ctx, cc, monkeypatch):
    """pcs_predict_pages_compact_submit / pcs_wait_pages: five batches of different pages submitted back to back over two
    rotating sets of host buffers (chunks of 2 pages: the staging buffers keep rotating across the call boundaries, each
    chained call's upload runs under the kernels of the call before).  Every batch: bit-identical to the blocking call,
    `binary` == the oracle's prepare_images, class maps against the fp64 oracle.  Then a blocking call and a submit of
    another shape (neither chains) in between, and the ticket errors."""
    from page_segmentation_b200._native import PcsError
    from page_segmentation_b200.runtime import PageBatchEngine
    monkeypatch.setenv("PCSEG_HOST_CHUNK", "2")
    n, H, W_ = 5, 393, 300
    nb = 5
    weights = synth.make_weights("fcn_skip", 3, seed=9)
    eng = PageBatchEngine("fcn_skip", weights, 3, precision="fp16")
    Hs, Ws = synth.scaled_shape(H, W_, 6 / 18)
    bw = (Hs * Ws + 31) // 32
    batches = [np.stack([synth.make_page(500 + 10 * k + s, H, W_, 18) for s in range(n)]) for k in range(nb)]
    blocking = []
    for k in range(nb):
        o = {"labels": np.zeros((n, Hs, Ws), np.uint8), "binary_bits": np.zeros((n, bw), np.uint32)}
        eng.run_host_compact(batches[k], 6 / 18, o, cc_majority=cc)
        blocking.append(o)
    h_in = [_pinned((n, H, W_), np.uint8) for _ in range(2)]
    h_out = [{"labels": _pinned((n, Hs, Ws), np.uint8), "binary_bits": _pinned((n, bw), np.uint32)} for _ in range(2)]
    got = [None] * nb
    tickets = [None] * nb

    def collect(k):
        eng.wait(tickets[k])
        got[k] = {key: v.copy() for key, v in h_out[k % 2].items()}

    for k in range(nb):
        if k >= 2:
            collect(k - 2)                              # the buffers of batch k - 2 are free again
        h_in[k % 2][...] = batches[k]
        for v in h_out[k % 2].values():
            v[...] = 0xA5 if v.dtype == np.uint8 else 0xA5A5A5A5
        tickets[k] = eng.submit_host_compact(h_in[k % 2], 6 / 18, h_out[k % 2], cc_majority=cc)
    collect(nb - 2)
    collect(nb - 1)
    assert tickets == sorted(tickets) and len(set(tickets)) == nb
    for k in range(nb):
        np.testing.assert_array_equal(got[k]["labels"], blocking[k]["labels"], err_msg=f"batch {k}")
        np.testing.assert_array_equal(got[k]["binary_bits"], blocking[k]["binary_bits"], err_msg=f"batch {k}")
    # the oracle on the first and the last batch (the blocking results equal the streamed ones)
    for k in (0, nb - 1):
        binary = unpack_bits_host(got[k]["binary_bits"], (Hs, Ws))
        raw = {"labels": np.zeros((n, Hs, Ws), np.uint8)}
        eng.run_host_compact(batches[k], 6 / 18, raw, cc_majority=False)
        for i in (0, n - 1):
            eimg, eb = opipe.prepare_images(batches[k][i], batches[k][i], 6, 18)
            np.testing.assert_array_equal(binary[i], eb)
            l64 = onet.Forward("fcn_skip", weights, 3, dtype=torch.float64).logits(eimg)[0]
            assert (raw["labels"][i] == l64.argmax(-1)).mean() >= 0.999
            exp = opipe.vote_connected_component_class(raw["labels"][i].astype(np.int64), eb) if cc else raw["labels"][i]
            np.testing.assert_array_equal(got[k]["labels"][i], exp)
    # submit, blocking call, submit of another shape, submit of the first shape again: no chain anywhere, same results
    t0 = eng.submit_host_compact(h_in[0], 6 / 18, h_out[0], cc_majority=cc)           # h_in[0] holds batch nb - 1 (nb odd)
    o = {"labels": np.zeros((n, Hs, Ws), np.uint8), "binary_bits": np.zeros((n, bw), np.uint32)}
    eng.run_host_compact(batches[1], 6 / 18, o, cc_majority=cc)
    np.testing.assert_array_equal(o["labels"], blocking[1]["labels"])
    small = batches[2][:3, :200, :160].copy()
    hs2, ws2 = synth.scaled_shape(200, 160, 6 / 18)
    so = {"labels": _pinned((3, hs2, ws2), np.uint8)}
    t1 = eng.submit_host_compact(small, 6 / 18, so, cc_majority=cc)
    t2 = eng.submit_host_compact(h_in[1], 6 / 18, h_out[1], cc_majority=cc)           # h_in[1] holds batch nb - 2
    eng.wait(t2)
    eng.wait(t0)                                        # an older ticket after a newer one
    eng.wait(t1)
    np.testing.assert_array_equal(h_out[0]["labels"], blocking[nb - 1]["labels"])
    np.testing.assert_array_equal(h_out[1]["labels"], blocking[nb - 2]["labels"])
    sb = {"labels": np.zeros((3, hs2, ws2), np.uint8)}
    eng.run_host_compact(small, 6 / 18, sb, cc_majority=cc)
    np.testing.assert_array_equal(so["labels"], sb["labels"])
    # the packed form, two chained submits
    pb = [pack_pages(batches[k]) for k in (0, 1)]
    po = [{"labels": _pinned((n, Hs, Ws), np.uint8), "binary_bits": _pinned((n, bw), np.uint32)} for _ in range(2)]
    tp = [eng.submit_host_packed(pb[k][0], pb[k][1], pb[k][2], H, W_, 6 / 18, po[k], cc_majority=cc) for k in (0, 1)]
    for k in (1, 0):
        eng.wait(tp[k])
        np.testing.assert_array_equal(po[k]["labels"], blocking[k]["labels"])
        np.testing.assert_array_equal(po[k]["binary_bits"], blocking[k]["binary_bits"])
    with pytest.raises(PcsError):
        eng.wait(tp[1] + 1)                             # never issued
    eng.ctx.synchronize()


def test_streaming_segment_call_matches_blocking(ctx, monkeypatch):
    """pcs_predict_pages_segments_compact_submit: three chained submits == the blocking segment call (class map, bit-packed
    binary, stats tables, component counts)."""
    from page_segmentation_b200.runtime import PageBatchEngine
    monkeypatch.setenv("PCSEG_HOST_CHUNK", "3")
    n, H, W_, maxc = 4, 393, 300, 512
    weights = synth.make_weights("fcn_skip", 3, seed=9)
    eng = PageBatchEngine("fcn_skip", weights, 3, precision="fp16")
    Hs, Ws = synth.scaled_shape(H, W_, 6 / 18)
    bw = (Hs * Ws + 31) // 32
    batches = [np.stack([synth.make_page(700 + 10 * k + s, H, W_, 18) for s in range(n)]) for k in range(3)]

    def outs(alloc):
        return {"labels": alloc((n, Hs, Ws), np.uint8), "binary_bits": alloc((n, bw), np.uint32),
                "stats": alloc((n, 3, maxc, 5), np.int32), "ncomp": alloc((n, 3), np.int32)}

    ref = []
    for k in range(3):
        o = outs(lambda s, d: np.zeros(s, d))
        eng.run_host_segments_compact(batches[k], 6 / 18, o, max_components=maxc, cc_majority=True)
        ref.append(o)
    h_in = [_pinned((n, H, W_), np.uint8) for _ in range(3)]
    h_out = [outs(_pinned) for _ in range(3)]
    tickets = []
    for k in range(3):
        h_in[k][...] = batches[k]
        tickets.append(eng.submit_host_compact(h_in[k], 6 / 18, h_out[k], cc_majority=True, max_components=maxc))
    for k in (2, 0, 1):
        eng.wait(tickets[k])
        for key in ("labels", "binary_bits", "ncomp"):
            np.testing.assert_array_equal(h_out[k][key], ref[k][key], err_msg=f"batch {k} {key}")
        for i in range(n):
            for c in range(3):
                m = min(int(ref[k]["ncomp"][i, c]), maxc)
                np.testing.assert_array_equal(h_out[k]["stats"][i, c, :m], ref[k]["stats"][i, c, :m], err_msg=f"batch {k} page {i} class {c}")
