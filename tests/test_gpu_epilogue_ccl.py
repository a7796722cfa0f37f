"""Bit-exact parity of the colour epilogue (output.py:44-60) and of the
connected-component post-processors (postprocess.py:9-42, cv2 4-connectivity)."""
import cv2
import numpy as np
import pytest

from oracle import pipeline as opipe
from page_segmentation_b200 import synth
from page_segmentation_b200.lib.colors import ColorMap, DEFAULT_COLOR_MAP
from page_segmentation_b200.lib.dataset import SingleData

pytestmark = pytest.mark.gpu


def _pred_and_binary(seed, h, w, n_classes=3):
    rng = np.random.default_rng(seed)
    page = synth.make_page(seed, h * 3, w * 3, 18)
    _, b = opipe.prepare_images(page, page, 6, 18)
    # blocky class map so that components see mixed classes
    coarse = rng.integers(0, n_classes, size=(h // 7 + 1, w // 9 + 1))
    pred = np.kron(coarse, np.ones((7, 9), dtype=np.int64))[:b.shape[0], :b.shape[1]]
    noise = rng.random(b.shape) < 0.15
    pred = np.where(noise, rng.integers(0, n_classes, size=b.shape), pred).astype(np.int64)
    return pred, b


@pytest.mark.parametrize("hw", [(2, 2), (61, 83), (389, 275), (1169, 827)])
def test_generate_output_masks_bit_exact(ctx, hw):
    from page_segmentation_b200.lib.output import generate_output_masks
    pred, b = _pred_and_binary(1, *hw)
    m = generate_output_masks(SingleData(binary=b), pred, DEFAULT_COLOR_MAP)
    lut = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}
    c, o, i, f = opipe.generate_output_masks(b, pred, lut)
    np.testing.assert_array_equal(m.color, c)
    np.testing.assert_array_equal(m.overlay, o)
    np.testing.assert_array_equal(m.inverted_overlay, i)
    np.testing.assert_array_equal(m.fg_color_mask, f)


def test_masks_unknown_label_is_black(ctx):
    from page_segmentation_b200.lib.output import generate_output_masks
    pred = np.array([[0, 1], [2, 3]], dtype=np.int64)
    b = np.array([[1, 0], [1, 1]], dtype=np.uint8)
    cm = ColorMap({(255, 255, 255): (0, "background"), (255, 0, 0): (1, "text"), (0, 0, 255): (3, "x")})
    m = generate_output_masks(SingleData(binary=b), pred, cm)
    assert m.color[1, 0].tolist() == [0, 0, 0]
    assert m.color[1, 1].tolist() == [0, 0, 255]
    assert m.overlay[0, 0].tolist() == [0, 0, 0] and m.overlay[0, 1].tolist() == [255, 0, 0]
    assert m.inverted_overlay[0, 1].tolist() == [0, 0, 0] and m.inverted_overlay[0, 0].tolist() == [255, 255, 255]


@pytest.mark.parametrize("hw", [(6, 8), (61, 83), (389, 275), (1169, 827)])
def test_ccl_matches_cv2(ctx, hw):
    from page_segmentation_b200.runtime import connected_components_with_stats
    _, b = _pred_and_binary(2, *hw)
    n, labels, stats = connected_components_with_stats(b)
    en, elabels, estats, _ = cv2.connectedComponentsWithStats(b, connectivity=4)
    assert n == en
    np.testing.assert_array_equal(labels, elabels)
    np.testing.assert_array_equal(stats, estats)


@pytest.mark.parametrize("pattern", ["empty", "full", "checker", "hstripes", "vstripes", "spiral", "wide"])
def test_ccl_patterns(ctx, pattern):
    from page_segmentation_b200.runtime import connected_components_with_stats
    h, w = 70, 131
    a = np.zeros((h, w), np.uint8)
    if pattern == "full":
        a[:] = 1
    elif pattern == "checker":
        a[::2, ::2] = 1
        a[1::2, 1::2] = 1
    elif pattern == "hstripes":
        a[::2, :] = 1
    elif pattern == "vstripes":
        a[:, ::2] = 1
    elif pattern == "spiral":
        for k in range(0, 30, 2):
            a[k, k:w - k] = 1
            a[k:h - k, w - 1 - k] = 1
            a[h - 1 - k, k:w - k] = 1
            a[k + 2:h - k, k] = 1
    elif pattern == "wide":
        a[10, :] = 1
        a[5:40, 64] = 1
        a[30, 3:100] = 1
    n, labels, stats = connected_components_with_stats(a)
    en, elabels, estats, _ = cv2.connectedComponentsWithStats(a, connectivity=4)
    assert n == en
    np.testing.assert_array_equal(labels, elabels)
    if pattern not in ("full",):          # cv2's background row is undefined when there is no background
        np.testing.assert_array_equal(stats, estats)
    else:
        np.testing.assert_array_equal(stats[1:], estats[1:])


@pytest.mark.parametrize("hw", [(61, 83), (389, 275), (1169, 827)])
def test_cc_majority_bit_exact(ctx, hw):
    from page_segmentation_b200.lib.postprocess import vote_connected_component_class
    pred, b = _pred_and_binary(3, *hw)
    exp = opipe.vote_connected_component_class(pred.copy(), b)
    got_in = pred.copy()
    got = vote_connected_component_class(got_in, SingleData(binary=b))
    assert got is got_in                        # the reference mutates in place
    np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("hw", [(61, 83), (200, 160), (97, 530)])         # the last one spans three labelling tiles across
def test_bounding_boxes(ctx, hw):
    from page_segmentation_b200.lib.postprocess import add_bounding_boxes
    pred, b = _pred_and_binary(4, *hw)
    pred = np.where(b > 0, pred, 0)
    exp = opipe.add_bounding_boxes(pred.copy())
    got = add_bounding_boxes(pred.copy(), SingleData(binary=b))
    np.testing.assert_array_equal(got, exp)


def test_find_postprocessor(ctx):
    from page_segmentation_b200.lib import postprocess as pp
    assert pp.find_postprocessor("cc_majority") is pp.vote_connected_component_class
    assert pp.find_postprocessor("CC-Vote") is pp.vote_connected_component_class
    assert pp.find_postprocessor("bounding_boxes") is pp.add_bounding_boxes
    with pytest.raises(KeyError):
        pp.find_postprocessor("nope")


@pytest.mark.parametrize("hw", [(61, 83), (200, 160), (97, 530), (389, 275), (1169, 827)])
def test_class_components_match_cv2_stats(ctx, hw):
    """pcs_class_components: per class c the label count and stats table of
    cv2.connectedComponentsWithStats(pred == c, connectivity=4) (postprocess.py:31-33), bit-exact incl. row 0."""
    from page_segmentation_b200.lib.postprocess import class_components
    pred, b = _pred_and_binary(6, *hw)
    pred = np.where(b > 0, pred, 0)
    got = class_components(pred, 3, max_components=64)          # small table first: exercises the regrow path
    for c in range(3):
        n, _, stats, _ = cv2.connectedComponentsWithStats((pred == c).astype(np.uint8), connectivity=4)
        assert got[c][0] == n, (c, got[c][0], n)
        np.testing.assert_array_equal(got[c][1], stats)


def test_class_components_batched_and_truncated(ctx):
    """n pages in one call; a class that is absent yields one label (the background) covering the page; components beyond
    max_components are dropped while the count still reports them."""
    import torch
    preds = [np.where(_pred_and_binary(20 + s, 150, 203)[1] > 0, _pred_and_binary(20 + s, 150, 203)[0], 0) for s in range(3)]
    preds[1][preds[1] == 2] = 1                                    # page 1 has no class 2
    d = torch.from_numpy(np.stack(preds).astype(np.uint8)).cuda()
    maxc = 16
    d_stats = torch.full((3, 3, maxc, 5), -7, dtype=torch.int32, device="cuda")
    d_ncomp = torch.zeros((3, 3), dtype=torch.int32, device="cuda")
    ctx.class_components(d, 3, 150, 203, 3, d_stats, maxc, d_ncomp)
    stats, ncomp = d_stats.cpu().numpy(), d_ncomp.cpu().numpy()
    for p in range(3):
        for c in range(3):
            n, _, exp, _ = cv2.connectedComponentsWithStats((preds[p] == c).astype(np.uint8), connectivity=4)
            assert ncomp[p, c] == n
            k = min(n, maxc)
            np.testing.assert_array_equal(stats[p, c, :k], exp[:k])
            assert not stats[p, c, k:].any()
    assert ncomp[1, 2] == 1 and stats[1, 2, 0].tolist() == [0, 0, 203, 150, 150 * 203]


def _cv2_class_tables(pred, n_classes):
    return [cv2.connectedComponentsWithStats((pred == c).astype(np.uint8), connectivity=4)[::2] for c in range(n_classes)]


def _device_class_tables(ctx, preds, n_classes, maxc):
    import torch
    d = torch.from_numpy(np.stack(preds).astype(np.uint8)).cuda()
    n, h, w = d.shape
    d_stats = torch.full((n, n_classes, maxc, 5), -7, dtype=torch.int32, device="cuda")
    d_ncomp = torch.zeros((n, n_classes), dtype=torch.int32, device="cuda")
    ctx.class_components(d, n, h, w, n_classes, d_stats, maxc, d_ncomp)
    return d_stats.cpu().numpy(), d_ncomp.cpu().numpy()


def _check_class_tables(ctx, preds, n_classes, maxc=1 << 17):
    stats, ncomp = _device_class_tables(ctx, preds, n_classes, maxc)
    for p, pred in enumerate(preds):
        for c, (n, exp) in enumerate(_cv2_class_tables(pred, n_classes)):
            assert ncomp[p, c] == n, (p, c, ncomp[p, c], n)
            k = min(n, maxc)
            if exp[0, 4] == 0:
                # class c covers the page: cv2 4.13 reports [-1, INT_MAX, 0, 0, 0] for the empty label 0 (its untouched
                # min / max initialisers); the reference never reads row 0 (postprocess.py:35), the device writes zeros
                assert not stats[p, c, 0].any()
                exp = exp.copy()
                exp[0] = 0
            np.testing.assert_array_equal(stats[p, c, :k], exp[:k], err_msg=f"page {p} class {c}")
            assert not stats[p, c, k:].any()


@pytest.mark.parametrize("hw", [(33, 257), (64, 512), (97, 530), (300, 700)])
def test_class_components_pure_noise(ctx, hw):
    """Uniform three-class noise: ~3 600 root candidates per 256 x 32 labelling tile, far beyond the 1 024 records a
    tile accumulates in shared memory, so most components take the global-atomics path of the one-pass labelling."""
    rng = np.random.default_rng(5)
    preds = [rng.integers(0, 3, size=hw).astype(np.uint8) for _ in range(2)]
    _check_class_tables(ctx, preds, 3)


def test_class_components_shapes_across_tiles(ctx):
    """Components that leave a labelling tile and come back (a spiral, a comb, nested frames): tile-local roots that
    merge through neighbouring tiles, whose records have to be folded into one."""
    h, w = 200, 700
    spiral = np.zeros((h, w), np.uint8)
    y0, y1, x0, x1 = 2, h - 3, 2, w - 3
    while y1 - y0 > 8 and x1 - x0 > 8:
        spiral[y0, x0:x1 + 1] = 1
        spiral[y0:y1 + 1, x1] = 1
        spiral[y1, x0 + 4:x1 + 1] = 1
        spiral[y0 + 4:y1 + 1, x0 + 4] = 1
        spiral[y0 + 4, x0 + 4:x1 - 3] = 1
        y0 += 4; y1 -= 4; x0 += 4; x1 -= 4                                   # noqa: E702
    comb = np.zeros((h, w), np.uint8)
    comb[h - 2, :] = 2
    comb[3:h - 2, ::3] = 2
    comb[5:h - 40:2, 1::3] = 1
    frames = np.zeros((h, w), np.uint8)
    for k in range(0, 90, 2):
        frames[k:h - k, k:w - k] = (k // 2) % 3
    _check_class_tables(ctx, [spiral, comb, frames, np.full((h, w), 2, np.uint8)], 3)


def test_class_components_other_bytes_and_class_counts(ctx, monkeypatch):
    """Bytes outside 0 .. n_classes-1 belong to no class but count as 'not c' in every row 0; nine classes take the
    per-class path (the one-pass labelling handles up to eight); PCSEG_SEGMENTS_PER_CLASS is read once per process,
    so the A/B switch itself is exercised by tools/probe_segments.py."""
    rng = np.random.default_rng(11)
    coarse = rng.integers(0, 6, size=(40, 60))
    pred = np.kron(coarse, np.ones((5, 9), np.uint8))[:170, :530].astype(np.uint8)
    pred[rng.random(pred.shape) < 0.1] = 250
    _check_class_tables(ctx, [pred], 3)                                      # classes 3, 4, 5 and 250 are 'other'
    _check_class_tables(ctx, [pred], 6)
    _check_class_tables(ctx, [pred % 9, (pred + 3) % 9], 9)
    _check_class_tables(ctx, [pred % 8], 8)
    _check_class_tables(ctx, [np.where(pred == 250, 1, 0).astype(np.uint8)], 1)


def test_class_components_a4_noise_like_truncated(ctx):
    """Full scaled page, class map with letter-sized and page-sized components, table shorter than the label count."""
    preds = []
    for s in range(2):
        pred, b = _pred_and_binary(40 + s, 1169, 827)
        preds.append(np.where(b > 0, pred, 0).astype(np.uint8))
    _check_class_tables(ctx, preds, 3, maxc=4096)


def test_bounding_boxes_batched_a4_and_many_classes(ctx):
    """pcs_bounding_boxes on a batch of scaled A4 class maps (one labelling for all classes, one scan of the per-class
    difference arrays) and with nine classes (one labelling per class)."""
    import torch
    preds = []
    for s in range(3):
        pred, b = _pred_and_binary(50 + s, 1169, 827, n_classes=4)
        preds.append(np.where(b > 0, pred, 0).astype(np.uint8))
    preds[2][:] = 3                                                          # one component covering the page
    d = torch.from_numpy(np.stack(preds)).cuda()
    out = torch.full_like(d, 77)
    ctx.bounding_boxes(d, 3, 1169, 827, 4, out)
    for i in range(3):
        np.testing.assert_array_equal(out[i].cpu().numpy(), opipe.add_bounding_boxes(preds[i].astype(np.int64)))
    rng = np.random.default_rng(3)
    p9 = np.kron(rng.integers(0, 9, size=(30, 40)), np.ones((6, 8), np.int64))[:170, :300].astype(np.uint8)
    d9 = torch.from_numpy(p9[None]).cuda()
    out9 = torch.full_like(d9, 77)
    ctx.bounding_boxes(d9, 1, 170, 300, 9, out9)
    np.testing.assert_array_equal(out9[0].cpu().numpy(), opipe.add_bounding_boxes(p9.astype(np.int64)))


def test_vote_noise_binary_and_shapes_across_tiles(ctx):
    """vote_connected_component_class with tile-local histograms: salt-and-pepper foreground (more root candidates per
    labelling tile than the shared-memory table holds), a spiral and a comb that leave tiles and come back, a page of
    foreground only; batched, against the restated reference."""
    import torch
    rng = np.random.default_rng(21)
    h, w = 200, 700
    noise = (rng.random((h, w)) < 0.4).astype(np.uint8)
    spiral = np.zeros((h, w), np.uint8)
    y0, y1, x0, x1 = 2, h - 3, 2, w - 3
    while y1 - y0 > 8 and x1 - x0 > 8:
        spiral[y0, x0:x1 + 1] = 1
        spiral[y0:y1 + 1, x1] = 1
        spiral[y1, x0 + 4:x1 + 1] = 1
        spiral[y0 + 4:y1 + 1, x0 + 4] = 1
        spiral[y0 + 4, x0 + 4:x1 - 3] = 1
        y0 += 4; y1 -= 4; x0 += 4; x1 -= 4                                   # noqa: E702
    comb = np.zeros((h, w), np.uint8)
    comb[h - 2, :] = 1
    comb[3:h - 2, ::3] = 1
    binaries = [noise, spiral, comb, np.ones((h, w), np.uint8)]
    preds = [rng.integers(0, 3, size=(h, w)).astype(np.uint8) for _ in binaries]
    d_pred = torch.from_numpy(np.stack(preds)).cuda()
    d_bin = torch.from_numpy(np.stack(binaries)).cuda()
    ctx.cc_majority(d_pred, d_bin, len(binaries), h, w, 3)
    got = d_pred.cpu().numpy()
    for i, (p, b) in enumerate(zip(preds, binaries)):
        np.testing.assert_array_equal(got[i], opipe.vote_connected_component_class(p.astype(np.int64), b), err_msg=f"page {i}")
    # five classes: the per-run global path
    p5 = rng.integers(0, 5, size=(h, w)).astype(np.uint8)
    d5 = torch.from_numpy(p5[None]).cuda()
    ctx.cc_majority(d5, torch.from_numpy(spiral[None]).cuda(), 1, h, w, 5)
    np.testing.assert_array_equal(d5[0].cpu().numpy(), opipe.vote_connected_component_class(p5.astype(np.int64), spiral))


@pytest.mark.parametrize("hw", [(1, 1), (1, 40), (40, 1), (2, 33), (32, 256), (31, 255), (64, 257)])
def test_class_components_and_boxes_degenerate_shapes(ctx, hw):
    """One-pixel pages, single rows / columns, pages that are exactly one labelling tile or one pixel more."""
    import torch
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    preds = [rng.integers(0, 3, size=hw).astype(np.uint8), np.zeros(hw, np.uint8),
             (np.add.outer(np.arange(hw[0]), np.arange(hw[1])) % 3).astype(np.uint8)]
    _check_class_tables(ctx, preds, 3)
    d = torch.from_numpy(np.stack(preds)).cuda()
    out = torch.full_like(d, 9)
    ctx.bounding_boxes(d, len(preds), hw[0], hw[1], 3, out)
    binaries = [(p > 0).astype(np.uint8) for p in preds]
    d_bin = torch.from_numpy(np.stack(binaries)).cuda()
    d_vote = d.clone()
    ctx.cc_majority(d_vote, d_bin, len(preds), hw[0], hw[1], 3)
    for i, p in enumerate(preds):
        np.testing.assert_array_equal(out[i].cpu().numpy(), opipe.add_bounding_boxes(p.astype(np.int64)))
        np.testing.assert_array_equal(d_vote[i].cpu().numpy(), opipe.vote_connected_component_class(p.astype(np.int64), binaries[i]))
