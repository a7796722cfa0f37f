"""End-to-end drop-in surface on the device: DatasetLoader -> Predictor -> Masks / files, the
host-buffer batch pipeline (pcs_predict_pages_host) and model loading from a Keras .h5."""
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import network as onet
from oracle import pipeline as opipe
from page_segmentation_b200 import synth
from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP

pytestmark = pytest.mark.gpu
LUT = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}


def _loaded(seed, h=450, w=330, lh=18):
    from page_segmentation_b200.lib.dataset import DatasetLoader, SingleData
    page = synth.make_page(seed, h, w, lh)
    data = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True).load_images(SingleData(image=page, line_height_px=lh))
    return page, data


def test_dataset_loader_quirks(ctx):
    from page_segmentation_b200.lib.dataset import DatasetLoader, SingleData
    page = synth.make_page(1, 300, 240, 18)
    other = synth.make_page(2, 300, 240, 18)
    # dataset.py:172: the binary is derived from `image`; a caller-supplied `binary` is ignored and overwritten
    d = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True).load_images(SingleData(image=page, binary=other, line_height_px=18))
    eimg, eb, eob = opipe.prepare_images(page, page, 6, 18, keep_orig_bin=True)
    np.testing.assert_array_equal(d.image, eimg)
    np.testing.assert_array_equal(d.binary, eb)
    np.testing.assert_array_equal(d.orig_binary, eob)
    assert d.original_shape == (300, 240) and d.mask is None


def test_predictor_predict_and_masks_with_postprocessing(ctx):
    from page_segmentation_b200.lib.dataset import Dataset
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.postprocess import find_postprocessor
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    W = synth.make_weights("fcn_skip", 3, seed=1)
    net = Network("Predict", n_classes=3, weights=W, precision="fp16")
    pages = [_loaded(s) for s in (3, 4)]
    settings = PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, post_process=[find_postprocessor("cc_majority")])
    pred = Predictor(settings, network=net)
    gen = pred.predict(Dataset([d for _, d in pages], DEFAULT_COLOR_MAP))
    assert hasattr(gen, "__next__")                                # lazy generator like predictor.py:27-30
    for (page, data), p in zip(pages, gen):
        logit, prob, raw = net.predict_single_data(data)
        exp = opipe.vote_connected_component_class(raw.copy(), data.binary)
        assert p.labels.dtype == np.int64 and p.probabilities.dtype == np.float32 and p.data is data
        np.testing.assert_array_equal(p.labels, exp)
        l64 = onet.Forward("fcn_skip", W, 3, dtype=torch.float64).logits(data.image)[0]
        assert (raw == l64.argmax(-1)).mean() >= 0.999
        m = pred.predict_masks(data)
        c, o, i, f = opipe.generate_output_masks(data.binary, exp, LUT)
        np.testing.assert_array_equal(m.color, c)
        np.testing.assert_array_equal(m.overlay, o)
        np.testing.assert_array_equal(m.inverted_overlay, i)
        np.testing.assert_array_equal(m.fg_color_mask, f)


def test_results_are_owned_by_the_caller(ctx):
    """Host copies of results come from a bounded pool of page-locked blocks (lazy.PinnedPool): an array a caller keeps
    must never be rewritten by a later call, blocks of dropped arrays go round, and the pageable path (pool exhausted
    or disabled) gives the same values."""
    import gc
    from page_segmentation_b200 import lazy
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    W = synth.make_weights("fcn_skip", 3, seed=2)
    net = Network("Predict", n_classes=3, weights=W, precision="fp16")
    pred = Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP), network=net)
    datas = [_loaded(s, h=1350, w=990)[1] for s in (5, 6, 7)]        # large enough for the page-locked path
    kept = [pred.predict_single(d) for d in datas]
    assert all(lazy.is_lazy(p.labels) and p.labels.dtype == np.int64 and p.probabilities.dtype == np.float32 for p in kept)
    snap = [(p.labels.copy(), p.probabilities.copy()) for p in kept]
    assert not any(lazy.is_lazy(p.labels) for p in kept)
    assert len({p.labels.ctypes.data for p in kept}) == 3 and len({p.probabilities.ctypes.data for p in kept}) == 3
    for _ in range(3):                                         # dropped results: their blocks go round
        for d in datas:
            p = pred.predict_single(d)
            q = np.asarray(p.labels), np.asarray(p.probabilities)
            m = pred.predict_masks(d)
            del p, q, m
        gc.collect()
    assert lazy.pinned_pool().stats["hits"] > 0
    for p, (l, q), d in zip(kept, snap, datas):
        np.testing.assert_array_equal(p.labels, l)
        np.testing.assert_array_equal(p.probabilities, q)
        assert p.labels.flags.writeable and p.labels.flags.c_contiguous
        again = pred.predict_single(d)
        np.testing.assert_array_equal(again.labels, l)
        np.testing.assert_array_equal(again.probabilities, q)
        logit, prob, raw = net.predict_single_data(d)          # the eager reference-named call gives the same
        np.testing.assert_array_equal(raw, l)
        np.testing.assert_array_equal(prob, q)
    saved = lazy._ENABLED
    try:
        lazy._ENABLED = False                                  # the pageable copy
        p = pred.predict_single(datas[0])
        np.testing.assert_array_equal(p.labels, snap[0][0])
        np.testing.assert_array_equal(p.probabilities, snap[0][1])
    finally:
        lazy._ENABLED = saved


def test_lazy_flow_matches_page_by_page(ctx, tmp_path):
    """load_data -> predict (+ cc_majority) -> output_data with everything staying on the device (pipeline.py) against the
    same pages taken one at a time through host arrays and the oracle: pages of two sizes (chunks break at a size
    change), more pages than one chunk, a foreign post-processor (runs on host arrays) after the registry's own."""
    import cv2
    from page_segmentation_b200 import lazy
    from page_segmentation_b200.lib.dataset import Dataset, DatasetLoader, SingleData
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.output import flush_outputs, output_data
    from page_segmentation_b200.lib.postprocess import find_postprocessor
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    W = synth.make_weights("fcn_skip", 3, seed=4)
    net = Network("Predict", n_classes=3, weights=W, precision="fp16")
    shapes = [(420, 300)] * 11 + [(390, 330)] * 2 + [(420, 300)]
    pages = [synth.make_page(40 + i, h, w, 18) for i, (h, w) in enumerate(shapes)]
    entries = [SingleData(image=p, line_height_px=18, output_path=f"p{i:02d}.png") for i, p in enumerate(pages)]
    loader = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True)
    ds = loader.load_data(entries)
    assert isinstance(ds, Dataset) and len(ds) == len(pages)
    assert all(lazy.is_lazy(lazy.peek(d, "image")) and lazy.is_lazy(lazy.peek(d, "binary")) and d.original_shape == s
               for d, s in zip(ds.data, shapes))
    seen = []

    def foreign(pred, data):
        assert isinstance(pred, np.ndarray) and pred.dtype == np.int64
        seen.append(pred.shape)
        return pred

    out = str(tmp_path)
    settings = PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, output=out, post_process=[find_postprocessor("cc_majority")])
    predictor = Predictor(settings, network=net)
    preds = []
    for p in predictor.predict(ds):
        assert lazy.is_lazy(p.labels)
        output_data(out, p.labels, p.data, DEFAULT_COLOR_MAP)
        preds.append(p)
    flush_outputs()
    for i, (page, p) in enumerate(zip(pages, preds)):
        eimg, ebin = opipe.prepare_images(page, page, 6, 18)
        np.testing.assert_array_equal(p.data.image, eimg)
        np.testing.assert_array_equal(p.data.binary, ebin)
        _, _, raw = net.predict_single_data(SingleData(image=eimg))
        exp = opipe.vote_connected_component_class(raw.copy(), ebin)
        np.testing.assert_array_equal(p.labels, exp)
        c, o, inv, _ = opipe.generate_output_masks(ebin, exp, LUT)
        for cat, want in (("color", c), ("overlay", o), ("inverted", inv)):
            got = cv2.imread(f"{out}/{cat}/p{i:02d}.png", cv2.IMREAD_COLOR)[..., ::-1]
            np.testing.assert_array_equal(got, want)
    settings2 = PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, post_process=[find_postprocessor("cc_majority"), foreign])
    ds2 = loader.load_data([SingleData(image=p, line_height_px=18) for p in pages[:3]])
    for p, q in zip(Predictor(settings2, network=net).predict(ds2), preds):
        assert isinstance(p.labels, np.ndarray)
        np.testing.assert_array_equal(p.labels, q.labels)
    assert len(seen) == 3
    np.testing.assert_array_equal(ds2.data[0].orig_binary, (pages[0] == 0).astype(np.uint8))      # lazily, on request


def test_high_res_output(ctx):
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    W = synth.make_weights("fcn_skip", 3, seed=2)
    net = Network("Predict", n_classes=3, weights=W)
    page, data = _loaded(5, 333, 241)
    p = Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, high_res_output=True), network=net).predict_single(data)
    _, _, raw = net.predict_single_data(data)
    eimg, ebin, epred = opipe.scale_to_original_shape(data.image, data.binary, data.orig_binary, data.original_shape, raw)
    assert p.labels.shape == page.shape and p.labels.dtype == np.int64
    np.testing.assert_array_equal(p.labels, epred)
    np.testing.assert_array_equal(p.data.image, eimg)
    np.testing.assert_array_equal(p.data.binary, data.orig_binary)      # output.py:70-71
    assert data.image.shape != page.shape                                # the input SingleData is not mutated (replace())


def test_output_data_writes_three_images(ctx, tmp_path):
    from page_segmentation_b200.lib.output import flush_outputs, output_data
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    from page_segmentation_b200.lib.network import Network
    net = Network("Predict", n_classes=3, weights=synth.make_weights("fcn_skip", 3, seed=3))
    _, data = _loaded(6, 240, 200)
    data.image_path = "/somewhere/page_0001.png"
    out = str(tmp_path / "out")
    Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP, output=out), network=net)
    for sub in ("color", "overlay", "inverted"):
        assert os.path.isdir(os.path.join(out, sub))                     # predictor.py:21-25
    _, _, pred = net.predict_single_data(data)
    output_data(out, pred[None], data, DEFAULT_COLOR_MAP)                 # leading batch dim is squeezed (output.py:21-23)
    flush_outputs()
    c, o, i, _ = opipe.generate_output_masks(data.binary, pred, LUT)
    for sub, exp in (("color", c), ("overlay", o), ("inverted", i)):
        img = cv2.imread(os.path.join(out, sub, "page_0001.png"), cv2.IMREAD_COLOR)[..., ::-1]
        np.testing.assert_array_equal(img, exp)


def test_model_from_keras_h5(ctx, tmp_path):
    from page_segmentation_b200.lib import h5
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    W = synth.make_weights("fcn", 3, seed=7)
    path = str(tmp_path / "model.h5")
    h5.write_keras_h5(path, W, "fcn", extra_layers=["lambda", "max_pooling2d"])
    _, data = _loaded(7, 200, 260)
    pred = Predictor(PredictSettings(network=path, n_classes=3, color_map=DEFAULT_COLOR_MAP))     # builds its own Network
    assert pred.network.model.name == "fcn"
    logit, _, labels = pred.network.predict_single_data(data)
    l32 = onet.Forward("fcn", W, 3).logits(data.image)[0]
    assert np.abs(logit - l32).max() <= 8e-3
    assert (labels == l32.argmax(-1)).mean() >= 0.995


@pytest.mark.parametrize("cc", [False, True])
def test_host_batch_pipeline_matches_stagewise(ctx, cc, monkeypatch):
    """pcs_predict_pages_host (sub-batched, copy/compute overlapped) == stage-by-stage device calls."""
    from page_segmentation_b200.runtime import PageBatchEngine
    monkeypatch.setenv("PCSEG_HOST_CHUNK", "2")
    n, H, W_ = 5, 360, 300
    pages = np.stack([synth.make_page(10 + s, H, W_, 18) for s in range(n)])
    weights = synth.make_weights("fcn_skip", 3, seed=9)
    lut = np.array([LUT[i] for i in range(3)], dtype=np.uint8)
    eng = PageBatchEngine("fcn_skip", weights, 3, precision="bf16", lut=lut)
    Hs, Ws = synth.scaled_shape(H, W_, 6 / 18)
    h_pages = torch.from_numpy(pages).pin_memory()
    out = {k: torch.empty((n, Hs, Ws) + ((3,) if k != "labels" else ()), dtype=torch.uint8).pin_memory().numpy()
           for k in ("labels", "color", "overlay", "inverted")}
    eng.run_host(h_pages.numpy(), 6 / 18, out, cc_majority=cc)
    dev = eng.run_device(torch.from_numpy(pages).cuda(), 6 / 18, cc_majority=cc)
    torch.cuda.synchronize()
    for k in out:
        np.testing.assert_array_equal(out[k], dev[k].cpu().numpy())
    for i in range(n):
        eimg, eb = opipe.prepare_images(pages[i], pages[i], 6, 18)
        np.testing.assert_array_equal(dev["image"][i].cpu().numpy(), eimg)
        c, o, inv, _ = opipe.generate_output_masks(eb, out["labels"][i].astype(np.int64), LUT)
        np.testing.assert_array_equal(out["color"][i], c)
        np.testing.assert_array_equal(out["overlay"][i], o)
        np.testing.assert_array_equal(out["inverted"][i], inv)
        if cc:      # voting is idempotent: a voted map is its own vote
            again = opipe.vote_connected_component_class(out["labels"][i].astype(np.int64), eb)
            np.testing.assert_array_equal(again, out["labels"][i])


def test_full_size_a4_page_properties(ctx):
    """BASELINE config 1 shape (2480x3508 -> 1169x827): size-independent properties at full size."""
    from page_segmentation_b200.lib.dataset import DatasetLoader, SingleData
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.postprocess import vote_connected_component_class
    page = synth.make_page(0)
    data = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True).load_images(SingleData(image=page, line_height_px=18))
    assert data.image.shape == (1169, 827)
    eimg, eb = opipe.prepare_images(page, page, 6, 18)
    np.testing.assert_array_equal(data.image, eimg)
    np.testing.assert_array_equal(data.binary, eb)
    W = synth.make_weights("fcn_skip", 3, seed=0)
    net = Network("Predict", n_classes=3, weights=W, precision="fp16")
    logit, prob, pred = net.predict_single_data(data)
    l32 = onet.Forward("fcn_skip", W, 3).logits(data.image)[0]
    assert np.abs(logit - l32).max() <= 1e-3
    assert (pred == l32.argmax(-1)).mean() >= 0.999
    np.testing.assert_allclose(prob.sum(-1), 1.0, atol=1e-5)
    voted = vote_connected_component_class(pred.copy(), data)
    np.testing.assert_array_equal(vote_connected_component_class(voted.copy(), data), voted)      # idempotent
    assert np.array_equal(voted[data.binary == 0], pred[data.binary == 0])                          # paper untouched


@pytest.mark.parametrize("cc", [False, True])
def test_host_batch_pipeline_with_segment_extraction(ctx, cc, monkeypatch):
    """pcs_predict_pages_segments (BASELINE configs[3]): the stats tables that come back with the class maps are cv2's
    for exactly those class maps (postprocess.py:31-33), page by page, over a chunked call."""
    from page_segmentation_b200.runtime import PageBatchEngine
    monkeypatch.setenv("PCSEG_HOST_CHUNK", "2")
    n, H, W_ = 5, 360, 300
    pages = np.stack([synth.make_page(30 + s, H, W_, 18) for s in range(n)])
    weights = synth.make_weights("fcn_skip", 3, seed=9)
    lut = np.array([LUT[i] for i in range(3)], dtype=np.uint8)
    eng = PageBatchEngine("fcn_skip", weights, 3, precision="fp16", lut=lut)
    Hs, Ws = synth.scaled_shape(H, W_, 6 / 18)
    maxc = 2048
    out = {"labels": np.zeros((n, Hs, Ws), np.uint8), "color": np.zeros((n, Hs, Ws, 3), np.uint8),
           "stats": np.full((n, 3, maxc, 5), -1, np.int32), "ncomp": np.zeros((n, 3), np.int32)}
    eng.run_host_segments(pages, 6 / 18, out, max_components=maxc, cc_majority=cc)
    ref = {"labels": np.zeros((n, Hs, Ws), np.uint8)}
    eng.run_host(pages, 6 / 18, ref, cc_majority=cc)
    np.testing.assert_array_equal(out["labels"], ref["labels"])
    for i in range(n):
        _, eb = opipe.prepare_images(pages[i], pages[i], 6, 18)
        c, _, _, _ = opipe.generate_output_masks(eb, out["labels"][i].astype(np.int64), LUT)
        np.testing.assert_array_equal(out["color"][i], c)
        for k in range(3):
            m, _, stats, _ = cv2.connectedComponentsWithStats((out["labels"][i] == k).astype(np.uint8), connectivity=4)
            assert out["ncomp"][i, k] == m
            kk = min(m, maxc)
            np.testing.assert_array_equal(out["stats"][i, k, :kk], stats[:kk])
            assert not out["stats"][i, k, kk:].any()
    # the same call with compact results (pcs_predict_pages_segments_compact): class map, bit-packed binary, tables
    from page_segmentation_b200.runtime import unpack_bits_host
    comp = {"labels": np.zeros((n, Hs, Ws), np.uint8), "binary_bits": np.zeros((n, (Hs * Ws + 31) // 32), np.uint32),
            "stats": np.full((n, 3, maxc, 5), -1, np.int32), "ncomp": np.zeros((n, 3), np.int32)}
    eng.run_host_segments_compact(pages, 6 / 18, comp, max_components=maxc, cc_majority=cc)
    np.testing.assert_array_equal(comp["labels"], out["labels"])
    np.testing.assert_array_equal(comp["stats"], out["stats"])
    np.testing.assert_array_equal(comp["ncomp"], out["ncomp"])
    binary = unpack_bits_host(comp["binary_bits"], (Hs, Ws))
    for i in range(n):
        _, eb = opipe.prepare_images(pages[i], pages[i], 6, 18)
        np.testing.assert_array_equal(binary[i], eb)


def test_loaded_pages_are_freed_without_the_cyclic_collector(ctx):
    """DatasetLoader.load_data -> Predictor.predict -> drop everything: the device tensors of the pages must go back to the
    allocator by reference counting alone.  (A closure of the loader once held its SingleData, which held the
    DeviceArray, which held the closure: 118 MB per 64 A4 pages piled up between collector passes.)"""
    import gc
    from page_segmentation_b200.lib.colors import DEFAULT_COLOR_MAP
    from page_segmentation_b200.lib.dataset import DatasetLoader, SingleData
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200.lib.predictor import Predictor
    from page_segmentation_b200.lib.predictor_data import PredictSettings
    pages = [synth.make_page(s, 600, 420, 18) for s in range(6)]
    net = Network("Predict", n_classes=3, weights=synth.make_weights("fcn_skip", 3, seed=0))
    predictor = Predictor(PredictSettings(n_classes=3, color_map=DEFAULT_COLOR_MAP), network=net)
    loader = DatasetLoader(6, DEFAULT_COLOR_MAP, prediction=True)

    def flow():
        dataset = loader.load_data([SingleData(image=p, line_height_px=18) for p in pages])
        for pred in predictor.predict(dataset):
            assert pred.labels.shape == pred.data.image.shape
    flow()
    flow()
    gc.collect()
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    gc.disable()
    try:
        flow()
        torch.cuda.synchronize()
        assert torch.cuda.memory_allocated() <= before + (1 << 20), (before, torch.cuda.memory_allocated())
    finally:
        gc.enable()
