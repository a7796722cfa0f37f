"""CPU-side checks: the drop-in API surface mirrors the reference's names/fields, the C-ABI
library loads and exports every symbol include/pcseg_b200.h declares, the Keras HDF5
reader/writer, and loud failure without a GPU (no CPU fallback)."""
import dataclasses
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from page_segmentation_b200 import _native
    header = open(os.path.join(ROOT, "include", "pcseg_b200.h")).read()
    declared = re.findall(r"PCS_API\s+[\w\s\*]+?\b(pcs_\w+)\s*\(", header)
    assert len(declared) >= 19
    lib = _native.load()
    for sym in declared:
        assert hasattr(lib, sym), f"libpcseg_b200.so does not export {sym}"
    assert sorted(declared) == sorted(_native.EXPORTS)
    assert lib.pcs_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from page_segmentation_b200 import _native, runtime
    with pytest.raises(_native.PcsError):
        _native.Context(0)
    with pytest.raises(_native.PcsError):
        runtime.get_context(0)
    from page_segmentation_b200.lib.dataset import prepare_images
    with pytest.raises(_native.PcsError):
        prepare_images(np.zeros((8, 8), np.uint8), np.zeros((8, 8), np.uint8), 6, 18)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "page_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py") or fn.endswith(".cu") or fn.endswith(".cuh"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{fn} imports the oracle"


def test_predict_settings_fields_match_reference():
    from page_segmentation_b200.lib.predictor_data import PredictSettings, Prediction
    fields = [(f.name, f.default) for f in dataclasses.fields(PredictSettings)]
    assert fields == [("network", None), ("output", None), ("high_res_output", False), ("color_map", None),
                      ("n_classes", -1), ("post_process", None), ("gpu_allow_growth", False)]   # predictor_data.py:18-26
    assert Prediction._fields == ("labels", "probabilities", "data")                               # :12-15


def test_single_data_and_dataset_match_reference():
    from page_segmentation_b200.lib.dataset import Dataset, SingleData
    names = [f.name for f in dataclasses.fields(SingleData)]
    assert names == ["image", "binary", "orig_binary", "mask", "image_path", "binary_path", "mask_path",
                     "line_height_px", "original_shape", "output_path", "user_data"]            # dataset.py:17-29
    assert SingleData().line_height_px == 1
    ds = Dataset([SingleData(), SingleData()], None)
    assert len(ds) == 2 and len(list(ds)) == 2


def test_masks_dataclass_and_postprocessor_registry():
    from page_segmentation_b200.lib.output import Masks
    from page_segmentation_b200.lib import postprocess as pp
    assert [f.name for f in dataclasses.fields(Masks)] == ["color", "overlay", "inverted_overlay", "fg_color_mask"]
    assert set(pp.POSTPROCESSORS) == {"ccmajority", "ccvote", "voteconnectedcomponents", "votecomponents",
                                      "boundingboxes", "bbox"}                                   # postprocess.py:57-64
    assert pp.find_postprocessor("CC_Majority") is pp.vote_connected_component_class
    assert pp.find_postprocessor("bounding-boxes") is pp.add_bounding_boxes
    with pytest.raises(KeyError):
        pp.find_postprocessor("unknown")
    assert "cc_majority" in pp.postprocess_help()


def test_architecture_enum_and_preprocess():
    from page_segmentation_b200.lib.architecture import Architecture, default_preprocess
    assert Architecture("fcn_skip") is Architecture.FCN_SKIP and Architecture.UNET.value == "unet"
    fn, rgb = Architecture.FCN_SKIP.preprocess()
    assert rgb is False and fn(np.array([255.0]))[0] == 1.0 and fn is default_preprocess
    with pytest.raises(NotImplementedError):
        Architecture.RES_NET.preprocess()


def test_util_helpers():
    from page_segmentation_b200.lib.util import gray_to_rgb, image_to_batch
    a = np.zeros((4, 5))
    assert image_to_batch(a).shape == (1, 4, 5, 1)
    assert gray_to_rgb(a).shape == (4, 5, 3)


def test_color_map_json_schema_roundtrip(tmp_path):
    from page_segmentation_b200.lib.colors import ColorMap
    raw = {"(255, 255, 255)": [0, "background"], "(255, 0, 0)": [1, "text"], "(0, 255, 0)": [2, "image"]}
    p = tmp_path / "cm.json"
    p.write_text(json.dumps(raw))
    cm = ColorMap.load(str(p))
    assert len(cm) == 3 and cm.color_for_label("text") == (255, 0, 0)
    assert cm.to_json() == raw
    lut = cm.lut(4)
    assert lut.shape == (4, 3) and lut[2].tolist() == [0, 255, 0] and lut[3].tolist() == [0, 0, 0]
    labels = np.array([[0, 2], [1, 7]])
    rgb = cm.to_rgb_array(labels)
    assert rgb[0, 1].tolist() == [0, 255, 0] and rgb[1, 1].tolist() == [0, 0, 0]
    assert cm.filter_label(rgb, "image").tolist() == [[False, True], [False, False]]
    assert cm.rgb_to_labels(rgb).tolist() == [[0, 2], [1, 0]]


def test_network_argument_errors():
    from page_segmentation_b200.lib.network import Network
    from page_segmentation_b200 import synth
    with pytest.raises(NotImplementedError):
        Network("train", n_classes=3, weights=synth.make_weights("fcn_skip", 3, 0))
    with pytest.raises(ValueError):
        Network("Predict", n_classes=3)
    with pytest.raises(FileNotFoundError):
        Network("Predict", n_classes=3, model="/nonexistent/model")                  # '.h5' is appended (network.py:59)
    with pytest.raises(ValueError):
        Network("Predict", n_classes=4, weights=synth.make_weights("fcn_skip", 3, 0))
    net = Network("Predict", weights=synth.make_weights("fcn_skip", 5, 0))
    assert net.n_classes == 5 and net.architecture == "fcn_skip" and net.model.name == "model"


def test_keras_h5_roundtrip_full_model_and_weights_only(tmp_path):
    from page_segmentation_b200 import synth
    from page_segmentation_b200.lib import h5
    from page_segmentation_b200.lib.network import Network
    for arch in ("fcn_skip", "fcn"):
        W = synth.make_weights(arch, 3, seed=4)
        full = str(tmp_path / f"{arch}.h5")
        h5.write_keras_h5(full, W, arch, extra_layers=["lambda", "lambda_1", "max_pooling2d", "concatenate"])
        m = h5.load_keras_model(full)
        assert m.name == arch and len(m.weights) == 13 and m.keras_version == "2.5.0"
        for (k, b), (k2, b2) in zip(W, m.weights):
            np.testing.assert_array_equal(k, k2)
            np.testing.assert_array_equal(b, b2)
        net = Network("Predict", n_classes=3, model=full[:-3])               # path without '.h5' (network.py:59)
        assert net.model.name == arch and net._arch == arch
    wo = str(tmp_path / "weights_only.h5")
    W = synth.make_weights("fcn_skip", 3, seed=5)
    h5.write_keras_h5(wo, W, None, weights_only=True)
    m = h5.load_keras_model(wo)
    assert m.name is None and len(m.weights) == 13
    net = Network("Predict", n_classes=3, model=wo)
    assert net.model.name == "model" and net._arch == "fcn_skip"             # falls back to model_constructor


def test_h5_reader_low_level_structures(tmp_path):
    from page_segmentation_b200.lib import h5
    W = [(np.arange(2 * 2 * 3 * 4, dtype=np.float32).reshape(2, 2, 3, 4), np.arange(4, dtype=np.float32))]
    p = str(tmp_path / "one.h5")
    h5.write_keras_h5(p, W, "fcn", layer_names=["logits"])
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0              # superblock version 0
    f = h5.H5File(p)
    root = f.root()
    assert root.is_group and "model_weights" in root.links
    assert json.loads(root.attrs["model_config"])["config"]["name"] == "fcn"
    mw = f.child(root, "model_weights")
    assert list(mw.attrs["layer_names"]) == ["logits"]
    ds = f.resolve(mw, "logits/logits/kernel:0")
    assert ds.dataspace == (2, 2, 3, 4) and ds.layout[0] == "contiguous"
    np.testing.assert_array_equal(f.read_dataset(ds), W[0][0])
    with pytest.raises(h5.H5Error):
        bad = tmp_path / "bad.h5"
        bad.write_bytes(b"not hdf5" * 10)
        h5.H5File(str(bad))


def test_scaled_and_padded_shapes():
    from page_segmentation_b200 import synth
    assert synth.scaled_shape(3508, 2480, 6 / 18) == (1169, 827)
    assert synth.padded_shape(1169, 827) == (1184, 832)
    assert abs(synth.A4_MPX - 8.69984) < 1e-9


def test_synthetic_page_is_deterministic_and_binarised():
    from page_segmentation_b200 import synth
    a, b = synth.make_page(3, 400, 300), synth.make_page(3, 400, 300)
    np.testing.assert_array_equal(a, b)
    assert set(np.unique(a)) <= {0, 255} and 0.02 < (a == 0).mean() < 0.5
    assert len(np.unique(synth.make_grey_page(3, 200, 150))) > 2


def test_page_locked_pool_is_bounded_and_recycles():
    """lazy.PinnedPool: blocks come back when the array handed out (and every view of it) has died; a caller that keeps
    everything runs into the byte budget and is refused (-> pageable copy); no timing is involved."""
    import gc
    from page_segmentation_b200.lazy import PinnedPool
    made = []

    import torch

    def block(n):                              # pageable here; the pool only needs `.numpy()` (base = the tensor)
        made.append(n)
        return torch.zeros(n, dtype=torch.uint8)

    pool = PinnedPool(budget_bytes=4 << 20, allocator=block)
    assert PinnedPool.size_class(1) == 1 << 16 and PinnedPool.size_class((1 << 20) + 1) == 2 << 20
    a = pool.alloc(1 << 20)
    view = a[:100].view(np.float32).reshape(5, 5)
    del a
    gc.collect()
    assert pool.outstanding == 1 << 20 and not pool.free            # the view keeps the block out
    del view
    gc.collect()
    assert pool.outstanding == 0 and len(pool.free[1 << 20]) == 1
    b = pool.alloc(900_000)                                          # same size class: the block comes round
    assert made == [1 << 20] and pool.stats["hits"] == 1
    kept = [b] + [pool.alloc(1 << 20) for _ in range(3)]             # keep-all caller: the budget (4 MB) is reached
    assert all(k is not None for k in kept) and pool.total == 4 << 20
    assert pool.alloc(1 << 20) is None and pool.alloc(1) is None and pool.stats["refused"] == 2
    del kept, b
    gc.collect()
    held = pool.alloc(1 << 20)
    assert pool.outstanding == 1 << 20 and held is not None
    pool.trim()
    assert pool.total == 1 << 20                                      # only the block still handed out counts


def test_device_array_is_an_ndarray_stand_in():
    """lazy.DeviceArray: metadata without data access, one host copy at the first look, reported dtype wider than the
    device dtype (class maps: uint8 on the device, int64 like np.argmax to the caller), numpy protocols."""
    import torch
    from page_segmentation_b200.lazy import DeviceArray, is_lazy
    calls = []
    t = torch.arange(12, dtype=torch.uint8).reshape(3, 4)

    def source():
        calls.append(1)
        return t

    a = DeviceArray((3, 4), np.int64, source, device=0)
    assert a.shape == (3, 4) and a.dtype == np.int64 and a.ndim == 2 and a.size == 12 and len(a) == 3 and not calls
    assert is_lazy(a) and "device" in repr(a)
    assert a.device_tensor() is t and calls == [1] and is_lazy(a)
    h = np.asarray(a)
    assert h.dtype == np.int64 and h.tolist() == t.tolist() and not is_lazy(a) and calls == [1]
    assert np.asarray(a) is h                                          # the host copy IS the array from now on
    a[0, 0] = 7
    assert h[0, 0] == 7 and int(a.max()) == 11 and a.astype(np.uint8).dtype == np.uint8
    assert (a + 1)[0, 0] == 8 and (a == h).all() and np.array_equal(a, h) and np.stack([a, a]).shape == (2, 3, 4)
    with pytest.raises(AttributeError):
        a._missing


def test_colour_lut_is_padded_and_checked_before_it_reaches_the_library():
    """pcs_forward / pcs_predict_pages_* read n_classes rows of the caller's LUT; the binding pads the table
    to 256 rows and refuses one that is shorter than the model's class count."""
    from page_segmentation_b200 import _native
    lut = _native._lut256([[255, 255, 255], [255, 0, 0]])
    assert lut.shape == (256, 3) and lut.dtype == np.uint8 and lut.flags["C_CONTIGUOUS"]
    assert lut[1].tolist() == [255, 0, 0] and not lut[2:].any()
    assert _native._lut256(None) is None
    with pytest.raises(_native.PcsError):
        _native._lut256([[0, 0, 0], [1, 1, 1]], n_classes=3)
    with pytest.raises(_native.PcsError):
        _native._lut256(np.zeros((257, 3), np.uint8))
