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


def test_page_locked_result_policy(monkeypatch):
    """runtime._pinned_empty: fast allocations (blocks recycled by the caching host allocator) stay page-locked; a run of
    slow ones (a caller that keeps every result: each block is a fresh cudaHostAlloc) switches to the pageable copy for a
    doubling number of results and comes back at the first fast one; one state per call site."""
    from page_segmentation_b200 import runtime
    now = [0.0]

    class FakeTorch:
        slow = False

        def empty(self, shape, dtype=None, pin_memory=False):
            assert pin_memory
            now[0] += runtime._PIN_SLOW_S * (4 if self.slow else 0.01)
            return ("block", shape)

    t = FakeTorch()
    monkeypatch.setattr(runtime, "_clock", lambda: now[0])
    monkeypatch.setattr(runtime, "_pin_sites", {})
    monkeypatch.setattr(runtime, "_PIN_ENABLED", True)
    assert runtime._pinned_empty(t, (4,), None, None) is None                     # site None: always pageable
    for _ in range(3 * runtime._PIN_MISSES):
        assert runtime._pinned_empty(t, (4,), None, "a") is not None              # streaming caller: hits
    t.slow = True
    got = [runtime._pinned_empty(t, (4,), None, "a") is not None for _ in range(runtime._PIN_MISSES + 64)]
    assert all(got[:runtime._PIN_MISSES]) and not any(got[runtime._PIN_MISSES:])  # then 64 pageable results
    t.slow = False
    assert runtime._pinned_empty(t, (4,), None, "b") is not None                  # another site is unaffected
    probes = [runtime._pinned_empty(t, (4,), None, "a") is not None for _ in range(4)]
    assert all(probes) and runtime._pin_sites["a"]["misses"] == 0 and runtime._pin_sites["a"]["pause"] == 32
    t.slow = True                                                                 # keep-all again: the pause doubles
    n_pinned = sum(runtime._pinned_empty(t, (4,), None, "a") is not None for _ in range(400))
    assert n_pinned <= runtime._PIN_MISSES + 3 * (runtime._PIN_MISSES - runtime._PIN_MISSES // 2)
    monkeypatch.setattr(runtime, "_PIN_ENABLED", False)
    assert runtime._pinned_empty(t, (4,), None, "b") is None


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
