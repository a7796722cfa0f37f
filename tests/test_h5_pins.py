"""Pins for the HDF5 reader (lib/h5.py) that do not come from its own writer.

1. `tests/golden/libhdf5_matlab_7.4.mat`: a file written by the real HDF5 library (MATLAB 7.4's -v7.3 writer; it
   ships with scipy as scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, BSD-licensed test data).  Superblock v0 behind a
   512-byte user block, version-1 object headers, an old-style root group, a contiguous float64 dataset and a
   fixed-length string attribute -- the same structural path a Keras file takes.
2. `tests/golden/keras_fcn_skip_tiny.h5`: assembled byte by byte from the format specification by
   `tests/golden/make_keras_h5_fixture.py` (no code shared with h5.py), laid out like an h5py-written
   `model.save('x.h5')` of the reference's fcn_skip graph (network.py:75-84).
"""
import json
import os
import sys

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)


def test_reader_on_a_file_written_by_libhdf5():
    from page_segmentation_b200.lib import h5
    f = h5.H5File(os.path.join(GOLDEN, "libhdf5_matlab_7.4.mat"))
    assert f.base == 512 and (f.O, f.L) == (8, 8)
    root = f.root()
    assert root.link_order == ["testdouble"]
    ds = f.child(root, "testdouble")
    assert not ds.is_group and ds.dataspace == (9, 1) and ds.layout[0] == "contiguous"
    assert ds.attrs == {"MATLAB_class": "double"}
    got = f.read_dataset(ds)
    assert got.dtype == np.float64
    assert np.array_equal(got, np.linspace(0, 2 * np.pi, 9).reshape(9, 1))      # what scipy's own test expects of it


def test_keras_layout_fixture_matches_its_generator():
    import make_keras_h5_fixture as mk
    with open(mk.OUT, "rb") as fh:
        assert fh.read() == mk.build(), "committed fixture and generator differ: regenerate or revert"


def test_reader_on_the_hand_assembled_keras_file():
    import make_keras_h5_fixture as mk
    from page_segmentation_b200.lib import h5
    m = h5.load_keras_model(mk.OUT)
    assert m.name == "fcn_skip" and m.keras_version == "2.6.0"
    W = mk.expected_weights()
    assert m.layer_names == [n for n in mk.LAYERS if n in W]                # layer_names0 + layer_names1, weighted only
    for ln, (k, b) in zip(m.layer_names, m.weights):
        assert k.shape == mk.KERNELS[ln]
        assert np.array_equal(k, W[ln][0]) and np.array_equal(b, W[ln][1]), ln

    f = h5.H5File(mk.OUT)
    root = f.root()
    assert json.loads(root.attrs["model_config"])["config"]["layers"][0]["config"]["note"] == "ü"     # vlen UTF-8 via GCOL
    assert root.attrs["backend"] == "tensorflow"
    mw = f.child(root, "model_weights")
    assert sorted(mw.link_order) == sorted(mk.LAYERS) and len(mw.link_order) == 24   # several symbol-table nodes
    assert list(mw.attrs["layer_names1"]) == mk.LAYERS[13:]
    assert mw.attrs["keras_version"] == "2.6.0"                             # lives in the continuation block
    pool = f.child(mw, "max_pooling2d")
    assert pool.is_group and pool.link_order == [] and np.asarray(pool.attrs["weight_names"]).shape == (0,)
    chunked = f.resolve(mw, f"{mk.CHUNKED}/{mk.CHUNKED}/kernel:0")
    assert chunked.layout[0] == "chunked" and chunked.layout[2][:4] == (3, 2, 4, 3)
    assert f.resolve(mw, f"{mk.COMPACT}/{mk.COMPACT}/bias:0").layout[0] == "compact"


def test_keras_fixture_weight_list_has_the_fcn_skip_structure():
    """The (kernel, bias) list the reader returns has the structure the forward pass expects: 13 weighted layers of
    fcn_skip in graph order, transposed kernels as (kh, kw, out, in)."""
    import make_keras_h5_fixture as mk
    from page_segmentation_b200.lib import h5
    m = h5.load_keras_model(mk.OUT)
    assert len(m.weights) == 13
    for (k, b), ln in zip(m.weights, m.layer_names):
        out_ch = k.shape[2] if "transpose" in ln else k.shape[3]
        assert b.shape == (out_ch,), ln
