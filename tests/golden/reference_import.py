"""Imports the REAL reference modules that run with numpy + cv2 alone.

`/root/reference` exists only in the build container.  tensorflow, scikit-image and ocr4all-pylib are not
installed, so the modules that need them at CALL time (network.py, model.py, the skimage part of dataset.py /
util.py) stay out of reach; but `xycut`, `pc_segmentation`, `postprocess`, `cc`, `image_ops`, `output` and the
dataclasses of `dataset` only need those packages at IMPORT time.  Empty stand-in modules satisfy the imports,
`ocr4all.colors.ColorMap` is replaced by this repo's stand-in (its surface is inferred from
the call sites), and every function exercised through this helper then executes the reference's own source
unmodified, against the cv2 / numpy installed here (4.13 / 2.3 instead of the pinned 4.5.5.62 / 1.x).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ocr4all_pixel_classifier", "lib"))


def _stub(name, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m


def load(module: str):
    """module: e.g. 'xycut' -> ocr4all_pixel_classifier.lib.xycut of the mounted reference."""
    if not available():
        raise RuntimeError("the reference tree is not mounted")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from page_segmentation_b200.lib.colors import ColorMap
    _stub("skimage")
    _stub("skimage.transform", resize=None, rescale=None)
    _stub("skimage.io", imsave=None)
    _stub("ocr4all")
    _stub("ocr4all.files", imread=None, imread_bin=None, random_indices=None, chunks=None, split_filename=None)
    _stub("ocr4all.colors", ColorMap=ColorMap)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    return importlib.import_module("ocr4all_pixel_classifier.lib." + module)
