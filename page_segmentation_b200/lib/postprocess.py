# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/postprocess.py (:9-64) on the device
connected-component kernels."""
from typing import Callable

import numpy as np

from .dataset import SingleData


def _n_classes(pred: np.ndarray) -> int:
    return int(pred.max()) + 1 if pred.size else 1


def _pred_to_device(pred: np.ndarray, device: int):
    """Class map of any integer dtype -> (uint8 device tensor, n_classes).  The array goes up as it is (from
    Predictor.predict it is page-locked int64) and is range-checked and narrowed on the device, not by host passes."""
    import torch
    from ..runtime import to_device_u8
    if pred.size == 0 or pred.dtype not in (np.int64, np.int32, np.int16, np.uint8, np.int8):
        return to_device_u8(pred, device), _n_classes(pred)
    t = torch.from_numpy(np.ascontiguousarray(pred)).to(f"cuda:{device}")
    lo, hi = (int(v) for v in torch.aminmax(t))
    if lo < 0 or hi > 255:
        raise ValueError("values outside 0..255 cannot be staged as uint8")
    return t.to(torch.uint8), hi + 1


def _store(pred: np.ndarray, d_pred) -> np.ndarray:
    """Device class map -> the caller's array, widened on the device and copied straight into its memory."""
    import torch
    if pred.size and pred.flags.c_contiguous and pred.flags.writeable and pred.dtype in (np.int64, np.int32, np.int16, np.uint8):
        torch.from_numpy(pred).copy_(d_pred.to(getattr(torch, pred.dtype.name)))
    else:
        pred[...] = d_pred.cpu().numpy().astype(pred.dtype)
    return pred


def _binary_to_device(data: SingleData, device):
    import torch
    from ..lazy import DeviceArray, device_tensor_of, peek
    if isinstance(peek(data, "binary"), DeviceArray):
        return device_tensor_of(peek(data, "binary"), device)
    binary = np.ascontiguousarray(data.binary)
    if binary.dtype == np.bool_:
        binary = binary.view(np.uint8)
    if binary.dtype != np.uint8:
        binary = np.ascontiguousarray(binary != 0).view(np.uint8)     # the kernels take non-zero as foreground
    return torch.from_numpy(binary).to(device)


def _lazy_pred(pred):
    """(device tensor uint8, n_classes) of a class map that is still on the device, else None."""
    from ..lazy import DeviceArray
    if isinstance(pred, DeviceArray) and pred.on_device:
        t = pred.device_tensor()
        return t, (int(t.max()) + 1 if t.numel() else 1)
    return None


def vote_connected_component_class(pred: np.ndarray, data: SingleData) -> np.ndarray:
    """postprocess.py:9-26; like the reference, writes into `pred` and returns it."""
    from ..runtime import get_context
    ctx = get_context()
    h, w = pred.shape
    lazy = _lazy_pred(pred)
    d_pred, n_classes = lazy if lazy else _pred_to_device(np.asarray(pred), ctx.device)
    d_bin = _binary_to_device(data, d_pred.device)
    ctx.cc_majority(d_pred, d_bin, 1, h, w, n_classes)
    return pred if lazy else _store(np.asarray(pred), d_pred)


def add_bounding_boxes(pred: np.ndarray, data: SingleData) -> np.ndarray:
    """postprocess.py:29-42.  The reference hands cv2 a bool array, which cv2
    rejects; this implements the evident intent (components of pred == c)."""
    import torch
    from ..runtime import get_context
    ctx = get_context()
    h, w = pred.shape
    lazy = _lazy_pred(pred)
    d_pred, n_classes = lazy if lazy else _pred_to_device(np.asarray(pred), ctx.device)
    d_out = torch.empty((h, w), dtype=torch.uint8, device=d_pred.device)
    ctx.bounding_boxes(d_pred, 1, h, w, n_classes, d_out)
    if lazy:
        from ..lazy import DeviceArray
        return DeviceArray((h, w), pred.dtype, (lambda t=d_out: t), ctx.device)
    return _store(np.empty_like(np.asarray(pred)), d_out)


def class_components(pred: np.ndarray, n_classes: int = None, max_components: int = 65536):
    """Segment extraction (BASELINE configs[3]): what add_bounding_boxes computes per class before it paints
    (postprocess.py:31-33) -- `cv2.connectedComponentsWithStats(pred == c, connectivity=4)` for every class c -- as a
    list of `(num_labels, stats)` pairs, `stats` int32 (num_labels, 5) in cv2's column order (cc.py:4-18).  Not part
    of the reference's API: the reference throws these tables away after painting."""
    import torch
    from ..runtime import get_context
    ctx = get_context()
    h, w = pred.shape
    d_pred, seen = _lazy_pred(pred) or _pred_to_device(np.asarray(pred), ctx.device)
    n_classes = seen if n_classes is None else int(n_classes)
    d_stats = torch.empty((1, n_classes, max_components, 5), dtype=torch.int32, device=d_pred.device)
    d_ncomp = torch.empty((1, n_classes), dtype=torch.int32, device=d_pred.device)
    ctx.class_components(d_pred, 1, h, w, n_classes, d_stats, max_components, d_ncomp)
    ncomp = d_ncomp.cpu().numpy()[0]
    if int(ncomp.max()) > max_components:
        return class_components(pred, n_classes, int(ncomp.max()))
    return [(int(ncomp[c]), d_stats[0, c, :int(ncomp[c])].cpu().numpy()) for c in range(n_classes)]


def find_postprocessor(key: str) -> Callable[[np.ndarray, SingleData], np.ndarray]:
    return POSTPROCESSORS[key.lower().replace('_', '').replace('-', '')]


def postprocess_help():
    return (
        "Postprocessors available:\n"
        "cc_majority:    classify all pixels of each connected component as most frequent class.\n"
        "bounding_boxes: replace each connected component in the prediction with its bounding box.\n"
    )


POSTPROCESSORS = {
    'ccmajority': vote_connected_component_class,
    'ccvote': vote_connected_component_class,
    'voteconnectedcomponents': vote_connected_component_class,
    'votecomponents': vote_connected_component_class,
    'boundingboxes': add_bounding_boxes,
    'bbox': add_bounding_boxes,
}
