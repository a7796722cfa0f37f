"""Mirror of ocr4all_pixel_classifier/lib/postprocess.py (:9-64) on the device
connected-component kernels."""
from typing import Callable

import numpy as np

from .dataset import SingleData


def _n_classes(pred: np.ndarray) -> int:
    return int(pred.max()) + 1 if pred.size else 1


def vote_connected_component_class(pred: np.ndarray, data: SingleData) -> np.ndarray:
    """postprocess.py:9-26; like the reference, writes into `pred` and returns it."""
    import torch
    from ..runtime import get_context, to_device_u8
    ctx = get_context()
    h, w = pred.shape
    d_pred = to_device_u8(pred, ctx.device)
    d_bin = to_device_u8(np.asarray(data.binary) != 0, ctx.device)
    ctx.cc_majority(d_pred, d_bin, 1, h, w, _n_classes(pred))
    from ..runtime import to_host
    pred[...] = to_host(d_pred.to(torch.int64)) if pred.dtype == np.int64 else to_host(d_pred).astype(pred.dtype)
    return pred


def add_bounding_boxes(pred: np.ndarray, data: SingleData) -> np.ndarray:
    """postprocess.py:29-42.  The reference hands cv2 a bool array, which cv2
    rejects; this implements the evident intent (components of pred == c)."""
    import torch
    from ..runtime import get_context, to_device_u8
    ctx = get_context()
    h, w = pred.shape
    d_pred = to_device_u8(pred, ctx.device)
    d_out = torch.empty((h, w), dtype=torch.uint8, device=d_pred.device)
    ctx.bounding_boxes(d_pred, 1, h, w, _n_classes(pred), d_out)
    return d_out.cpu().numpy().astype(pred.dtype)


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
