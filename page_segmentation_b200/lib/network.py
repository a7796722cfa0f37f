# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/network.py for prediction:
`Network.__init__` model loading (:19-107) and `predict_single_data` (:248-260).
Training (`train_dataset`, :167-242) is out of scope."""
from __future__ import annotations

import logging
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .architecture import Architecture
from .dataset import Dataset, SingleData
from .colors import ColorMap

logger = logging.getLogger(__name__)
_model_tokens = __import__("itertools").count(1)     # id() values are recycled by the allocator; tokens are not

# tensor-core operand type: fp16 and bf16 run at the same tcgen05 rate; fp16's 11-bit significand keeps the class
# map within the 99.9 % agreement bar against the fp32 reference even on random-init weights,
# bf16 stays selectable for models whose activations could leave the fp16 range (stores saturate at 65504).
DEFAULT_PRECISION = os.environ.get("PCSEG_PRECISION", "fp16")


class _ModelHandle:
    """What `self.model` of the reference exposes to its callers: `.name`."""

    def __init__(self, name: str, weights):
        self.name = name
        self.weights = weights


class Network:
    def __init__(self,
                 type: str,
                 n_classes: int = -1,
                 model_constructor: Architecture = Architecture.FCN_SKIP,
                 l_rate: float = 1e-4,
                 has_binary: bool = False,
                 foreground_masks: bool = False,
                 model: str = None,
                 continue_training: bool = False,
                 input_image_dimension: int = 1,
                 optimizer=None,
                 optimizer_norm_clipping: bool = True,
                 optimizer_norm_clip_value: float = 1.0,
                 optimizer_clipping=False,
                 optimizer_clip_value=1,
                 loss_func=None,
                 weights: Optional[Sequence[Tuple[np.ndarray, np.ndarray]]] = None,
                 precision: Optional[str] = None,
                 device: Optional[int] = None,
                 ):
        """Same leading parameters as the reference (network.py:19-35).  Extra
        keyword-only-in-practice parameters: `weights` (Keras-ordered list of
        (kernel, bias) instead of a file), `precision` ('bf16' | 'fp16' tensor-core
        operand type) and `device`."""
        if type.lower() == "train":
            raise NotImplementedError("training is outside the B200 inference hot path")
        self.architecture = model_constructor.value
        self._data: Dataset = Dataset([], ColorMap({}))
        self.type = type
        self.has_binary = has_binary
        self.foreground_masks = foreground_masks
        self.n_classes = n_classes
        self.precision = precision or DEFAULT_PRECISION
        self.device = device
        Architecture(self.architecture).preprocess()   # raises for out-of-scope architectures

        name = 'model'
        if weights is None:
            if not model:
                raise ValueError("Network needs `model` (a Keras .h5 path) or `weights`")
            model = model if '.' in model else model + '.h5'                 # network.py:59
            if not os.path.exists(model):
                raise FileNotFoundError(f"model file {model} does not exist (TF1 .meta migration, network.py:60-68, "
                                        "is not available)")
            from . import h5
            # network.py:75-84 load_model, falling back to load_weights (:106-107): both layouts are read
            loaded = h5.load_keras_model(model)
            weights = loaded.weights
            if loaded.name:
                name = loaded.name
            if loaded.name in ('fcn_skip', 'fcn', 'unet'):
                self.architecture = loaded.name                               # network.py:251
        self.model = _ModelHandle(name, list(weights))
        arch = self.architecture if self.model.name == 'model' else self.model.name
        n_from_weights = int(self.model.weights[-1][0].shape[-1])
        if self.n_classes is None or self.n_classes < 0:
            self.n_classes = n_from_weights
        elif self.n_classes != n_from_weights:
            raise ValueError(f"n_classes={self.n_classes} but the logits layer has {n_from_weights} outputs")
        self._arch = arch
        self._token = next(_model_tokens)
        self.engine = os.environ.get("PCSEG_ENGINE", "umma")      # 'direct' = the CUDA-core numerics twin

    # -- device state ----------------------------------------------------------
    def _context(self):
        from ..runtime import get_context
        ctx = get_context(self.device)
        want = (self._arch, self.n_classes, self.precision, self._token)
        if ctx.loaded_key != want:                  # the context holds one model; whoever loaded last owns it
            ctx.load_model(self._arch, self.n_classes, self.model.weights, self.precision, key=want)
        ctx.set_engine(self.engine)
        return ctx

    def predict_single_data(self, data: SingleData):
        """network.py:248-260 -> (logit f32 HWC, prob f32 HWC, pred int64 HW)."""
        return self._predict(data, want_logits=True)

    def _device_image(self, data: SingleData, ctx):
        from ..lazy import DeviceArray, device_tensor_of, peek
        image = peek(data, "image")
        if not isinstance(image, DeviceArray):
            image = np.ascontiguousarray(image)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise ValueError("data.image must be a 2-D uint8 array (DatasetLoader output)")
        return device_tensor_of(image, ctx.device)

    def _predict(self, data: SingleData, want_logits: bool):
        """`want_logits=False` (Predictor: predictor.py:33 drops the logits at once) skips their 12 bytes per pixel of
        device-to-host traffic and returns None in their place."""
        import torch
        from ..runtime import results_to_host
        ctx = self._context()
        ctx.use_torch_stream()
        d_image = self._device_image(data, ctx)
        h, w = d_image.shape
        dev = d_image.device
        d_labels = torch.empty((h, w), dtype=torch.uint8, device=dev)
        d_logits = torch.empty((h, w, self.n_classes), dtype=torch.float32, device=dev) if want_logits else None
        d_prob = torch.empty((h, w, self.n_classes), dtype=torch.float32, device=dev)
        ctx.forward(d_image, None, 1, h, w, d_labels, d_logits, d_prob)
        # np.argmax yields int64; widened on the device, not by a host pass
        logit, prob, pred = results_to_host(d_logits, d_prob, d_labels.to(torch.int64))
        return logit, prob, pred

    def _probabilities_device(self, data: SingleData):
        """softmax(logits) of one page as an (H, W, n_classes) float32 device tensor: what a lazily evaluated
        `Prediction.probabilities` runs when somebody reads it (the batched predict path does not keep 12 bytes per
        pixel per page around for callers that never look)."""
        import torch
        ctx = self._context()
        ctx.use_torch_stream()
        d_image = self._device_image(data, ctx)
        h, w = d_image.shape
        d_labels = torch.empty((h, w), dtype=torch.uint8, device=d_image.device)
        d_prob = torch.empty((h, w, self.n_classes), dtype=torch.float32, device=d_image.device)
        ctx.forward(d_image, None, 1, h, w, d_labels, None, d_prob)
        return d_prob

    def predict_labels_device(self, d_image, d_labels):
        """Device-resident variant: (n,H,W) uint8 CUDA tensors in/out."""
        ctx = self._context()
        n, h, w = d_image.shape
        ctx.forward(d_image, None, n, h, w, d_labels)
        return d_labels


def tf_backend_allow_growth():
    """network.py:263-268 configures TensorFlow's allocator; nothing to do here."""
    return None
