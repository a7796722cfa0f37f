# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/predictor.py:10-54."""
import os
from typing import Generator

from .predictor_data import Prediction, PredictSettings
from .dataset import Dataset, SingleData
from .network import Network, tf_backend_allow_growth
from .output import Masks, scale_to_original_shape, generate_output_masks


class Predictor:
    def __init__(self, settings: PredictSettings, network: Network = None):
        self.settings = settings
        self.network = network

        if settings.gpu_allow_growth:
            tf_backend_allow_growth()

        if not network:
            self.network = Network("Predict", n_classes=settings.n_classes,
                                   model=os.path.abspath(self.settings.network))
        if settings.output:
            output_dir = settings.output
            os.makedirs(os.path.join(output_dir, "overlay"), exist_ok=True)
            os.makedirs(os.path.join(output_dir, "color"), exist_ok=True)
            os.makedirs(os.path.join(output_dir, "inverted"), exist_ok=True)

    def predict(self, dataset: Dataset) -> Generator[Prediction, None, None]:
        """predictor.py:27-30.  Same generator contract; underneath, consecutive pages of one size run as one batch with
        look-ahead and the results stay on the device until read (pipeline.predict_stream)."""
        from ..pipeline import predict_stream
        return predict_stream(self, dataset.data)

    def _post(self, data: SingleData, pred):
        if self.settings.high_res_output:
            data, pred = scale_to_original_shape(data, pred)
        if self.settings.post_process:
            for processor in self.settings.post_process:
                pred = processor(pred, data)
        return data, pred

    def _forward(self, data: SingleData):
        # the logits are dropped right away (predictor.py:33), so the device path does not copy them back; a
        # caller-supplied network object without that shortcut is used through the reference's method
        fast = getattr(self.network, "_predict", None)
        return fast(data, want_logits=False) if fast else self.network.predict_single_data(data)

    def predict_single(self, data: SingleData) -> Prediction:
        """predictor.py:32-42."""
        if self.settings.high_res_output or not hasattr(self.network, "_context"):
            logit, prob, pred = self._forward(data)
            data, pred = self._post(data, pred)
            return Prediction(pred, prob, data)
        from ..pipeline import predict_stream
        return next(iter(predict_stream(self, [data])))

    def predict_masks(self, data: SingleData) -> Masks:
        logit, prob, pred = self._forward(data)
        data, pred = self._post(data, pred)
        return generate_output_masks(data, pred, self.settings.color_map)
