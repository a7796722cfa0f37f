# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/predictor_data.py:12-26."""
from dataclasses import dataclass
from typing import Callable, List, NamedTuple, Optional

import numpy as np

from .colors import ColorMap
from .dataset import SingleData


class Prediction(NamedTuple):
    labels: np.ndarray
    probabilities: np.ndarray
    data: SingleData


@dataclass
class PredictSettings:
    network: str = None
    output: str = None
    high_res_output: bool = False
    color_map: Optional[ColorMap] = None  # Only needed for generating colored images
    n_classes: int = -1
    post_process: Optional[List[Callable[[np.ndarray, SingleData], np.ndarray]]] = None
    gpu_allow_growth: bool = False
