"""Stand-in for `ocr4all.colors.ColorMap` (ocr4all-pylib 0.2.6 is not installed).

Surface inferred from the reference's call sites:
`ColorMap(mapping)` network.py:46, `.to_rgb_array(labels)` output.py:45,
`.imread_labels(path)` dataset.py:181, `.color_for_label(name)`
pc_segmentation.py:71, `.filter_label(rgb, name)` pc_segmentation.py:48,56.
JSON schema as written by pagexml.py:114-129: {"(r, g, b)": [label, name]}.
"""
from __future__ import annotations

import ast
import json
from typing import Dict, Tuple

import numpy as np


class ColorMap:
    def __init__(self, mapping: Dict):
        self.mapping: Dict[Tuple[int, int, int], Tuple[int, str]] = {}
        for k, v in (mapping or {}).items():
            if isinstance(k, str):
                k = ast.literal_eval(k)
            label, name = (v[0], v[1]) if isinstance(v, (list, tuple)) else (int(v), str(v))
            self.mapping[tuple(int(c) for c in k)] = (int(label), str(name))

    @classmethod
    def load(cls, path: str) -> "ColorMap":
        with open(path, "r") as f:
            return cls(json.load(f))

    def to_json(self) -> Dict[str, list]:
        return {str(tuple(k)): [v[0], v[1]] for k, v in self.mapping.items()}

    def __len__(self):
        return len(self.mapping)

    def labels(self):
        return sorted(v[0] for v in self.mapping.values())

    def lut(self, n_classes: int = None) -> np.ndarray:
        """(n, 3) uint8 table label -> rgb; labels without a colour map to black."""
        n = max([v[0] for v in self.mapping.values()] + [-1]) + 1
        if n_classes is not None:
            n = max(n, int(n_classes))
        out = np.zeros((n, 3), dtype=np.uint8)
        for rgb, (label, _name) in self.mapping.items():
            out[label] = rgb
        return out

    def to_rgb_array(self, labels: np.ndarray) -> np.ndarray:
        lut = self.lut()
        labels = np.asarray(labels)
        out = np.zeros(labels.shape + (3,), dtype=np.uint8)
        ok = (labels >= 0) & (labels < lut.shape[0])
        out[ok] = lut[labels[ok]]
        return out

    def color_for_label(self, name: str) -> Tuple[int, int, int]:
        for rgb, (_label, n) in self.mapping.items():
            if n == name:
                return rgb
        raise KeyError(name)

    def filter_label(self, rgb_image: np.ndarray, name: str) -> np.ndarray:
        rgb = np.asarray(self.color_for_label(name), dtype=rgb_image.dtype)
        return np.all(rgb_image[..., :3] == rgb, axis=-1)

    def rgb_to_labels(self, rgb_image: np.ndarray) -> np.ndarray:
        out = np.zeros(rgb_image.shape[:2], dtype=np.uint8)
        for rgb, (label, _name) in self.mapping.items():
            out[np.all(rgb_image[..., :3] == np.asarray(rgb, dtype=rgb_image.dtype), axis=-1)] = label
        return out

    def imread_labels(self, path: str) -> np.ndarray:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_COLOR)
        if img is None:
            raise FileNotFoundError(path)
        return self.rgb_to_labels(img[..., ::-1])


DEFAULT_COLOR_MAP = ColorMap({(255, 255, 255): (0, "background"), (255, 0, 0): (1, "text"), (0, 255, 0): (2, "image")})
