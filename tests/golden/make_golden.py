"""Generates tests/golden/*.npz with the CPU oracle.

The reference itself cannot run in this environment (tensorflow, scikit-image,
ocr4all-pylib and h5py are absent and not installable), so these vectors are produced
by the oracle restatement, not by the reference: they pin the oracle against drift and
give the device path fixed, committed inputs/outputs.   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import network as onet  # noqa: E402
from oracle import pipeline as opipe  # noqa: E402
from page_segmentation_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LUT = {0: (255, 255, 255), 1: (255, 0, 0), 2: (0, 255, 0)}


def case(name, arch, page_seed, weight_seed, h, w, lh, grey=False):
    page = (synth.make_grey_page if grey else synth.make_page)(page_seed, h, w, lh)
    img, b, ob = opipe.prepare_images(page, page, 6, lh, keep_orig_bin=True)
    W = synth.make_weights(arch, 3, seed=weight_seed)
    l32, _ = onet.Forward(arch, W, 3).logits(img)
    l64, _ = onet.Forward(arch, W, 3, dtype=torch.float64).logits(img)
    prob, pred = opipe.softmax_argmax(l32)
    voted = opipe.vote_connected_component_class(pred.copy(), b)
    boxes = opipe.add_bounding_boxes(np.where(b > 0, pred, 0))
    color, overlay, inverted, _ = opipe.generate_output_masks(b, voted, LUT)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), page=page, image=img, binary=b, orig_binary=ob,
                        logits32=l32.astype(np.float32), margin64=np.sort(l64, -1)[..., -1] - np.sort(l64, -1)[..., -2],
                        pred64=l64.argmax(-1).astype(np.uint8), pred32=pred.astype(np.uint8), voted=voted.astype(np.uint8),
                        boxes=boxes.astype(np.uint8), color=color, overlay=overlay, inverted=inverted,
                        meta=np.array([page_seed, weight_seed, lh, int(grey)]), arch=np.array(arch))


if __name__ == "__main__":
    case("fcn_skip_small", "fcn_skip", 11, 3, 210, 168, 18)
    case("fcn_small", "fcn", 12, 4, 150, 201, 15)
    case("fcn_skip_grey", "fcn_skip", 13, 5, 180, 150, 18, grey=True)
    print(sorted(os.listdir(HERE)))
