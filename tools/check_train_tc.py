"""Per-variable comparison of the tensor-core training step (csrc/train_tc.cu) with the fp32 CUDA-core engine and the
torch-autograd CPU oracle: loss (forward path), bias gradients (input-gradient chain), kernel gradients (wgrad kernel).
python tools/check_train_tc.py [--arch fcn_skip] [--size 64x96] [--no-oracle]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="fcn_skip")
    ap.add_argument("--size", default="64x96")
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    h, w = (int(v) for v in a.size.split("x"))
    from page_segmentation_b200 import synth
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    W = synth.make_weights(a.arch, 3, seed=5)
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    img[: h // 2, : w // 3] = 255
    lab = rng.integers(0, 3, (h, w)).astype(np.uint8)
    ref = FcnTrainStep(a.arch, W, 3, engine="fp32")
    loss_ref = ref.forward_backward(img, lab)
    g_ref = ref.gradients()
    out = {"arch": a.arch, "size": [h, w], "loss_fp32": loss_ref}
    if not a.no_oracle:
        from oracle import train as otr
        loss_o, g_o, _ = otr.loss_and_grads(a.arch, W, img, lab, 3)
        out["loss_oracle"] = loss_o
    for swap in ("0", "1"):
        os.environ["PCSEG_WGRAD_SWAP"] = swap
        eng = FcnTrainStep(a.arch, W, 3, engine="tensor")
        loss = eng.forward_backward(img, lab)
        rows = {}
        for (name, *_r), (gk, gb), (ek, eb) in zip(eng.table, eng.gradients(), g_ref):
            rel = lambda g, e: float(np.linalg.norm((g - e).ravel()) / max(np.linalg.norm(e.ravel()), 1e-30))      # noqa: E731
            rows[name] = {"kernel": round(rel(gk, ek), 5), "bias": round(rel(gb, eb), 5), "finite": bool(np.isfinite(gk).all())}
        out[f"tensor_swap{swap}"] = {"loss": loss, "rel_err_vs_fp32": rows}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
