"""Diagnostic: one-hot conv2 through the tcgen05 path, prints the mismatch pattern per tile."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from page_segmentation_b200 import synth
from page_segmentation_b200.lib.network import Network
from page_segmentation_b200.lib.dataset import SingleData
from oracle import pipeline as opipe

page = synth.make_page(4, 70 * 3, 300 * 3, 18)
img, _ = opipe.prepare_images(page, page, 6, 18)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for (ty, tx) in [(2, 2), (1, 3), (0, 0)]:
        W = [(np.zeros_like(k), np.zeros_like(b)) for k, b in synth.make_weights("fcn_skip", 3, seed=0)]
        for c in range(20):
            W[0][0][2, 2, 0, c] = (c + 1) / 32.0
            W[1][0][ty, tx, c, c] = 1.0
        net = Network("Predict", n_classes=3, weights=W, precision="bf16")
        net.engine = "umma"; ctx = net._context()
        net.predict_single_data(SingleData(image=img))
        conv1 = ctx.debug_activation("conv1")[0]; conv2 = ctx.debug_activation("conv2")[0]
        H, Wd, _ = conv1.shape
        exp = np.zeros((H, Wd, 30), np.float32)
        ys, xs = np.arange(H)[:, None] + ty - 2, np.arange(Wd)[None, :] + tx - 2
        ok = (ys >= 0) & (ys < H) & (xs >= 0) & (xs < Wd)
        src = conv1[np.clip(ys, 0, H - 1), np.clip(xs, 0, Wd - 1), :]
        exp[..., :20] = np.where(ok[..., None], src, 0.0)
        bad = (conv2 != exp).any(-1)
        print(f"rep {rep} tap {(ty,tx)} bad px {bad.sum()} of {bad.size}")
        if bad.any():
            # per tile (8 rows x 124 px) summary
            for rb in range(H // 8):
                row = []
                for st in range((Wd + 123) // 124):
                    blk = bad[rb * 8:(rb + 1) * 8, st * 124:(st + 1) * 124]
                    row.append(f"{int(blk.sum()):4d}/{blk.size}")
                print("   rb", rb, " ".join(row))
            yy, xx = np.nonzero(bad)
            print("   rows", np.unique(yy)[:40], "cols min/max", xx.min(), xx.max())
            y, x = yy[0], xx[0]
            print("   first bad", y, x, "got", conv2[y, x, :6], "exp", exp[y, x, :6])
            # is the bad output equal to some other shift?
            zero = (conv2[bad] == 0).all()
            print("   bad outputs all zero:", zero)
