"""The opt-in fused conv1 + conv2 kernel (conv12_fused.cu, PCSEG_FUSE12=1) against the oracle and against the default pair of
kernels.  The switch is read once per process, so the check runs in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r'''
import sys
import numpy as np
import torch
sys.path.insert(0, %r)
from oracle import network as onet
from oracle import pipeline as opipe
from page_segmentation_b200 import synth
from page_segmentation_b200.lib.architecture import Architecture
from page_segmentation_b200.lib.network import Network

W = synth.make_weights(ARCH, 3, seed=0)
for (n, H, Wd) in ((3, 216, 390), (1, 1128, 400), (2, 96, 1000)):       # strips and CTA ranges that cross pages; one tall page; a wide one
    imgs = []
    for s in range(n):
        page = synth.make_page(40 + s, H * 3, Wd * 3, 18)
        imgs.append(opipe.prepare_images(page, page, 6, 18)[0])
    imgs = np.stack(imgs)
    h, w = imgs.shape[1:]
    for precision, tol, agree in (("fp16", 1e-3, 0.999), ("bf16", 8e-3, 0.997)):
        net = Network("Predict", n_classes=3, model_constructor=Architecture(ARCH), weights=W, precision=precision)
        c = net._context()
        d_img = torch.from_numpy(imgs).cuda()
        d_labels = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
        d_logits = torch.empty((n, h, w, 3), dtype=torch.float32, device="cuda")
        c.forward(d_img, None, n, h, w, d_labels, d_logits, None)
        torch.cuda.synchronize()
        for i in range(n):
            l32 = onet.Forward(ARCH, W, 3).logits(imgs[i])[0]
            l64 = onet.Forward(ARCH, W, 3, dtype=torch.float64).logits(imgs[i])[0]
            lg = d_logits[i].cpu().numpy()
            assert np.isfinite(lg).all(), (n, H, Wd, precision, i)
            err = np.abs(lg - l32).max()
            assert err <= tol, (n, H, Wd, precision, i, err)
            a = (d_labels[i].cpu().numpy() == l64.argmax(-1)).mean()
            assert a >= agree, (n, H, Wd, precision, i, a)
print("FUSED12_OK")
'''


@pytest.mark.parametrize("arch", ["fcn_skip", "fcn"])
def test_fused_conv12_matches_the_oracle(arch):
    code = ("ARCH = %r\n" % arch) + (CHECK % ROOT)
    env = dict(os.environ, PCSEG_FUSE12="1", PCSEG_C12_STATS="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FUSED12_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert "[conv12]" in r.stderr, "the fused kernel did not run"
