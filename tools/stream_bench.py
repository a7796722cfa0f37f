"""Blocking against streaming host-buffer calls (pcs_predict_pages_compact vs ..._submit / pcs_wait_pages), 64 A4 pages per
call, pinned host buffers, CUDA events around K calls.  Development tool: python tools/stream_bench.py [K] [segments]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
from page_segmentation_b200 import synth
import torch
from page_segmentation_b200.runtime import PageBatchEngine
K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
SEG = "segments" in sys.argv[1:]
n, maxc = 64, 4096
eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, seed=0), 3)
base = np.stack([synth.make_page(s) for s in range(8)])
h_pages = torch.from_numpy(np.concatenate([base] * 8)).pin_memory().numpy()
Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 1 / 3)
bw = (Hs * Ws + 31) // 32


def outs():
    o = {"labels": torch.empty((n, Hs, Ws), dtype=torch.uint8).pin_memory().numpy(),
         "binary_bits": torch.empty((n, bw), dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
    if SEG:
        o["stats"] = torch.empty((n, 3, maxc, 5), dtype=torch.int32).pin_memory().numpy()
        o["ncomp"] = torch.empty((n, 3), dtype=torch.int32).pin_memory().numpy()
    return o


o = [outs(), outs()]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def blocking(k):
    for i in range(k):
        if SEG:
            eng.run_host_segments_compact(h_pages, 1 / 3, o[i % 2], max_components=maxc, cc_majority=True)
        else:
            eng.run_host_compact(h_pages, 1 / 3, o[i % 2])


def streaming(k):
    t = []
    for i in range(k):
        if i >= 2:
            eng.wait(t[i - 2])
        t.append(eng.submit_host_compact(h_pages, 1 / 3, o[i % 2], cc_majority=SEG, max_components=maxc if SEG else 0))
    for x in t[-2:]:
        eng.wait(x)


for name, fn in (("blocking", blocking), ("streaming", streaming), ("blocking", blocking), ("streaming", streaming)):
    fn(2)
    torch.cuda.synchronize()
    e0.record()
    fn(K)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"mode": name, "segments": SEG, "calls": K, "ms_per_call": round(ms / K, 3), "pages_per_s": round(n * K / ms * 1e3, 1)}), flush=True)
ref = outs()
eng.run_host_compact(h_pages, 1 / 3, ref) if not SEG else eng.run_host_segments_compact(h_pages, 1 / 3, ref, max_components=maxc, cc_majority=True)
print(json.dumps({"streamed == blocking": bool((ref["labels"] == o[(K - 1) % 2]["labels"]).all())}))
