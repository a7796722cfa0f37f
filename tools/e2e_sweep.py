#!/usr/bin/env python
"""Sweep of the host-buffer pipeline of pcs_predict_pages_host (chunk schedule, number of staging
buffers) on one GPU, plus the device-resident time per batch size.  Development tool, not a bench:

    python tools/e2e_sweep.py [--trace]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from page_segmentation_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--mode", default="raw", choices=["raw", "compact", "packed"], help="raw masks out / class map + bit-packed binary out")
    ap.add_argument("--scheds", default="8,8,8,2;2,8,2,2;2,8,2,3;2,8,2,4;2,16,2,3;4,16,4,3;1,8,1,3;2,12,2,3;2,10,2,4;2,16,2,4")
    args = ap.parse_args()
    import torch
    from page_segmentation_b200.runtime import PageBatchEngine
    lut = np.array([[255, 255, 255], [255, 0, 0], [0, 255, 0]], dtype=np.uint8)
    eng = PageBatchEngine("fcn_skip", synth.make_weights("fcn_skip", 3, seed=0), 3, precision=args.precision, lut=lut)
    n = args.pages
    base = np.stack([synth.make_page(s) for s in range(8)])
    h_pages = torch.empty((n, synth.A4_H, synth.A4_W), dtype=torch.uint8).pin_memory()
    for i in range(n):
        h_pages[i] = torch.from_numpy(base[i % 8])
    Hs, Ws = synth.scaled_shape(synth.A4_H, synth.A4_W, 1 / 3)
    h_out = {k: torch.empty((n, Hs, Ws) + ((3,) if k != "labels" else ()), dtype=torch.uint8).pin_memory().numpy()
             for k in ("labels", "color", "overlay", "inverted")}
    hp = h_pages.numpy()
    if args.mode != "raw":
        h_out = {"labels": h_out["labels"],
                 "binary_bits": torch.empty((n, (Hs * Ws + 31) // 32), dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
    run_host = (lambda: eng.run_host(hp, 1 / 3, h_out)) if args.mode == "raw" else (lambda: eng.run_host_compact(hp, 1 / 3, h_out))
    if args.mode == "packed":
        from page_segmentation_b200.runtime import pack_pages
        pbits, l0, l1 = pack_pages(base)
        h_bits = torch.empty((n, pbits.shape[1]), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        for i in range(n):
            h_bits[i] = pbits[i % 8]
        run_host = lambda: eng.run_host_packed(h_bits, l0, l1, synth.A4_H, synth.A4_W, 1 / 3, h_out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # device-resident time per batch size
    d_pages = h_pages.cuda()
    for m in (2, 4, 8, 16, 32, 64):
        if m > n:
            break
        for _ in range(3):
            eng.run_device(d_pages[:m], 1 / 3)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            eng.run_device(d_pages[:m], 1 / 3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"device_resident_pages": m, "ms": ms, "ms_per_page": ms / m}), flush=True)

    ref = None
    for sched in args.scheds.split(";"):
        os.environ.pop("PCSEG_HOST_CHUNKS", None)
        if sched.startswith("L") and ":" in sched:          # "La,b,c:nbuf"
            sched, nb = sched.split(":")
            os.environ["PCSEG_HOST_SCHED"] = f"8,32,8,{nb}"
        os.environ["PCSEG_HOST_CHUNKS" if sched.startswith("L") else "PCSEG_HOST_SCHED"] = sched.lstrip("L")
        os.environ.pop("PCSEG_TRACE_HOST", None)
        for _ in range(2):
            run_host()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            run_host()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        ok = True
        if ref is None:
            ref = {k: v.copy() for k, v in h_out.items()}
        else:
            ok = all(np.array_equal(ref[k], h_out[k]) for k in ref)
        print(json.dumps({"sched": sched, "ms_per_step": ms, "pages_per_s": n / ms * 1e3, "same_output": ok}), flush=True)
        if args.trace:
            os.environ["PCSEG_TRACE_HOST"] = "1"
            sys.stderr.write(f"--- trace sched {sched}\n")
            sys.stderr.flush()
            run_host()
            os.environ.pop("PCSEG_TRACE_HOST", None)


if __name__ == "__main__":
    main()
