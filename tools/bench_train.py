#!/usr/bin/env python
"""Training-step time of the FCN on one A4 page per rank per step (BASELINE configs[4]): forward + loss + backward +
(NCCL) gradient all-reduce + Adam with clipnorm.  Under torchrun every rank trains on its own page (data parallel).

    python tools/bench_train.py [--steps 10]            torchrun --nproc-per-node N tools/bench_train.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from page_segmentation_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--arch", default="fcn_skip")
    ap.add_argument("--lazy-loss", action="store_true", help="do not read the loss back after every step (what Trainer.train does)")
    ap.add_argument("--engine", default=None, help="tensor | fp32 (default: FcnTrainStep's)")
    ap.add_argument("--cpu-steps", type=int, default=1, help="steps of the torch-CPU oracle timed beside it (rank 0, single process)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from oracle import pipeline as opipe
    from page_segmentation_b200.lib.trainer import FcnTrainStep
    page = synth.make_page(rank)
    img, binary = opipe.prepare_images(page, page, 6, 18)                  # 1169 x 827, the scaled page the network sees
    labels = binary.astype(np.uint8)                                        # ink / paper as a two-class target of three
    W = synth.make_weights(args.arch, 3, seed=0)
    eng = FcnTrainStep(args.arch, W, 3, l_rate=1e-4, device=local, engine=args.engine)
    for _ in range(args.warmup):
        eng.step(img, labels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = [eng.step(img, labels, lazy_loss=args.lazy_loss) for _ in range(args.steps)]
    e1.record()
    torch.cuda.synchronize()
    losses = [float(v) for v in losses]
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if rank == 0:
        cpu = None
        if world == 1 and args.cpu_steps > 0:
            from oracle import train as otr
            otr.loss_and_grads(args.arch, W, img[:256, :256], labels[:256, :256], 3)
            t0 = time.perf_counter()
            for _ in range(args.cpu_steps):
                otr.loss_and_grads(args.arch, W, img, labels, 3)
            cpu = {"s_per_step": (time.perf_counter() - t0) / args.cpu_steps, "threads": torch.get_num_threads(),
                   "what": "torch-CPU autograd forward + backward of the same page (oracle), no optimizer"}
        print(json.dumps({"metric": "train_steps_per_sec", "value": world * 1e3 / ms, "unit": "pages/s (one page per rank per step)",
                          "n_gpus": world, "ms_per_step": ms, "steps": args.steps, "dtype": eng.describe()["dtype"], "engine": eng.engine, "arch": args.arch,
                          "page": list(img.shape), "allreduce_floats": int(eng.params.numel()), "loss_first_last": [losses[0], losses[-1]],
                          "gflop_per_step": 3 * 112.0, "tflops": 3 * 112.0 / ms, "cpu_oracle": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
