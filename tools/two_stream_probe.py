#!/usr/bin/env python
"""Probe: does running two independent forwards on two CUDA streams raise throughput (the scan is
issue-bound, the projections / conv / norm are HBM-bound)?  python tools/two_stream_probe.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

args = bench.parse_args()
dev = torch.device("cuda")
model = bench.build_model(args, torch.bfloat16, dev)
xs = [torch.rand(args.batch, 3, args.frames, args.img, args.img, device=dev).to(torch.bfloat16) for _ in range(2)]


def run(nstreams, steps):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        s = streams[i % nstreams]
        with torch.cuda.stream(s), torch.no_grad():
            model(xs[i % 2])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


for n in (1, 2, 3):
    run(n, 4)
    print(f"{n} stream(s): {run(n, 12):.2f} ms per forward of {args.batch} clips", flush=True)
