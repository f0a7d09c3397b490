#!/usr/bin/env python
"""Latency of the fused scan on short sequences at small batch (the streaming regime: one launch per layer, nothing
else to overlap with): 24 dependent launches replayed as one CUDA graph, us per launch, for the unsplit two-warp
kernel, the automatic choice and the sequence split with minimum segments of 128 / 96 / 64 / 32 tokens (tune / 100).
    python tools/scan_split_latency.py > profiles/rNN_scan_split_latency.jsonl"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import ops  # noqa: E402

dev, bf = "cuda", torch.bfloat16
D, N = 384, 16
Di, R = 2 * D, 24
Xp = ops.xdbl_pitch(R, N)
g = torch.Generator(device=dev).manual_seed(0)
A = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(Di, 1)
               + 0.1 * torch.randn(Di, N, generator=g, device=dev))
A2 = (A * ops.LOG2E).contiguous()
w_dt = (torch.randn(Di, R, generator=g, device=dev) * R ** -0.5).to(bf)
Dp = torch.ones(Di, device=dev)
bias = torch.full((Di,), -3.0, device=dev)
LAYERS, REPLAYS = 24, 30


def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(LAYERS):
                fn()
        for _ in range(3):
            gr.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(REPLAYS):
            gr.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) / (REPLAYS * LAYERS) * 1e3


for B in (1, 2, 4):
    for L in (196, 392, 784, 1568, 3137, 6273, 12544):
        u = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
        z = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
        xdbl = torch.randn(B, L, Xp, generator=g, device=dev).to(bf)
        h0 = torch.randn(B, Di, N, generator=g, device=dev)
        rec = {"B": B, "L": L}
        ref = None
        for name, tune in (("unsplit_two_warps", 20), ("automatic", 0), ("min128", 400), ("min96", 300), ("min64", 200), ("min32", 100)):
            run = lambda: ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, h0, want_last=True,
                                                          z_gate=True, tune=tune)
            y, h = run()
            if ref is None:
                ref = (y.float(), h)
            else:
                rec[name + "_err"] = round(max(((y.float() - ref[0]).abs().max() / ref[0].abs().max()).item(),
                                               ((h - ref[1]).abs().max() / ref[1].abs().max()).item()), 6)
            rec[name + "_us"] = round(graph_time(run), 1)
        print(json.dumps(rec), flush=True)
