#!/usr/bin/env python
"""A/B of the fused scan's decay-factor evaluators and layouts (needs a measurement build of the
library: VMB_NVCC_EXTRA=-DVMB_SCAN_LAB python -m videomamba_b200.build --force).

For every evaluator (MUFU only; 1-4 of a lane's four state pairs on the FMA-pipe polynomial; geometric A)
and both layouts (one warp / two warps per unit): CUDA-event time per launch at the bench shape
(32 x 3137 x 768) alone and with two launches sharing the GPU on two streams, and the error against
float64 on identical bf16 inputs for step sizes spread over [1e-3, 1e-1] (trained-checkpoint-like slow
decays, where an inexact exponential shows first).  One JSON line per case.
    python tools/scan_lab.py [--iters K] > profiles/rNN_scan_evaluators.jsonl"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--tokens", type=int, default=3137)
args = ap.parse_args()
dev, bf = "cuda", torch.bfloat16
D, N = 384, 16
Di, R = 2 * D, 24
Xp = ops.xdbl_pitch(R, N)
g = torch.Generator(device=dev).manual_seed(0)
NAMES = {1: "mufu", 2: "poly1", 3: "poly2", 4: "poly3", 5: "poly4", 9: "geometric"}


def weights(geometric):
    if geometric:
        A = -torch.arange(1, N + 1, device=dev).float().repeat(Di, 1)
    else:
        A = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(Di, 1)
                       + 0.1 * torch.randn(Di, N, generator=g, device=dev))
    # softplus(bias) log-uniform in [1e-3, 1e-1] (the reference's dt init, mamba_simple.py:251-263)
    dt = torch.exp(torch.rand(Di, generator=g, device=dev) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
    bias = dt + torch.log(-torch.expm1(-dt))
    w_dt = (torch.randn(Di, R, generator=g, device=dev) * R ** -0.5).to(bf)
    return A, (A * ops.LOG2E).contiguous(), torch.ones(Di, device=dev), bias, w_dt


def inputs(B, L):
    u = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
    z = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
    xdbl = torch.randn(B, L, Xp, generator=g, device=dev).to(bf)
    return u, z, xdbl


def ref64(u, z, xdbl, w_dt, A, Dp, bias):
    u, z, xd = u.double(), z.double(), xdbl.double()
    dt = torch.nn.functional.softplus(xd[..., :R] @ w_dt.double().t() + bias.double())
    Bm, Cm = xd[..., R:R + N], xd[..., R + N:R + 2 * N]
    h = torch.zeros(u.shape[0], Di, N, dtype=torch.float64, device=dev)
    ys = torch.empty_like(u)
    Ad = A.double()
    for t in range(u.shape[1]):
        h = torch.exp(dt[:, t, :, None] * Ad) * h + (dt[:, t] * u[:, t])[:, :, None] * Bm[:, t, None, :]
        ys[:, t] = (h * Cm[:, t, None, :]).sum(-1)
    return (ys + u * Dp.double()) * (z * torch.sigmoid(z)), h


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def timed(fn, streams=1):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    def go(n):
        for i in range(n):
            with torch.cuda.stream(ss[i % streams]):
                fn()
    go(3 * streams)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in ss:
        s.wait_stream(torch.cuda.current_stream())
    go(args.iters * streams)
    for s in ss:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (args.iters * streams) * 1e3      # us per launch


def main():
    B, L = args.batch, args.tokens
    big = inputs(B, L)
    small = inputs(4, 2000)
    for geometric in (False, True):
        A, A2, Dp, bias, w_dt = weights(geometric)
        want, want_h = ref64(*small, w_dt, A, Dp, bias)
        for ev in ((9,) if geometric else (1, 2, 3, 4, 5)):
            for layout in (1, 2):
                tune = 10 * layout + ev
                run = lambda x: ops.selective_scan_fused_tokens(
                    *x, w_dt, A2, R, N, Dp, bias, want_last=True, a_geometric=geometric, tune=tune, allow_split=False)
                try:
                    y, h = run(small)
                except RuntimeError as e:
                    print(json.dumps({"evaluator": NAMES[ev], "layout": layout, "skipped": str(e)[:120]}), flush=True)
                    continue
                rec = {"evaluator": NAMES[ev], "layout": "one warp" if layout == 1 else "two warps",
                       "shape": [B, L, Di], "err_y_vs_f64": rel(y, want), "err_h_vs_f64": rel(h, want_h),
                       "us_alone": round(timed(lambda: run(big)), 1)}
                if layout == 1:
                    rec["us_two_launches_sharing"] = round(timed(lambda: run(big), streams=2), 1)
                    rec["us_three_launches_sharing"] = round(timed(lambda: run(big), streams=3), 1)
                print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
