#!/usr/bin/env python
"""Op-level selective-scan sweep of SURVEY.md section 8(d) / BASELINE.json configs[4]: the fused
dt_proj + scan kernel at d_state 16, d_inner 768, L in {1569, 3137, 6273, 12544, 25089}, batch in
{1, 4, 32}, both directions, with and without an initial state; plus the op-level fp32
selective_scan_fn (un-fused signature) at the shorter lengths.  One JSON line per case:
CUDA-event time per launch and algorithmic GB/s against the measured HBM peak.
    python tools/scan_sweep.py [--iters K] > profiles/rNN_scan_sweep.jsonl"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
dev, bf = "cuda", torch.bfloat16
D, N = 384, 16
Di, R = 2 * D, (D + 15) // 16
Xp = ops.xdbl_pitch(R, N)
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters


A2 = (-torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(Di, 1)
                 + 0.1 * torch.randn(Di, N, generator=g, device=dev)) * ops.LOG2E).contiguous()
Dp = torch.ones(Di, device=dev)
bias = torch.full((Di,), -3.0, device=dev)
w_dt = (torch.randn(Di, R, generator=g, device=dev) * R ** -0.5).to(bf)
for L in (1569, 3137, 6273, 12544, 25089):
    for B in (1, 4, 32):
        if B * L > 32 * 12544:
            continue                       # 25 089 x 32 does not add information (already saturated)
        u = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
        z = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
        xdbl = torch.randn(B, L, Xp, generator=g, device=dev).to(bf)
        h0 = torch.randn(B, Di, N, generator=g, device=dev)
        nbytes = B * L * (3 * Di + Xp) * 2
        for reverse in (False, True):
            for init in (False, True):
                ms = timeit(lambda: ops.selective_scan_fused_tokens(
                    u, z, xdbl, w_dt, A2, R, N, Dp, bias, h0=h0 if init else None, want_last=True,
                    reverse=reverse))
                print(json.dumps({"op": "fused dt_proj+scan (bf16)", "L": L, "B": B, "reverse": reverse,
                                  "initial_state": init, "us_per_launch": round(ms * 1e3, 1),
                                  "tokens_per_s": round(B * L / ms * 1e3),
                                  "algorithmic_GBps": round(nbytes / ms / 1e6),
                                  "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 4)}), flush=True)
        del u, z, xdbl
# op-level drop-in signature in fp32 (the 1e-5 parity path): (B, Di, L) operands, delta given
for L in (1569, 3137):
    for B in (1, 4):
        f32 = torch.float32
        u = torch.randn(B, Di, L, generator=g, device=dev, dtype=f32)
        delta = torch.randn(B, Di, L, generator=g, device=dev, dtype=f32)
        z = torch.randn(B, Di, L, generator=g, device=dev, dtype=f32)
        Bm = torch.randn(B, N, L, generator=g, device=dev, dtype=f32)
        Cm = torch.randn(B, N, L, generator=g, device=dev, dtype=f32)
        A = A2 / ops.LOG2E
        ms = timeit(lambda: ops.selective_scan_fn(u, delta, A, Bm, Cm, Dp, z=z, delta_bias=bias,
                                                  delta_softplus=True, return_last_state=True))
        nbytes = B * L * (4 * Di + 2 * N) * 4
        print(json.dumps({"op": "selective_scan_fn (fp32, un-fused signature)", "L": L, "B": B,
                          "us_per_launch": round(ms * 1e3, 1), "tokens_per_s": round(B * L / ms * 1e3),
                          "algorithmic_GBps": round(nbytes / ms / 1e6),
                          "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 4)}), flush=True)
