#!/usr/bin/env python
"""Runs single operators of the hot path at the bench shapes (VideoMamba-S 16f, batch 32) so that
one kernel can be captured with ncu or timed alone:   python tools/prof_ops.py scan|conv|norm|gemm [iters]
(PGATE=1: the scan multiplies by a stored gate, as the inference mixer runs it; PACT=<column>: the projection applies
SiLU from that output column on, PN=1536 PACT=768 = in_proj of the inference mixer.)
Prints CUDA-event time per launch and achieved algorithmic GB/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "scan"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, L, D = int(os.environ.get("PB", 32)), int(os.environ.get("PL", 3137)), int(os.environ.get("PD", 384))
Di, N, R = 2 * D, 16, (D + 15) // 16
Xp = ops.xdbl_pitch(R, N)
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device=dev) * scale).to(bf)


def timeit(fn, nbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{which}: {ms * 1e3:.1f} us/launch, {nbytes / ms / 1e6:.0f} GB/s algorithmic "
          f"({nbytes / 1e6:.0f} MB)", flush=True)


if which == "scan":
    xz = rn(B, L, 2 * Di)
    u, z = rn(B, L, Di), xz[..., Di:]
    xdbl = rn(B, L, Xp)
    w_dt = rn(Di, R, scale=R ** -0.5)
    A2 = (-torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(Di, 1)
                     + 0.1 * torch.randn(Di, N, generator=g, device=dev)) * ops.LOG2E).contiguous()
    Dp = torch.ones(Di, device=dev)
    bias = torch.full((Di,), -3.0, device=dev)
    split = os.environ.get("PSPLIT", "1") != "0"      # tool-only switch: time the unsplit walk
    gate = os.environ.get("PGATE", "0") != "0"
    timeit(lambda: ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, allow_split=split,
                                                   z_gate=gate),
           B * L * (3 * Di + Xp) * 2)
elif which == "conv":
    xz = rn(B, L, 2 * Di)
    w, b = rn(Di, 4, scale=0.5), rn(Di)
    timeit(lambda: ops.causal_conv1d_tokens(xz[..., :Di], w, b), B * L * 2 * Di * 2)
elif which == "norm":
    x = rn(B, L, D)
    res = torch.randn(B, L, D, generator=g, device=dev)
    w = torch.ones(D, device=dev, dtype=bf)
    timeit(lambda: ops.add_norm(x, w, None, res, 1e-5, True, True, True), B * L * D * (2 + 4 + 4 + 2))
elif which == "gemm":
    n, k = int(os.environ.get("PN", 4 * D)), int(os.environ.get("PK", D))
    a, w = rn(B * L, k), rn(n, k, scale=k ** -0.5)
    act = os.environ.get("PACT")
    timeit(lambda: ops.linear_raw(a, w, None, silu_from=None if act is None else int(act)),
           (B * L * (k + n) + n * k) * 2)
