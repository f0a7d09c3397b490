#!/usr/bin/env python
"""Probe: two fused-scan launches (bench shapes) issued back to back on ONE stream vs on TWO streams
(so their one-warp CTAs share the SMs): what co-residency is worth for this kernel.
    python tools/scan_pair_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import ops  # noqa: E402

B, L, D = int(os.environ.get("PB", 32)), 3137, 384
Di, N, R = 2 * D, 16, 24
Xp = ops.xdbl_pitch(R, N)
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device=dev) * scale).to(bf)
sets = []
for _ in range(2):
    sets.append((rn(B, L, Di), rn(B, L, Di), rn(B, L, Xp)))
w_dt = rn(Di, R, scale=R ** -0.5)
A2 = (-torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(Di, 1)
                 + 0.1 * torch.randn(Di, N, generator=g, device=dev)) * ops.LOG2E).contiguous()
Dp, bias = torch.ones(Di, device=dev), torch.full((Di,), -3.0, device=dev)


def scan(i):
    u, z, x = sets[i]
    return ops.selective_scan_fused_tokens(u, z, x, w_dt, A2, R, N, Dp, bias, allow_split=False)


def timed(two_streams, reps=20):
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(3):
        scan(0); scan(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if two_streams:
            for i in range(2):
                s[i].wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s[i]):
                    scan(i)
            for i in range(2):
                torch.cuda.current_stream().wait_stream(s[i])
        else:
            scan(0); scan(1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def single_double():
    u = torch.cat([sets[0][0], sets[1][0]]); z = torch.cat([sets[0][1], sets[1][1]]); x = torch.cat([sets[0][2], sets[1][2]])
    f = lambda: ops.selective_scan_fused_tokens(u, z, x, w_dt, A2, R, N, Dp, bias, allow_split=False)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3


print(f"batch {B}: ONE launch of batch {2 * B}: {single_double():.1f} us")
print(f"batch {B}: two launches on one stream {timed(False):.1f} us, on two streams {timed(True):.1f} us", flush=True)
