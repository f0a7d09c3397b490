#!/usr/bin/env python
"""vmb_linear_wgrad at the projection shapes of VideoMamba-S (batch 32, 16 frames): us per call and TFLOP/s.
    python tools/wgrad_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import autograd as ag  # noqa: E402

M = 32 * 3137
shapes = {"in_proj": (1536, 384), "out_proj": (384, 768), "x_proj": (64, 768), "dt_proj": (768, 24 + 8)}
for name, (N, K) in shapes.items():
    dy = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        ag.linear_wgrad(dy, x, torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ag.linear_wgrad(dy, x, torch.bfloat16)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"{name:9s} N={N:5d} K={K:4d}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s  "
          f"{(M * (N + K) * 2) / us / 1e3:6.0f} GB/s operand reads")
