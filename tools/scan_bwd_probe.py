#!/usr/bin/env python
"""The selective-scan backward alone at the bench shape (for ncu captures / timing):
    python tools/scan_bwd_probe.py [B] [L] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import autograd as ag  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3137
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
Di, N, R = 768, 16, 24
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
u = torch.randn(B, L, Di, device=dev).to(bf)
z = torch.randn(B, L, Di, device=dev).to(bf)
delta = (0.5 * torch.randn(B, L, Di, device=dev) - 3).to(bf)
bc = torch.randn(B, L, 64, device=dev).to(bf)
A2 = -(torch.rand(Di, N, device=dev) * 16 + 0.5) * 1.4427
D = torch.ones(Di, device=dev)
bias = torch.zeros(Di, device=dev)
dout = torch.randn(B, L, Di, device=dev).to(bf)
for _ in range(2):
    ag._scan_bwd(u, delta, A2, bc, R, R + N, N, D, z, bias, True, None, dout, None, False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ag._scan_bwd(u, delta, A2, bc, R, R + N, N, D, z, bias, True, None, dout, None, False)
e1.record()
torch.cuda.synchronize()
print(f"scan backward (ckpt + reverse + reduce) B={B} L={L} Di={Di}: {e0.elapsed_time(e1) / iters:.2f} ms per call")
