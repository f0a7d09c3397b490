#!/usr/bin/env python
"""The bf16 selective-scan backward (csrc/scan_bwd_fast.cu) against the true-fp32 kernels of csrc/scan_bwd.cu
on the same (bf16-representable) inputs, per output:  python tools/scan_bwd_check.py [B L Di]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import autograd as ag  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def run(B, L, Di, with_z=True, with_h0=True, seed=0):
    N, R = 16, 24
    dev, bf = "cuda", torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    u, z, dout = rn(B, L, Di).to(bf), rn(B, L, Di).to(bf), rn(B, L, Di).to(bf)
    delta = (0.5 * rn(B, L, Di) - 2).to(bf)
    bc = rn(B, L, 64).to(bf)
    A2 = -(torch.rand(Di, N, device=dev, generator=g) * 8 + 0.1) * 1.4427
    D, bias = rn(Di), 0.3 * rn(Di)
    h0 = rn(B, Di, N) if with_h0 else None
    dlast = rn(B, Di, N) if with_h0 else None
    zz = z if with_z else None
    fast = ag._scan_bwd(u, delta, A2, bc, R, R + N, N, D, zz, bias, True, h0, dout, dlast, with_h0)
    f32 = lambda t: None if t is None else t.float()
    ref = ag._scan_bwd(f32(u), f32(delta), A2, f32(bc), R, R + N, N, D, f32(zz), bias, True, h0, f32(dout), dlast,
                       with_h0)
    names = ["du", "ddelta", "dz", "dbc", "dA", "dD", "dbias", "dh0"]
    out = {}
    for n, a, b in zip(names, fast, ref):
        if a is not None:
            out[n] = rel(a, b)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 3:
        shapes = [tuple(int(v) for v in sys.argv[1:4])]
    else:
        shapes = [(2, 37, 48), (1, 3, 16), (2, 64, 32), (3, 131, 64), (2, 1000, 128)]
    worst = 0.0
    for sh in shapes:
        for wz, wh in ((True, True), (False, False)):
            r = run(*sh, with_z=wz, with_h0=wh)
            worst = max(worst, max(r.values()))
            print(sh, "z" if wz else "-", "h0" if wh else "-", " ".join(f"{k}={v:.2e}" for k, v in r.items()), flush=True)
    print("worst", f"{worst:.3e}")
    sys.exit(0 if worst < 3e-2 else 1)
