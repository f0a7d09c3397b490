#!/usr/bin/env python
"""Random small geometries of the bf16 scan backward against the true-fp32 kernels (edge lengths around the
4-token sub-chunk and 16-token tile boundaries, segment splits, all option combinations).
    python tools/scan_bwd_fuzz.py [cases] [seed]"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import autograd as ag  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    dev, bf = "cuda", torch.bfloat16
    worst = 0.0
    for case in range(cases):
        B = rng.choice([1, 1, 2, 3])
        L = rng.choice([1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65, 100, 383, 384, 385, 400, 777])
        Di = rng.choice([16, 32, 48, 80])
        R = rng.choice([12, 24, 36, 0])                     # B / C column offsets: x_dbl layouts and a plain (B | C) tensor
        N = 16
        cols = (R + 2 * N + 15) // 16 * 16 if R else 2 * N
        with_z, with_h0, with_D, with_last = (rng.random() < 0.7 for _ in range(4))
        g = torch.Generator(device=dev).manual_seed(case)
        rn = lambda *s: torch.randn(*s, device=dev, generator=g)
        u, z, dout = rn(B, L, Di).to(bf), rn(B, L, Di).to(bf), rn(B, L, Di).to(bf)
        delta = (0.5 * rn(B, L, Di) - 2).to(bf)
        bc = rn(B, L, cols).to(bf)
        A2 = -(torch.rand(Di, N, device=dev, generator=g) * 8 + 0.1) * 1.4427
        D = rn(Di) if with_D else None
        bias = 0.3 * rn(Di)
        h0 = rn(B, Di, N) if with_h0 else None
        dlast = rn(B, Di, N) if with_last else None
        zz = z if with_z else None
        args = (A2, None, R, R + N, N, D)
        f32 = lambda t: None if t is None else t.float()
        fast = ag._scan_bwd(u, delta, A2, bc, R, R + N, N, D, zz, bias, True, h0, dout, dlast, with_h0)
        ref = ag._scan_bwd(f32(u), f32(delta), A2, f32(bc), R, R + N, N, D, f32(zz), bias, True, h0, f32(dout), dlast,
                           with_h0)
        names = ["du", "ddelta", "dz", "dbc", "dA", "dD", "dbias", "dh0"]
        errs = {n: rel(a, b) for n, a, b in zip(names, fast, ref) if a is not None}
        w = max(errs.values())
        worst = max(worst, w)
        flag = "" if w < 2e-2 else "   <-- FAIL"
        print(f"case {case:3d} B={B} L={L:4d} Di={Di:3d} R={R:2d} z={int(with_z)} h0={int(with_h0)} D={int(with_D)} "
              f"last={int(with_last)}: worst {max(errs, key=errs.get)} {w:.2e}{flag}", flush=True)
    print("worst", f"{worst:.3e}")
    sys.exit(0 if worst < 2e-2 else 1)


if __name__ == "__main__":
    main()
