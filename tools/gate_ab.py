#!/usr/bin/env python
"""A/B of where SiLU(z) is applied (bench shape 32 x 3137, VideoMamba-S): inside the fused scan (the reference's
position, mamba_simple.py:423-435) or in the in_proj epilogue (`vmb_linear_fwd_act`, scan with `z_gate`).
CUDA-event time per launch of the scan (alone / three launches sharing the SMs) and of in_proj, one JSON line each.
    python tools/gate_ab.py [--iters K] > profiles/rNN_gate_ab.jsonl"""
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
    return round(e0.elapsed_time(e1) / (args.iters * streams) * 1e3, 1)      # us per launch


def main():
    B, L = args.batch, args.tokens
    u = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
    z = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
    xdbl = torch.randn(B, L, Xp, generator=g, device=dev).to(bf)
    w_dt = (torch.randn(Di, R, generator=g, device=dev) * R ** -0.5).to(bf)
    Dp = torch.ones(Di, device=dev)
    dt = torch.exp(torch.rand(Di, generator=g, device=dev) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
    bias = dt + torch.log(-torch.expm1(-dt))
    for geometric in (False, True):
        if geometric:
            A = -torch.arange(1, N + 1, device=dev).float().repeat(Di, 1)
        else:
            A = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(Di, 1)
                           + 0.1 * torch.randn(Di, N, generator=g, device=dev))
        A2 = (A * ops.LOG2E).contiguous()
        for rounds in range(2):                      # alternate so clock drift shows up as spread, not as a winner
            for gate in (False, True):
                fn = lambda: ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias,
                                                             a_geometric=geometric, z_gate=gate)
                print(json.dumps({"op": "fused scan", "evaluator": "geometric" if geometric else "general",
                                  "gate": "stored (z_gate)" if gate else "SiLU in the scan", "round": rounds,
                                  "us_alone": timed(fn), "us_three_launches_sharing": timed(fn, 3)}), flush=True)
    hidden = torch.randn(B * L, D, generator=g, device=dev).to(bf)
    w_in = (torch.randn(2 * Di, D, generator=g, device=dev) * D ** -0.5).to(bf)
    for rounds in range(3):
        for frm in (None, Di):
            fn = lambda: ops.linear_raw(hidden, w_in, None, silu_from=frm)
            print(json.dumps({"op": "in_proj (100384 x 1536 x 384)", "round": rounds,
                              "epilogue": "plain" if frm is None else "SiLU on columns >= 768",
                              "us_alone": timed(fn)}), flush=True)


if __name__ == "__main__":
    main()
