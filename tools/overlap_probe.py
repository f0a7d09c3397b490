#!/usr/bin/env python
"""How well do the kernels of the hot path overlap with a running scan?  Two streams: stream 1 loops the fused
scan, stream 2 loops one other kernel (conv / add_norm / in_proj / out_proj / x_proj / a second scan), bench
shapes (32 x 3137, VideoMamba-S).  Prints the time of each loop alone and of both together: `together` close to
max(alone) = the second kernel hides under the scan, close to the sum = they serialise.
    python tools/overlap_probe.py [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200 import ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
TUNE = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # scan layout: 0 auto, 10 one warp, 20 two warps
B, L, D = 32, 3137, 384
Di, N, R = 2 * D, 16, 24
Xp = ops.xdbl_pitch(R, N)
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device=dev) * scale).to(bf)

xz = rn(B, L, 2 * Di)
u, z, xdbl = rn(B, L, Di), xz[..., Di:], rn(B, L, Xp)
w_dt = rn(Di, R, scale=R ** -0.5)
A2 = (-torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(Di, 1)
                 + 0.1 * torch.randn(Di, N, generator=g, device=dev)) * ops.LOG2E).contiguous()
Dp, bias = torch.ones(Di, device=dev), torch.full((Di,), -3.0, device=dev)
hid, res = rn(B, L, D), torch.randn(B, L, D, generator=g, device=dev)
nw = torch.ones(D, device=dev, dtype=bf)
w_in, w_out, w_x = rn(2 * Di, D, scale=D ** -0.5), rn(D, Di, scale=Di ** -0.5), rn(Xp, Di, scale=Di ** -0.5)
cw, cb = rn(Di, 4, scale=0.5), rn(Di)
y = rn(B, L, Di)

KERNELS = {
    "scan": lambda: ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, tune=TUNE),
    "conv": lambda: ops.causal_conv1d_tokens(xz[..., :Di], cw, cb),
    "add_norm": lambda: ops.add_norm(hid, nw, None, res, 1e-5, True, True, True),
    "in_proj": lambda: ops.linear(hid, w_in),
    "out_proj": lambda: ops.linear(y, w_out),
    "x_proj": lambda: ops.linear(u, w_x),
}


def run(fns):
    streams = [torch.cuda.Stream() for _ in fns]
    main = torch.cuda.current_stream()

    def go(n):
        for s in streams:
            s.wait_stream(main)
        for _ in range(n):
            for s, f in zip(streams, fns):
                with torch.cuda.stream(s):
                    f()
        for s in streams:
            main.wait_stream(s)

    go(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    go(iters)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


with torch.no_grad():
    alone = {k: run([f]) for k, f in KERNELS.items()}
    print("alone (us per launch): " + ", ".join(f"{k} {v:.0f}" for k, v in alone.items()), flush=True)
    for k, f in KERNELS.items():
        t = run([KERNELS["scan"], f])
        print(f"scan + {k:9s}: together {t:6.0f} us per pair   sum {alone['scan'] + alone[k]:6.0f}   "
              f"max {max(alone['scan'], alone[k]):6.0f}   hidden {alone['scan'] + alone[k] - t:6.0f} us", flush=True)
    # the non-scan kernels of one layer, in order, beside a scan (what a second lane does)
    chain = lambda: [KERNELS[k]() for k in ("add_norm", "in_proj", "conv", "x_proj", "out_proj")]
    t_chain = run([chain])
    t_both = run([KERNELS["scan"], chain])
    print(f"scan + [add_norm, in_proj, conv, x_proj, out_proj]: together {t_both:.0f}   chain alone {t_chain:.0f}   "
          f"sum {alone['scan'] + t_chain:.0f}   hidden {alone['scan'] + t_chain - t_both:.0f} us", flush=True)
    t3 = run([KERNELS["scan"], KERNELS["scan"], chain])
    print(f"2 x scan + chain: together {t3:.0f}   sum {2 * alone['scan'] + t_chain:.0f}", flush=True)
