#!/usr/bin/env python
"""A/B of the fused conv + x_proj kernel against the separate kernels: run once per setting
(VMB_CONV_XPROJ=1 / 0, read at library load) and compare the saved mixer outputs bit for bit.
    VMB_CONV_XPROJ=1 python tools/fused_conv_check.py /tmp/a.pt; VMB_CONV_XPROJ=0 python tools/fused_conv_check.py /tmp/b.pt /tmp/a.pt"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomamba_b200.mixer import Mamba
torch.manual_seed(0)
dev, bf = "cuda", torch.bfloat16
outs = {}
for (B, L, D) in ((2, 777, 384), (3, 1569, 384), (1, 700, 384), (32, 3137, 384)):
    m = Mamba(d_model=D, d_state=16, d_conv=4, expand=2, use_fast_path=False).eval()
    with torch.no_grad():
        m.A_log.add_(0.1 * torch.randn_like(m.A_log))
    m = m.to(bf).to(dev)
    x = torch.randn(B, L, D, generator=torch.Generator().manual_seed(L)).to(bf).to(dev)
    with torch.no_grad():
        y = m(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            y = m(x)
        torch.cuda.synchronize()
    print(f"B={B} L={L}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per mixer forward", flush=True)
    outs[(B, L, D)] = y.float().cpu()
torch.save(outs, sys.argv[1])
if len(sys.argv) > 2:
    ref = torch.load(sys.argv[2])
    for k in outs:
        same = torch.equal(outs[k], ref[k])
        print(k, "bit-identical" if same else f"DIFFERENT max abs {float((outs[k] - ref[k]).abs().max()):.3e}")
