#!/usr/bin/env python
"""Latency per continuation chunk: eager `model(x, ssm_state=..., temporal_pos_offset=...)` chain vs
`videomamba_b200.graphed.GraphedStream` (one CUDA graph launch per chunk), VideoMamba-S.
    python tools/graphed_stream_latency.py"""
import sys, time, torch
sys.path.insert(0, ".")
import video_mamba
from videomamba_b200.graphed import GraphedStream
dev, bf = torch.device("cuda"), torch.bfloat16
torch.manual_seed(0)
m = video_mamba.PretrainVideoMamba(img_size=224, patch_size=16, depth=24, embed_dim=384, channels=3,
        ssm_cfg={"use_fast_path": False}, num_frames=4096, pool_type="avg").eval().to(bf).to(dev)
for B, T in ((1, 1), (1, 8), (32, 1)):
    x = torch.rand(B, 3, T, 224, 224, device=dev).to(bf)
    r = GraphedStream(m); r.first(x)
    for _ in range(3): r.step(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(30): r.step(x)
    torch.cuda.synchronize(); g = (time.perf_counter() - t0) / 30 * 1e3
    with torch.no_grad():
        st = m.allocate_state(B, dtype=bf, device=dev)
        out = m(x, ssm_state=st, temporal_pos_offset=0); st = out[2]; off = T
        for _ in range(3): out = m(x, ssm_state=st, temporal_pos_offset=off); st = out[2]; off += T
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(30): out = m(x, ssm_state=st, temporal_pos_offset=off); st = out[2]; off += T
        torch.cuda.synchronize(); e = (time.perf_counter() - t0) / 30 * 1e3
    print(f"B={B} T={T}: eager {e:.2f} ms/chunk, GraphedStream {g:.2f} ms/chunk", flush=True)
