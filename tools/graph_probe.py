#!/usr/bin/env python
"""Small-chunk streaming latency: eager launches vs one CUDA graph of the same continuation-chunk
forward (VideoMamba-S, state carried, pool_type="avg").  The library only enqueues on the caller's
stream and allocates nothing, so `torch.cuda.graph` captures a whole forward as is.
    python tools/graph_probe.py"""
import sys, time, torch
sys.path.insert(0, ".")
import video_mamba
dev, bf = torch.device("cuda"), torch.bfloat16
torch.manual_seed(0)
m = video_mamba.PretrainVideoMamba(img_size=224, patch_size=16, depth=24, embed_dim=384, channels=3,
        ssm_cfg={"use_fast_path": False}, num_frames=64, pool_type="avg").eval().to(bf).to(dev)
for B, T in ((32, 1), (32, 4), (1, 1), (1, 8)):
    x = torch.rand(B, 3, T, 224, 224, device=dev).to(bf)
    st = m.allocate_state(B, dtype=bf, device=dev)
    with torch.no_grad():
        out = m(x, ssm_state=st, temporal_pos_offset=0)          # first chunk (has CLS)
        st1 = out[2]
        def step():
            return m(x, ssm_state=st1, temporal_pos_offset=T)
        for _ in range(3): step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): step()
        torch.cuda.synchronize()
        eager = (time.perf_counter() - t0) / 20 * 1e3
        # graph capture of the same call
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2): step()
        torch.cuda.current_stream().wait_stream(s)
        try:
            with torch.cuda.graph(g):
                y = step()
            for _ in range(3): g.replay()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20): g.replay()
            torch.cuda.synchronize()
            graphed = (time.perf_counter() - t0) / 20 * 1e3
        except Exception as e:
            graphed = float("nan"); print("capture failed:", repr(e)[:300])
    print(f"B={B} T={T} tokens={B*T*196}: eager {eager:.2f} ms, graphed {graphed:.2f} ms", flush=True)
