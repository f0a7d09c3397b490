#!/usr/bin/env python
"""One training step (forward + backward) of a VideoMamba model on the libvmb200 backward kernels:
    python tools/train_step.py [--model small] [--depth 24] [--batch 8] [--frames 16] [--iters 5] [--ckpt]
Prints CUDA-event times of the forward and the backward half and the peak memory.  Under
`ncu --metrics gpu__time_duration.sum` (use --depth 2 --iters 1) the launch list shows where a layer's
backward goes (tools/launch_list.py)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_mamba  # noqa: E402

WIDTH = {"tiny": 192, "small": 384, "middle": 576}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="small", choices=list(WIDTH))
    ap.add_argument("--depth", type=int, default=24)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ckpt", action="store_true", help="activation checkpointing of every mixer")
    a = ap.parse_args()
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    m = video_mamba.PretrainVideoMamba(
        img_size=224, patch_size=16, depth=a.depth, embed_dim=WIDTH[a.model], channels=3,
        ssm_cfg={"use_fast_path": False}, num_frames=a.frames, use_checkpoint=a.ckpt,
        checkpoint_num=a.depth if a.ckpt else 0).to(dt).cuda().train()
    with torch.no_grad():
        for layer in m.layers:      # general A (trained checkpoints are not geometric)
            layer.mixer.A_log.add_(0.1 * torch.randn_like(layer.mixer.A_log))
    x = torch.rand(a.batch, 3, a.frames, 224, 224, device="cuda").to(dt)
    fwd_ms, bwd_ms = [], []
    for it in range(a.warmup + a.iters):
        m.zero_grad(set_to_none=True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        vis, pool = m(x)
        loss = pool.float().square().mean() + vis.float().mean()
        e[1].record()
        loss.backward()
        e[2].record()
        torch.cuda.synchronize()
        if it >= a.warmup:
            fwd_ms.append(e[0].elapsed_time(e[1]))
            bwd_ms.append(e[1].elapsed_time(e[2]))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    f, b = sum(fwd_ms) / len(fwd_ms), sum(bwd_ms) / len(bwd_ms)
    print(json.dumps({
        "what": f"VideoMamba-{a.model} depth {a.depth}, {a.frames}f@224 {a.dtype}, batch {a.batch}: one training step "
                "(forward in training mode + backward), CUDA events",
        "forward_ms": round(f, 2), "backward_ms": round(b, 2), "step_ms": round(f + b, 2),
        "clips_per_s": round(a.batch / (f + b) * 1e3, 1), "checkpointing": a.ckpt,
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))


if __name__ == "__main__":
    main()
