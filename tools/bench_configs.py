#!/usr/bin/env python
"""Throughput of the OTHER BASELINE.json configurations (bench.py measures configs[1], the headline):
    middle     configs[2]: VideoMamba-Middle (embed 576, depth 32) 32 frames @224 bf16, batch 8 per GPU
    streaming  configs[3]: VideoMamba-S, 64-frame chunks with (conv_state, ssm_state) carry and
               temporal_pos_offset, 32 concurrent streams per GPU (256 over 8)
    longclip   configs[4]: VideoMamba-S 128 frames @224 (25 089 tokens), batch 1 / 2 / 4
One JSON line per case: CUDA-event time over `--steps` forwards after warm-up, inputs resident in HBM.
    python tools/bench_configs.py [middle] [streaming] [longclip] [--steps K]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_mamba  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("cases", nargs="*", default=["middle", "streaming", "longclip"])
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
dev, bf = torch.device("cuda"), torch.bfloat16


def build(embed, depth, frames, pool="cls+avg"):
    torch.manual_seed(0)
    m = video_mamba.PretrainVideoMamba(img_size=224, patch_size=16, depth=depth, embed_dim=embed,
                                       channels=3, ssm_cfg={"use_fast_path": False}, num_frames=frames,
                                       pool_type=pool).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():                      # general-A weights, as in bench.py
        for blk in m.layers:
            blk.mixer.A_log.add_(0.1 * torch.randn(blk.mixer.A_log.shape, generator=g))
    return m.to(bf).to(dev)


def timed(fn, steps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def emit(**kw):
    print(json.dumps(kw), flush=True)


with torch.no_grad():
    if "middle" in args.cases:
        model = build(576, 32, 32)
        x = torch.rand(8, 3, 32, 224, 224, device=dev).to(bf)
        ms = timed(lambda: model(x), args.steps)
        emit(case="middle", workload="VideoMamba-Middle 32f@224 bf16 batch 8", ms_per_step=ms,
             clips_per_s=8 / ms * 1e3, tokens_per_step=8 * (1 + 32 * 196))
        del model, x
    if "streaming" in args.cases:
        streams, chunk, total = 32, 64, 256
        model = build(384, 24, total, pool="avg")
        x = torch.rand(streams, 3, chunk, 224, 224, device=dev).to(bf)
        state = model.allocate_state(streams, dtype=bf, device=dev)
        holder = {"state": state, "k": 0}

        def step():
            k = holder["k"] % (total // chunk)
            st = holder["state"] if k else model.allocate_state(streams, dtype=bf, device=dev)
            _, _, holder["state"] = model(x, ssm_state=st, temporal_pos_offset=k * chunk)
            holder["k"] += 1

        ms = timed(step, args.steps)
        emit(case="streaming", workload="VideoMamba-S, 32 streams x 64-frame chunks, state carry",
             ms_per_step=ms, chunks_per_s=streams / ms * 1e3, frames_per_s=streams * chunk / ms * 1e3,
             tokens_per_step=streams * chunk * 196)
        del model, x, state, holder
    if "longclip" in args.cases:
        model = build(384, 24, 128)
        for B in (1, 2, 4):
            x = torch.rand(B, 3, 128, 224, 224, device=dev).to(bf)
            ms = timed(lambda: model(x), args.steps)
            emit(case="longclip", workload=f"VideoMamba-S 128f@224 bf16 batch {B}", ms_per_step=ms,
                 clips_per_s=B / ms * 1e3, tokens_per_step=B * (1 + 128 * 196))
