#!/usr/bin/env python
"""Which torch (ATen) ops are left in one training step: torch.profiler table, grouped by input shape.
    python tools/train_profile.py [--batch 32] [--depth 2]"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_mamba  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--depth", type=int, default=2)
a = ap.parse_args()
torch.manual_seed(0)
m = video_mamba.PretrainVideoMamba(img_size=224, patch_size=16, depth=a.depth, embed_dim=384, channels=3,
                                   ssm_cfg={"use_fast_path": False}, num_frames=16).to(torch.bfloat16).cuda().train()
x = torch.rand(a.batch, 3, 16, 224, 224, device="cuda").to(torch.bfloat16)


def step():
    m.zero_grad(set_to_none=True)
    vis, pool = m(x)
    (pool.float().square().mean() + vis.float().mean()).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60,
                                                        max_shapes_column_width=70))
print(prof.key_averages(group_by_stack_n=6).table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=50))
