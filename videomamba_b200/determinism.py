"""Seeding / determinism switches (same surface as the reference's video_mamba/determinism.py).

The libvmb200 kernels are run-to-run deterministic by construction (no atomics, fixed reduction
order), so these switches only configure PyTorch itself (patch-embed conv, pooling).
"""
from __future__ import annotations

import argparse
import random
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass(frozen=True)
class DeterminismConfig:
    seed: int = 0
    deterministic: bool = False
    warn_only: bool = True
    cudnn_benchmark: bool = True
    allow_tf32: bool = True


def configure_determinism(seed: int, deterministic: bool, warn_only: bool = True,
                          cudnn_benchmark: Optional[bool] = None,
                          allow_tf32: Optional[bool] = None) -> DeterminismConfig:
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    deterministic = bool(deterministic)
    benchmark = (not deterministic) if cudnn_benchmark is None else bool(cudnn_benchmark)
    tf32 = (not deterministic) if allow_tf32 is None else bool(allow_tf32)
    torch.backends.cudnn.benchmark = benchmark
    torch.backends.cudnn.deterministic = deterministic
    torch.use_deterministic_algorithms(deterministic, warn_only=warn_only)
    if hasattr(torch.backends, "cuda") and hasattr(torch.backends.cuda, "matmul"):
        torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    return DeterminismConfig(seed=seed, deterministic=deterministic, warn_only=bool(warn_only),
                             cudnn_benchmark=benchmark, allow_tf32=tf32)


_TRI = {"on": True, "off": False, "auto": None}


def add_determinism_args(parser: argparse.ArgumentParser) -> argparse.ArgumentParser:
    parser.add_argument("--seed", type=int, default=0, help="Random seed.")
    parser.add_argument("--deterministic", action="store_true",
                        help="Enable deterministic PyTorch algorithms and cuDNN mode.")
    parser.add_argument("--deterministic-warn-only", action="store_true",
                        help="Use warn-only mode for deterministic algorithm enforcement.")
    parser.add_argument("--cudnn-benchmark", choices=sorted(_TRI), default="auto",
                        help="cuDNN benchmark mode. auto => inverse of --deterministic.")
    parser.add_argument("--allow-tf32", choices=sorted(_TRI), default="auto",
                        help="TF32 matmul/convolution mode. auto => inverse of --deterministic.")
    return parser


def configure_determinism_from_args(args: argparse.Namespace) -> DeterminismConfig:
    return configure_determinism(seed=int(args.seed), deterministic=bool(args.deterministic),
                                 warn_only=bool(args.deterministic_warn_only),
                                 cudnn_benchmark=_TRI[args.cudnn_benchmark],
                                 allow_tf32=_TRI[args.allow_tf32])
