"""videomamba_b200 -- B200 (sm_100a) implementation of the VideoMamba Mamba-mixer hot path.

Public surface = the reference's ``video_mamba`` package (video_mamba/__init__.py:24-42); the
top-level ``video_mamba`` package of this repo re-exports it for drop-in imports.
"""
from .block import Block, RMSNorm, create_block
from .determinism import (DeterminismConfig, add_determinism_args, configure_determinism,
                          configure_determinism_from_args)
from .mixer import InferenceParamsLike, Mamba
from .model import PretrainVideoMamba, build_videomamba, load_state_dict
from .refiner import BiMambaRefinerBlock
from .streaming import (STREAMING_CONTRACT_VERSION, ForwardReturnSemantics, LayerState,
                        StateShape, StreamingState, allocate_state, expected_state_shapes,
                        forward_return_semantics, model_forward_return_semantics, validate_state)

__all__ = [
    "DeterminismConfig", "ForwardReturnSemantics", "LayerState", "BiMambaRefinerBlock",
    "PretrainVideoMamba", "STREAMING_CONTRACT_VERSION", "StateShape", "StreamingState",
    "add_determinism_args", "allocate_state", "build_videomamba", "configure_determinism",
    "configure_determinism_from_args", "expected_state_shapes", "forward_return_semantics",
    "model_forward_return_semantics", "validate_state",
    "Mamba", "InferenceParamsLike", "Block", "RMSNorm", "create_block", "load_state_dict",
]
