"""Stable import path of the reference (``import video_mamba``), served by videomamba_b200."""
from videomamba_b200 import (STREAMING_CONTRACT_VERSION, BiMambaRefinerBlock, DeterminismConfig,
                             ForwardReturnSemantics, LayerState, PretrainVideoMamba, StateShape,
                             StreamingState, add_determinism_args, allocate_state,
                             build_videomamba, configure_determinism,
                             configure_determinism_from_args, expected_state_shapes,
                             forward_return_semantics, model_forward_return_semantics,
                             validate_state)

__all__ = [
    "DeterminismConfig", "ForwardReturnSemantics", "LayerState", "BiMambaRefinerBlock",
    "PretrainVideoMamba", "STREAMING_CONTRACT_VERSION", "StateShape", "StreamingState",
    "add_determinism_args", "allocate_state", "build_videomamba", "configure_determinism",
    "configure_determinism_from_args", "expected_state_shapes", "forward_return_semantics",
    "model_forward_return_semantics", "validate_state",
]
