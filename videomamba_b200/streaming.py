"""Streaming state contract (version 1.0.0) -- shapes, allocation, validation.

Drop-in for the reference's models/videomamba/streaming.py: same public names, same return
strings (:25-34), same exception types and messages (:54-133).  Per layer the state is
``(conv_state (B, d_inner, d_conv), ssm_state (B, d_inner, d_state))``; containers may be a
list, a tuple, or a dict keyed by layer index.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Protocol, Sequence, Tuple, Union

import torch
from torch import Tensor

STREAMING_CONTRACT_VERSION = "1.0.0"

LayerState = Tuple[Tensor, Tensor]
StreamingState = Union[List[LayerState], Tuple[LayerState, ...], Dict[int, LayerState]]


@dataclass(frozen=True)
class StateShape:
    conv_state: Tuple[int, int, int]
    ssm_state: Tuple[int, int, int]


@dataclass(frozen=True)
class ForwardReturnSemantics:
    without_state: str
    with_state: str


class _LayerLike(Protocol):
    mixer: object


class _ModelLike(Protocol):
    layers: Sequence[_LayerLike]
    add_pool_norm: bool


def forward_return_semantics(add_pool_norm: bool) -> ForwardReturnSemantics:
    """What ``model.forward`` returns, with and without a streaming state argument."""
    if bool(add_pool_norm):
        return ForwardReturnSemantics("(x_vis, x_pool)", "(x_vis, x_pool, next_state)")
    return ForwardReturnSemantics("x_vis", "(x_vis, next_state)")


def model_forward_return_semantics(model: _ModelLike) -> ForwardReturnSemantics:
    return forward_return_semantics(bool(getattr(model, "add_pool_norm", True)))


def _mixer_dims(idx: int, layer) -> Tuple[int, int, int]:
    mixer = getattr(layer, "mixer", None)
    if mixer is None:
        raise TypeError(f"Layer {idx} does not expose a mixer attribute.")
    try:
        return int(getattr(mixer, "d_inner")), int(getattr(mixer, "d_conv")), \
            int(getattr(mixer, "d_state"))
    except (AttributeError, TypeError, ValueError) as exc:
        raise TypeError(
            f"Layer {idx} mixer does not expose integer d_inner/d_conv/d_state.") from exc


def expected_state_shapes(model: _ModelLike, batch_size: int) -> Dict[int, StateShape]:
    if batch_size <= 0:
        raise ValueError("batch_size must be a positive integer.")
    out: Dict[int, StateShape] = {}
    for idx, layer in enumerate(model.layers):
        d_inner, d_conv, d_state = _mixer_dims(idx, layer)
        out[idx] = StateShape((batch_size, d_inner, d_conv), (batch_size, d_inner, d_state))
    return out


def allocate_state(model: object, batch_size: int, dtype=None, device=None,
                   as_dict: bool = False) -> StreamingState:
    for name in ("allocate_state", "init_state"):
        fn = getattr(model, name, None)
        if callable(fn):
            return fn(batch_size, dtype=dtype, device=device, as_dict=as_dict)
    raise TypeError("Model does not expose allocate_state(...) or init_state(...).")


def validate_state(model: _ModelLike, state: StreamingState, batch_size: int) -> None:
    shapes = expected_state_shapes(model, batch_size)
    depth = len(shapes)
    if isinstance(state, dict):
        want, got = set(range(depth)), set(state.keys())
        if got != want:
            raise ValueError(
                f"State dict keys mismatch: expected {sorted(want)}, got {sorted(got)}.")
        per_layer = [state[i] for i in range(depth)]
    elif isinstance(state, (list, tuple)):
        if len(state) != depth:
            raise ValueError(f"State length mismatch: expected {depth}, got {len(state)}.")
        per_layer = list(state)
    else:
        raise TypeError("State must be a list, tuple, or dict indexed by layer id.")

    for idx, item in enumerate(per_layer):
        if not isinstance(item, (list, tuple)) or len(item) != 2:
            raise TypeError("Each layer state must be a 2-tuple: (conv_state, ssm_state).")
        conv, ssm = item
        if not (torch.is_tensor(conv) and torch.is_tensor(ssm)):
            raise TypeError("conv_state and ssm_state must both be tensors.")
        want = shapes[idx]
        if tuple(conv.shape) != want.conv_state:
            raise ValueError(f"Layer {idx} conv_state shape mismatch: expected "
                             f"{want.conv_state}, got {tuple(conv.shape)}.")
        if tuple(ssm.shape) != want.ssm_state:
            raise ValueError(f"Layer {idx} ssm_state shape mismatch: expected "
                             f"{want.ssm_state}, got {tuple(ssm.shape)}.")
