"""Add -> Norm -> Mixer block and its norm layer, on the libvmb200 kernels.

Drop-in for ``Block`` / ``create_block`` of the reference (models/videomamba/videomamba.py:87-291)
and for ``mamba_ssm.ops.triton.layer_norm.RMSNorm`` which the reference instantiates at
videomamba.py:281, :469.  The residual add and the norm always run as ONE kernel
(``vmb_add_norm_fwd``); ``fused_add_norm=False`` only changes the rounding points, exactly as in
the reference (:141-150 vs :151-166).
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Dict, Optional, Protocol, Tuple, Union, cast

import torch
import torch.nn as nn
import torch.utils.checkpoint as checkpoint
from torch import Tensor

from . import ops
from .mixer import InferenceParamsLike, Mamba


class RMSNorm(nn.Module):
    """``weight * x / sqrt(mean(x^2) + eps)`` -- parameters: ``weight`` only (``bias`` is None)."""

    def __init__(self, hidden_size: int, eps: float = 1e-5, dropout_p: float = 0.0,
                 device=None, dtype=None):
        super().__init__()
        if dropout_p:
            raise NotImplementedError("RMSNorm dropout is not used by VideoMamba")
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)

    def forward(self, x: Tensor, residual: Optional[Tensor] = None, prenorm: bool = False,
                residual_in_fp32: bool = False):
        return ops.add_norm(x, self.weight, self.bias, residual, self.eps, True, prenorm,
                            residual_in_fp32)


class DropPath(nn.Module):
    """Stochastic depth per sample (identity in eval mode / at rate 0)."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x: Tensor) -> Tensor:
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


class NormLayerProtocol(Protocol):
    weight: Tensor
    bias: Optional[Tensor]
    eps: float

    def __call__(self, __x: Tensor) -> Tensor: ...


class MixerProtocol(Protocol):
    def __call__(self, hidden_states: Tensor,
                 inference_params: Optional[InferenceParamsLike] = None,
                 ssm_state: Optional[Tensor] = None,
                 state: Optional[Tuple[Tensor, Tensor]] = None,
                 return_state: bool = False
                 ) -> Union[Tensor, Tuple[Tensor, Tuple[Tensor, Tensor]]]: ...

    def allocate_inference_cache(self, batch_size: int, max_seqlen: int, dtype=None,
                                 **kwargs) -> Tuple[Tensor, Tensor]: ...

    def allocate_state(self, batch_size: int, dtype=None,
                       device=None) -> Tuple[Tensor, Tensor]: ...


def apply_norm(norm: nn.Module, x: Tensor, residual: Optional[Tensor], prenorm: bool,
               residual_in_fp32: bool):
    """One fused add+norm launch for either norm flavour (LayerNorm keeps its nn.LayerNorm type
    for state_dict / isinstance compatibility, only its parameters are used)."""
    is_rms = isinstance(norm, RMSNorm)
    return ops.add_norm(x, norm.weight, norm.bias, residual, norm.eps, is_rms, prenorm,
                        residual_in_fp32)


class Block(nn.Module):
    def __init__(self, dim: int, mixer_cls: Callable[[int], MixerProtocol],
                 norm_cls: Callable[[int], nn.Module] = nn.LayerNorm,
                 fused_add_norm: bool = False, residual_in_fp32: bool = False,
                 drop_path: float = 0.0):
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.fused_add_norm = fused_add_norm
        self.mixer = cast(MixerProtocol, mixer_cls(dim))
        self.norm = cast(NormLayerProtocol, norm_cls(dim))
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        if fused_add_norm and not isinstance(self.norm, (nn.LayerNorm, RMSNorm)):
            raise AssertionError("Only LayerNorm and RMSNorm are supported for fused_add_norm")

    def _add_norm(self, hidden: Tensor, residual: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        if self.fused_add_norm:
            # videomamba.py:151-166: fp32 sum, residual stream dtype kept (fp32 if requested)
            branch = hidden if residual is None else self.drop_path(hidden)
            return apply_norm(self.norm, branch, residual, True, self.residual_in_fp32)
        # videomamba.py:141-150: the sum is formed in torch's promoted dtype, rounded to the
        # norm-weight dtype, normalised; the returned residual is the un-rounded sum.
        summed = hidden if residual is None else residual + self.drop_path(hidden)
        normed = apply_norm(self.norm, summed.to(dtype=self.norm.weight.dtype), None, False, False)
        if self.residual_in_fp32:
            summed = summed.to(torch.float32)
        return normed, summed

    def forward(self, hidden_states: Tensor, residual: Optional[Tensor] = None,
                inference_params: Optional[InferenceParamsLike] = None,
                use_checkpoint: bool = False, ssm_state: Optional[Tensor] = None,
                state: Optional[Tuple[Tensor, Tensor]] = None, return_state: bool = False):
        """Returns ``(hidden, residual)`` or ``(hidden, residual, new_state)`` when a full
        ``state`` is given together with ``return_state`` (videomamba.py:238-246)."""
        if state is not None and ssm_state is not None:
            raise ValueError("Pass either state or ssm_state, not both.")
        hidden_states, residual = self._add_norm(hidden_states, residual)

        if state is not None:
            call = partial(self.mixer, inference_params=inference_params, state=state,
                           return_state=return_state)
        else:
            call = partial(self.mixer, inference_params=inference_params, ssm_state=ssm_state)
        if use_checkpoint:
            # forward-only kernels: checkpointing has nothing to recompute, but the argument is
            # part of the reference signature (videomamba.py:168-206) and must stay callable.
            result = checkpoint.checkpoint(call, hidden_states, use_reentrant=False)
        else:
            result = call(hidden_states)
        if state is not None and return_state:
            hidden_states, new_state = result
            return hidden_states, residual, new_state
        return result, residual

    def allocate_inference_cache(self, batch_size: int, max_seqlen: int, dtype=None, **kwargs):
        return self.mixer.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kwargs)


def create_block(d_model: int, ssm_cfg: Optional[Dict[str, object]] = None,
                 norm_epsilon: float = 1e-5, drop_path: float = 0.0, rms_norm: bool = True,
                 residual_in_fp32: bool = True, fused_add_norm: bool = True,
                 layer_idx: Optional[int] = None, bimamba: bool = True,
                 device: Optional[torch.device] = None,
                 dtype: Optional[torch.dtype] = None) -> Block:
    fk: Dict[str, object] = {}
    if device is not None:
        fk["device"] = device
    if dtype is not None:
        fk["dtype"] = dtype
    # The mixer itself is always unidirectional (videomamba.py:276-280); bidirectional use is
    # composed outside (BiMambaRefinerBlock).
    mixer_cls = partial(Mamba, layer_idx=layer_idx, bimamba=False, **(ssm_cfg or {}), **fk)
    norm_cls = partial(RMSNorm if rms_norm else nn.LayerNorm, eps=norm_epsilon)
    block = Block(d_model, mixer_cls, norm_cls=norm_cls, drop_path=drop_path,
                  fused_add_norm=fused_add_norm, residual_in_fp32=residual_in_fp32)
    object.__setattr__(block, "layer_idx", layer_idx)
    return block
