"""``BiMambaRefinerBlock``: forward block + time-reversed backward block + sigmoid fusion gate.

Drop-in for the reference's models/refiner_backbone.py:13-135.  The reference materialises the
reversed sequence with ``torch.flip`` before and after the backward block; here the backward
block's conv and scan WALK the tokens back to front (``reverse`` flag of ``vmb_mixer_fwd``), so a
3-D input needs no flip copies at all.  A 4-D ``(B, T, N, C)`` input flips the frame axis only
(intra-frame order is kept, refiner_backbone.py:61-68): the same kernels walk the frames back to
front and the tokens of a frame front to back (``frame_len`` of ``vmb_mixer_fwd``), again without
a copy.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .block import Block, apply_norm, create_block

LayerState = Tuple[Tensor, Tensor]
PackedShape = Optional[Tuple[int, int, int]]


class BiMambaRefinerBlock(nn.Module):
    """Bidirectional wrapper around two independent unidirectional VideoMamba blocks."""

    def __init__(self, dim: int, ssm_cfg: Optional[Dict[str, object]] = None, **block_kwargs):
        super().__init__()
        layer_idx = block_kwargs.pop("layer_idx", None)
        self.block_fwd = create_block(d_model=dim, ssm_cfg=ssm_cfg, layer_idx=layer_idx,
                                      bimamba=False, **block_kwargs)
        bwd_idx = None if layer_idx is None else int(layer_idx) + 1_000_000
        self.block_bwd = create_block(d_model=dim, ssm_cfg=ssm_cfg, layer_idx=bwd_idx,
                                      bimamba=False, **block_kwargs)
        self.fusion_gate = nn.Sequential(nn.Linear(dim * 2, dim), nn.Sigmoid())
        self.out_proj = nn.Linear(dim, dim)

    @staticmethod
    def _pack_tokens(x: Tensor) -> Tuple[Tensor, PackedShape]:
        if x.ndim == 3:
            return x, None
        if x.ndim == 4:
            b, t, n, c = x.shape
            return x.reshape(b, t * n, c), (b, t, n)
        raise ValueError("Expected x to be [B, L, C] or [B, T, N, C].")

    @staticmethod
    def _unpack_tokens(x: Tensor, packed_shape: PackedShape) -> Tensor:
        if packed_shape is None:
            return x
        b, t, n = packed_shape
        return x.reshape(b, t, n, x.shape[-1])

    @staticmethod
    def _flip_time(x: Tensor, packed_shape: PackedShape) -> Tensor:
        if packed_shape is None:
            return torch.flip(x, dims=[1])
        b, t, n = packed_shape
        return torch.flip(x.reshape(b, t, n, x.shape[-1]), dims=[1]).reshape(b, t * n, x.shape[-1])

    @staticmethod
    def _ensure_state(block: Block, state: Optional[LayerState], batch_size: int,
                      device: torch.device) -> LayerState:
        if state is not None:
            return state
        return block.mixer.allocate_state(batch_size=batch_size, device=device)

    def allocate_state(self, batch_size: int, dtype=None, device=None):
        return (self.block_fwd.mixer.allocate_state(batch_size=batch_size, dtype=dtype, device=device),
                self.block_bwd.mixer.allocate_state(batch_size=batch_size, dtype=dtype, device=device))

    def _backward_block(self, x_seq: Tensor, state: LayerState, packed: PackedShape) -> Tensor:
        """block_bwd on the time-reversed sequence, result in ORIGINAL token order."""
        blk = self.block_bwd
        mixer = blk.mixer
        grad = torch.is_grad_enabled() and (x_seq.requires_grad or any(
            t.requires_grad for t in state) or any(q.requires_grad for q in blk.parameters()))
        if hasattr(mixer, "_kernel_weights") and not grad:
            # norm is per token, so only the mixer needs the reversed walk (3-D input: token axis;
            # 4-D input: frame axis, frames of packed[2] tokens)
            normed, _ = blk._add_norm(x_seq, None)
            out, _, _ = ops.mixer_fwd(mixer._kernel_weights(), normed, state[0], state[1],
                                      want_conv_state=False, want_ssm_state=False, reverse=True,
                                      frame_len=0 if packed is None else packed[2])
            return out
        # training (the reversed walks have no backward kernel): flip copies, as the reference does
        # (refiner_backbone.py:61-68, :112-121)
        out_rev, _, _ = blk(self._flip_time(x_seq, packed), state=state, return_state=True)
        return self._flip_time(out_rev, packed)

    def forward(self, x: Tensor, state_fwd: Optional[LayerState] = None,
                state_bwd_init: Optional[LayerState] = None,
                use_checkpoint: bool = False) -> Tuple[Tensor, LayerState]:
        """Returns ``(out, new_state_fwd)``; only the forward state is carried
        (refiner_backbone.py:135)."""
        x_seq, packed = self._pack_tokens(x)
        bsz = x_seq.shape[0]
        fwd_state = self._ensure_state(self.block_fwd, state_fwd, bsz, x_seq.device)
        out_fwd, _, new_state_fwd = self.block_fwd(x_seq, state=fwd_state, return_state=True,
                                                   use_checkpoint=use_checkpoint)
        bwd_state = self._ensure_state(self.block_bwd, state_bwd_init, bsz, x_seq.device)
        out_bwd = self._backward_block(x_seq, bwd_state, packed)

        out = self._fuse(out_fwd, out_bwd)
        return self._unpack_tokens(out, packed), new_state_fwd

    def _fuse(self, out_fwd: Tensor, out_bwd: Tensor) -> Tensor:
        """``out_proj(gate * out_fwd + (1 - gate) * out_bwd)``, ``gate = sigmoid(Linear(cat))``
        (refiner_backbone.py:129-134) through the library: the gate projection runs as two
        projections over the two halves of its weight (no concatenated tensor), the sigmoid and
        the blend are one kernel, ``out_proj`` is the tensor-core projection."""
        lin = self.fusion_gate[0]
        dim = out_fwd.shape[-1]
        w = lin.weight
        g1 = ops.linear(out_fwd, w[:, :dim], lin.bias)
        g2 = ops.linear(out_bwd, w[:, dim:], None)
        mixed = ops.gate_blend(g1, g2, out_fwd, out_bwd)
        return ops.linear(mixed, self.out_proj.weight, self.out_proj.bias)
