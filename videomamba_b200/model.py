"""``PretrainVideoMamba`` backbone on the libvmb200 mixer kernels.

Drop-in for the reference's models/videomamba/videomamba.py (:371-1200): constructor arguments,
state_dict names, ``forward`` / ``forward_features`` return contracts, streaming state handling
(list / tuple / dict containers, CLS only on the first chunk, ``temporal_pos_offset``), mask
handling and error messages are kept; the layer loop runs on the B200 kernels.  Patch embedding,
position embeddings, masking and pooling are host glue in plain torch (< 2 % of the bytes).
"""
from __future__ import annotations

import logging
import math
from functools import partial
from typing import Any, Dict, List, Optional, Tuple, Union, cast

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .block import Block, DropPath, RMSNorm, apply_norm, create_block
from .streaming import (STREAMING_CONTRACT_VERSION, ForwardReturnSemantics, StateShape,
                        forward_return_semantics as _return_semantics)

logger = logging.getLogger(__name__)

LayerState = Union[Tensor, Tuple[Tensor, Tensor]]
StateCollection = Union[List[LayerState], Tuple[LayerState, ...], Dict[int, LayerState]]


def _pair(v) -> Tuple[int, int]:
    if isinstance(v, (tuple, list)):
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _default_cfg(url: str = "", **kwargs) -> Dict[str, Any]:
    """timm-style default_cfg dict (the reference attaches timm's ``_cfg()``, :1189)."""
    cfg = {"url": url, "num_classes": 1000, "input_size": (3, 224, 224), "pool_size": None,
           "crop_pct": 0.9, "interpolation": "bicubic", "fixed_input_size": True,
           "mean": (0.5, 0.5, 0.5), "std": (0.5, 0.5, 0.5),
           "first_conv": "patch_embed.proj", "classifier": "head"}
    cfg.update(kwargs)
    return cfg


def _infer_spatial_grid(token_count: int, reference_grid: Tuple[int, int]) -> Tuple[int, int]:
    """Factor ``token_count`` into (h, w) with the aspect ratio closest to ``reference_grid``
    (ties: closest in absolute size) -- videomamba.py:32-55."""
    if token_count <= 0:
        raise ValueError("Position embedding must contain at least one spatial token.")
    ref_h, ref_w = reference_grid
    want = ref_h / ref_w
    candidates = []
    for h in range(1, math.isqrt(token_count) + 1):
        if token_count % h == 0:
            w = token_count // h
            candidates += [(h, w), (w, h)]
    if not candidates:
        raise ValueError(f"Unable to infer spatial grid from token count {token_count}.")
    return min(candidates, key=lambda hw: (abs(hw[0] / hw[1] - want),
                                           abs(hw[0] - ref_h) + abs(hw[1] - ref_w)))


def segm_init_weights(m: nn.Module) -> None:
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


def _init_weights(module: nn.Module, n_layer: int, initializer_range: float = 0.02,
                  rescale_prenorm_residual: bool = True, n_residuals_per_layer: int = 1) -> None:
    """GPT-2 style init (videomamba.py:295-324): zero Linear biases unless ``_no_reinit``; residual
    output projections ~ kaiming_uniform / sqrt(n_residuals_per_layer * n_layer)."""
    if isinstance(module, nn.Linear):
        if module.bias is not None and not getattr(module.bias, "_no_reinit", False):
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.Embedding):
        nn.init.normal_(module.weight, std=initializer_range)
    if rescale_prenorm_residual:
        for name, p in module.named_parameters():
            if name in ("out_proj.weight", "fc2.weight"):
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                with torch.no_grad():
                    p /= math.sqrt(n_residuals_per_layer * n_layer)


class PatchEmbed(nn.Module):
    """Clip -> patch tokens: Conv3d with kernel = stride = (tubelet, patch_h, patch_w)."""

    def __init__(self, img_size=224, patch_size=16, kernel_size: int = 1, in_chans: int = 3,
                 embed_dim: int = 768):
        super().__init__()
        self.img_size = _pair(img_size)
        self.patch_size = _pair(patch_size)
        self.num_patches = (self.img_size[1] // self.patch_size[1]) * \
            (self.img_size[0] // self.patch_size[0])
        self.tubelet_size = kernel_size
        k = (kernel_size, self.patch_size[0], self.patch_size[1])
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=k, stride=k)

    def tokens(self, x: Tensor) -> Tensor:
        """(B, C, T, H, W) -> patch tokens (B, T', h, w, D), frame-major.  Kernel == stride, so the
        Conv3d of the reference (videomamba.py:359-368) is one dense projection over
        non-overlapping patches: an im2col permute followed by the library's linear kernel (true
        fp32 in fp32 mode -- cuDNN would use TF32 -- and tensor cores in bf16)."""
        k = self.tubelet_size
        ph, pw = self.patch_size
        B, C, T, H, W = x.shape
        t, h, w = T // k, H // ph, W // pw
        weight = self.proj.weight.reshape(self.proj.weight.shape[0], -1)
        cols = ops.patchify(x.to(weight.dtype), k, ph, pw)      # one gather kernel (csrc/embed.cu)
        out = ops.linear(cols, weight, self.proj.bias)
        return out.reshape(B, t, h, w, -1)

    def forward(self, x: Tensor) -> Tensor:
        """(B, D, T', h, w), the layout the reference's Conv3d returns (a view of ``tokens``)."""
        return self.tokens(x).permute(0, 4, 1, 2, 3)


class PretrainVideoMamba(nn.Module):
    streaming_contract_version: str = STREAMING_CONTRACT_VERSION

    def __init__(self, img_size=224, patch_size=16, depth: int = 24, embed_dim: int = 192,
                 channels: int = 3, drop_path_rate: float = 0.0,
                 ssm_cfg: Optional[Dict[str, object]] = None, norm_epsilon: float = 1e-5,
                 initializer_cfg: Optional[Dict[str, object]] = None, fused_add_norm: bool = True,
                 rms_norm: bool = True, residual_in_fp32: bool = True, bimamba: bool = True,
                 pool_type: str = "cls+avg", kernel_size: int = 1, num_frames: int = 8,
                 device: Optional[torch.device] = None, dtype: Optional[torch.dtype] = None,
                 use_checkpoint: bool = False, checkpoint_num: int = 0,
                 add_pool_norm: bool = True):
        super().__init__()
        if not bimamba:
            raise NotImplementedError("This minimal VideoMamba package only supports bimamba=True.")
        fk: Dict[str, Any] = {}
        if device is not None:
            fk["device"] = device
        if dtype is not None:
            fk["dtype"] = dtype
        self.residual_in_fp32 = residual_in_fp32
        self.fused_add_norm = fused_add_norm
        self.use_checkpoint = use_checkpoint
        self.checkpoint_num = checkpoint_num
        self.depth = depth
        self.pool_type = pool_type
        logger.info("Use checkpoint: %s, checkpoint number: %s, pool type: %s",
                    use_checkpoint, checkpoint_num, pool_type)
        self.d_model = self.num_features = self.embed_dim = embed_dim

        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size,
                                      kernel_size=kernel_size, in_chans=channels,
                                      embed_dim=embed_dim)
        n_patch = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patch + 1, embed_dim))
        self.temporal_pos_embedding = nn.Parameter(
            torch.zeros(1, num_frames // kernel_size, embed_dim))

        rates = [0.0] + [r.item() for r in torch.linspace(0, drop_path_rate, depth)]
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0.0 else nn.Identity()
        self.layers = nn.ModuleList([
            create_block(embed_dim, ssm_cfg=ssm_cfg, norm_epsilon=norm_epsilon, rms_norm=rms_norm,
                         residual_in_fp32=residual_in_fp32, fused_add_norm=fused_add_norm,
                         layer_idx=i, bimamba=bimamba, drop_path=rates[i], **fk)
            for i in range(depth)])
        self.norm = (RMSNorm if rms_norm else nn.LayerNorm)(embed_dim, eps=norm_epsilon, **fk)
        self.add_pool_norm = add_pool_norm
        if add_pool_norm:
            self.pool_norm = nn.LayerNorm(embed_dim)

        # init order of the reference (:478-489): segm init, pos_embed, then the GPT-2 rescale
        self.apply(segm_init_weights)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        self.apply(partial(_init_weights, n_layer=depth, **(initializer_cfg or {})))

    # ---- state allocation (videomamba.py:491-578) ---------------------------------------------
    def _blocks(self) -> List[Block]:
        return [cast(Block, layer) for layer in self.layers]

    def allocate_inference_cache(self, batch_size: int, max_seqlen: int, dtype=None,
                                 **kwargs) -> Dict[int, Tuple[Tensor, Tensor]]:
        return {i: blk.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kwargs)
                for i, blk in enumerate(self._blocks())}

    def init_ssm_state(self, batch_size: int, dtype=None, device=None, as_dict: bool = False):
        """Legacy ssm-only per-layer tensors (updated in place by ``forward``)."""
        states = []
        for blk in self._blocks():
            _, ssm = blk.allocate_inference_cache(batch_size, max_seqlen=1, dtype=dtype)
            states.append(ssm if device is None else ssm.to(device=device))
        return dict(enumerate(states)) if as_dict else states

    def allocate_state(self, batch_size: int, dtype=None, device=None, as_dict: bool = False):
        """Per-layer ``(conv_state, ssm_state)`` zeros for chunked execution
        (contract version ``self.streaming_contract_version``)."""
        states = [blk.mixer.allocate_state(batch_size, dtype=dtype, device=device)
                  for blk in self._blocks()]
        return dict(enumerate(states)) if as_dict else states

    def init_state(self, batch_size: int, dtype=None, device=None, as_dict: bool = False):
        """Backward-compatible alias for ``allocate_state``."""
        return self.allocate_state(batch_size=batch_size, dtype=dtype, device=device,
                                   as_dict=as_dict)

    def expected_state_shapes(self, batch_size: int) -> Dict[int, StateShape]:
        if batch_size <= 0:
            raise ValueError("batch_size must be a positive integer.")
        shapes: Dict[int, StateShape] = {}
        for i, blk in enumerate(self._blocks()):
            mx = blk.mixer
            di, dc, ds = int(getattr(mx, "d_inner")), int(getattr(mx, "d_conv")), \
                int(getattr(mx, "d_state"))
            shapes[i] = StateShape((batch_size, di, dc), (batch_size, di, ds))
        return shapes

    def forward_return_semantics(self) -> ForwardReturnSemantics:
        return _return_semantics(self.add_pool_norm)

    @torch.jit.ignore()
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "temporal_pos_embedding"}

    def get_num_layers(self) -> int:
        return len(self.layers)

    @torch.jit.ignore()
    def load_pretrained(self, checkpoint_path, prefix=""):
        """The reference hands the model to timm's ViT ``.npz`` loader (videomamba.py:587-589); do the
        same when timm is installed.  Plain state_dict checkpoints go through ``load_state_dict(path,
        model, ckpt_num_frame, num_frames)`` below."""
        try:
            from timm.models.vision_transformer import _load_weights
        except ImportError as e:        # timm is an optional dependency of this package
            raise ImportError("load_pretrained needs timm (the reference delegates to "
                              "timm.models.vision_transformer._load_weights); for plain state_dict "
                              "checkpoints use load_state_dict(path, model, ckpt_num_frame, num_frames)") from e
        _load_weights(self, checkpoint_path, prefix)

    # ---- helpers ------------------------------------------------------------------------------
    def _get_layer_state(self, state: Optional[StateCollection], idx: int) -> Optional[LayerState]:
        if state is None:
            return None
        if isinstance(state, dict):
            return state.get(idx)
        if isinstance(state, (list, tuple)):
            return state[idx]
        raise TypeError("state must be a list, tuple, or dict indexed by layer id")

    @staticmethod
    def _is_full_state(layer_state) -> bool:
        return isinstance(layer_state, (list, tuple)) and len(layer_state) == 2

    def _validate_temporal_length(self, frame_count: int) -> int:
        tubelet = self.patch_embed.tubelet_size
        if frame_count <= 0:
            raise ValueError("Input must contain at least one frame.")
        if frame_count % tubelet != 0:
            raise ValueError(f"Input frame count ({frame_count}) must be divisible by "
                             f"tubelet size ({tubelet}).")
        return frame_count // tubelet

    def _spatial_token_grid(self, height: int, width: int) -> Tuple[int, int]:
        ph, pw = self.patch_embed.patch_size
        if height < ph or width < pw:
            raise ValueError("Input spatial size must be at least one patch: "
                             f"got ({height}, {width}) with patch size ({ph}, {pw}).")
        return height // ph, width // pw

    def _get_spatial_pos_embedding(self, grid_h: int, grid_w: int, dtype=None, device=None):
        """Patch position table, bicubically resized when the runtime grid differs (:621-644)."""
        device = self.pos_embed.device if device is None else device
        dtype = self.pos_embed.dtype if dtype is None else dtype
        table = self.pos_embed[:, 1:]
        base = (self.patch_embed.img_size[0] // self.patch_embed.patch_size[0],
                self.patch_embed.img_size[1] // self.patch_embed.patch_size[1])
        if base[0] * base[1] != table.shape[1]:
            base = _infer_spatial_grid(table.shape[1], base)
        if (grid_h, grid_w) == base:
            return table.to(device=device, dtype=dtype)
        grid = table.reshape(1, base[0], base[1], self.embed_dim).permute(0, 3, 1, 2).float()
        grid = F.interpolate(grid, size=(grid_h, grid_w), mode="bicubic", align_corners=False)
        grid = grid.permute(0, 2, 3, 1).reshape(1, grid_h * grid_w, self.embed_dim)
        return grid.to(device=device, dtype=dtype)

    def _has_cls_token_for_forward(self, ssm_state: Optional[StateCollection],
                                   temporal_pos_offset: int) -> bool:
        """CLS is present unless this is a continuation chunk (offset > 0) of a full-state
        stream (:646-653)."""
        if ssm_state is None or temporal_pos_offset <= 0:
            return True
        return not self._is_full_state(self._get_layer_state(ssm_state, 0))

    def _get_temporal_pos_embedding(self, seqlen: int, offset: int = 0, dtype=None,
                                    device=None) -> Tensor:
        """Rows [offset, offset+seqlen) of the temporal table; beyond its end the table is first
        linearly stretched to length offset+seqlen (the reference's behaviour, :655-675)."""
        if offset < 0:
            raise ValueError("temporal_pos_offset must be non-negative.")
        override = getattr(self, "_temporal_rows_override", None)
        if override is not None:        # graphed streaming: rows live in a replay-stable buffer
            if override.shape[1] != seqlen:
                raise ValueError("temporal row buffer does not match the chunk length.")
            return override
        device = self.temporal_pos_embedding.device if device is None else device
        dtype = self.temporal_pos_embedding.dtype if dtype is None else dtype
        table = self.temporal_pos_embedding.to(device=device, dtype=dtype)
        end = offset + seqlen
        if end <= table.shape[1]:
            return table[:, offset:end]
        stretched = F.interpolate(table.permute(0, 2, 1).float(), size=end, mode="linear",
                                  align_corners=False)
        return stretched.permute(0, 2, 1).to(dtype=dtype)[:, offset:end]

    def _normalize_mask(self, mask: Optional[Tensor], batch_size: int, token_count: int,
                        device: torch.device, require_cls_visible: bool) -> Optional[Tensor]:
        if mask is None:
            return None
        if mask.ndim != 2:
            raise ValueError("mask must be 2D with shape [B, N].")
        if mask.shape[0] != batch_size:
            raise ValueError(
                f"mask batch size mismatch: expected {batch_size}, got {mask.shape[0]}.")
        mask = mask.to(device=device, dtype=torch.bool)
        if mask.shape[1] != token_count:
            raise ValueError(
                f"mask token length mismatch: expected {token_count}, got {mask.shape[1]}.")
        if require_cls_visible and token_count > 0 and torch.any(mask[:, 0]):
            raise ValueError("mask must keep CLS token visible (mask[:, 0] must be False).")
        return mask

    def _visible_token_positions(self, mask: Optional[Tensor], batch_size: int, token_count: int,
                                 device: torch.device, require_cls_visible: bool
                                 ) -> Tuple[Optional[Tensor], Optional[Tensor]]:
        """(normalised mask, sorted visible indices per sample) -- :753-784."""
        mask = self._normalize_mask(mask, batch_size, token_count, device, require_cls_visible)
        if mask is None:
            return None, None
        visible = ~mask
        counts = visible.sum(dim=1)
        num_visible = 0
        if counts.numel() > 0:
            if not torch.all(counts == counts[0]):
                raise ValueError("mask must keep the same number of visible tokens per sample; "
                                 f"got per-sample counts: {counts.tolist()}.")
            num_visible = int(counts[0].item())
            if num_visible <= 0:
                raise ValueError("mask must keep at least one visible token per sample.")
        positions = torch.arange(token_count, device=device).unsqueeze(0).expand(batch_size, -1)
        positions = positions.masked_fill(~visible, token_count)
        return mask, torch.sort(positions, dim=1).values[:, :num_visible]

    def _masked_temporal_average(self, patch_tokens: Tensor, visible_positions: Tensor,
                                 temporal_tokens: int, tokens_per_frame: int,
                                 has_cls_token: bool) -> Tensor:
        """Per-frame mean over the visible patch tokens (:702-751)."""
        if patch_tokens.ndim != 3:
            raise ValueError("patch_tokens must have shape [B, N, C].")
        if visible_positions.ndim != 2:
            raise ValueError("visible_positions must have shape [B, N_total_visible].")
        if patch_tokens.shape[0] != visible_positions.shape[0]:
            raise ValueError("Batch size mismatch between patch_tokens and visible_positions.")
        if visible_positions.shape[1] != patch_tokens.shape[1] + (1 if has_cls_token else 0):
            raise ValueError("visible_positions and patch_tokens lengths are inconsistent.")
        if has_cls_token and visible_positions.numel() > 0 \
                and not torch.all(visible_positions[:, 0] == 0):
            raise ValueError("mask must keep CLS token visible for temporal pooling.")
        pos = visible_positions[:, 1:] - 1 if has_cls_token else visible_positions
        frame = torch.div(pos, tokens_per_frame, rounding_mode="floor").to(torch.long)
        B, n, C = patch_tokens.shape
        sums = patch_tokens.new_zeros(B, temporal_tokens, C)
        sums.scatter_add_(1, frame.unsqueeze(-1).expand(-1, -1, C), patch_tokens)
        counts = patch_tokens.new_zeros(B, temporal_tokens, 1)
        counts.scatter_add_(1, frame.unsqueeze(-1), patch_tokens.new_ones(B, n, 1))
        if torch.any(counts == 0):
            raise ValueError("keep_temporal with masking requires at least one visible patch "
                             "token for each temporal slice.")
        return sums / counts

    # ---- forward ------------------------------------------------------------------------------
    def forward_features(self, x: Tensor, mask: Optional[Tensor] = None, use_image: bool = False,
                         ssm_state: Optional[StateCollection] = None,
                         temporal_pos_offset: int = 0):
        """Tokens after the final norm, ``(B, N_vis, C)``; with ``ssm_state`` also the next state in
        the same container type (legacy ssm-only tensors are updated in place and returned)."""
        if x.ndim != 5:
            raise ValueError("x must have shape [B, C, T, H, W].")
        self._validate_temporal_length(x.shape[2])
        patches = self.patch_embed.tokens(x)                    # (B, T', h, w, C), token-major
        B, T, H, W, C = patches.shape
        has_cls = self._has_cls_token_for_forward(ssm_state, temporal_pos_offset)
        cls_row = None
        if has_cls:
            cls_row = (self.cls_token + self.pos_embed[:, :1].to(
                device=patches.device, dtype=patches.dtype)).reshape(C)
        # position embeddings + CLS placement in one pass over the tokens (csrc/embed.cu)
        tokens = ops.embed_tokens(
            patches.reshape(B, T, H * W, C),
            self._get_spatial_pos_embedding(H, W, patches.dtype, patches.device).reshape(H * W, C),
            self._get_temporal_pos_embedding(T, offset=temporal_pos_offset, dtype=patches.dtype,
                                             device=patches.device).reshape(T, C),
            cls_row)

        _, visible = self._visible_token_positions(mask, B, tokens.shape[1], tokens.device,
                                                   require_cls_visible=has_cls)
        if visible is not None:
            tokens = ops.gather_rows(tokens, visible)          # visible-token gather (csrc/pool.cu)

        hidden, residual = tokens, None
        new_states: Optional[Union[Dict[int, LayerState], List[Optional[LayerState]]]] = None
        for idx, blk in enumerate(self._blocks()):
            layer_state = self._get_layer_state(ssm_state, idx)
            full = self._is_full_state(layer_state)
            if full and new_states is None:
                new_states = {} if isinstance(ssm_state, dict) else [None] * len(self.layers)
            ckpt = self.use_checkpoint and idx < self.checkpoint_num
            if full:
                hidden, residual, layer_state = blk(hidden, residual, inference_params=None,
                                                    use_checkpoint=ckpt, state=layer_state,
                                                    return_state=True)
            else:
                hidden, residual = blk(hidden, residual, inference_params=None,
                                       use_checkpoint=ckpt, ssm_state=layer_state)
            if new_states is not None:
                new_states[idx] = layer_state

        if self.fused_add_norm:
            hidden = apply_norm(self.norm, self.drop_path(hidden), residual, False,
                                self.residual_in_fp32)
        else:
            summed = hidden if residual is None else residual + self.drop_path(hidden)
            hidden = apply_norm(self.norm, summed.to(dtype=self.norm.weight.dtype), None, False,
                                False)

        if ssm_state is None:
            return hidden
        if new_states is None:
            return hidden, ssm_state
        if isinstance(new_states, dict):
            return hidden, new_states
        if any(item is None for item in new_states):
            raise ValueError("Expected full state for all layers.")
        return hidden, (tuple(new_states) if isinstance(ssm_state, tuple) else list(new_states))

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, use_image: bool = False,
                keep_temporal: bool = False, ssm_state: Optional[StateCollection] = None,
                temporal_pos_offset: int = 0):
        """Returns per ``forward_return_semantics()``: ``(x_vis, x_pool[, next_state])`` with
        ``add_pool_norm`` (x_vis without CLS), else ``x_vis[, next_state]`` (CLS kept)."""
        if x.ndim != 5:
            raise ValueError("x must have shape [B, C, T, H, W].")
        grid_h, grid_w = self._spatial_token_grid(x.shape[-2], x.shape[-1])
        per_frame = grid_h * grid_w
        temporal_tokens = self._validate_temporal_length(x.shape[2])
        has_cls = self._has_cls_token_for_forward(ssm_state, temporal_pos_offset)
        feats = self.forward_features(x, mask, use_image, ssm_state=ssm_state,
                                      temporal_pos_offset=temporal_pos_offset)
        if ssm_state is None:
            x_vis = cast(Tensor, feats)
        else:
            x_vis, ssm_state = feats
        if not self.add_pool_norm:
            return x_vis if ssm_state is None else (x_vis, ssm_state)

        cls_token = x_vis[:, :1] if has_cls else None
        patches = x_vis[:, 1:] if has_cls else x_vis
        if self.pool_type in {"cls", "cls+avg", "cls_cat_avg"} and cls_token is None:
            raise ValueError(
                f"pool_type='{self.pool_type}' requires a CLS token, but continuation "
                "streaming chunks (temporal_pos_offset > 0 with full state) do not include CLS. "
                "Use pool_type='avg' for chunked streaming.")
        if self.pool_type != "cls" and patches.shape[1] == 0:
            raise ValueError("mask must keep at least one patch token visible when using "
                             f"pool_type='{self.pool_type}'.")
        if self.pool_type not in ops.POOL_MODES:
            raise ValueError(f"Unsupported pool_type: {self.pool_type}")
        if self.pool_type == "cls" or not (keep_temporal and mask is not None):
            # mean over the patch tokens (all, or per frame) + CLS combine + pool_norm: csrc/pool.cu
            groups = temporal_tokens if (keep_temporal and self.pool_type != "cls") else 1
            pooled = ops.pool_norm(x_vis, has_cls, groups, patches.shape[1] // groups, self.pool_type,
                                   self.pool_norm.weight, self.pool_norm.bias, self.pool_norm.eps)
        else:
            # keep_temporal under a mask: per-frame mean over the VISIBLE patch tokens (host-checked like
            # the reference, videomamba.py:702-751), then the reference's own combine + pool_norm
            total = (1 if has_cls else 0) + temporal_tokens * per_frame
            _, visible = self._visible_token_positions(mask, patches.shape[0], total,
                                                       x.device, require_cls_visible=has_cls)
            avg = self._masked_temporal_average(patches, visible, temporal_tokens, per_frame, has_cls)
            if self.pool_type == "cls+avg":
                pooled = self.pool_norm(cls_token + avg)
            elif self.pool_type == "cls_cat_avg":
                pooled = self.pool_norm(torch.cat([cls_token, avg], dim=1))
            else:
                pooled = self.pool_norm(avg)
        if ssm_state is None:
            return patches, pooled
        return patches, pooled, ssm_state


def load_state_dict(pretrained_path, model: PretrainVideoMamba, ckpt_num_frame, num_frames):
    """Load a PLAIN state_dict checkpoint, resizing the spatial table (bicubic) and the temporal
    table (linear) to the model's geometry -- videomamba.py:1070-1147."""
    logger.info("Loading pretrained weights from %s", pretrained_path)
    try:
        ckpt = torch.load(pretrained_path, map_location="cpu", weights_only=True)
    except TypeError:  # very old torch without weights_only
        ckpt = torch.load(pretrained_path, map_location="cpu")
    if not isinstance(ckpt, dict):
        raise TypeError("Expected a plain state_dict (dict) checkpoint.")
    if "model" in ckpt or "module" in ckpt:
        raise ValueError("Checkpoint wrapper keys ('model'/'module') are not supported. "
                         "Pass a plain state_dict checkpoint.")

    pos = ckpt["pos_embed"]
    width = pos.shape[-1]
    n_patch = model.patch_embed.num_patches
    n_extra = model.pos_embed.shape[-2] - n_patch
    new_hw = (model.patch_embed.img_size[0] // model.patch_embed.patch_size[0],
              model.patch_embed.img_size[1] // model.patch_embed.patch_size[1])
    if new_hw[0] * new_hw[1] != n_patch:
        raise ValueError("Model patch grid size mismatch: "
                         f"{new_hw[0]}x{new_hw[1]} != num_patches({n_patch}).")
    old_hw = _infer_spatial_grid(pos.shape[-2] - n_extra, new_hw)
    if old_hw != new_hw:
        logger.info("Position interpolate from %dx%d to %dx%d", *old_hw, *new_hw)
        grid = pos[:, n_extra:].reshape(-1, old_hw[0], old_hw[1], width).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=new_hw, mode="bicubic", align_corners=False)
        grid = grid.permute(0, 2, 3, 1).flatten(1, 2)
        ckpt["pos_embed"] = torch.cat((pos[:, :n_extra], grid), dim=1)

    if ckpt_num_frame is None or ckpt_num_frame <= 0:
        raise ValueError("ckpt_num_frame must be a positive integer when loading pretrained weights.")
    tubelet = model.patch_embed.tubelet_size
    old_t, new_t = ckpt_num_frame // tubelet, num_frames // tubelet
    if old_t != new_t:
        logger.info("Temporal interpolate from %d to %d", old_t, new_t)
        table = ckpt["temporal_pos_embedding"].permute(0, 2, 1)
        table = F.interpolate(table, size=(new_t,), mode="linear", align_corners=False)
        ckpt["temporal_pos_embedding"] = table.permute(0, 2, 1)
    logger.info(model.load_state_dict(ckpt, strict=True))


def build_videomamba(config, add_pool_norm: bool = True) -> PretrainVideoMamba:
    """Model factory: reads ``config.vision_encoder.*`` (attribute access; every field the
    reference reads is required, ``channels`` included) -- videomamba.py:1150-1200."""
    v = config.vision_encoder
    model = PretrainVideoMamba(
        img_size=v.img_size, patch_size=v.patch_size, depth=v.depth, embed_dim=v.embed_dim,
        channels=v.channels, drop_path_rate=v.drop_path_rate, ssm_cfg=v.ssm_cfg,
        norm_epsilon=v.norm_epsilon, fused_add_norm=v.fused_add_norm, rms_norm=v.rms_norm,
        residual_in_fp32=v.residual_in_fp32, bimamba=v.bimamba, pool_type=v.pool_type,
        kernel_size=v.kernel_size, num_frames=v.num_frames, use_checkpoint=v.use_checkpoint,
        checkpoint_num=v.checkpoint_num, add_pool_norm=add_pool_norm)
    object.__setattr__(model, "default_cfg", _default_cfg())
    if v.pretrained is not None:
        load_state_dict(pretrained_path=v.pretrained, model=model,
                        ckpt_num_frame=v.ckpt_num_frame, num_frames=v.num_frames)
    else:
        logger.info("No pretrained weights!!!")
    return model
