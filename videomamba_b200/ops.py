"""Tensor-level operators of the B200 mixer path (thin marshalling over the C ABI).

Two groups:

* token-major primitives used by the modules of this package (``add_norm``, ``linear``,
  ``mixer_fwd`` ...): activations are (batch, token, channel);
* drop-in replacements for the third-party operators the reference imports, with THEIR
  signatures and channel-major layouts (``causal_conv1d_fn``, ``selective_scan_fn``,
  ``rms_norm_fn`` ...; reference call sites listed in SURVEY.md section 8b).

Everything enqueues on the current CUDA stream of the tensors' device and returns immediately.
PyTorch is used for memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import VMB_BF16, VMB_F32, FusedScanArgs, MixerArgs, ScanArgs, StepArgs

LOG2E = 1.4426950408889634

_CUDA_ONLY_MSG = ("VideoMamba requires CUDA tensors in this package because "
                  "the mixer kernels are CUDA-only (sm_100a).")


def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return VMB_F32
    if t.dtype == torch.bfloat16:
        return VMB_BF16
    raise TypeError(f"videomamba_b200 kernels support float32 and bfloat16, got {t.dtype}")


def _p(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _require_cuda(t: Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError(_CUDA_ONLY_MSG)


class _on_device:
    """Make the tensor's device current for the duration of a launch (no-op when it already is)."""

    __slots__ = ("idx", "prev")

    def __init__(self, t: Tensor):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is not None and cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


class _ForwardOnly(torch.autograd.Function):
    """Entry points WITHOUT a backward kernel (the fused inference mixer / scan / decode step, the
    reversed token walks).  Outputs computed from tensors that require grad get this node, so
    ``loss.backward()`` fails HERE with a clear message instead of silently dropping gradients.
    The differentiable operators live in ``autograd.py``."""

    @staticmethod
    def forward(ctx, out, *deps):
        return out.view_as(out)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(
            "this libvmb200 entry point is forward-only (fused inference kernel or reversed token walk). "
            "Call it under torch.no_grad(), or go through the differentiable operators "
            "(videomamba_b200.autograd / the modules in training mode).")


def forward_only(out, *deps):
    """``out`` (a tensor, or a tuple whose first element is the main output) tagged so that autograd
    cannot flow through it silently; no-op under ``no_grad`` or when nothing requires grad."""
    if not torch.is_grad_enabled():
        return out
    deps = tuple(d for d in deps if isinstance(d, Tensor) and d.requires_grad)
    if not deps:
        return out
    if isinstance(out, tuple):
        return (_ForwardOnly.apply(out[0], *deps),) + tuple(out[1:])
    return _ForwardOnly.apply(out, *deps)


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, Tensor) and t.requires_grad for t in tensors)


def aligned_rows(t: Tensor, elems: int = 4) -> Tensor:
    """A view whose base address is a multiple of ``elems`` elements, or a contiguous copy."""
    return t if t.data_ptr() % (elems * t.element_size()) == 0 else t.clone(memory_format=torch.contiguous_format)


def xdbl_pitch(dt_rank: int, d_state: int) -> int:
    """Row pitch (elements) of the x_dbl buffer: [dt_low | B | C] padded to a multiple of 16."""
    return (dt_rank + 2 * d_state + 15) // 16 * 16


# ----------------------------------------------------------------------------------------------
# token-major primitives
# ----------------------------------------------------------------------------------------------
def add_norm_raw(x: Tensor, weight: Tensor, bias: Optional[Tensor], residual: Optional[Tensor],
                 eps: float, is_rms: bool, prenorm: bool, residual_in_fp32: bool):
    """One ``vmb_add_norm_fwd`` launch.  Returns ``(y, residual_out | None, x2, res2)``: ``y`` in x's
    shape, ``residual_out`` in x's shape (None when the sum is x itself), and the 2-D operands the
    kernel read (what the backward needs)."""
    _require_cuda(x)
    lib = _lib.load()
    dim = x.shape[-1]
    if x.stride(-1) != 1:
        x = x.contiguous()
    x2 = x.reshape(-1, dim)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < dim):
        x2 = x2.contiguous()
    x2 = aligned_rows(x2)       # e.g. a last-dim slice big[..., 2:2+dim]: vector loads need 4-element alignment
    rows = x2.shape[0]
    res2 = None
    if residual is not None:
        if residual.shape != x.shape:
            raise ValueError("residual must have the same shape as x")
        res2 = residual.reshape(-1, dim).contiguous()
        res_dtype = residual.dtype
    else:
        res_dtype = torch.float32 if residual_in_fp32 else x.dtype
    y = torch.empty((rows, dim), dtype=x.dtype, device=x.device)
    need_res_out = prenorm and (residual is not None or res_dtype != x.dtype)
    res_out = torch.empty((rows, dim), dtype=res_dtype, device=x.device) if need_res_out else None
    weight = weight.contiguous()
    if bias is not None:
        bias = bias.to(weight.dtype).contiguous()
    with _on_device(x):
        rc = lib.vmb_add_norm_fwd(
            _p(x2), _dt(x2), x2.stride(0) if rows > 1 else dim,
            _p(res2), _dt(res2) if res2 is not None else VMB_F32,
            _p(weight), _p(bias), _dt(weight), _p(y),
            _p(res_out), _dt(res_out) if res_out is not None else VMB_F32,
            rows, dim, float(eps), 1 if is_rms else 0, _stream(x))
    _lib.check(rc, "vmb_add_norm_fwd")
    return y.reshape(x.shape), None if res_out is None else res_out.reshape(x.shape), x2, res2


def add_norm(x: Tensor, weight: Tensor, bias: Optional[Tensor], residual: Optional[Tensor],
             eps: float, is_rms: bool, prenorm: bool, residual_in_fp32: bool):
    """Fused residual add + norm over the last dim.  Returns ``y`` or ``(y, residual_out)``.
    Differentiable (``autograd.AddNormFn``) when an input requires grad."""
    res_dtype = residual.dtype if residual is not None else (torch.float32 if residual_in_fp32 else x.dtype)
    alias = prenorm and residual is None and res_dtype == x.dtype   # the residual stream IS x
    if _wants_grad(x, weight, bias, residual):
        from . import autograd as ag
        if alias or not prenorm:
            y = ag.AddNormFn.apply(x, weight, bias, residual, eps, is_rms, False, residual_in_fp32)
            return (y, x) if prenorm else y
        return ag.AddNormFn.apply(x, weight, bias, residual, eps, is_rms, True, residual_in_fp32)
    y, res_out, _, _ = add_norm_raw(x, weight, bias, residual, eps, is_rms, prenorm, residual_in_fp32)
    if not prenorm:
        return y
    return y, (x if res_out is None else res_out)


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """``x @ weight.T (+ bias)`` over the last dim (fp32 accumulate, one rounding).  Differentiable
    (``autograd.LinearFn``) when an input requires grad."""
    if _wants_grad(x, weight, bias):
        from . import autograd as ag
        return ag.LinearFn.apply(x, weight, bias)
    return linear_raw(x, weight, bias)


def linear_raw(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, silu_from: Optional[int] = None) -> Tensor:
    """``silu_from``: SiLU on the output columns ``[silu_from, N)`` in the projection's epilogue
    (``vmb_linear_fwd_act``: tensor-core kernel only, a multiple of 64)."""
    _require_cuda(x)
    lib = _lib.load()
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise ValueError("linear: weight / input size mismatch")
    if weight.dtype != x.dtype:
        weight = weight.to(x.dtype)
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < K):
        x2 = x2.contiguous()
    weight = weight if weight.stride(-1) == 1 else weight.contiguous()
    M = x2.shape[0]
    out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    with _on_device(x):
        if silu_from is None:
            rc = lib.vmb_linear_fwd(_p(x2), x2.stride(0) if M > 1 else K, _p(weight), weight.stride(0),
                                    _p(bias), _p(out), N, M, N, K, _dt(x2), _stream(x))
        else:
            rc = lib.vmb_linear_fwd_act(_p(x2), x2.stride(0) if M > 1 else K, _p(weight), weight.stride(0),
                                        _p(bias), _p(out), N, M, N, K, _dt(x2), int(silu_from), _stream(x))
    _lib.check(rc, "vmb_linear_fwd" if silu_from is None else "vmb_linear_fwd_act")
    return out.reshape(*x.shape[:-1], N)


def linear_act(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, silu_from: int = 0) -> Tensor:
    """``x @ weight.T (+ bias)`` with SiLU on the output columns ``[silu_from, N)`` applied in the epilogue of the
    tensor-core projection (forward only).  ``in_proj`` with ``silu_from = d_inner`` gives ``x | SiLU(z)``."""
    return forward_only(linear_raw(x, weight, bias, silu_from=silu_from), x, weight, bias)


def conv_xproj_tokens(x: Tensor, conv_weight: Tensor, conv_bias: Optional[Tensor], w_x_pad: Tensor):
    """Stateless causal conv (d_conv 4) + SiLU over ``x (B, L, Di)`` (may be a strided view) fused with
    the x_proj projection: returns ``(xc (B, L, Di), x_dbl (B, L, N))``, bit-identical to
    ``causal_conv1d_tokens`` followed by ``linear`` (reference mamba_simple.py:381-409)."""
    _require_cuda(x)
    lib = _lib.load()
    B, L, Di = x.shape
    if x.stride(-1) != 1 or x.stride(0) != L * x.stride(1):
        x = x.contiguous()
    N = w_x_pad.shape[0]
    conv_weight = conv_weight.reshape(Di, -1).to(x.dtype).contiguous()
    if conv_bias is not None:
        conv_bias = conv_bias.to(x.dtype).contiguous()
    w_x_pad = w_x_pad.to(x.dtype).contiguous()
    xc = torch.empty((B, L, Di), dtype=x.dtype, device=x.device)
    xdbl = torch.empty((B, L, N), dtype=x.dtype, device=x.device)
    with _on_device(x):
        rc = lib.vmb_conv_xproj_fwd(_p(x), x.stride(1), _p(conv_weight), _p(conv_bias), _p(w_x_pad), Di,
                                    _p(xc), Di, _p(xdbl), N, B * L, N, Di, L, _stream(x))
    _lib.check(rc, "vmb_conv_xproj_fwd")
    return xc, xdbl


def gate_blend(g1: Tensor, g2: Optional[Tensor], fwd: Tensor, bwd: Tensor) -> Tensor:
    """``s * fwd + (1 - s) * bwd`` with ``s = sigmoid(g1 (+ g2))`` in one pass (refiner fusion gate,
    reference models/refiner_backbone.py:129-134)."""
    _require_cuda(fwd)
    lib = _lib.load()
    if not (g1.shape == fwd.shape == bwd.shape) or (g2 is not None and g2.shape != fwd.shape):
        raise ValueError("gate_blend: shape mismatch")
    if _wants_grad(g1, g2, fwd, bwd):      # autograd glue (element-wise, not on the mixer path)
        s = torch.sigmoid((g1.float() + g2.float()) if g2 is not None else g1.float())
        return (s * fwd.float() + (1.0 - s) * bwd.float()).to(fwd.dtype)
    g1, fwd, bwd = aligned_rows(g1.contiguous()), aligned_rows(fwd.contiguous()), aligned_rows(bwd.contiguous())
    if g2 is not None:
        g2 = aligned_rows(g2.contiguous())
    out = torch.empty_like(fwd)
    with _on_device(fwd):
        rc = lib.vmb_gate_blend_fwd(_p(g1), _p(g2), _p(fwd), _p(bwd), _p(out), fwd.numel(), _dt(fwd),
                                    _stream(fwd))
    _lib.check(rc, "vmb_gate_blend_fwd")
    return forward_only(out, g1, g2, fwd, bwd)


def patchify(x: Tensor, tubelet: int, ph: int, pw: int) -> Tensor:
    """Clip ``(B, C, T, H, W)`` -> patch rows ``(B*t*h*w, C*tubelet*ph*pw)`` (im2col of a Conv3d whose
    kernel equals its stride; rows ordered (b, t, y, x), columns (c, dt, dy, dx))."""
    _require_cuda(x)
    lib = _lib.load()
    x = x.contiguous()
    B, C, T, H, W = x.shape
    t, h, w = T // tubelet, H // ph, W // pw
    if _wants_grad(x):        # autograd glue (gradient with respect to the clip): a permuting view + copy
        v = x[:, :, :t * tubelet, :h * ph, :w * pw].reshape(B, C, t, tubelet, h, ph, w, pw)
        return v.permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B * t * h * w, C * tubelet * ph * pw)
    cols = torch.empty((B * t * h * w, C * tubelet * ph * pw), dtype=x.dtype, device=x.device)
    with _on_device(x):
        rc = lib.vmb_patchify(_p(x), _p(cols), B, C, T, H, W, tubelet, ph, pw, _dt(x), _stream(x))
    _lib.check(rc, "vmb_patchify")
    return cols


def embed_tokens(patches: Tensor, spatial: Tensor, temporal: Tensor,
                 cls_row: Optional[Tensor] = None) -> Tensor:
    """``(B, t, hw, D)`` patch tokens + spatial ``(hw, D)`` + temporal ``(t, D)`` position embeddings
    (each add rounded to the token dtype), CLS row ``(D,)`` written at position 0 when given:
    ``(B, has_cls + t*hw, D)``."""
    _require_cuda(patches)
    lib = _lib.load()
    B, t, hw, D = patches.shape
    dt = patches.dtype
    if _wants_grad(patches, spatial, temporal, cls_row):   # autograd glue: the same two rounded adds + cat
        tok = patches + spatial.to(dt).reshape(1, 1, hw, D)
        tok = (tok + temporal.to(dt).reshape(1, t, 1, D)).reshape(B, t * hw, D)
        if cls_row is not None:
            tok = torch.cat([cls_row.to(dt).reshape(1, 1, D).expand(B, -1, -1), tok], dim=1)
        return tok
    patches = patches.contiguous()
    spatial = spatial.to(dt).reshape(hw, D).contiguous()
    temporal = temporal.to(dt).reshape(t, D).contiguous()
    if cls_row is not None:
        cls_row = cls_row.to(dt).reshape(D).contiguous()
    out = torch.empty((B, (cls_row is not None) + t * hw, D), dtype=dt, device=patches.device)
    with _on_device(patches):
        rc = lib.vmb_embed_tokens(_p(patches), _p(spatial), _p(temporal), _p(cls_row), _p(out), B, t, hw,
                                  D, _dt(patches), _stream(patches))
    _lib.check(rc, "vmb_embed_tokens")
    return forward_only(out, patches, spatial, temporal, cls_row)


POOL_MODES = {"cls": 0, "cls+avg": 1, "cls_cat_avg": 2, "avg": 3}


def pool_norm(x_vis: Tensor, has_cls: bool, groups: int, per: int, pool_type: str,
              ln_weight: Optional[Tensor], ln_bias: Optional[Tensor], eps: float) -> Tensor:
    """Pooling over the patch tokens + ``pool_norm`` (reference videomamba.py:983-1063) in two launches.
    ``x_vis (B, has_cls + groups * per, C)``: tokens after the final norm.  ``groups = 1`` averages all
    patch tokens, ``groups = T`` with ``per`` tokens per frame is ``keep_temporal``."""
    _require_cuda(x_vis)
    lib = _lib.load()
    mode = POOL_MODES[pool_type]
    x_vis = _token_major(x_vis)
    B, L, Cdim = x_vis.shape
    if L != (1 if has_cls else 0) + groups * per:
        raise ValueError("pool_norm: token count does not match has_cls + groups * per")
    if _wants_grad(x_vis, ln_weight, ln_bias):     # autograd glue: the reference's own torch ops
        cls = x_vis[:, :1] if has_cls else None
        pat = x_vis[:, 1:] if has_cls else x_vis
        if mode == 0:
            pooled = cls
        else:
            avg = pat.reshape(B, groups, per, Cdim).mean(dim=2)
            pooled = cls + avg if mode == 1 else (torch.cat([cls, avg], dim=1) if mode == 2 else avg)
        return torch.nn.functional.layer_norm(pooled, (Cdim,), ln_weight, ln_bias, eps)
    cast = lambda t: None if t is None else t.to(x_vis.dtype).contiguous()
    ln_weight, ln_bias = cast(ln_weight), cast(ln_bias)
    rows = 1 if mode == 0 else (groups + 1 if mode == 2 else groups)
    out = torch.empty((B, rows, Cdim), dtype=x_vis.dtype, device=x_vis.device)
    nbytes = lib.vmb_pool_norm_workspace_bytes(B, groups, per, Cdim) if mode != 0 else 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x_vis.device) if nbytes else None
    with _on_device(x_vis):
        rc = lib.vmb_pool_norm_fwd(_p(x_vis), x_vis.stride(0), x_vis.stride(1), B, groups, per, Cdim,
                                   1 if has_cls else 0, mode, _p(ln_weight), _p(ln_bias), float(eps), _p(out),
                                   _p(ws), nbytes, _dt(x_vis), _stream(x_vis))
    _lib.check(rc, "vmb_pool_norm_fwd")
    return forward_only(out, x_vis, ln_weight, ln_bias)


def gather_rows(src: Tensor, index: Tensor) -> Tensor:
    """``src (B, L, C)``, ``index (B, n)`` -> ``(B, n, C)`` rows ``src[b, index[b, i]]`` (the visible-token
    gather of the masked path, reference videomamba.py:826-836)."""
    _require_cuda(src)
    lib = _lib.load()
    src = _token_major(src)
    B, L, Cdim = src.shape
    index = index.to(device=src.device, dtype=torch.int64).contiguous()
    if _wants_grad(src):                           # autograd glue
        return torch.gather(src, 1, index.unsqueeze(-1).expand(-1, -1, Cdim))
    n = index.shape[1]
    out = torch.empty((B, n, Cdim), dtype=src.dtype, device=src.device)
    with _on_device(src):
        rc = lib.vmb_gather_rows(_p(src), src.stride(0), src.stride(1), _p(index), B, n, Cdim, _p(out),
                                 _dt(src), _stream(src))
    _lib.check(rc, "vmb_gather_rows")
    return forward_only(out, src)


def _token_major(t: Tensor) -> Tensor:
    """(B, L, C) tensor with channel stride 1 (copy only when needed)."""
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def causal_conv1d_tokens(x: Tensor, weight: Tensor, bias: Optional[Tensor],
                         conv_state: Optional[Tensor] = None, want_state: bool = False,
                         silu: bool = True, reverse: bool = False, frame_len: int = 0):
    """Depthwise causal conv on token-major ``x (B, L, Di)``; ``weight (Di, W)``.
    ``reverse``: walk the tokens back to front; with ``frame_len > 0`` only the FRAME axis is reversed
    (frames of ``frame_len`` tokens back to front, tokens inside a frame front to back).
    Returns ``y`` or ``(y, new_conv_state (B, Di, W))``.  The forward walk is differentiable
    (``autograd.ConvFn``); the reversed walks are forward-only."""
    if not reverse and _wants_grad(x, weight, bias, conv_state):
        from . import autograd as ag
        y, cs = ag.ConvFn.apply(x, weight, bias, conv_state, want_state, silu)
        return (y, cs) if want_state else y
    out = causal_conv1d_tokens_raw(x, weight, bias, conv_state, want_state, silu, reverse, frame_len)
    return forward_only(out, x, weight, bias, conv_state)


def causal_conv1d_tokens_raw(x: Tensor, weight: Tensor, bias: Optional[Tensor],
                             conv_state: Optional[Tensor] = None, want_state: bool = False,
                             silu: bool = True, reverse: bool = False, frame_len: int = 0):
    _require_cuda(x)
    lib = _lib.load()
    x = _token_major(x)
    B, L, Di = x.shape
    W = weight.shape[-1]
    weight = weight.reshape(Di, W).to(x.dtype).contiguous()
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    y = torch.empty((B, L, Di), dtype=x.dtype, device=x.device)
    cs_in = None
    if conv_state is not None:
        if tuple(conv_state.shape) != (B, Di, W):
            raise ValueError(f"conv_state must have shape {(B, Di, W)}, got {tuple(conv_state.shape)}")
        cs_in = conv_state.contiguous()
    cs_out = None
    if want_state:
        cs_dtype = x.dtype if conv_state is None else torch.promote_types(conv_state.dtype, x.dtype)
        cs_out = torch.empty((B, Di, W), dtype=cs_dtype, device=x.device)
    with _on_device(x):
        rc = lib.vmb_causal_conv1d_fwd(
            _p(x), x.stride(0), x.stride(1), _p(weight), _p(bias),
            _p(cs_in), _dt(cs_in) if cs_in is not None else VMB_F32,
            _p(y), y.stride(0), y.stride(1),
            _p(cs_out), _dt(cs_out) if cs_out is not None else VMB_F32,
            B, L, Di, W, 1 if silu else 0, 1 if reverse else 0, int(frame_len) if reverse else 0,
            _dt(x), _stream(x))
    _lib.check(rc, "vmb_causal_conv1d_fwd")
    return (y, cs_out) if want_state else y


def selective_scan_tokens(u: Tensor, delta: Tensor, A2: Tensor, bc: Tensor, b_off: int, c_off: int,
                          d_state: int, D: Optional[Tensor] = None, z: Optional[Tensor] = None,
                          dt_bias: Optional[Tensor] = None, softplus: bool = True,
                          h0: Optional[Tensor] = None, want_last: bool = False,
                          reverse: bool = False, frame_len: int = 0):
    """Token-major selective scan.  ``u, delta, z: (B, L, Di)``; ``bc: (B, L, >=c_off+N)`` holds
    B_t at ``[b_off, b_off+N)`` and C_t at ``[c_off, c_off+N)``; ``A2 = A*log2(e)`` fp32 (Di, N).
    Forward-only in this form (``A2`` is a derived operand); the differentiable operator is
    ``autograd.ScanFn`` (natural ``A``), used by ``selective_scan_fn`` and the training-mode mixer."""
    out = selective_scan_tokens_raw(u, delta, A2, bc, b_off, c_off, d_state, D, z, dt_bias, softplus, h0,
                                    want_last, reverse, frame_len)
    return forward_only(out, u, delta, A2, bc, D, z, dt_bias, h0)


def selective_scan_tokens_raw(u: Tensor, delta: Tensor, A2: Tensor, bc: Tensor, b_off: int, c_off: int,
                              d_state: int, D: Optional[Tensor] = None, z: Optional[Tensor] = None,
                              dt_bias: Optional[Tensor] = None, softplus: bool = True,
                              h0: Optional[Tensor] = None, want_last: bool = False,
                              reverse: bool = False, frame_len: int = 0):
    _require_cuda(u)
    lib = _lib.load()
    u, delta, bc = _token_major(u), _token_major(delta), _token_major(bc)
    if z is not None:
        z = _token_major(z)
    B, L, Di = u.shape
    y = torch.empty((B, L, Di), dtype=u.dtype, device=u.device)
    h_last = torch.empty((B, Di, d_state), dtype=torch.float32, device=u.device) if want_last else None
    if h0 is not None:
        h0 = h0.contiguous()
    a = ScanArgs()
    a.u, a.u_bstride, a.u_tstride = u.data_ptr(), u.stride(0), u.stride(1)
    a.delta, a.d_bstride, a.d_tstride = delta.data_ptr(), delta.stride(0), delta.stride(1)
    if z is not None:
        a.z, a.z_bstride, a.z_tstride = z.data_ptr(), z.stride(0), z.stride(1)
    a.bc, a.bc_bstride, a.bc_tstride = bc.data_ptr(), bc.stride(0), bc.stride(1)
    a.b_off, a.c_off = b_off, c_off
    a.A2 = A2.data_ptr()
    a.D = None if D is None else D.data_ptr()
    a.dt_bias = None if dt_bias is None else dt_bias.data_ptr()
    if h0 is not None:
        a.h0, a.h0_dtype = h0.data_ptr(), _dt(h0)
    a.y, a.y_bstride, a.y_tstride = y.data_ptr(), y.stride(0), y.stride(1)
    a.h_last = None if h_last is None else h_last.data_ptr()
    a.B, a.L, a.Di, a.N = B, L, Di, d_state
    a.dtype, a.softplus, a.reverse = _dt(u), 1 if softplus else 0, 1 if reverse else 0
    a.frame_len = int(frame_len) if reverse else 0
    with _on_device(u):
        rc = lib.vmb_selective_scan_fwd(C.byref(a), _stream(u))
    _lib.check(rc, "vmb_selective_scan_fwd")
    return (y, h_last) if want_last else y


def selective_scan_fused_tokens(u: Tensor, z: Tensor, xdbl: Tensor, w_dt: Tensor, A2: Tensor,
                                dt_rank: int, d_state: int, D: Optional[Tensor] = None,
                                dt_bias: Optional[Tensor] = None, h0: Optional[Tensor] = None,
                                want_last: bool = False, reverse: bool = False,
                                allow_split: bool = True, a_geometric: bool = False, tune: int = 0,
                                frame_len: int = 0, z_gate: bool = False):
    """Fused dt_proj + softplus + scan + D skip + SiLU(z) gate (bf16, d_state 16).
    ``u, z: (B, L, Di)``; ``xdbl: (B, L, Xp)`` rows ``[dt_low (R) | B (N) | C (N) | pad]`` with
    ``Xp = xdbl_pitch(R, N)``; ``w_dt: (Di, R)`` bf16, or already zero-padded to ``(Di, Rp)`` with
    ``Rp = round_up(R, 16)``; ``A2 = A*log2(e)`` fp32.  ``a_geometric``: the caller's promise that
    ``A2[d, n] == (n+1) * A2[d, 0]`` (see ``is_geometric``).  ``z_gate``: ``z`` already holds ``SiLU(z)``
    (``linear(..., silu_from=Di)``).  Raises when the shape is not covered.
    Forward-only in this form; ``autograd.FusedScanFn`` is the differentiable operator over it."""
    out = selective_scan_fused_tokens_raw(u, z, xdbl, w_dt, A2, dt_rank, d_state, D, dt_bias, h0, want_last,
                                          reverse, allow_split, a_geometric, tune, frame_len, z_gate=z_gate)
    return forward_only(out, u, z, xdbl, w_dt, A2, D, dt_bias, h0)


def selective_scan_fused_tokens_raw(u: Tensor, z: Tensor, xdbl: Tensor, w_dt: Tensor, A2: Tensor,
                                    dt_rank: int, d_state: int, D: Optional[Tensor] = None,
                                    dt_bias: Optional[Tensor] = None, h0: Optional[Tensor] = None,
                                    want_last: bool = False, reverse: bool = False,
                                    allow_split: bool = True, a_geometric: bool = False, tune: int = 0,
                                    frame_len: int = 0, bwd_ckpt: Optional[Tensor] = None,
                                    z_gate: bool = False):
    """``bwd_ckpt``: a ``scan_bwd_ckpt_bytes(B, L, Di)`` byte buffer that receives the state records the
    backward kernel needs (training forward; forward walk only)."""
    _require_cuda(u)
    lib = _lib.load()
    u, z, xdbl = _token_major(u), _token_major(z), _token_major(xdbl)
    rp = (dt_rank + 15) // 16 * 16      # the kernel reads whole 16-wide k-steps: zero-pad the rank
    if w_dt.shape[1] != rp or w_dt.stride(-1) != 1:   # decided on shapes only (no device sync)
        padded = torch.zeros((w_dt.shape[0], rp), dtype=w_dt.dtype, device=w_dt.device)
        padded[:, :dt_rank] = w_dt[:, :dt_rank]
        w_dt = padded
    B, L, Di = u.shape
    y = torch.empty((B, L, Di), dtype=u.dtype, device=u.device)
    h_last = torch.empty((B, Di, d_state), dtype=torch.float32, device=u.device) if want_last else None
    if h0 is not None:
        h0 = h0.contiguous()
    a = FusedScanArgs()
    a.u, a.u_bstride, a.u_tstride = u.data_ptr(), u.stride(0), u.stride(1)
    a.z, a.z_bstride, a.z_tstride = z.data_ptr(), z.stride(0), z.stride(1)
    a.xdbl, a.x_bstride, a.x_tstride = xdbl.data_ptr(), xdbl.stride(0), xdbl.stride(1)
    a.w_dt, a.A2 = w_dt.data_ptr(), A2.data_ptr()
    a.D = None if D is None else D.data_ptr()
    a.dt_bias = None if dt_bias is None else dt_bias.data_ptr()
    if h0 is not None:
        a.h0, a.h0_dtype = h0.data_ptr(), _dt(h0)
    a.y, a.y_bstride, a.y_tstride = y.data_ptr(), y.stride(0), y.stride(1)
    a.h_last = None if h_last is None else h_last.data_ptr()
    a.B, a.L, a.Di, a.N, a.R = B, L, Di, d_state, dt_rank
    a.Rp, a.Xp, a.reverse = w_dt.stride(0), xdbl.shape[-1], 1 if reverse else 0
    a.a_geometric, a.tune = 1 if a_geometric else 0, int(tune)
    a.frame_len = int(frame_len) if reverse else 0
    if bwd_ckpt is not None:
        a.bwd_ckpt = bwd_ckpt.data_ptr()
    a.z_gate = 1 if z_gate else 0
    if u.dtype != torch.bfloat16 or z.dtype != u.dtype or xdbl.dtype != u.dtype \
            or w_dt.dtype != u.dtype:
        raise TypeError("the fused scan is a bf16 kernel")
    with _on_device(u):     # the split plan depends on the SM count of the tensors' device
        ws_bytes = lib.vmb_fused_scan_workspace_bytes(B, L, Di, d_state) if allow_split else 0
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=u.device) if ws_bytes > 0 else None
        if ws is not None:
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
        rc = lib.vmb_selective_scan_fused_fwd(C.byref(a), _stream(u))
    _lib.check(rc, "vmb_selective_scan_fused_fwd")
    return (y, h_last) if want_last else y


def scan_bwd_ckpt_bytes(B: int, L: int, Di: int) -> int:
    return int(_lib.load().vmb_scan_bwd_ckpt_bytes(B, L, Di))


def is_geometric(A2: Tensor, rtol: float = 1e-6) -> bool:
    """``A2[d, n] == (n+1) * A2[d, 0]`` for every channel, to ``rtol`` of the exponent: the exact
    S4D-real structure of the reference's initialisation (mamba_simple.py:265-272, ``A = -(1..N)``).
    Holds for fp32 ``A_log`` parameters at their initial value; a model cast to bf16 rounds
    ``log(n)`` and loses it, as trained checkpoints generally do.  One device sync: call it when
    weights are loaded, not per forward."""
    target = torch.arange(1, A2.shape[1] + 1, device=A2.device, dtype=torch.float32) * A2[:, :1].float()
    return bool(torch.all((A2.float() - target).abs() <= rtol * target.abs()))


class MixerWeights:
    """Kernel-ready view of one mixer's parameters (built once per weight version)."""

    __slots__ = ("w_in", "b_in", "w_conv", "b_conv", "w_x", "w_dt", "w_out", "b_out", "A2", "Dskip",
                 "dt_bias", "w_x_pad", "w_dt_pad", "Xp", "Rp", "D", "Di", "N", "R", "W", "dtype",
                 "a_geometric", "raw")

    def __init__(self, in_w, in_b, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dp, out_w, out_b):
        self.dtype = in_w.dtype
        _dt(in_w)
        # the parameters this view was built from, under their state_dict names
        self.raw = {"in_proj.weight": in_w, "in_proj.bias": in_b, "conv1d.weight": conv_w,
                    "conv1d.bias": conv_b, "x_proj.weight": x_w, "dt_proj.weight": dt_w,
                    "dt_proj.bias": dt_b, "A_log": A_log, "D": Dp, "out_proj.weight": out_w,
                    "out_proj.bias": out_b}
        dev = in_w.device
        self.Di, self.W = conv_w.shape[0], conv_w.shape[-1]
        self.D = in_w.shape[1]
        self.N = A_log.shape[1]
        self.R = dt_w.shape[1]
        cast = lambda t: None if t is None else t.detach().to(self.dtype).contiguous()
        self.w_in, self.b_in = cast(in_w), cast(in_b)
        self.w_conv, self.b_conv = cast(conv_w.reshape(self.Di, self.W)), cast(conv_b)
        self.w_x, self.w_dt = cast(x_w), cast(dt_w)
        self.w_out, self.b_out = cast(out_w), cast(out_b)
        A = -torch.exp(A_log.detach().float())                      # mamba_simple.py:341
        self.A2 = (A * LOG2E).contiguous()
        self.Dskip = Dp.detach().float().contiguous()               # mamba_simple.py:429
        self.dt_bias = (dt_b.detach().float().contiguous() if dt_b is not None
                        else torch.zeros(self.Di, device=dev))      # mamba_simple.py:431
        # exact S4D-real structure -> the scan's geometric evaluator (checked once per weight version)
        self.a_geometric = self.dtype == torch.bfloat16 and self.N == 16 and is_geometric(self.A2)
        self.Xp = xdbl_pitch(self.R, self.N)
        self.Rp = (self.R + 15) // 16 * 16
        self.w_x_pad = self.w_dt_pad = None
        if self.dtype == torch.bfloat16:
            X = self.R + 2 * self.N
            self.w_x_pad = torch.zeros(self.Xp, self.Di, dtype=self.dtype, device=dev)
            self.w_x_pad[:X] = self.w_x
            self.w_dt_pad = torch.zeros(self.Di, self.Rp, dtype=self.dtype, device=dev)
            self.w_dt_pad[:, :self.R] = self.w_dt


def mixer_fwd(w: MixerWeights, hidden: Tensor, conv_state: Optional[Tensor] = None,
              ssm_state: Optional[Tensor] = None, want_conv_state: bool = False,
              want_ssm_state: bool = False, reverse: bool = False, path: int = 0,
              scan_tune: int = 0, fuse_conv_xproj: bool = False, frame_len: int = 0,
              gate_in_proj: bool = True):
    """Whole Mamba mixer on token-major ``hidden (B, L, D)``.
    Returns ``(out, new_conv_state | None, last_ssm_state | None)``."""
    _require_cuda(hidden)
    lib = _lib.load()
    if hidden.dtype != w.dtype:
        raise TypeError(f"hidden dtype {hidden.dtype} does not match the mixer weights ({w.dtype})")
    B, L, D = hidden.shape
    if D != w.D:
        raise ValueError(f"hidden size {D} does not match d_model {w.D}")
    if hidden.stride(-1) != 1 or (B > 1 and hidden.stride(0) != L * hidden.stride(1)) \
            or hidden.stride(1) < D:
        hidden = hidden.contiguous()
    dev = hidden.device
    out = torch.empty((B, L, D), dtype=w.dtype, device=dev)
    cs_in = None
    if conv_state is not None:
        if tuple(conv_state.shape) != (B, w.Di, w.W):
            raise ValueError(f"conv_state must have shape {(B, w.Di, w.W)}, "
                             f"got {tuple(conv_state.shape)}")
        cs_in = conv_state.contiguous()
    ss_in = None
    if ssm_state is not None:
        if tuple(ssm_state.shape) != (B, w.Di, w.N):
            raise ValueError(f"ssm_state must have shape {(B, w.Di, w.N)}, "
                             f"got {tuple(ssm_state.shape)}")
        ss_in = ssm_state.contiguous()
    cs_out = ss_out = None
    if want_conv_state:
        cs_dtype = w.dtype if cs_in is None else torch.promote_types(cs_in.dtype, w.dtype)
        cs_out = torch.empty((B, w.Di, w.W), dtype=cs_dtype, device=dev)
    if want_ssm_state:
        ss_out = torch.empty((B, w.Di, w.N), dtype=torch.float32, device=dev)
    if B == 0 or L == 0:
        if L == 0 and B > 0:
            if cs_out is not None:
                cs_out.copy_(cs_in) if cs_in is not None else cs_out.zero_()
            if ss_out is not None:
                ss_out.copy_(ss_in) if ss_in is not None else ss_out.zero_()
        return out, cs_out, ss_out
    dt = _dt(hidden)
    with _on_device(hidden):    # the scan's split plan depends on the SM count of hidden's device
        nbytes = lib.vmb_mixer_workspace_bytes(B, L, D, w.Di, w.N, w.R, dt)
    if nbytes < 0:
        raise RuntimeError("vmb_mixer_workspace_bytes rejected the shape")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    a = MixerArgs()
    a.hidden, a.h_bstride, a.h_tstride = hidden.data_ptr(), L * hidden.stride(1), hidden.stride(1)
    a.out, a.o_bstride, a.o_tstride = out.data_ptr(), out.stride(0), out.stride(1)
    a.w_in, a.b_in = w.w_in.data_ptr(), None if w.b_in is None else w.b_in.data_ptr()
    a.w_conv, a.b_conv = w.w_conv.data_ptr(), None if w.b_conv is None else w.b_conv.data_ptr()
    a.w_x, a.w_dt = w.w_x.data_ptr(), w.w_dt.data_ptr()
    a.w_out, a.b_out = w.w_out.data_ptr(), None if w.b_out is None else w.b_out.data_ptr()
    a.A2, a.Dskip, a.dt_bias = w.A2.data_ptr(), w.Dskip.data_ptr(), w.dt_bias.data_ptr()
    if w.w_x_pad is not None:
        a.w_x_pad, a.w_dt_pad = w.w_x_pad.data_ptr(), w.w_dt_pad.data_ptr()
    a.Xp, a.Rp = w.Xp, w.Rp
    if cs_in is not None:
        a.conv_state_in, a.cs_in_dtype = cs_in.data_ptr(), _dt(cs_in)
    if cs_out is not None:
        a.conv_state_out, a.cs_out_dtype = cs_out.data_ptr(), _dt(cs_out)
    if ss_in is not None:
        a.ssm_state_in, a.ss_in_dtype = ss_in.data_ptr(), _dt(ss_in)
    if ss_out is not None:
        a.ssm_state_out = ss_out.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    a.B, a.L, a.D, a.Di, a.N, a.R, a.W = B, L, D, w.Di, w.N, w.R, w.W
    a.dtype, a.reverse, a.path = dt, 1 if reverse else 0, path
    a.a_geometric, a.scan_tune = 1 if w.a_geometric else 0, int(scan_tune)
    a.fuse_conv_xproj = 1 if fuse_conv_xproj else 0
    a.gate_in_proj = 1 if gate_in_proj else 0
    a.frame_len = int(frame_len) if reverse else 0
    with _on_device(hidden):
        rc = lib.vmb_mixer_fwd(C.byref(a), _stream(hidden))
    _lib.check(rc, "vmb_mixer_fwd")
    return forward_only((out, cs_out, ss_out), hidden, conv_state, ssm_state, *w.raw.values())


def mixer_step(w: MixerWeights, xz: Tensor, conv_state: Tensor, ssm_state: Tensor) -> Tensor:
    """Single-token decode between in_proj and out_proj (reference mamba_simple.py:466-494) as ONE
    kernel: conv update, x_proj, dt_proj, state update, D skip, gate.  ``xz (B, 2Di)``; both states are
    updated in place.  Returns the gated ``y (B, Di)``."""
    _require_cuda(xz)
    lib = _lib.load()
    B = xz.shape[0]
    if xz.shape[1] != 2 * w.Di or xz.dtype != w.dtype:
        raise ValueError("mixer_step: xz must be (B, 2 * d_inner) in the weights' dtype")
    if tuple(conv_state.shape) != (B, w.Di, w.W) or tuple(ssm_state.shape) != (B, w.Di, w.N):
        raise ValueError("mixer_step: state shapes do not match the batch / mixer")
    if not conv_state.is_contiguous() or not ssm_state.is_contiguous():
        raise ValueError("conv_state / ssm_state must be contiguous (they are updated in place)")
    if xz.stride(-1) != 1:
        xz = xz.contiguous()
    y = torch.empty((B, w.Di), dtype=w.dtype, device=xz.device)
    a = StepArgs()
    a.xz, a.xz_bstride = xz.data_ptr(), xz.stride(0)
    a.conv_state, a.cs_dtype = conv_state.data_ptr(), _dt(conv_state)
    a.ssm_state, a.ss_dtype = ssm_state.data_ptr(), _dt(ssm_state)
    a.w_conv, a.b_conv = w.w_conv.data_ptr(), None if w.b_conv is None else w.b_conv.data_ptr()
    a.w_x, a.w_dt = w.w_x.data_ptr(), w.w_dt.data_ptr()
    a.A2, a.Dskip, a.dt_bias = w.A2.data_ptr(), w.Dskip.data_ptr(), w.dt_bias.data_ptr()
    a.y, a.y_bstride = y.data_ptr(), y.stride(0)
    a.B, a.Di, a.N, a.R, a.W, a.dtype = B, w.Di, w.N, w.R, w.W, _dt(xz)
    with _on_device(xz):
        rc = lib.vmb_mixer_step_fwd(C.byref(a), _stream(xz))
    _lib.check(rc, "vmb_mixer_step_fwd")
    return forward_only(y, xz, conv_state, ssm_state, *w.raw.values())


def state_gather(pool: Tensor, index: Tensor) -> Tensor:
    """rows ``pool[index]`` of a resident per-stream state pool (leading dim = stream slot)."""
    _require_cuda(pool)
    lib = _lib.load()
    pool = pool.contiguous()
    index = index.to(device=pool.device, dtype=torch.int32).contiguous()
    n = index.numel()
    row = pool[0].numel() if pool.shape[0] > 0 else 0
    out = torch.empty((n, *pool.shape[1:]), dtype=pool.dtype, device=pool.device)
    with _on_device(pool):
        rc = lib.vmb_state_gather(_p(pool), _p(index), _p(out), n, row, _dt(pool), _stream(pool))
    _lib.check(rc, "vmb_state_gather")
    return out


def state_scatter(pool: Tensor, index: Tensor, batch: Tensor) -> None:
    """``pool[index] = batch`` in place (index entries must be distinct)."""
    _require_cuda(pool)
    lib = _lib.load()
    if not pool.is_contiguous():
        raise ValueError("state pool must be contiguous")
    batch = batch.to(pool.dtype).contiguous()
    index = index.to(device=pool.device, dtype=torch.int32).contiguous()
    n = index.numel()
    row = pool[0].numel() if pool.shape[0] > 0 else 0
    with _on_device(pool):
        rc = lib.vmb_state_scatter(_p(pool), _p(index), _p(batch), n, row, _dt(pool), _stream(pool))
    _lib.check(rc, "vmb_state_scatter")


# ----------------------------------------------------------------------------------------------
# drop-in operator signatures (channel-major, as the reference calls them)
# ----------------------------------------------------------------------------------------------
def causal_conv1d_fn(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, seq_idx=None,
                     initial_states=None, return_final_states: bool = False, final_states_out=None,
                     activation: Optional[str] = None):
    """``causal_conv1d.causal_conv1d_fn`` (reference mamba_simple.py:383-399): x (B, D, L)."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    if seq_idx is not None or initial_states is not None or return_final_states:
        raise NotImplementedError("seq_idx / initial_states are not used by VideoMamba")
    y = causal_conv1d_tokens(x.transpose(1, 2), weight, bias, None, False,
                             silu=activation is not None)
    return y.transpose(1, 2)


def causal_conv1d_update(x: Tensor, conv_state: Tensor, weight: Tensor,
                         bias: Optional[Tensor] = None, activation: Optional[str] = None) -> Tensor:
    """``causal_conv1d.causal_conv1d_update`` (mamba_simple.py:468-474): x (B, D), state in place."""
    _require_cuda(x)
    lib = _lib.load()
    if not conv_state.is_contiguous():
        raise ValueError("conv_state must be contiguous (it is updated in place)")
    B, Di = x.shape
    W = weight.shape[-1]
    if x.stride(-1) != 1:
        x = x.contiguous()
    weight = weight.reshape(Di, W).to(x.dtype).contiguous()
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    y = torch.empty((B, Di), dtype=x.dtype, device=x.device)
    with _on_device(x):
        rc = lib.vmb_causal_conv1d_update(_p(x), x.stride(0), _p(conv_state), _dt(conv_state),
                                          _p(weight), _p(bias), _p(y), y.stride(0), B, Di, W,
                                          1 if activation in ("silu", "swish") else 0, _dt(x),
                                          _stream(x))
    _lib.check(rc, "vmb_causal_conv1d_update")
    return y


def selective_scan_fn(u: Tensor, delta: Tensor, A: Tensor, B: Tensor, C_: Tensor,
                      D: Optional[Tensor] = None, z: Optional[Tensor] = None,
                      delta_bias: Optional[Tensor] = None, delta_softplus: bool = False,
                      return_last_state: bool = False, initial_state: Optional[Tensor] = None):
    """``mamba_ssm...selective_scan_fn`` with ``initial_state`` support (mamba_simple.py:125-152).
    u, delta, z: (B, D, L); A: (D, N) real; B, C: (B, N, L)."""
    if A.is_complex() or B.dim() != 3 or C_.dim() != 3:
        raise NotImplementedError("only real A and (B, N, L) input-dependent B / C are supported")
    N = A.shape[1]
    bc = torch.cat([B.transpose(1, 2), C_.transpose(1, 2)], dim=-1).to(u.dtype).contiguous()
    if _wants_grad(u, delta, A, B, C_, D, z, delta_bias, initial_state):
        from . import autograd as ag
        y, last = ag.ScanFn.apply(u.transpose(1, 2), delta.to(u.dtype).transpose(1, 2), A, bc, 0, N, N, D,
                                  None if z is None else z.to(u.dtype).transpose(1, 2), delta_bias,
                                  delta_softplus, initial_state, return_last_state)
        return (y.transpose(1, 2), last) if return_last_state else y.transpose(1, 2)
    A2 = (A.float() * LOG2E).contiguous()
    out = selective_scan_tokens(
        u.transpose(1, 2), delta.to(u.dtype).transpose(1, 2), A2, bc, 0, N, N,
        None if D is None else D.float().contiguous(),
        None if z is None else z.to(u.dtype).transpose(1, 2),
        None if delta_bias is None else delta_bias.float().contiguous(),
        delta_softplus, initial_state, return_last_state)
    if return_last_state:
        return out[0].transpose(1, 2), out[1]
    return out.transpose(1, 2)


def selective_state_update(state: Tensor, x: Tensor, dt: Tensor, A: Tensor, B: Tensor, C_: Tensor,
                           D: Optional[Tensor] = None, z: Optional[Tensor] = None,
                           dt_bias: Optional[Tensor] = None, dt_softplus: bool = False) -> Tensor:
    """``mamba_ssm...selective_state_update`` (mamba_simple.py:483-494): state (B, D, N) in place."""
    _require_cuda(x)
    lib = _lib.load()
    if not state.is_contiguous():
        raise ValueError("state must be contiguous (it is updated in place)")
    Bsz, Di = x.shape
    N = A.shape[1]
    rowmajor = lambda t: t if t.stride(-1) == 1 else t.contiguous()
    x, dt = rowmajor(x), rowmajor(dt.to(x.dtype))
    B, C_ = rowmajor(B.to(x.dtype)), rowmajor(C_.to(x.dtype))
    if z is not None:
        z = rowmajor(z.to(x.dtype))
    A2 = (A.float() * LOG2E).contiguous()
    Df = None if D is None else D.float().contiguous()
    bias = None if dt_bias is None else dt_bias.float().contiguous()
    y = torch.empty((Bsz, Di), dtype=x.dtype, device=x.device)
    with _on_device(x):
        rc = lib.vmb_selective_state_update(
            _p(state), _dt(state), _p(x), x.stride(0), _p(dt), dt.stride(0), _p(A2),
            _p(B), B.stride(0), _p(C_), C_.stride(0), _p(Df), _p(z),
            0 if z is None else z.stride(0), _p(bias), 1 if dt_softplus else 0,
            _p(y), y.stride(0), Bsz, Di, N, _dt(x), _stream(x))
    _lib.check(rc, "vmb_selective_state_update")
    return y


def _norm_fn(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms):
    return add_norm(x, weight, bias, residual, eps, is_rms, prenorm, residual_in_fp32)


def rms_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False,
                **_unused):
    """``mamba_ssm.ops.triton.layer_norm.rms_norm_fn`` (videomamba.py:157-165, :909-917)."""
    return _norm_fn(x, weight, bias, residual, eps, prenorm, residual_in_fp32, True)


def layer_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False,
                  **_unused):
    """``mamba_ssm.ops.triton.layer_norm.layer_norm_fn``."""
    return _norm_fn(x, weight, bias, residual, eps, prenorm, residual_in_fp32, False)


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                   out_proj_weight, out_proj_bias, A, B=None, C_=None, D=None, delta_bias=None,
                   delta_softplus=True):
    """``mamba_ssm...mamba_inner_fn`` (mamba_simple.py:352-366): xz (B, 2Di, L) -> (B, L, D).
    conv -> x_proj -> dt_proj -> scan -> out_proj on an already in-projected ``xz``."""
    if B is not None or C_ is not None or not delta_softplus:
        raise NotImplementedError("only input-dependent B / C with softplus are supported")
    Di = xz.shape[1] // 2
    N, R = A.shape[1], delta_proj_weight.shape[1]
    xz_t = xz.transpose(1, 2)
    if _wants_grad(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                   out_proj_bias, A, D, delta_bias):
        from . import autograd as ag
        xc, _ = ag.ConvFn.apply(xz_t[..., :Di], conv1d_weight, conv1d_bias, None, False, True)
        x_dbl = linear(xc, x_proj_weight)
        delta = linear(x_dbl[..., :R], delta_proj_weight)
        y, _ = ag.ScanFn.apply(xc, delta, A, x_dbl, R, R + N, N, D, xz_t[..., Di:], delta_bias, True,
                               None, False)
        return linear(y, out_proj_weight, out_proj_bias)
    xc = causal_conv1d_tokens(xz_t[..., :Di], conv1d_weight, conv1d_bias)
    x_dbl = linear(xc, x_proj_weight)
    delta = linear(x_dbl[..., :R], delta_proj_weight)
    y = selective_scan_tokens(xc, delta, (A.float() * LOG2E).contiguous(), x_dbl, R, R + N, N,
                              None if D is None else D.float().contiguous(), xz_t[..., Di:],
                              None if delta_bias is None else delta_bias.float().contiguous(), True)
    return linear(y, out_proj_weight, out_proj_bias)
