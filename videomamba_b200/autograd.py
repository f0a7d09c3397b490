"""Differentiable operators: autograd nodes over the libvmb200 forward AND backward kernels.

The reference trains through the third-party operators' own autograd (streaming-training check
``scripts/check_streaming_state.py:47-60``: gradients flow through the carried
``(conv_state, ssm_state)``; ``Block.forward`` wraps the mixer in activation checkpointing,
``models/videomamba/videomamba.py:168-206``).  Here every operator of the block has an
``autograd.Function`` whose backward is a library kernel (``csrc/backward.cu``, ``csrc/scan_bwd.cu``)
or, for the projections, two calls of the forward GEMM on transposed operands.  The training-mode
mixer (``mixer_train``) is the reference's slow path op for op (mamba_simple.py:333-446), so the
rounding points are the reference's and torch's checkpoint wrapper can recompute it.

Not differentiable: the fused inference entry points (``vmb_mixer_fwd``, the fused scan, the decode
step) and the reversed token walks -- they stay tagged forward-only (``ops.forward_only``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib, ops
from ._lib import VMB_F32, ScanBwdArgs
from .ops import LOG2E, _dt, _on_device, _p, _stream


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, Tensor) and t.requires_grad for t in tensors)


def _ws(nbytes: int, like: Tensor) -> Optional[Tensor]:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=like.device)


def transpose2d(x2: Tensor, pad_to: int = 8) -> Tensor:
    """``x2 (rows, cols)`` (row stride >= cols) -> its transpose ``(cols, rows)`` as a view of a buffer
    whose row pitch is a multiple of ``pad_to`` elements (zero tail), the pitch the GEMM's TMA needs."""
    lib = _lib.load()
    rows, cols = x2.shape
    pitch = (rows + pad_to - 1) // pad_to * pad_to
    out = torch.zeros((cols, pitch), dtype=x2.dtype, device=x2.device) if pitch != rows else \
        torch.empty((cols, pitch), dtype=x2.dtype, device=x2.device)
    with _on_device(x2):
        rc = lib.vmb_transpose_2d(_p(x2), x2.stride(0) if rows > 1 else cols, _p(out), pitch, rows, cols,
                                  _dt(x2), _stream(x2))
    _lib.check(rc, "vmb_transpose_2d")
    return out            # (cols, pitch): columns >= rows are zero


def colsum(x2: Tensor, out_dtype: torch.dtype) -> Tensor:
    lib = _lib.load()
    M, N = x2.shape
    out = torch.empty(N, dtype=torch.float32, device=x2.device)
    with _on_device(x2):
        nbytes = lib.vmb_colsum_workspace_bytes(M, N)
    ws = _ws(nbytes, x2)
    with _on_device(x2):
        rc = lib.vmb_colsum(_p(x2), x2.stride(0) if M > 1 else N, M, N, _dt(x2), _p(out), VMB_F32, _p(ws),
                            nbytes, _stream(x2))
    _lib.check(rc, "vmb_colsum")
    return out.to(out_dtype)


def linear_wgrad(dy2: Tensor, x2: Tensor, out_dtype: torch.dtype) -> Tensor:
    """``dy2^T @ x2`` -> ``(N, K)``: the weight gradient of ``y = x W^T`` summed over the M rows.  bf16
    operands with aligned rows run the split-token tensor-core kernel (``vmb_linear_wgrad``, no transposed
    copies); everything else goes through two transposes and the forward projection."""
    lib = _lib.load()
    M, N = dy2.shape
    K = x2.shape[1]
    lds = (dy2.stride(0) if M > 1 else N, x2.stride(0) if M > 1 else K)
    direct = (dy2.dtype == torch.bfloat16 and x2.dtype == torch.bfloat16 and M > 0
              and N % 8 == 0 and K % 8 == 0 and lds[0] % 8 == 0 and lds[1] % 8 == 0
              and dy2.data_ptr() % 16 == 0 and x2.data_ptr() % 16 == 0
              and out_dtype in (torch.bfloat16, torch.float32))
    if direct:
        dw = torch.empty((N, K), dtype=out_dtype, device=dy2.device)
        with _on_device(dy2):          # the split over the tokens follows the SM count of the tensors' device
            nbytes = lib.vmb_linear_wgrad_workspace_bytes(M, N, K)
        ws = _ws(nbytes, dy2)
        with _on_device(dy2):
            rc = lib.vmb_linear_wgrad(_p(dy2), lds[0], _p(x2), lds[1], _p(dw), _dt(dw), M, N, K, _p(ws), nbytes,
                                      _stream(dy2))
        _lib.check(rc, "vmb_linear_wgrad")
        return dw
    if M == 0:
        return torch.zeros((N, K), dtype=out_dtype, device=dy2.device)
    return ops.linear_raw(transpose2d(dy2), transpose2d(x2)).to(out_dtype)   # zero tails: exact


def _rows(t: Tensor, width: int) -> Tensor:
    t2 = t.reshape(-1, width)
    if t2.stride(-1) != 1 or (t2.shape[0] > 1 and t2.stride(0) < width):
        t2 = t2.contiguous()
    return t2


class LinearFn(torch.autograd.Function):
    """``x @ weight.T (+ bias)``; backward: dX = dY W, dW = dY^T X (both on the forward GEMM), db = colsum(dY)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = None if bias is None else bias.dtype
        return ops.linear_raw(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        N, K = weight.shape
        dy2 = _rows(dy.to(x.dtype), N)
        x2 = _rows(x, K)
        M = x2.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            w_t = transpose2d(_rows(weight.to(x.dtype), K))[:, :N]             # (K, N)
            dx = ops.linear_raw(dy2, w_t).reshape(x.shape)
        if ctx.needs_input_grad[1]:
            dw = linear_wgrad(dy2, x2, weight.dtype)                           # (N, K): sum over the M rows
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2, ctx.bias_dtype)
        return dx, dw, db


class AddNormFn(torch.autograd.Function):
    """Fused residual add + RMSNorm / LayerNorm (``vmb_add_norm_fwd`` / ``vmb_add_norm_bwd``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, eps, is_rms, prenorm, residual_in_fp32):
        ctx.set_materialize_grads(False)
        y, res_out, x2, res2 = ops.add_norm_raw(x, weight, bias, residual, eps, is_rms, prenorm,
                                                residual_in_fp32)
        ctx.save_for_backward(x2, res2, weight)
        ctx.meta = (eps, is_rms, x.shape, None if residual is None else residual.dtype,
                    None if bias is None else bias.dtype)
        if not prenorm:
            return y
        # (ops.add_norm never asks for prenorm when the residual stream would be x itself)
        return y, res_out

    @staticmethod
    def backward(ctx, dy, dres_out=None):
        lib = _lib.load()
        x2, res2, weight = ctx.saved_tensors
        eps, is_rms, shape, res_dtype, bias_dtype = ctx.meta
        rows, dim = x2.shape
        if dy is None:
            dy2 = torch.zeros_like(x2)
        else:
            dy2 = dy.reshape(rows, dim).to(x2.dtype).contiguous()
        dro = None
        if dres_out is not None:
            dro = dres_out.reshape(rows, dim).contiguous()
            if dro.dtype not in (torch.float32, torch.bfloat16):
                dro = dro.float()
        dx = torch.empty((rows, dim), dtype=x2.dtype, device=x2.device)
        dres = torch.empty((rows, dim), dtype=res2.dtype, device=x2.device) if res2 is not None else None
        dw = torch.empty(dim, dtype=torch.float32, device=x2.device)
        db = torch.empty(dim, dtype=torch.float32, device=x2.device) if bias_dtype is not None else None
        with _on_device(x2):
            nbytes = lib.vmb_add_norm_bwd_workspace_bytes(rows, dim)
        ws = _ws(nbytes, x2)
        w = weight.contiguous()
        with _on_device(x2):
            rc = lib.vmb_add_norm_bwd(
                _p(x2), _dt(x2), x2.stride(0) if rows > 1 else dim,
                _p(res2), _dt(res2) if res2 is not None else VMB_F32,
                _p(w), _dt(w), _p(dy2), _p(dro), _dt(dro) if dro is not None else VMB_F32,
                _p(dx), _p(dres), _p(dw), _p(db), rows, dim, float(eps), 1 if is_rms else 0,
                _p(ws), nbytes, _stream(x2))
        _lib.check(rc, "vmb_add_norm_bwd")
        return (dx.reshape(shape), dw.to(weight.dtype), None if db is None else db.to(bias_dtype),
                None if dres is None else dres.reshape(shape), None, None, None, None)


# The fused training forward stores the scan state before every 4-token group for the backward (12 KB per
# token and layer at Di 768: 1.2 GB per layer at batch 32 x 3137 tokens).  False: the backward recomputes them
# with a forward pass of its own (+0.45 ms per layer at that size, no extra memory between the passes).
SAVE_SCAN_STATES = True


class XZGrad:
    """The d(xz) buffer of one mixer call, shared by the backward nodes of its two halves: the conv backward
    writes dx into ``[..., :Di]`` and the scan backward dz into ``[..., Di:]`` (both kernels take strided
    outputs), so ``SplitXZ.backward`` returns the buffer as it is instead of copying two (B, L, Di) tensors
    side by side.  Anything else that arrives (a gradient torch accumulated, a missing half) is copied."""

    __slots__ = ("buf", "di")

    def __init__(self, di: int):
        self.buf, self.di = None, di

    def half(self, which: int, B: int, L: int, dtype, device) -> Tensor:
        if self.buf is None or self.buf.shape != (B, L, 2 * self.di) or self.buf.dtype != dtype:
            self.buf = torch.empty((B, L, 2 * self.di), dtype=dtype, device=device)
        return self.buf[..., :self.di] if which == 0 else self.buf[..., self.di:]

    def owns(self, which: int, grad: Tensor) -> bool:
        b = self.buf
        return (b is not None and grad is not None and grad.dtype == b.dtype and grad.shape == (*b.shape[:2], self.di)
                and grad.stride() == b.stride()
                and grad.data_ptr() == b.data_ptr() + which * self.di * b.element_size())


class SplitXZ(torch.autograd.Function):
    """``xz (.., 2 Di) -> (x, z)`` as two strided views (mamba_simple.py:369).  torch's own slice backward
    materialises one zero-filled (.., 2 Di) tensor per half and adds them (five passes over the largest
    activation of the block); here the two incoming gradients are written side by side into one buffer."""

    @staticmethod
    def forward(ctx, xz, di, arena=None):
        ctx.set_materialize_grads(False)
        ctx.di = di
        ctx.arena = arena
        ctx.meta = (xz.shape, xz.dtype, xz.device)
        return xz[..., :di], xz[..., di:]

    @staticmethod
    def backward(ctx, dx, dz):
        shape, dtype, dev = ctx.meta
        di = ctx.di
        arena = ctx.arena
        if arena is not None:
            buf = arena.buf
            arena.buf = None                                  # one backward pass owns it
            if buf is not None and buf.shape == shape:
                arena.buf = buf
                ok = arena.owns(0, dx) and arena.owns(1, dz)
                arena.buf = None
                if ok:
                    return buf, None, None
        out = torch.empty(shape, dtype=dtype, device=dev)
        for part, grad in ((out[..., :di], dx), (out[..., di:], dz)):
            if grad is None:
                part.zero_()
            else:
                part.copy_(grad)
        return out, None, None


class ConvFn(torch.autograd.Function):
    """Depthwise causal conv + SiLU with streaming history, token-major (``vmb_causal_conv1d_fwd`` /
    ``_bwd``).  Gradients flow into ``conv_state`` and arrive through the returned state."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv_state, want_state, silu, arena=None):
        ctx.set_materialize_grads(False)
        out = ops.causal_conv1d_tokens_raw(x, weight, bias, conv_state, want_state, silu)
        ctx.save_for_backward(x, weight, bias, conv_state)
        ctx.silu = silu
        ctx.arena = arena
        if want_state:
            return out
        return out, None

    @staticmethod
    def backward(ctx, dy, dcs_out=None):
        lib = _lib.load()
        x, weight, bias, conv_state = ctx.saved_tensors
        x = ops._token_major(x)
        B, L, Di = x.shape
        W = weight.shape[-1]
        w2 = weight.reshape(Di, W).to(x.dtype).contiguous()
        b2 = None if bias is None else bias.to(x.dtype).contiguous()
        cs = None if conv_state is None else conv_state.contiguous()
        dy = torch.zeros_like(x, memory_format=torch.contiguous_format) if dy is None \
            else dy.to(x.dtype).contiguous()
        if dcs_out is not None:
            dcs_out = dcs_out.contiguous()
            if dcs_out.dtype not in (torch.float32, torch.bfloat16):
                dcs_out = dcs_out.float()
        if ctx.arena is not None and ctx.arena.di == Di:
            dx = ctx.arena.half(0, B, L, x.dtype, x.device)   # the x half of d(xz), written in place
        else:
            dx = torch.empty((B, L, Di), dtype=x.dtype, device=x.device)
        dcs_in = torch.empty_like(cs) if cs is not None and ctx.needs_input_grad[3] else None
        dw = torch.empty((Di, W), dtype=torch.float32, device=x.device)
        db = torch.empty(Di, dtype=torch.float32, device=x.device) if bias is not None else None
        with _on_device(x):
            nbytes = lib.vmb_causal_conv1d_bwd_workspace_bytes(B, L, Di, W)
        ws = _ws(nbytes, x)
        with _on_device(x):
            rc = lib.vmb_causal_conv1d_bwd(
                _p(x), x.stride(0), x.stride(1), _p(w2), _p(b2), _p(cs),
                _dt(cs) if cs is not None else VMB_F32, _p(dy), _p(dcs_out),
                _dt(dcs_out) if dcs_out is not None else VMB_F32, _p(dx), dx.stride(0), dx.stride(1), _p(dcs_in),
                _p(dw), _p(db),
                B, L, Di, W, 1 if ctx.silu else 0, _dt(x), _p(ws), nbytes, _stream(x))
        _lib.check(rc, "vmb_causal_conv1d_bwd")
        return (dx, dw.reshape(weight.shape).to(weight.dtype),
                None if db is None else db.to(bias.dtype), dcs_in, None, None, None)


class ScanFn(torch.autograd.Function):
    """Selective scan, token-major, explicit ``delta_raw`` (``vmb_selective_scan_fwd`` / ``_bwd``).
    ``A`` is the natural ``A = -exp(A_log)`` (fp32); B_t / C_t are columns of ``bc``."""

    @staticmethod
    def forward(ctx, u, delta, A, bc, b_off, c_off, d_state, D, z, dt_bias, softplus, h0, want_last, arena=None):
        ctx.set_materialize_grads(False)
        ctx.arena = arena
        A2 = (A.float() * LOG2E).contiguous()
        Df = None if D is None else D.float().contiguous()
        bias = None if dt_bias is None else dt_bias.float().contiguous()
        u, delta, bc = ops._token_major(u), ops._token_major(delta), ops._token_major(bc)
        z = None if z is None else ops._token_major(z)
        out = ops.selective_scan_tokens_raw(u, delta, A2, bc, b_off, c_off, d_state, Df, z, bias,
                                            softplus, h0, want_last)
        ctx.save_for_backward(u, delta, A2, bc, Df, z, bias, h0)
        ctx.meta = (b_off, c_off, d_state, softplus, A.dtype, None if D is None else D.dtype,
                    None if dt_bias is None else dt_bias.dtype)
        if want_last:
            return out
        return out, None

    @staticmethod
    def backward(ctx, dout, dh_last=None):
        u, delta, A2, bc, Df, z, bias, h0 = ctx.saved_tensors
        b_off, c_off, N, softplus, A_dtype, D_dtype, bias_dtype = ctx.meta
        du, dd, dz, dbc, dA, dD, dbias, dh0 = _scan_bwd(u, delta, A2, bc, b_off, c_off, N, Df, z, bias,
                                                         softplus, h0, dout, dh_last,
                                                         ctx.needs_input_grad[11], ctx.arena)
        return (du, dd, dA.to(A_dtype), dbc, None, None, None,
                None if dD is None else dD.to(D_dtype), dz,
                None if dbias is None else dbias.to(bias_dtype), None, dh0, None, None)


def _scan_bwd(u, delta, A2, bc, b_off, c_off, N, Df, z, bias, softplus, h0, dout, dh_last, want_dh0, arena=None,
              fwd_ckpt=None):
    """One ``vmb_selective_scan_bwd`` call.  Returns (du, ddelta_raw, dz, dbc, dA, dD, dbias, dh0);
    ``dbc`` is zero outside the B / C columns.  With an ``XZGrad`` arena dz is its z half (written in place)."""
    lib = _lib.load()
    B, L, Di = u.shape
    dev = u.device
    dout = torch.zeros_like(u, memory_format=torch.contiguous_format) if dout is None \
        else ops._token_major(dout.to(u.dtype))
    du = torch.empty((B, L, Di), dtype=u.dtype, device=dev)
    dd = torch.empty((B, L, Di), dtype=u.dtype, device=dev)
    if z is None:
        dz = None
    elif arena is not None and arena.di == Di:
        dz = arena.half(1, B, L, u.dtype, dev)
    else:
        dz = torch.empty((B, L, Di), dtype=u.dtype, device=dev)
    dbc = torch.zeros(bc.shape, dtype=bc.dtype, device=dev)
    dA = torch.empty((Di, N), dtype=torch.float32, device=dev)
    dD = torch.empty(Di, dtype=torch.float32, device=dev) if Df is not None else None
    dbias = torch.empty(Di, dtype=torch.float32, device=dev) if bias is not None else None
    dh0 = torch.empty((B, Di, N), dtype=torch.float32, device=dev) if h0 is not None and want_dh0 else None
    if dh_last is not None:
        dh_last = dh_last.float().contiguous()
    h0c = None if h0 is None else h0.contiguous()
    with _on_device(u):                # the segment plan follows the SM count of the tensors' device
        nbytes = lib.vmb_selective_scan_bwd_workspace_bytes(B, L, Di, N)
    ws = _ws(nbytes, u)
    a = ScanBwdArgs()
    a.u, a.u_bstride, a.u_tstride = u.data_ptr(), u.stride(0), u.stride(1)
    a.delta, a.d_bstride, a.d_tstride = delta.data_ptr(), delta.stride(0), delta.stride(1)
    if z is not None:
        a.z, a.z_bstride, a.z_tstride = z.data_ptr(), z.stride(0), z.stride(1)
    a.bc, a.bc_bstride, a.bc_tstride = bc.data_ptr(), bc.stride(0), bc.stride(1)
    a.b_off, a.c_off = b_off, c_off
    a.A2 = A2.data_ptr()
    a.D = None if Df is None else Df.data_ptr()
    a.dt_bias = None if bias is None else bias.data_ptr()
    if h0c is not None:
        a.h0, a.h0_dtype = h0c.data_ptr(), _dt(h0c)
    a.dout, a.dout_bstride, a.dout_tstride = dout.data_ptr(), dout.stride(0), dout.stride(1)
    a.dh_last = None if dh_last is None else dh_last.data_ptr()
    a.du, a.ddelta = du.data_ptr(), dd.data_ptr()
    if dz is not None:
        a.dz, a.dz_bstride, a.dz_tstride = dz.data_ptr(), dz.stride(0), dz.stride(1)
    a.dbc, a.dbc_tstride = dbc.data_ptr(), dbc.stride(1)
    a.dA = dA.data_ptr()
    a.dD = None if dD is None else dD.data_ptr()
    a.ddt_bias = None if dbias is None else dbias.data_ptr()
    a.dh0 = None if dh0 is None else dh0.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    a.B, a.L, a.Di, a.N = B, L, Di, N
    a.dtype, a.softplus = _dt(u), 1 if softplus else 0
    if fwd_ckpt is not None:
        a.fwd_ckpt = fwd_ckpt.data_ptr()
    with _on_device(u):
        rc = lib.vmb_selective_scan_bwd(C.byref(a), _stream(u))
    _lib.check(rc, "vmb_selective_scan_bwd")
    if dh0 is not None and h0.dtype != torch.float32:
        dh0 = dh0.to(h0.dtype)
    return du, dd, dz, dbc, dA, dD, dbias, dh0


class FusedScanFn(torch.autograd.Function):
    """dt_proj + softplus + scan + D skip + SiLU(z) gate on the FUSED inference kernel (bf16, d_state 16:
    delta is never materialised in the forward).  The backward recomputes ``delta_raw`` with one projection
    (rounded to bf16 where the reference rounds it, mamba_simple.py:413-414), runs the scan backward kernel
    and folds the dt_proj backward in (``d dt_low = d delta W_dt``, ``d W_dt = d delta^T dt_low``)."""

    @staticmethod
    def forward(ctx, u, z, xdbl, w_dt, A, D, dt_bias, h0, want_last, R, N, arena=None):
        ctx.set_materialize_grads(False)
        ctx.arena = arena
        A2 = (A.float() * LOG2E).contiguous()
        Df = None if D is None else D.float().contiguous()
        bias = None if dt_bias is None else dt_bias.float().contiguous()
        u, z, xdbl = ops._token_major(u), ops._token_major(z), ops._token_major(xdbl)
        w_dt = w_dt.contiguous()
        # the forward also writes the state before every 4-token group (1 KB per 16 channels): the backward
        # then needs no forward pass of its own over the sequence
        ckpt = None
        if SAVE_SCAN_STATES:
            B, L, Di = u.shape
            try:
                ckpt = torch.empty(ops.scan_bwd_ckpt_bytes(B, L, Di), dtype=torch.uint8, device=u.device)
            except torch.cuda.OutOfMemoryError:          # no room for the records: the backward recomputes them
                ckpt = None
        out = ops.selective_scan_fused_tokens_raw(u, z, xdbl, w_dt, A2, R, N, Df, bias, h0, want_last, bwd_ckpt=ckpt)
        ctx.save_for_backward(u, z, xdbl, w_dt, A2, Df, bias, h0, ckpt)
        ctx.meta = (R, N, A.dtype, None if D is None else D.dtype, None if dt_bias is None else dt_bias.dtype)
        if want_last:
            return out
        return out, None

    @staticmethod
    def backward(ctx, dout, dh_last=None):
        u, z, xdbl, w_dt, A2, Df, bias, h0, ckpt = ctx.saved_tensors
        R, N, A_dtype, D_dtype, bias_dtype = ctx.meta
        B, L, Di = u.shape
        # dt_rank 12 / 36 (Tiny, Middle) is not a multiple of 8, which the tensor-core kernels need for their
        # 16-byte rows: run the three dt projections on R8 = round_up(R, 8) columns -- the x_dbl columns
        # [R, R8) belong to B_t, so the weight gets zero columns there (exact) and the weight gradient's extra
        # columns are dropped
        R8 = (R + 7) // 8 * 8
        w8 = w_dt if R8 == R else torch.nn.functional.pad(w_dt, (0, R8 - R))    # (Di, R8)
        dt_low8 = xdbl[..., :R8]
        delta = ops.linear_raw(dt_low8, w8)                                    # (B, L, Di), bf16
        du, dd, dz, dbc, dA, dD, dbias, dh0 = _scan_bwd(u, delta, A2, xdbl, R, R + N, N, Df, z, bias, True, h0,
                                                         dout, dh_last, ctx.needs_input_grad[7], ctx.arena, ckpt)
        dd2 = dd.reshape(B * L, Di)
        w_t = transpose2d(w8)[:, :Di]                                          # (R8, Di), rows >= R zero
        dbc[..., :R] = ops.linear_raw(dd2, w_t).reshape(B, L, R8)[..., :R]     # d dt_low
        dw_dt = None
        if ctx.needs_input_grad[3]:
            dw_dt = linear_wgrad(dd2, _rows(dt_low8, R8), w_dt.dtype)[:, :R]   # (Di, R)
        return (du, dz, dbc, dw_dt, dA.to(A_dtype), None if dD is None else dD.to(D_dtype),
                None if dbias is None else dbias.to(bias_dtype), dh0, None, None, None, None)


def fused_scan_covers(dtype, Di: int, N: int, R: int) -> bool:
    """Shapes of the fused bf16 scan kernel (csrc/scan_fast.cu: scan_fast_supported)."""
    return dtype == torch.bfloat16 and N == 16 and R in (12, 24, 36) and Di % 16 == 0


# ----------------------------------------------------------------------------------------------
# training-mode mixer: the reference's slow path, op for op (mamba_simple.py:333-446)
# ----------------------------------------------------------------------------------------------
def mixer_train(in_w, in_b, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dp, out_w, out_b,
                hidden: Tensor, conv_state: Optional[Tensor] = None, ssm_state: Optional[Tensor] = None,
                want_conv_state: bool = False, want_ssm_state: bool = False):
    """Returns ``(out, new_conv_state | None, last_ssm_state | None)``; every step is differentiable,
    including the two state inputs and the two state outputs."""
    Di = conv_w.shape[0]
    N, R = A_log.shape[1], dt_w.shape[1]
    xz = ops.linear(hidden, in_w, in_b)                                  # :333-339
    arena = XZGrad(Di)                                                   # dx / dz are written into one d(xz) buffer
    x_in, z = SplitXZ.apply(xz, Di, arena)                               # :369
    xc, new_conv = ConvFn.apply(x_in, conv_w, conv_b, conv_state, want_conv_state, True, arena)   # :381-404
    A = -torch.exp(A_log.float())                                        # :341
    if fused_scan_covers(hidden.dtype, Di, N, R) and x_w.dtype == hidden.dtype and dt_w.dtype == hidden.dtype:
        # production bf16 shapes: the forward runs the fused inference scan (dt_proj inside, delta never in HBM)
        Xp = ops.xdbl_pitch(R, N)
        x_dbl = ops.linear(xc, torch.nn.functional.pad(x_w, (0, 0, 0, Xp - x_w.shape[0])))   # rows [dt_low | B | C | 0]
        y, last = FusedScanFn.apply(xc, z, x_dbl, dt_w, A, Dp, dt_b, ssm_state, want_ssm_state, R, N, arena)
    else:
        x_dbl = ops.linear(xc, x_w)                                      # :409
        delta = ops.linear(x_dbl[..., :R], dt_w)                         # :413-414 (rounded to the model dtype)
        y, last = ScanFn.apply(xc, delta, A, x_dbl, R, R + N, N, Dp.float(), z,
                               None if dt_b is None else dt_b.float(), True, ssm_state, want_ssm_state, arena)
    out = ops.linear(y, out_w, out_b)                                    # :445-446
    return out, new_conv, last
