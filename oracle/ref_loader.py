"""TEST INFRASTRUCTURE ONLY -- import the LIVE reference host code from /root/reference.

Used in this container only (the GPU box has no /root/reference): by tools/make_golden.py to
generate tests/golden/*, and by tests/test_oracle_vs_reference.py to pin oracle/ against the
reference's own Python.  Nothing is copied: the reference modules are imported from where
they lie.  Two things are needed to run them here:

1. The un-vendored leaves (``causal_conv1d``, ``mamba_ssm.ops.*``, ``timm.*``; reference
   pyproject.toml:13-17) are absent, so stand-ins are injected into ``sys.modules``.  The scan
   stand-in delegates to the reference's OWN in-tree ``_selective_scan_ref``
   (models/videomamba/mamba_simple.py:30-106); conv / norm / state-update stand-ins restate
   the wheels' published ``*_ref`` functions in plain torch.
2. The reference refuses CPU tensors (mamba_simple.py:304-308, :456-460).  The two guards are
   neutralised IN MEMORY when the module source is compiled (an import hook rewrites
   ``if not hidden_states.is_cuda:`` to ``if False:``); no file is written.
"""
from __future__ import annotations

import collections.abc
import importlib.abc
import importlib.util
import os
import sys
import types
from itertools import repeat

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("VIDEOMAMBA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "videomamba", "mamba_simple.py"))


def _module(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = []  # behave as a package so submodule imports resolve
    sys.modules[name] = mod
    return mod


def _install_leaf_stubs() -> None:
    # ---- causal_conv1d (published causal_conv1d_ref / causal_conv1d_update_ref)
    cc = _module("causal_conv1d")

    def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, initial_states=None,
                         return_final_states=False, final_states_out=None, activation=None):
        width = weight.shape[1]
        y = F.conv1d(x.float(), weight.float()[:, None, :], None if bias is None else bias.float(),
                     padding=width - 1, groups=weight.shape[0])[..., : x.shape[-1]]
        if activation in ("silu", "swish"):
            y = F.silu(y)
        return y.to(x.dtype)

    def causal_conv1d_update(x, conv_state, weight, bias=None, activation=None):
        conv_state.copy_(torch.roll(conv_state, -1, -1))
        conv_state[..., -1] = x
        y = (conv_state.float() * weight.float()).sum(-1)
        if bias is not None:
            y = y + bias.float()
        if activation in ("silu", "swish"):
            y = F.silu(y)
        return y.to(x.dtype)

    cc.causal_conv1d_fn = causal_conv1d_fn
    cc.causal_conv1d_update = causal_conv1d_update

    # ---- mamba_ssm
    _module("mamba_ssm")
    _module("mamba_ssm.ops")
    _module("mamba_ssm.ops.triton")
    ssi = _module("mamba_ssm.ops.selective_scan_interface")

    def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                          delta_softplus=False, return_last_state=False, initial_state=None):
        ref_scan = sys.modules["models.videomamba.mamba_simple"]._selective_scan_ref
        return ref_scan(u, delta, A, B, C, D, z, delta_bias, delta_softplus, initial_state,
                        return_last_state)

    def mamba_inner_fn(*args, **kwargs):
        raise NotImplementedError("reference fast path is not available without mamba_ssm")

    ssi.selective_scan_fn = selective_scan_fn
    ssi.mamba_inner_fn = mamba_inner_fn

    ssu = _module("mamba_ssm.ops.triton.selective_state_update")

    def selective_state_update(state, x, dt, A, B, C, D=None, z=None, dt_bias=None,
                               dt_softplus=False):
        dt = dt.float() if dt_bias is None else dt.float() + dt_bias.float()
        if dt_softplus:
            dt = F.softplus(dt)
        nxt = state.float() * torch.exp(dt[..., None] * A.float()) \
            + dt[..., None] * B.float()[:, None] * x.float()[..., None]
        state.copy_(nxt)
        out = torch.einsum("bdn,bn->bd", nxt, C.float())
        if D is not None:
            out = out + x.float() * D.float()
        if z is not None:
            out = out * F.silu(z.float())
        return out.to(x.dtype)

    ssu.selective_state_update = selective_state_update

    ln = _module("mamba_ssm.ops.triton.layer_norm")

    def _fused(x, weight, bias, residual=None, eps=1e-6, prenorm=False,
               residual_in_fp32=False, is_rms_norm=False, **_unused):
        s = x.float() if residual is None else x.float() + residual.float()
        if residual is not None:
            rdt = residual.dtype
        else:
            rdt = torch.float32 if residual_in_fp32 else x.dtype
        if is_rms_norm:
            y = s * torch.rsqrt(s.pow(2).mean(-1, keepdim=True) + eps) * weight.float()
            if bias is not None:
                y = y + bias.float()
        else:
            y = F.layer_norm(s, s.shape[-1:], weight.float(),
                             None if bias is None else bias.float(), eps)
        y = y.to(x.dtype)
        return (y, s.to(rdt)) if prenorm else y

    def rms_norm_fn(x, weight, bias, **kw):
        return _fused(x, weight, bias, is_rms_norm=True, **kw)

    def layer_norm_fn(x, weight, bias, **kw):
        return _fused(x, weight, bias, is_rms_norm=False, **kw)

    class RMSNorm(nn.Module):
        def __init__(self, hidden_size, eps=1e-5, dropout_p=0.0, device=None, dtype=None):
            super().__init__()
            self.eps = eps
            self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
            self.register_parameter("bias", None)

        def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
            return rms_norm_fn(x, self.weight, self.bias, residual=residual, eps=self.eps,
                               prenorm=prenorm, residual_in_fp32=residual_in_fp32)

    ln.RMSNorm = RMSNorm
    ln.rms_norm_fn = rms_norm_fn
    ln.layer_norm_fn = layer_norm_fn

    # ---- timm (init helpers only)
    _module("timm")
    _module("timm.layers")
    _module("timm.models")
    drop = _module("timm.layers.drop")

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            raise NotImplementedError("stochastic depth is training-only")

    drop.DropPath = DropPath
    helpers = _module("timm.layers.helpers")
    helpers.to_2tuple = lambda v: tuple(v) if isinstance(v, collections.abc.Iterable) \
        and not isinstance(v, str) else tuple(repeat(v, 2))
    winit = _module("timm.layers.weight_init")
    winit.trunc_normal_ = lambda t, mean=0.0, std=1.0, a=-2.0, b=2.0: \
        nn.init.trunc_normal_(t, mean, std, a, b)
    vit = _module("timm.models.vision_transformer")
    vit._cfg = lambda url="", **kw: {"url": url, **kw}

    def _no_weights(*a, **k):
        raise NotImplementedError

    vit._load_weights = _no_weights
    tc = _module("termcolor")
    tc.colored = lambda s, *a, **k: s


class _GuardPatchingFinder(importlib.abc.MetaPathFinder):
    """Compile models/videomamba/mamba_simple.py with its two is_cuda guards disabled."""

    TARGET = "models.videomamba.mamba_simple"

    def find_spec(self, fullname, path=None, target=None):
        if fullname != self.TARGET:
            return None
        src_path = os.path.join(REFERENCE_ROOT, "models", "videomamba", "mamba_simple.py")
        return importlib.util.spec_from_loader(fullname, _GuardPatchingLoader(src_path),
                                               origin=src_path)


class _GuardPatchingLoader(importlib.abc.Loader):
    def __init__(self, src_path):
        self.src_path = src_path

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(self.src_path, "r", encoding="utf-8") as fh:
            source = fh.read()
        needle = "if not hidden_states.is_cuda:"
        if source.count(needle) != 2:
            raise RuntimeError("reference layout changed: expected two is_cuda guards")
        source = source.replace(needle, "if False:")
        module.__file__ = self.src_path
        module.__package__ = "models.videomamba"
        exec(compile(source, self.src_path, "exec"), module.__dict__)


_loaded = False


def load_reference(patch_cuda_guard: bool = True):
    """Return the reference's ``models.videomamba`` package, importable on CPU."""
    global _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if not _loaded:
        for name in list(sys.modules):
            if name == "models" or name.startswith("models.") or name == "video_mamba_ref":
                del sys.modules[name]
        _install_leaf_stubs()
        if patch_cuda_guard:
            sys.meta_path.insert(0, _GuardPatchingFinder())
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            import models.videomamba  # noqa: F401
            import models.refiner_backbone  # noqa: F401
        finally:
            sys.path.remove(REFERENCE_ROOT)
        _loaded = True
    return sys.modules["models.videomamba"]


def reference_modules():
    """(mamba_simple, videomamba, streaming, refiner_backbone) modules of the live reference."""
    load_reference()
    return (sys.modules["models.videomamba.mamba_simple"],
            sys.modules["models.videomamba.videomamba"],
            sys.modules["models.videomamba.streaming"],
            sys.modules["models.refiner_backbone"])
