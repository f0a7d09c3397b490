"""``Mamba`` mixer module backed by the sm_100a kernels of libvmb200.

Drop-in for the reference's ``models/videomamba/mamba_simple.py:175-590``: same constructor
arguments, same parameter names / shapes / init (state_dict compatible), same forward
contract (``state`` / ``ssm_state`` / ``inference_params`` / ``return_state``), same error
messages.  The arithmetic matches the reference's ``use_fast_path=False`` path
(``_selective_scan_ref`` semantics).  ``use_fast_path`` is accepted and stored, but both values
run the same kernels (no multi-backend dispatch).  This fork's mixer is unidirectional
(mamba_simple.py:209-211): there are no ``*_b`` parameters.
"""
from __future__ import annotations

import math
import os
from typing import Any, MutableMapping, Optional, Protocol, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import ops

_TRUE_STRINGS = {"1", "true", "yes", "y", "on"}


class InferenceParamsLike(Protocol):
    seqlen_offset: int
    key_value_memory_dict: MutableMapping[int, Tuple[Tensor, Tensor]]


class Mamba(nn.Module):
    def __init__(
        self,
        d_model: int,
        d_state: int = 16,
        d_conv: int = 4,
        expand: int = 2,
        dt_rank: Union[int, str] = "auto",
        dt_min: float = 0.001,
        dt_max: float = 0.1,
        dt_init: str = "random",
        dt_scale: float = 1.0,
        dt_init_floor: float = 1e-4,
        conv_bias: bool = True,
        bias: bool = False,
        use_fast_path: bool = True,
        layer_idx: Optional[int] = None,
        bimamba: bool = True,
        device: Optional[torch.device] = None,
        dtype: Optional[torch.dtype] = None,
        **_: Any,
    ):
        super().__init__()
        fk: dict = {}
        if device is not None:
            fk["device"] = device
        if dtype is not None:
            fk["dtype"] = dtype
        self.d_model = d_model
        self.d_state = d_state
        self.d_conv = d_conv
        self.expand = expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else int(dt_rank)
        self.bimamba = bool(bimamba)          # config-surface compatibility only
        if os.getenv("VIDEOMAMBA_DISABLE_FUSED", "").lower() in _TRUE_STRINGS:
            use_fast_path = False
        self.use_fast_path = use_fast_path
        self.layer_idx = layer_idx
        self.activation = "silu"

        # Parameter set and creation order follow mamba_simple.py:218-281 (same RNG stream).
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner, bias=bias, **fk)
        self.conv1d = nn.Conv1d(self.d_inner, self.d_inner, kernel_size=d_conv, groups=self.d_inner,
                                padding=d_conv - 1, bias=conv_bias, **fk)
        self.act = nn.SiLU()
        self.x_proj = nn.Linear(self.d_inner, self.dt_rank + 2 * d_state, bias=False, **fk)
        self.dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True, **fk)

        std = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError
        # softplus(dt_bias) log-uniform in [dt_min, dt_max]
        log_lo, log_hi = math.log(dt_min), math.log(dt_max)
        dt = torch.exp(torch.rand(self.d_inner, **fk) * (log_hi - log_lo) + log_lo)
        dt = dt.clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))   # softplus^-1
        setattr(self.dt_proj.bias, "_no_reinit", True)

        A = torch.arange(1, d_state + 1, dtype=torch.float32, device=device)
        self.A_log = nn.Parameter(torch.log(A).repeat(self.d_inner, 1).contiguous())
        setattr(self.A_log, "_no_weight_decay", True)
        self.D = nn.Parameter(torch.ones(self.d_inner, device=device))
        setattr(self.D, "_no_weight_decay", True)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)

        self._weights_key: Optional[tuple] = None
        self._weights: Optional[ops.MixerWeights] = None
        # load_state_dict copies into the parameters in place: drop the derived copies with it
        self.register_load_state_dict_post_hook(lambda module, _incompatible: module.refresh_weights())

    # ------------------------------------------------------------------------------------------
    #: debugging aid: when True every forward re-derives the kernel operands and compares them with
    #: the cached ones (one device sync per call); catches writes that bypass autograd's version
    #: counter (``param.data.copy_()``, flat-parameter optimisers) without ``refresh_weights()``.
    verify_weights: bool = False

    #: stateless forward walks run conv + x_proj as one kernel (vmb_conv_xproj_fwd): bit-identical, less
    #: HBM traffic, faster when ONE forward is in flight, slower when several share the GPU
    fuse_conv_xproj: bool = False

    #: the in_proj epilogue stores SiLU(z) and the fused scan multiplies by it as it is (bf16 fast path; the gate is
    #: rounded after the activation instead of before it).  False = the scan applies SiLU to the stored z.
    gate_in_proj: bool = True

    def refresh_weights(self) -> None:
        """Drop the kernel-ready copies derived from the parameters (``A2 = -exp(A_log)*log2(e)``,
        fp32 ``D`` / ``dt_proj.bias``, zero-padded ``x_proj`` / ``dt_proj`` weights, the geometric-A
        flag).  They are rebuilt on the next forward.  Called automatically after
        ``load_state_dict`` and whenever a parameter's storage, dtype or autograd version changes;
        call it yourself after writing through ``param.data`` (EMA / SWA averaging, flat-parameter
        frameworks), which bumps no version counter."""
        self._weights_key = None
        self._weights = None

    def _apply(self, fn, *args, **kwargs):          # .to() / .cuda() / .bfloat16() replace the storages
        self.refresh_weights()
        return super()._apply(fn, *args, **kwargs)

    def _params(self):
        return (self.in_proj.weight, self.in_proj.bias, self.conv1d.weight, self.conv1d.bias,
                self.x_proj.weight, self.dt_proj.weight, self.dt_proj.bias, self.A_log, self.D,
                self.out_proj.weight, self.out_proj.bias)

    def weights_key(self) -> tuple:
        """Identity of the parameter storages the cached kernel operands were derived from."""
        return tuple(None if p is None else (p.data_ptr(), p._version, p.dtype) for p in self._params())

    def _kernel_weights(self) -> ops.MixerWeights:
        """Kernel-ready weights, rebuilt when any parameter changed (storage, version or dtype)."""
        key = self.weights_key()
        if key != self._weights_key:
            self._weights = ops.MixerWeights(*self._params())
            self._weights_key = key
        elif self.verify_weights:
            fresh = ops.MixerWeights(*self._params())
            for name in ("A2", "Dskip", "dt_bias", "w_x_pad", "w_dt_pad"):
                a, b = getattr(self._weights, name), getattr(fresh, name)
                if a is not None and not torch.equal(a, b):
                    raise RuntimeError(
                        f"Mamba(layer {self.layer_idx}): cached kernel operand {name} is stale -- a parameter "
                        "was written through .data; call refresh_weights() after such writes.")
        return self._weights

    def _wants_grad(self, *tensors) -> bool:
        if not torch.is_grad_enabled():
            return False
        return any(isinstance(t, Tensor) and t.requires_grad for t in tensors) or \
            any(p is not None and p.requires_grad for p in self._params())

    @staticmethod
    def _require_cuda(t: Tensor) -> None:
        if not t.is_cuda:
            raise RuntimeError(
                "VideoMamba requires CUDA tensors in this package because "
                "the mixer kernels are CUDA-only (sm_100a)."
            )

    # ------------------------------------------------------------------------------------------
    def forward(
        self,
        hidden_states: Tensor,
        inference_params: Optional[InferenceParamsLike] = None,
        ssm_state: Optional[Tensor] = None,
        state: Optional[Tuple[Tensor, Tensor]] = None,
        return_state: bool = False,
    ) -> Union[Tensor, Tuple[Tensor, Tuple[Tensor, Tensor]]]:
        """hidden_states (B, L, D) -> (B, L, D), or ``(out, (conv_state, ssm_state))`` when
        ``return_state``.  Control flow mirrors mamba_simple.py:300-330, :419-451."""
        if state is not None and ssm_state is not None:
            raise ValueError("Pass either state or ssm_state, not both.")
        if inference_params is not None and state is not None:
            raise ValueError("state is not supported with inference_params.")
        self._require_cuda(hidden_states)
        batch = hidden_states.shape[0]
        w = self._kernel_weights()

        conv_state = None
        if state is not None:
            conv_state, ssm_state = state

        if inference_params is not None:
            cache_conv, cache_ssm = self._get_states_from_cache(inference_params, batch)
            if ssm_state is None:
                ssm_state = cache_ssm
            if inference_params.seqlen_offset > 0:
                out, _, _ = self.step(hidden_states, cache_conv, ssm_state)
                return out
            # prefill: conv without history, cache_conv <- last d_conv inputs, cached ssm state
            # is the initial state and receives the final one (mamba_simple.py:372-378, :436-440)
            out, new_conv, last = ops.mixer_fwd(w, hidden_states, None, ssm_state, True, True)
            cache_conv.copy_(new_conv)
            ssm_state.copy_(last)
            return out

        inplace_ssm = ssm_state is not None and state is None and not return_state
        want = return_state or inplace_ssm
        if self._wants_grad(hidden_states, conv_state, ssm_state):
            # training: the reference's slow path op for op, every operator with a backward kernel
            # (autograd.py); the fused inference entry point below has none
            from .autograd import mixer_train
            out, new_conv, last = mixer_train(*self._params(), hidden_states, conv_state, ssm_state,
                                              want_conv_state=return_state, want_ssm_state=want)
        else:
            out, new_conv, last = ops.mixer_fwd(w, hidden_states, conv_state, ssm_state,
                                                want_conv_state=return_state, want_ssm_state=want,
                                                fuse_conv_xproj=self.fuse_conv_xproj,
                                                gate_in_proj=self.gate_in_proj)
        if inplace_ssm:
            with torch.no_grad():
                ssm_state.copy_(last)
        if return_state:
            return out, (new_conv, last)
        return out

    def step(self, hidden_states: Tensor, conv_state: Tensor,
             ssm_state: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """Single-token decode; both states are updated in place (mamba_simple.py:453-497)."""
        self._require_cuda(hidden_states)
        assert hidden_states.shape[1] == 1, "Only support decoding with 1 token at a time for now"
        w = self._kernel_weights()
        # three launches: in_proj, the fused step kernel (conv update + x_proj + dt_proj + state update +
        # gate, csrc/step.cu), out_proj
        xz = ops.linear(hidden_states[:, 0], w.w_in, w.b_in)
        y = ops.mixer_step(w, xz, conv_state, ssm_state)
        out = ops.linear(y, w.w_out, w.b_out)
        return out.unsqueeze(1), conv_state, ssm_state

    # ------------------------------------------------------------------------------------------
    def _zero_states(self, batch_size: int, dtype, device) -> Tuple[Tensor, Tensor]:
        conv_dtype = self.conv1d.weight.dtype if dtype is None else dtype
        ssm_dtype = self.dt_proj.weight.dtype if dtype is None else dtype
        conv = torch.zeros(batch_size, self.d_inner, self.d_conv, device=device, dtype=conv_dtype)
        ssm = torch.zeros(batch_size, self.d_inner, self.d_state, device=device, dtype=ssm_dtype)
        return conv, ssm

    def allocate_inference_cache(self, batch_size: int, max_seqlen: int, dtype=None,
                                 **kwargs) -> Tuple[Tensor, Tensor]:
        return self._zero_states(batch_size, dtype, self.out_proj.weight.device)

    def allocate_state(self, batch_size: int, dtype=None, device=None) -> Tuple[Tensor, Tensor]:
        """Zero ``(conv_state (B, d_inner, d_conv), ssm_state (B, d_inner, d_state))``."""
        if device is None:
            device = self.out_proj.weight.device
        return self._zero_states(batch_size, dtype, device)

    def _get_states_from_cache(self, inference_params: InferenceParamsLike, batch_size: int,
                               initialize_states: bool = False) -> Tuple[Tensor, Tensor]:
        """Per-layer cache entry; re-allocated when the batch size changed
        (mamba_simple.py:546-590)."""
        assert self.layer_idx is not None
        cache = inference_params.key_value_memory_dict
        entry = cache.get(self.layer_idx)
        if entry is None or entry[0].shape[0] != batch_size or entry[1].shape[0] != batch_size:
            entry = self._zero_states(batch_size, None, self.conv1d.weight.device)
            cache[self.layer_idx] = entry
        elif initialize_states:
            entry[0].zero_()
            entry[1].zero_()
        return entry
