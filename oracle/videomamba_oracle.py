"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the VideoMamba mixer hot path (see oracle/__init__.py).

A plain-torch, CPU restatement of the algorithm the reference runs on this path.  Every
function cites the reference lines it follows (paths relative to /root/reference).  It works
on a flat ``state_dict`` (name -> tensor) so it shares no module classes with the product.

Arithmetic convention (reference Appendix: slow path, ``use_fast_path=False``): tensors move
between ops in the model dtype ``T`` (fp32 or bf16); inside an op the math is fp32 and the
result is rounded to ``T`` once -- that is what the upstream CUDA/Triton leaves do.  The scan
is fp32 throughout and returns its last state in fp32.

Parity pinning: see oracle/__init__.py ("pinned to the live reference host code and to
tests/golden; the third-party leaves cross-checked against vllm's copies of the upstream kernels").
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor

LayerState = Tuple[Tensor, Tensor]


# --------------------------------------------------------------------------------------
# leaf ops
# --------------------------------------------------------------------------------------
def causal_conv1d_ref(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None,
                      activation: Optional[str] = None) -> Tensor:
    """Depthwise causal conv, channel-major ``x (B, D, L)``, ``weight (D, W)``.

    Call sites: models/videomamba/mamba_simple.py:383-399; the module declares the same
    operator as ``nn.Conv1d(groups=D, padding=W-1)`` at :222-230.  fp32 accumulate, one
    rounding to the input dtype (upstream causal-conv1d 1.6.2 ``causal_conv1d_ref``).
    """
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    dtype_in = x.dtype
    seqlen = x.shape[-1]
    dim, width = weight.shape
    acc = F.conv1d(x.float(), weight.float().unsqueeze(1),
                   None if bias is None else bias.float(),
                   padding=width - 1, groups=dim)[..., :seqlen]
    if activation is not None:
        acc = acc * torch.sigmoid(acc)
    return acc.to(dtype_in)


def causal_conv1d_update_ref(x: Tensor, conv_state: Tensor, weight: Tensor,
                             bias: Optional[Tensor] = None,
                             activation: Optional[str] = None) -> Tensor:
    """Single-token conv step; rolls ``conv_state (B, D, W)`` IN PLACE.

    Call site: models/videomamba/mamba_simple.py:468-474.
    """
    dtype_in = x.dtype
    conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
    conv_state[:, :, -1] = x.to(conv_state.dtype)
    acc = (conv_state.float() * weight.float()).sum(-1)
    if bias is not None:
        acc = acc + bias.float()
    if activation is not None:
        acc = acc * torch.sigmoid(acc)
    return acc.to(dtype_in)


def selective_scan_ref(u: Tensor, delta: Tensor, A: Tensor, B: Tensor, C: Tensor,
                       D: Optional[Tensor] = None, z: Optional[Tensor] = None,
                       delta_bias: Optional[Tensor] = None, delta_softplus: bool = False,
                       initial_state: Optional[Tensor] = None,
                       return_last_state: bool = False):
    """Selective scan, real A, input-dependent B/C of shape (B, N, L).

    Follows models/videomamba/mamba_simple.py:30-106 (``_selective_scan_ref``):
      :43-49   fp32 upcast, bias add, softplus
      :65-70   h0 = initial_state.float() or zeros
      :72,77   dA = exp(delta*A), dBu = delta*B*u
      :84-94   h = dA*h + dBu ; y = <h, C> ; last_state = h at the final step
      :99-102  out = (y + u*D) * silu(z), cast to the input dtype
    Restated step by step (no (B,D,L,N) materialisation) so it also runs at full sizes.
    """
    dtype_in = u.dtype
    uf = u.float()
    df = delta.float()
    if delta_bias is not None:
        df = df + delta_bias.float()[None, :, None]
    if delta_softplus:
        df = F.softplus(df)
    bsz, dim, seqlen = uf.shape
    nstate = A.shape[1]
    Af = A.float()
    Bf = B.float()
    Cf = C.float()
    if initial_state is None:
        h = torch.zeros(bsz, dim, nstate, dtype=torch.float32, device=u.device)
    else:
        h = initial_state.float().clone()
    ys = torch.empty(bsz, dim, seqlen, dtype=torch.float32, device=u.device)
    du = df * uf
    for t in range(seqlen):
        decay = torch.exp(df[:, :, t, None] * Af[None])
        h = decay * h + du[:, :, t, None] * Bf[:, None, :, t]
        ys[:, :, t] = (h * Cf[:, None, :, t]).sum(-1)
    out = ys if D is None else ys + uf * D.float()[None, :, None]
    if z is not None:
        zf = z.float()
        out = out * (zf * torch.sigmoid(zf))
    out = out.to(dtype_in)
    if return_last_state:
        return out, h
    return out


def selective_state_update_ref(state: Tensor, x: Tensor, dt: Tensor, A: Tensor, B: Tensor,
                               C: Tensor, D: Optional[Tensor] = None,
                               z: Optional[Tensor] = None, dt_bias: Optional[Tensor] = None,
                               dt_softplus: bool = False) -> Tensor:
    """One recurrent step; ``state (B, D, N)`` is updated IN PLACE.

    Call sites: models/videomamba/mamba_simple.py:160-171 and :483-494.
    """
    dtf = dt.float()
    if dt_bias is not None:
        dtf = dtf + dt_bias.float()
    if dt_softplus:
        dtf = F.softplus(dtf)
    xf = x.float()
    decay = torch.exp(dtf[..., None] * A.float()[None])
    drive = dtf[..., None] * B.float()[:, None, :] * xf[..., None]
    new_state = state.float() * decay + drive
    state.copy_(new_state.to(state.dtype))
    out = (new_state * C.float()[:, None, :]).sum(-1)
    if D is not None:
        out = out + xf * D.float()
    if z is not None:
        zf = z.float()
        out = out * (zf * torch.sigmoid(zf))
    return out.to(x.dtype)


def add_norm_ref(x: Tensor, weight: Tensor, bias: Optional[Tensor], residual: Optional[Tensor],
                 eps: float, prenorm: bool, residual_in_fp32: bool, is_rms: bool):
    """Fused residual-add + RMSNorm / LayerNorm.

    Call sites: models/videomamba/videomamba.py:151-166 (prenorm=True) and :902-918
    (prenorm=False).  fp32 sum of x and residual; residual_out keeps the incoming residual's
    dtype, or fp32 when ``residual_in_fp32`` and there is no incoming residual; the norm is
    computed from the fp32 sum and rounded to x's dtype (upstream mamba-ssm 2.3.2
    ``_layer_norm_fwd`` semantics).
    """
    out_dtype = x.dtype
    acc = x.float()
    if residual is not None:
        acc = acc + residual.float()
        res_dtype = residual.dtype
    else:
        res_dtype = torch.float32 if residual_in_fp32 else out_dtype
    if is_rms:
        rstd = torch.rsqrt(acc.square().mean(-1, keepdim=True) + eps)
        y = acc * rstd * weight.float()
    else:
        mean = acc.mean(-1, keepdim=True)
        var = (acc - mean).square().mean(-1, keepdim=True)
        y = (acc - mean) * torch.rsqrt(var + eps) * weight.float()
    if bias is not None:
        y = y + bias.float()
    y = y.to(out_dtype)
    if prenorm:
        return y, acc.to(res_dtype)
    return y


# --------------------------------------------------------------------------------------
# mixer (models/videomamba/mamba_simple.py:283-497)
# --------------------------------------------------------------------------------------
def _linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """``x @ w.T`` with fp32 accumulation and one rounding to x's dtype."""
    y = x.float() @ w.float().t()
    if b is not None:
        y = y + b.float()
    return y.to(x.dtype)


def mixer_ref(p: Dict[str, Tensor], hidden: Tensor,
              conv_state: Optional[Tensor] = None, ssm_state: Optional[Tensor] = None,
              want_state: bool = False):
    """Mamba mixer slow path on ``hidden (B, L, D)``.

    ``p`` holds this mixer's tensors keyed as in the reference state_dict
    (``in_proj.weight`` ... ``out_proj.weight``, mamba_simple.py:218-281).
    Data flow and rounding points: mamba_simple.py:333-339 (in_proj), :369 (split),
    :381-404 (conv, optional conv_state history, new conv state = last W pre-conv inputs),
    :409-416 (x_proj, split, dt_proj), :423-435 (scan with initial state), :445-446 (out_proj).
    Returns ``out`` or ``(out, (new_conv_state, last_ssm_state))``.
    """
    T = hidden.dtype
    conv_w = p["conv1d.weight"]
    d_inner, _, d_conv = conv_w.shape
    d_state = p["A_log"].shape[1]
    dt_rank = p["dt_proj.weight"].shape[1]
    bsz, seqlen, _ = hidden.shape

    xz = _linear(hidden, p["in_proj.weight"], p.get("in_proj.bias"))      # (B, L, 2Di)
    x_in = xz[..., :d_inner].transpose(1, 2)                                 # (B, Di, L)
    z = xz[..., d_inner:].transpose(1, 2)

    w2 = conv_w.reshape(d_inner, d_conv)
    if conv_state is not None:
        x_cat = torch.cat([conv_state.to(T), x_in], dim=-1)
        xc = causal_conv1d_ref(x_cat, w2, p.get("conv1d.bias"), "silu")[..., -seqlen:]
        new_conv = x_cat[..., -d_conv:]
    else:
        xc = causal_conv1d_ref(x_in, w2, p.get("conv1d.bias"), "silu")
        new_conv = F.pad(x_in, (d_conv - seqlen, 0))

    x_dbl = _linear(xc.transpose(1, 2).reshape(bsz * seqlen, d_inner), p["x_proj.weight"])
    dt_low, Bm, Cm = torch.split(x_dbl, [dt_rank, d_state, d_state], dim=-1)
    delta = _linear(dt_low, p["dt_proj.weight"])                             # (BL, Di), rounded to T
    delta = delta.reshape(bsz, seqlen, d_inner).transpose(1, 2)
    Bm = Bm.reshape(bsz, seqlen, d_state).transpose(1, 2)
    Cm = Cm.reshape(bsz, seqlen, d_state).transpose(1, 2)

    A = -torch.exp(p["A_log"].float())
    y, last = selective_scan_ref(
        xc, delta, A, Bm, Cm, p["D"].float(), z=z,
        delta_bias=p["dt_proj.bias"].float(), delta_softplus=True,
        initial_state=ssm_state, return_last_state=True)
    out = _linear(y.transpose(1, 2), p["out_proj.weight"], p.get("out_proj.bias"))
    if want_state:
        return out, (new_conv.contiguous(), last)
    return out


def mixer_step_ref(p: Dict[str, Tensor], hidden: Tensor, conv_state: Tensor,
                   ssm_state: Tensor) -> Tensor:
    """Single-token decode (models/videomamba/mamba_simple.py:453-497); states in place."""
    d_inner, _, d_conv = p["conv1d.weight"].shape
    d_state = p["A_log"].shape[1]
    dt_rank = p["dt_proj.weight"].shape[1]
    xz = _linear(hidden[:, 0], p["in_proj.weight"], p.get("in_proj.bias"))
    x, z = xz[:, :d_inner], xz[:, d_inner:]
    x = causal_conv1d_update_ref(x, conv_state, p["conv1d.weight"].reshape(d_inner, d_conv),
                                 p.get("conv1d.bias"), "silu")
    x_db = _linear(x, p["x_proj.weight"])
    dt_low, Bm, Cm = torch.split(x_db, [dt_rank, d_state, d_state], dim=-1)
    dt = _linear(dt_low, p["dt_proj.weight"])
    A = -torch.exp(p["A_log"].float())
    y = selective_state_update_ref(ssm_state, x, dt, A, Bm, Cm, p["D"], z=z,
                                   dt_bias=p["dt_proj.bias"], dt_softplus=True)
    return _linear(y, p["out_proj.weight"], p.get("out_proj.bias")).unsqueeze(1)


def _sub(sd: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------------------
# whole-model forward (models/videomamba/videomamba.py:786-1067)
# --------------------------------------------------------------------------------------
class OracleVideoMamba:
    """Functional restatement of ``PretrainVideoMamba.forward`` on a flat state_dict.

    ``cfg`` keys: img_size, patch_size, depth, embed_dim, kernel_size, num_frames,
    norm_epsilon, rms_norm, fused_add_norm, residual_in_fp32, pool_type, add_pool_norm.
    Covers the unmasked path, masks, all pooling variants, and list/tuple/dict streaming
    state (videomamba.py:646-653 CLS rule, :655-675 temporal embedding slice/interp).
    """

    def __init__(self, cfg: dict, sd: Dict[str, Tensor]):
        self.cfg = dict(cfg)
        self.sd = sd
        self.depth = int(cfg["depth"])
        self.dim = int(cfg["embed_dim"])
        ps = cfg["patch_size"]
        self.patch = (ps, ps) if isinstance(ps, int) else tuple(ps)
        im = cfg["img_size"]
        self.img = (im, im) if isinstance(im, int) else tuple(im)
        self.tubelet = int(cfg.get("kernel_size", 1))
        self.eps = float(cfg.get("norm_epsilon", 1e-5))
        self.rms = bool(cfg.get("rms_norm", True))
        self.fused = bool(cfg.get("fused_add_norm", True))
        self.res_fp32 = bool(cfg.get("residual_in_fp32", True))
        self.pool_type = cfg.get("pool_type", "cls+avg")
        self.add_pool_norm = bool(cfg.get("add_pool_norm", True))
        self.layers = [_sub(sd, f"layers.{i}.mixer.") for i in range(self.depth)]

    # -- embeddings ------------------------------------------------------------------
    def _spatial_pos(self, gh: int, gw: int, dtype) -> Tensor:
        # videomamba.py:621-644
        pos = self.sd["pos_embed"][:, 1:]
        bh, bw = self.img[0] // self.patch[0], self.img[1] // self.patch[1]
        if (gh, gw) == (bh, bw):
            return pos.to(dtype)
        grid = pos.reshape(1, bh, bw, self.dim).permute(0, 3, 1, 2).float()
        grid = F.interpolate(grid, size=(gh, gw), mode="bicubic", align_corners=False)
        return grid.permute(0, 2, 3, 1).reshape(1, gh * gw, self.dim).to(dtype)

    def _temporal_pos(self, t: int, offset: int, dtype) -> Tensor:
        # videomamba.py:655-675 (interpolates to size offset+t, then slices)
        table = self.sd["temporal_pos_embedding"].to(dtype)
        end = offset + t
        if end <= table.shape[1]:
            return table[:, offset:end]
        stretched = F.interpolate(table.permute(0, 2, 1).float(), size=end, mode="linear",
                                  align_corners=False)
        return stretched.permute(0, 2, 1).to(dtype)[:, offset:end]

    @staticmethod
    def _layer_state(state, idx):
        if state is None:
            return None
        if isinstance(state, dict):
            return state.get(idx)
        return state[idx]

    def _has_cls(self, state, offset: int) -> bool:
        # videomamba.py:646-653
        if state is None or offset <= 0:
            return True
        first = self._layer_state(state, 0)
        return not (isinstance(first, (list, tuple)) and len(first) == 2)

    def _norm(self, x, w, b, residual, prenorm):
        # fused (videomamba.py:151-166, :902-918) and unfused (:141-150, :896-901) twins
        if self.fused:
            return add_norm_ref(x, w, b, residual, self.eps, prenorm, self.res_fp32, self.rms)
        summed = x if residual is None else residual + x
        y = add_norm_ref(summed.to(w.dtype), w, b, None, self.eps, False, False, self.rms)
        if not prenorm:
            return y
        return y, (summed.float() if self.res_fp32 else summed)

    # -- forward ---------------------------------------------------------------------
    def forward_features(self, x: Tensor, mask: Optional[Tensor] = None, ssm_state=None,
                         temporal_pos_offset: int = 0):
        sd = self.sd
        w = sd["patch_embed.proj.weight"]
        tok = F.conv3d(x.to(w.dtype), w, sd["patch_embed.proj.bias"],
                       stride=(self.tubelet, self.patch[0], self.patch[1]))
        bsz, C, T, H, W = tok.shape
        tok = tok.permute(0, 2, 3, 4, 1).reshape(bsz, T, H * W, C)
        tok = tok + self._spatial_pos(H, W, tok.dtype).unsqueeze(1)
        tok = tok + self._temporal_pos(T, temporal_pos_offset, tok.dtype).unsqueeze(2)
        tok = tok.reshape(bsz, T * H * W, C)
        has_cls = self._has_cls(ssm_state, temporal_pos_offset)
        if has_cls:
            cls = (sd["cls_token"] + sd["pos_embed"][:, :1]).to(tok.dtype).expand(bsz, -1, -1)
            tok = torch.cat([cls, tok], dim=1)
        if mask is not None:
            vis = ~mask.to(torch.bool)
            nvis = int(vis[0].sum())
            idx = torch.arange(tok.shape[1]).unsqueeze(0).expand(bsz, -1)
            idx = idx.masked_fill(~vis, tok.shape[1]).sort(dim=1).values[:, :nvis]
            tok = tok.gather(1, idx.unsqueeze(-1).expand(-1, -1, C))

        hidden, residual = tok, None
        new_states: Optional[Union[dict, list]] = None
        for i in range(self.depth):
            st = self._layer_state(ssm_state, i)
            full = isinstance(st, (list, tuple)) and len(st) == 2
            if full and new_states is None:
                new_states = {} if isinstance(ssm_state, dict) else [None] * self.depth
            hidden, residual = self._norm(hidden, sd[f"layers.{i}.norm.weight"],
                                          sd.get(f"layers.{i}.norm.bias"), residual, True)
            if full:
                hidden, st = mixer_ref(self.layers[i], hidden, st[0], st[1], want_state=True)
            elif st is not None:
                # legacy ssm-only tensor: initial state in, updated in place
                # (mamba_simple.py:419-421, :436-440)
                hidden, (_, last) = mixer_ref(self.layers[i], hidden, None, st, want_state=True)
                st.copy_(last)
            else:
                hidden = mixer_ref(self.layers[i], hidden)
            if new_states is not None:
                new_states[i] = st
        out = self._norm(hidden, sd["norm.weight"], sd.get("norm.bias"), residual, False)
        if ssm_state is None:
            return out
        if new_states is None:
            return out, ssm_state
        if isinstance(ssm_state, tuple):
            return out, tuple(new_states)
        return out, new_states

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, keep_temporal: bool = False,
                ssm_state=None, temporal_pos_offset: int = 0):
        # videomamba.py:943-1067
        gh, gw = x.shape[-2] // self.patch[0], x.shape[-1] // self.patch[1]
        per_frame = gh * gw
        t_tokens = x.shape[2] // self.tubelet
        has_cls = self._has_cls(ssm_state, temporal_pos_offset)
        feats = self.forward_features(x, mask, ssm_state, temporal_pos_offset)
        if ssm_state is None:
            x_vis, nxt = feats, None
        else:
            x_vis, nxt = feats
        if not self.add_pool_norm:
            return x_vis if ssm_state is None else (x_vis, nxt)
        cls = x_vis[:, :1] if has_cls else None
        patches = x_vis[:, 1:] if has_cls else x_vis
        pn = lambda v: F.layer_norm(v, (self.dim,), self.sd["pool_norm.weight"],
                                    self.sd["pool_norm.bias"], 1e-5)
        if self.pool_type == "cls":
            pooled = pn(cls)
        else:
            if keep_temporal:
                if mask is None:
                    avg = patches.reshape(patches.shape[0], t_tokens, per_frame, -1).mean(2)
                else:
                    vis = ~mask.to(torch.bool)
                    n_tot = vis.shape[1]
                    idx = torch.arange(n_tot).unsqueeze(0).expand(vis.shape[0], -1)
                    idx = idx.masked_fill(~vis, n_tot).sort(dim=1).values[:, :int(vis[0].sum())]
                    ppos = idx[:, 1:] - 1 if has_cls else idx
                    frame = torch.div(ppos, per_frame, rounding_mode="floor")
                    sums = torch.zeros(patches.shape[0], t_tokens, self.dim, dtype=patches.dtype)
                    sums.scatter_add_(1, frame.unsqueeze(-1).expand(-1, -1, self.dim), patches)
                    cnt = torch.zeros(patches.shape[0], t_tokens, 1, dtype=patches.dtype)
                    cnt.scatter_add_(1, frame.unsqueeze(-1),
                                     torch.ones(patches.shape[0], patches.shape[1], 1,
                                                dtype=patches.dtype))
                    avg = sums / cnt
            else:
                avg = patches.mean(1, keepdim=True)
            if self.pool_type == "cls+avg":
                pooled = pn(cls + avg)
            elif self.pool_type == "cls_cat_avg":
                pooled = pn(torch.cat([cls, avg], dim=1))
            elif self.pool_type == "avg":
                pooled = pn(avg)
            else:
                raise ValueError(f"Unsupported pool_type: {self.pool_type}")
        if ssm_state is None:
            return patches, pooled
        return patches, pooled, nxt


# --------------------------------------------------------------------------------------
# deterministic synthetic weights (no RNG-order dependence on any model class)
# --------------------------------------------------------------------------------------
def synthetic_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32,
                         perturbed: bool = False) -> Dict[str, Tensor]:
    """Random weights with the reference's parameter names / shapes / init statistics
    (mamba_simple.py:218-281; videomamba.py:434-438, :327-334, :295-324).  ``perturbed``
    breaks the S4D-real structure of A, uses the stand-alone inverse-softplus dt bias and a
    non-zero temporal embedding."""
    g = torch.Generator().manual_seed(seed)
    D = int(cfg["embed_dim"])
    depth = int(cfg["depth"])
    Di = 2 * D
    N = int(cfg.get("d_state", 16))
    W = int(cfg.get("d_conv", 4))
    R = math.ceil(D / 16)
    ps = cfg["patch_size"]
    ph, pw = (ps, ps) if isinstance(ps, int) else ps
    im = cfg["img_size"]
    ih, iw = (im, im) if isinstance(im, int) else im
    k = int(cfg.get("kernel_size", 1))
    chans = int(cfg.get("channels", 3))
    npatch = (ih // ph) * (iw // pw)
    rn = lambda *s, std=0.02: (torch.randn(*s, generator=g) * std).clamp_(-2 * std, 2 * std)
    sd: Dict[str, Tensor] = {}
    sd["cls_token"] = torch.zeros(1, 1, D)
    sd["pos_embed"] = rn(1, npatch + 1, D)
    t_len = int(cfg["num_frames"]) // k
    sd["temporal_pos_embedding"] = rn(1, t_len, D) if perturbed else torch.zeros(1, t_len, D)
    fan_in = chans * k * ph * pw
    bound = 1.0 / math.sqrt(fan_in)
    sd["patch_embed.proj.weight"] = (torch.rand(D, chans, k, ph, pw, generator=g) * 2 - 1) * bound
    sd["patch_embed.proj.bias"] = (torch.rand(D, generator=g) * 2 - 1) * bound
    for i in range(depth):
        pre = f"layers.{i}."
        sd[pre + "mixer.in_proj.weight"] = rn(2 * Di, D)
        cb = 1.0 / math.sqrt(W)
        sd[pre + "mixer.conv1d.weight"] = (torch.rand(Di, 1, W, generator=g) * 2 - 1) * cb
        sd[pre + "mixer.conv1d.bias"] = (torch.rand(Di, generator=g) * 2 - 1) * cb
        sd[pre + "mixer.x_proj.weight"] = rn(R + 2 * N, Di)
        sd[pre + "mixer.dt_proj.weight"] = rn(Di, R)
        if perturbed:
            dt = torch.exp(torch.rand(Di, generator=g) * (math.log(0.1) - math.log(0.001))
                           + math.log(0.001)).clamp(min=1e-4)
            sd[pre + "mixer.dt_proj.bias"] = dt + torch.log(-torch.expm1(-dt))
        else:
            sd[pre + "mixer.dt_proj.bias"] = torch.zeros(Di)
        a_log = torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(Di, 1)
        if perturbed:
            a_log = a_log + 0.1 * torch.randn(Di, N, generator=g)
        sd[pre + "mixer.A_log"] = a_log
        sd[pre + "mixer.D"] = torch.ones(Di)
        ob = math.sqrt(6.0 / ((1 + 5.0) * Di)) / math.sqrt(depth)
        sd[pre + "mixer.out_proj.weight"] = (torch.rand(D, Di, generator=g) * 2 - 1) * ob
        sd[pre + "norm.weight"] = torch.ones(D)
        if not cfg.get("rms_norm", True):
            sd[pre + "norm.bias"] = torch.zeros(D)
    sd["norm.weight"] = torch.ones(D)
    if not cfg.get("rms_norm", True):
        sd["norm.bias"] = torch.zeros(D)
    if cfg.get("add_pool_norm", True):
        sd["pool_norm.weight"] = torch.ones(D)
        sd["pool_norm.bias"] = torch.zeros(D)
    return {k_: v.to(dtype) for k_, v in sd.items()}


def rel_err(a: Tensor, b: Tensor) -> float:
    """max|a-b| / max(|b|, tiny): the "relative" error the parity tests use."""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# --------------------------------------------------------------------------------------
# callers either side of the mixer ("next" rows: refiner block, inference cache)
# --------------------------------------------------------------------------------------
def mixer_prefill_cache_ref(p: Dict[str, Tensor], hidden: Tensor, conv_state: Tensor,
                            ssm_state: Tensor) -> Tensor:
    """``inference_params`` prefill (mamba_simple.py:316-330, :372-378, :419-440): the conv runs
    WITHOUT history, ``conv_state`` is overwritten in place with the last W pre-conv inputs, the
    cached ssm state is the scan's initial state and receives the last state in place."""
    out, (new_conv, last) = mixer_ref(p, hidden, None, ssm_state, want_state=True)
    conv_state.copy_(new_conv)
    ssm_state.copy_(last)
    return out


def refiner_ref(sd: Dict[str, Tensor], x: Tensor, state_fwd: Optional[LayerState] = None,
                state_bwd: Optional[LayerState] = None, eps: float = 1e-5):
    """``BiMambaRefinerBlock.forward`` (models/refiner_backbone.py:92-135): forward block on x,
    backward block on the time-flipped x (4-D input flips the frame axis only, :61-68), sigmoid
    fusion gate, output projection.  Blocks are called with ``residual=None`` so their add+norm
    is norm-only (videomamba.py:151-166).  Returns ``(out, new_state_fwd)``."""
    packed = None
    if x.ndim == 4:
        b, t, n, c = x.shape
        packed = (b, t, n)
        seq = x.reshape(b, t * n, c)
    else:
        seq = x

    def flip(v):
        if packed is None:
            return torch.flip(v, dims=[1])
        b, t, n = packed
        return torch.flip(v.reshape(b, t, n, v.shape[-1]), dims=[1]).reshape(b, t * n, v.shape[-1])

    def run(prefix, inp, st):
        p = _sub(sd, prefix + "mixer.")
        d_inner, _, d_conv = p["conv1d.weight"].shape
        if st is None:
            st = (torch.zeros(inp.shape[0], d_inner, d_conv, dtype=p["conv1d.weight"].dtype),
                  torch.zeros(inp.shape[0], d_inner, p["A_log"].shape[1],
                              dtype=p["dt_proj.weight"].dtype))
        normed = add_norm_ref(inp, sd[prefix + "norm.weight"], sd.get(prefix + "norm.bias"),
                              None, eps, False, True, prefix + "norm.bias" not in sd)
        return mixer_ref(p, normed, st[0], st[1], want_state=True)

    out_f, new_f = run("block_fwd.", seq, state_fwd)
    out_b_rev, _ = run("block_bwd.", flip(seq), state_bwd)
    out_b = flip(out_b_rev)
    gate = torch.sigmoid(F.linear(torch.cat([out_f, out_b], dim=-1),
                                  sd["fusion_gate.0.weight"], sd["fusion_gate.0.bias"]))
    mixed = gate * out_f + (1.0 - gate) * out_b
    out = F.linear(mixed, sd["out_proj.weight"], sd["out_proj.bias"])
    if packed is not None:
        out = out.reshape(packed[0], packed[1], packed[2], out.shape[-1])
    return out, new_f
