"""Test helper: run the package's HOST logic on CPU by standing the oracle in for the kernels.

Only tests use this (``-m "not gpu"`` suite): it monkeypatches the marshalling layer
(``videomamba_b200.ops``) with oracle-backed functions so the module / model / streaming code
paths can be exercised without a GPU.  The product itself has no such path.
"""
import torch

from oracle import videomamba_oracle as orc


def _params(w):
    return {k: v.detach() for k, v in w.raw.items() if v is not None}


def mixer_fwd(w, hidden, conv_state=None, ssm_state=None, want_conv_state=False,
              want_ssm_state=False, reverse=False, path=0, frame_len=0, **_tuning):
    p = _params(w)

    def flip(t):            # whole-sequence reversal, or frame-axis reversal (frames of frame_len tokens)
        if not reverse:
            return t
        if not frame_len:
            return torch.flip(t, dims=[1])
        b, l, c = t.shape
        return torch.flip(t.reshape(b, l // frame_len, frame_len, c), dims=[1]).reshape(b, l, c)

    out, (new_conv, last) = orc.mixer_ref(p, flip(hidden), conv_state, ssm_state, want_state=True)
    out = flip(out)
    if want_conv_state and conv_state is not None:
        new_conv = new_conv.to(torch.promote_types(conv_state.dtype, hidden.dtype))
    return out, (new_conv if want_conv_state else None), (last if want_ssm_state else None)


def mixer_train(in_w, in_b, conv_w, conv_b, x_w, dt_w, dt_b, A_log, Dp, out_w, out_b, hidden,
                conv_state=None, ssm_state=None, want_conv_state=False, want_ssm_state=False):
    # the training-mode mixer (videomamba_b200.autograd.mixer_train) on the oracle: live parameters,
    # so torch autograd differentiates it on the CPU
    p = {"in_proj.weight": in_w, "in_proj.bias": in_b, "conv1d.weight": conv_w, "conv1d.bias": conv_b,
         "x_proj.weight": x_w, "dt_proj.weight": dt_w, "dt_proj.bias": dt_b, "A_log": A_log, "D": Dp,
         "out_proj.weight": out_w, "out_proj.bias": out_b}
    p = {k: v for k, v in p.items() if v is not None}
    out, (new_conv, last) = orc.mixer_ref(p, hidden, conv_state, ssm_state, want_state=True)
    return out, (new_conv if want_conv_state else None), (last if want_ssm_state else None)


def add_norm(x, weight, bias, residual, eps, is_rms, prenorm, residual_in_fp32):
    return orc.add_norm_ref(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms)


def linear(x, weight, bias=None):
    return orc._linear(x, weight.to(x.dtype), bias)


def gate_blend(g1, g2, fwd, bwd):
    s = torch.sigmoid((g1.float() + g2.float()) if g2 is not None else g1.float())
    return (s * fwd.float() + (1.0 - s) * bwd.float()).to(fwd.dtype)


def mixer_step(w, xz, conv_state, ssm_state):
    # reference mamba_simple.py:466-494 between in_proj and out_proj; states in place
    p = _params(w)
    d_inner, _, d_conv = p["conv1d.weight"].shape
    d_state, dt_rank = p["A_log"].shape[1], p["dt_proj.weight"].shape[1]
    x, z = xz[:, :d_inner], xz[:, d_inner:]
    x = orc.causal_conv1d_update_ref(x, conv_state, p["conv1d.weight"].reshape(d_inner, d_conv),
                                     p.get("conv1d.bias"), "silu")
    x_db = orc._linear(x, p["x_proj.weight"])
    dt_low, Bm, Cm = torch.split(x_db, [dt_rank, d_state, d_state], dim=-1)
    dt = orc._linear(dt_low, p["dt_proj.weight"])
    A = -torch.exp(p["A_log"].float())
    return orc.selective_state_update_ref(ssm_state, x, dt, A, Bm, Cm, p["D"], z=z,
                                          dt_bias=p["dt_proj.bias"], dt_softplus=True)


def patchify(x, tubelet, ph, pw):
    # im2col of a Conv3d whose kernel equals its stride (reference videomamba.py:359-368)
    B, C, T, H, W = x.shape
    t, h, w = T // tubelet, H // ph, W // pw
    x = x[:, :, :t * tubelet, :h * ph, :w * pw]
    cols = x.reshape(B, C, t, tubelet, h, ph, w, pw).permute(0, 2, 4, 6, 1, 3, 5, 7)
    return cols.reshape(B * t * h * w, C * tubelet * ph * pw)


def embed_tokens(patches, spatial, temporal, cls_row=None):
    # reference videomamba.py:806-823: two adds in the model dtype, then CLS in front
    B, t, hw, D = patches.shape
    tokens = patches + spatial.to(patches.dtype).reshape(1, 1, hw, D)
    tokens = tokens + temporal.to(patches.dtype).reshape(1, t, 1, D)
    tokens = tokens.reshape(B, t * hw, D)
    if cls_row is not None:
        tokens = torch.cat((cls_row.to(patches.dtype).reshape(1, 1, D).expand(B, -1, -1), tokens), dim=1)
    return tokens


def pool_norm(x_vis, has_cls, groups, per, pool_type, ln_weight, ln_bias, eps):
    # reference videomamba.py:983-1063 with torch ops in the model dtype
    F = torch.nn.functional
    C = x_vis.shape[-1]
    ln = lambda t: F.layer_norm(t, (C,), ln_weight, ln_bias, eps)
    cls = x_vis[:, :1] if has_cls else None
    patches = x_vis[:, 1:] if has_cls else x_vis
    if pool_type == "cls":
        return ln(cls)
    avg = patches.reshape(x_vis.shape[0], groups, per, C).mean(2)
    if pool_type == "cls+avg":
        return ln(cls + avg)
    if pool_type == "cls_cat_avg":
        return ln(torch.cat([cls, avg], dim=1))
    return ln(avg)


def gather_rows(src, index):
    return src.gather(1, index.unsqueeze(-1).expand(-1, -1, src.shape[-1]))


def install(monkeypatch):
    import videomamba_b200.autograd as autograd_mod
    import videomamba_b200.mixer as mixer_mod
    import videomamba_b200.ops as ops

    monkeypatch.setattr(autograd_mod, "mixer_train", mixer_train)
    monkeypatch.setattr(ops, "mixer_fwd", mixer_fwd)
    monkeypatch.setattr(ops, "add_norm", add_norm)
    monkeypatch.setattr(ops, "linear", linear)
    monkeypatch.setattr(ops, "gate_blend", gate_blend)
    monkeypatch.setattr(ops, "patchify", patchify)
    monkeypatch.setattr(ops, "embed_tokens", embed_tokens)
    monkeypatch.setattr(ops, "mixer_step", mixer_step)
    monkeypatch.setattr(ops, "pool_norm", pool_norm)
    monkeypatch.setattr(ops, "gather_rows", gather_rows)
    monkeypatch.setattr(mixer_mod.Mamba, "_require_cuda", staticmethod(lambda t: None))
