"""GPU: gradients of the CUDA backward kernels (csrc/backward.cu, csrc/scan_bwd.cu, autograd.py) against
torch autograd through the CPU oracle on the same seeded inputs.

The reference obtains its gradients from autograd through the third-party operators
(scripts/check_streaming_state.py:47-60 differentiates through the carried state;
models/videomamba/videomamba.py:168-206 wraps the mixer in activation checkpointing).  Bars:
fp32 1e-4 relative (max|a-b| / max|b|; the kernels are true fp32, the difference is summation
order), bf16 3e-2."""
import pytest
import torch
import torch.nn.functional as F

import video_mamba
from oracle import videomamba_oracle as orc
from oracle.videomamba_oracle import rel_err
from video_mamba.mamba_simple import Mamba
from videomamba_b200 import autograd as ag
from videomamba_b200 import ops
from videomamba_b200.refiner import BiMambaRefinerBlock

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float32: 1e-4, torch.bfloat16: 3e-2}


def _leaf(t, dev=None):
    t = t.detach().clone()
    if dev is not None:
        t = t.to(dev)
    return t.requires_grad_(True)


def _check(got, want, tol, what=""):
    assert got is not None, f"{what}: no gradient"
    assert tuple(got.shape) == tuple(want.shape), (what, got.shape, want.shape)
    err = rel_err(got, want)
    assert err <= tol, (what, err)


def _gen(seed):
    return torch.Generator().manual_seed(seed)


# ---- projections -------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,shape", [(torch.float32, (3, 50, 40, 24)), (torch.float32, (1, 7, 24, 130)),
                                         (torch.bfloat16, (2, 517, 384, 768)), (torch.bfloat16, (2, 517, 768, 64)),
                                         (torch.bfloat16, (1, 99, 56, 24))])
def test_linear_backward(dtype, shape):
    B, L, K, N = shape
    g = _gen(K + N)
    x0 = torch.randn(B, L, K + 8, generator=g).to(dtype)
    w0 = (torch.randn(N, K, generator=g) * 0.1).to(dtype)
    b0 = torch.randn(N, generator=g).to(dtype)
    dy = torch.randn(B, L, N, generator=g).to(dtype)
    # reference: fp32 math on the same (rounded) operands
    xr, wr, br = _leaf(x0.float()), _leaf(w0.float()), _leaf(b0.float())
    (F.linear(xr[..., :K], wr, br) * dy.float()).sum().backward()
    xg, wg, bg = _leaf(x0, DEV), _leaf(w0, DEV), _leaf(b0, DEV)
    out = ops.linear(xg[..., :K], wg, bg)           # strided input view (row pitch K + 8)
    assert out.requires_grad
    (out.float() * dy.to(DEV).float()).sum().backward()
    tol = TOL[dtype]
    _check(xg.grad, xr.grad, tol, "dx")
    _check(wg.grad, wr.grad, tol, "dw")
    _check(bg.grad, br.grad, tol, "db")
    assert wg.grad.dtype == dtype and xg.grad.dtype == dtype


# ---- add + norm ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("is_rms", [True, False])
@pytest.mark.parametrize("mode", ["prenorm_res", "prenorm_first", "final_res", "plain"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_add_norm_backward(is_rms, mode, dtype):
    rows, dim = (5, 37), 192
    g = _gen(7)
    x0 = torch.randn(*rows, dim, generator=g).to(dtype)
    r0 = torch.randn(*rows, dim, generator=g)                    # fp32 residual stream
    w0 = (1 + 0.3 * torch.randn(dim, generator=g)).to(dtype)
    b0 = None if is_rms else (0.2 * torch.randn(dim, generator=g)).to(dtype)
    prenorm = mode.startswith("prenorm")
    with_res = mode.endswith("res")
    gy = torch.randn(*rows, dim, generator=g)
    gr = torch.randn(*rows, dim, generator=g)

    def run(dev, fn):
        x, w = _leaf(x0, dev), _leaf(w0, dev)
        b = None if b0 is None else _leaf(b0, dev)
        r = _leaf(r0, dev) if with_res else None
        out = fn(x, w, b, r)
        if prenorm:
            y, res = out
            loss = (y.float() * gy.to(y.device)).sum() + (res.float() * gr.to(y.device)).sum()
        else:
            loss = (out.float() * gy.to(out.device)).sum()
        loss.backward()
        return x.grad, w.grad, None if b is None else b.grad, None if r is None else r.grad

    want = run(None, lambda x, w, b, r: orc.add_norm_ref(x, w, b, r, 1e-5, prenorm, True, is_rms))
    got = run(DEV, lambda x, w, b, r: ops.add_norm(x, w, b, r, 1e-5, is_rms, prenorm, True))
    tol = TOL[dtype]
    for name, a, b in zip(("dx", "dw", "db", "dres"), got, want):
        if b is not None:
            _check(a, b, tol, name)


# ---- causal conv ---------------------------------------------------------------------------------------
def _conv_ref(x, w, b, cs, want_state):
    """token-major x (B, L, Di); the state handling of mamba_simple.py:381-404."""
    xin = x.transpose(1, 2)
    W = w.shape[-1]
    L = xin.shape[-1]
    if cs is not None:
        x_cat = torch.cat([cs.to(x.dtype), xin], dim=-1)
        y = orc.causal_conv1d_ref(x_cat, w.reshape(-1, W), b, "silu")[..., -L:]
        new = x_cat[..., -W:]
    else:
        y = orc.causal_conv1d_ref(xin, w.reshape(-1, W), b, "silu")
        new = F.pad(xin, (W - L, 0))
    return (y.transpose(1, 2), new) if want_state else y.transpose(1, 2)


@pytest.mark.parametrize("geom", [(2, 150, 40, 4, True, True), (2, 3, 40, 4, True, True), (1, 64, 130, 4, False, False),
                                  (2, 70, 40, 3, False, True), (3, 1, 24, 4, True, True), (1, 129, 768, 4, True, False),
                                  # >= 32 per-CTA partial rows (the tall reduction), channel tail of a 256-wide CTA
                                  (5, 300, 264, 4, True, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_backward(geom, dtype):
    B, L, Di, W, with_state, want_state = geom
    g = _gen(L)
    xz0 = torch.randn(B, L, 2 * Di, generator=g).to(dtype)          # conv reads the x half of xz in place
    w0 = (0.5 * torch.randn(Di, 1, W, generator=g)).to(dtype)
    b0 = (0.1 * torch.randn(Di, generator=g)).to(dtype)
    cs0 = torch.randn(B, Di, W, generator=g).to(dtype) if with_state else None
    gy = torch.randn(B, L, Di, generator=g)
    gs = torch.randn(B, Di, W, generator=g)

    def run(dev, fn):
        xz, w, b = _leaf(xz0, dev), _leaf(w0, dev), _leaf(b0, dev)
        cs = None if cs0 is None else _leaf(cs0, dev)
        out = fn(xz[..., :Di], w, b, cs)
        if want_state:
            y, new = out
            loss = (y.float() * gy.to(y.device)).sum() + (new.float() * gs.to(y.device)).sum()
        else:
            loss = (out.float() * gy.to(out.device)).sum()
        loss.backward()
        return xz.grad, w.grad, b.grad, None if cs is None else cs.grad

    want = run(None, lambda x, w, b, cs: _conv_ref(x, w, b, cs, want_state))
    got = run(DEV, lambda x, w, b, cs: ops.causal_conv1d_tokens(x, w, b, cs, want_state))
    tol = TOL[dtype]
    for name, a, b in zip(("dx", "dw", "db", "dstate"), got, want):
        if b is not None:
            _check(a, b, tol, name)


# ---- selective scan ------------------------------------------------------------------------------------
@pytest.mark.parametrize("geom", [
    # B, L, Di, N, z, D, h0, last
    (2, 37, 40, 16, True, True, True, True),
    (1, 100, 130, 8, True, True, False, False),
    (3, 8, 32, 16, False, False, True, True),
    (2, 1, 16, 16, True, True, True, True),
    (1, 65, 24, 4, False, True, False, True),
    # d_state 16, Di % 16 == 0: in bf16 these run scan_bwd_fast.cu (ragged tiles / sub-chunks, every option)
    (2, 37, 48, 16, True, True, True, True),
    (1, 100, 32, 16, True, False, False, False),
    (2, 16, 16, 16, True, True, False, True),
    (1, 3, 16, 16, True, True, True, True),
    (2, 65, 64, 16, False, True, True, False),
    (1, 131, 32, 16, True, True, True, True),
    # few units, long sequence: the reverse walk is split into segments (carry pass + chained g)
    (1, 700, 32, 16, True, True, True, True),
    (2, 1030, 16, 16, False, True, False, False),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_scan_backward(geom, dtype):
    B, L, Di, N, with_z, with_D, with_h0, with_last = geom
    g = _gen(L * 7 + Di)
    u0 = torch.randn(B, Di, L, generator=g).to(dtype)
    d0 = (0.5 * torch.randn(B, Di, L, generator=g)).to(dtype)
    A0 = -(torch.rand(Di, N, generator=g) * 2 + 0.2)
    B0 = torch.randn(B, N, L, generator=g).to(dtype)
    C0 = torch.randn(B, N, L, generator=g).to(dtype)
    D0 = torch.randn(Di, generator=g) if with_D else None
    z0 = torch.randn(B, Di, L, generator=g).to(dtype) if with_z else None
    bias0 = 0.3 * torch.randn(Di, generator=g)
    h00 = torch.randn(B, Di, N, generator=g) if with_h0 else None
    gy = torch.randn(B, Di, L, generator=g)
    gl = torch.randn(B, Di, N, generator=g)

    def run(dev, fn):
        names = ["u", "delta", "A", "B", "C", "D", "z", "bias", "h0"]
        leaves = [None if t is None else _leaf(t, dev) for t in (u0, d0, A0, B0, C0, D0, z0, bias0, h00)]
        u, d, A, Bm, Cm, D, z, bias, h0 = leaves
        out = fn(u, d, A, Bm, Cm, D, z, bias, h0)
        if with_last:
            y, last = out
            loss = (y.float() * gy.to(y.device)).sum() + (last.float() * gl.to(y.device)).sum()
        else:
            loss = (out.float() * gy.to(out.device)).sum()
        loss.backward()
        return {n: (None if t is None else t.grad) for n, t in zip(names, leaves)}

    want = run(None, lambda u, d, A, Bm, Cm, D, z, bias, h0: orc.selective_scan_ref(
        u, d, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=True, initial_state=h0,
        return_last_state=with_last))
    got = run(DEV, lambda u, d, A, Bm, Cm, D, z, bias, h0: ops.selective_scan_fn(
        u, d, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=True, return_last_state=with_last,
        initial_state=h0))
    tol = TOL[dtype]
    for name in want:
        if want[name] is not None:
            _check(got[name], want[name], tol, name)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_scan_backward_bf16_kernels_against_fp32_kernels_on_random_geometries(seed):
    """The bf16 production kernels (csrc/scan_bwd_fast.cu) against the true-fp32 kernels (csrc/scan_bwd.cu) on the
    same bf16-representable inputs: lengths around the 4-token sub-chunk / 16-token tile / segment boundaries,
    B / C columns behind dt_rank 12 / 24 / 36 and in a plain (B | C) tensor, every option combination."""
    import random
    rng = random.Random(seed)
    bf = torch.bfloat16
    for case in range(12):
        B = rng.choice([1, 2, 3])
        L = rng.choice([1, 2, 3, 4, 5, 15, 16, 17, 33, 63, 64, 65, 100, 385, 400, 777])
        Di = rng.choice([16, 32, 48, 80])
        R = rng.choice([12, 24, 36, 0])
        N = 16
        cols = (R + 2 * N + 15) // 16 * 16 if R else 2 * N
        with_z, with_h0, with_D, with_last = (rng.random() < 0.7 for _ in range(4))
        g = torch.Generator(device=DEV).manual_seed(100 * seed + case)
        rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
        u, z, dout = rn(B, L, Di).to(bf), rn(B, L, Di).to(bf), rn(B, L, Di).to(bf)
        delta = (0.5 * rn(B, L, Di) - 2).to(bf)
        bc = rn(B, L, cols).to(bf)
        A2 = -(torch.rand(Di, N, device=DEV, generator=g) * 8 + 0.1) * 1.4427
        D = rn(Di) if with_D else None
        bias = 0.3 * rn(Di)
        h0 = rn(B, Di, N) if with_h0 else None
        dlast = rn(B, Di, N) if with_last else None
        zz = z if with_z else None
        f32 = lambda t: None if t is None else t.float()
        fast = ag._scan_bwd(u, delta, A2, bc, R, R + N, N, D, zz, bias, True, h0, dout, dlast, with_h0)
        ref = ag._scan_bwd(f32(u), f32(delta), A2, f32(bc), R, R + N, N, D, f32(zz), bias, True, h0, f32(dout), dlast,
                           with_h0)
        for name, a, b in zip(("du", "ddelta", "dz", "dbc", "dA", "dD", "dbias", "dh0"), fast, ref):
            if a is not None:
                assert rel_err(a, b) <= 2e-2, (case, B, L, Di, R, name, rel_err(a, b))


def test_scan_backward_is_deterministic_and_chunk_invariant():
    """No atomics: two runs are bit-identical; splitting the sequence in two calls with the state (and its
    gradient) carried between them gives the gradients of the single call."""
    torch.manual_seed(3)
    B, L, Di, N = 2, 45, 64, 16
    mk = lambda *s: torch.randn(*s, device=DEV)
    u, d, z = mk(B, L, Di), 0.5 * mk(B, L, Di), mk(B, L, Di)
    bc = mk(B, L, 2 * N)
    A = -(torch.rand(Di, N, device=DEV) + 0.3)
    Dp, bias = mk(Di), 0.2 * mk(Di)
    gy = mk(B, L, Di)

    def run(split):
        leaves = [t.clone().requires_grad_(True) for t in (u, d, A, bc, Dp, z, bias)]
        uu, dd, AA, bb, DD, zz, bi = leaves
        if split is None:
            y, _ = ag.ScanFn.apply(uu, dd, AA, bb, 0, N, N, DD, zz, bi, True, None, False)
        else:
            y1, h = ag.ScanFn.apply(uu[:, :split], dd[:, :split], AA, bb[:, :split], 0, N, N, DD, zz[:, :split],
                                    bi, True, None, True)
            y2, _ = ag.ScanFn.apply(uu[:, split:], dd[:, split:], AA, bb[:, split:], 0, N, N, DD, zz[:, split:],
                                    bi, True, h, False)
            y = torch.cat([y1, y2], dim=1)
        (y * gy).sum().backward()
        return [t.grad for t in leaves]

    a, b = run(None), run(None)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    c = run(19)
    for x, y in zip(a, c):
        assert rel_err(y, x) <= 1e-5


# ---- mixer with the state carried between chunks (reference scripts/check_streaming_state.py) ----------
@pytest.mark.parametrize("dtype,d_model,d_state", [(torch.float32, 64, 8), (torch.float32, 96, 16),
                                                   (torch.bfloat16, 384, 16), (torch.bfloat16, 192, 16),
                                                   (torch.bfloat16, 576, 16)])
def test_mixer_backward_streaming_state(dtype, d_model, d_state):
    torch.manual_seed(0)
    mx = Mamba(d_model=d_model, d_state=d_state, d_conv=4, expand=2, use_fast_path=False).to(dtype).to(DEV)
    with torch.no_grad():
        mx.A_log.add_(0.1 * torch.randn_like(mx.A_log))
    B, L, split = 2, 50, 21
    x0 = torch.randn(B, L, d_model, generator=_gen(5)).to(dtype)
    gy = torch.randn(B, L, d_model, generator=_gen(6))
    gs = torch.randn(B, mx.d_inner, d_state, generator=_gen(7))

    # CUDA: full sequence, and two chunks with the state carried
    x = _leaf(x0, DEV)
    out_full = mx(x)
    x1, x2 = x[:, :split], x[:, split:]
    out1, state = mx(x1, return_state=True)
    out2, (cs2, ss2) = mx(x2, state=state, return_state=True)
    out_chunked = torch.cat([out1, out2], dim=1)
    torch.testing.assert_close(out_full.float(), out_chunked.float(), rtol=1e-4 if dtype == torch.float32 else 3e-2,
                               atol=1e-4 if dtype == torch.float32 else 3e-2)
    ((out_chunked.float() * gy.to(DEV)).sum() + (ss2 * gs.to(DEV)).sum()).backward()
    assert x.grad is not None
    got = {n: p.grad.detach().cpu() for n, p in mx.named_parameters()}
    got_x = x.grad.detach().cpu()

    # oracle: the same two chunks through mixer_ref on the CPU
    p = {n: _leaf(t.detach().cpu()) for n, t in mx.state_dict().items()}
    xo = _leaf(x0)
    o1, st = orc.mixer_ref(p, xo[:, :split], want_state=True)
    o2, (c2, s2) = orc.mixer_ref(p, xo[:, split:], conv_state=st[0], ssm_state=st[1], want_state=True)
    ((torch.cat([o1, o2], dim=1).float() * gy).sum() + (s2 * gs).sum()).backward()
    tol = 2e-4 if dtype == torch.float32 else 4e-2
    _check(got_x, xo.grad, tol, "dx")
    for n in p:
        _check(got[n], p[n].grad, tol, n)


@pytest.mark.parametrize("dtype,d_model", [(torch.float32, 128), (torch.bfloat16, 384), (torch.bfloat16, 576)])
def test_mixer_training_forward_matches_inference_path(dtype, d_model):
    """The training-mode mixer (op by op, differentiable; bf16 production widths run the fused scan in the
    forward) and the fused inference entry point agree -- bit for bit where both run the fused scan with SiLU(z)
    inside it; the inference default (SiLU(z) in the in_proj epilogue, the training forward keeps the raw z for its
    backward) moves one bf16 rounding of the gate: same states, outputs within that rounding."""
    torch.manual_seed(1)
    mx = Mamba(d_model=d_model, use_fast_path=False).to(dtype).to(DEV)
    x = torch.randn(2, 70, d_model, device=DEV).to(dtype)
    a, (ca, sa) = mx(x, return_state=True)
    with torch.no_grad():
        b, (cb, sb) = mx(x, return_state=True)
    assert a.requires_grad and not b.requires_grad
    if dtype == torch.bfloat16:
        assert torch.equal(ca, cb) and torch.equal(sa, sb)
        assert rel_err(a, b) <= 8e-3, rel_err(a, b)
        mx.gate_in_proj = False
        with torch.no_grad():
            c, (cc, sc) = mx(x, return_state=True)
        assert torch.equal(a, c) and torch.equal(ca, cc) and torch.equal(sa, sc)
    else:
        assert rel_err(a, b) <= 1e-5 and rel_err(sa, sb) <= 1e-5


@pytest.mark.parametrize("geom", [(2, 70, 384), (1, 700, 128), (40, 50, 768), (3, 131, 576)])
def test_scan_states_saved_by_the_forward_equal_a_recomputing_backward(geom, monkeypatch):
    """The fused training forward writes the state before every 4-token group (one-warp kernel, two-warp
    kernel, sequence split); the backward that reads those records and the backward that recomputes them with
    its own pass agree (the records differ only by delta being rounded to bf16 in the recomputation)."""
    B, L, d_model = geom
    torch.manual_seed(4)
    mx = Mamba(d_model=d_model, use_fast_path=False).to(torch.bfloat16).to(DEV)
    x0 = torch.randn(B, L, d_model, device=DEV).to(torch.bfloat16)
    gy = torch.randn(B, L, d_model, device=DEV)
    gs = torch.randn(B, mx.d_inner, mx.d_state, device=DEV)

    def run(save):
        monkeypatch.setattr(ag, "SAVE_SCAN_STATES", save)
        mx.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        out, (_, ss) = mx(x, return_state=True)
        ((out.float() * gy).sum() + (ss * gs).sum()).backward()
        return out.detach(), {"x": x.grad, **{n: p.grad.clone() for n, p in mx.named_parameters()}}

    out_a, ga = run(True)
    out_b, gb = run(False)
    assert torch.equal(out_a, out_b)
    for n in ga:
        assert rel_err(ga[n], gb[n]) <= 1e-2, (n, rel_err(ga[n], gb[n]))


# ---- whole model -----------------------------------------------------------------------------------------
def _cfg(**over):
    cfg = dict(img_size=32, patch_size=16, depth=2, embed_dim=64, kernel_size=1, num_frames=2,
               norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True, residual_in_fp32=True,
               pool_type="cls+avg", add_pool_norm=True)
    cfg.update(over)
    return cfg


def _build(cfg, sd, dtype, **over):
    kw = dict(img_size=cfg["img_size"], patch_size=cfg["patch_size"], depth=cfg["depth"],
              embed_dim=cfg["embed_dim"], channels=3, ssm_cfg={"use_fast_path": False},
              rms_norm=cfg["rms_norm"], fused_add_norm=cfg["fused_add_norm"],
              residual_in_fp32=cfg["residual_in_fp32"], kernel_size=cfg["kernel_size"],
              num_frames=cfg["num_frames"], pool_type=cfg["pool_type"])
    kw.update(over)
    m = video_mamba.PretrainVideoMamba(**kw).to(dtype)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("variant", ["rms_fused", "ln_unfused", "masked"])
def test_model_backward_matches_oracle_autograd(variant):
    cfg = _cfg(rms_norm=variant != "ln_unfused", fused_add_norm=variant != "ln_unfused")
    sd = orc.synthetic_state_dict(cfg, seed=4, dtype=torch.float32, perturbed=True)
    x = torch.rand(2, 3, 2, 32, 32, generator=_gen(9))
    mask = None
    if variant == "masked":
        mask = torch.zeros(2, 1 + 2 * 4, dtype=torch.bool)
        mask[0, [2, 5]] = True
        mask[1, [3, 8]] = True
    p = {k: _leaf(v) for k, v in sd.items()}
    vis_o, pool_o = orc.OracleVideoMamba(cfg, p).forward(x, mask=mask)
    gv = torch.randn(vis_o.shape, generator=_gen(10))
    gp = torch.randn(pool_o.shape, generator=_gen(11))
    ((vis_o * gv).sum() + (pool_o * gp).sum()).backward()

    m = _build(cfg, sd, torch.float32).train()
    vis, pool = m(x.to(DEV), mask=None if mask is None else mask.to(DEV))
    assert rel_err(vis, vis_o) <= 1e-5 and rel_err(pool, pool_o) <= 1e-5
    ((vis * gv.to(DEV)).sum() + (pool * gp.to(DEV)).sum()).backward()
    checked = 0
    for name, prm in m.named_parameters():
        want = p[name].grad
        if want is None:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, name
            continue
        _check(prm.grad, want, 5e-4, name)
        checked += 1
    assert checked >= 25


def test_checkpoint_wrapper_recomputes_the_mixer():
    """use_checkpoint (videomamba.py:168-206): same outputs and gradients, the mixer is recomputed."""
    cfg = _cfg()
    sd = orc.synthetic_state_dict(cfg, seed=5, dtype=torch.float32, perturbed=True)
    x = torch.rand(2, 3, 2, 32, 32, generator=_gen(12)).to(DEV)
    grads = []
    for ckpt in (False, True):
        m = _build(cfg, sd, torch.float32, use_checkpoint=ckpt, checkpoint_num=2).train()
        vis, pool = m(x)
        ((vis * torch.randn(vis.shape, generator=_gen(20)).to(DEV)).sum() + pool.sum()).backward()
        grads.append({n: q.grad.clone() for n, q in m.named_parameters() if q.grad is not None})
    assert grads[0].keys() == grads[1].keys()
    for n in grads[0]:
        assert rel_err(grads[1][n], grads[0][n]) <= 1e-6, n


def test_streaming_model_backward_through_carried_state():
    """Two chunks with the full state carried: the gradient of the second chunk's output reaches the first
    chunk's input through (conv_state, ssm_state) -- "streaming training"."""
    cfg = _cfg(pool_type="avg", num_frames=4)
    sd = orc.synthetic_state_dict(cfg, seed=6, dtype=torch.float32, perturbed=True)
    m = _build(cfg, sd, torch.float32).train()
    x = torch.rand(1, 3, 4, 32, 32, generator=_gen(13))
    p = {k: _leaf(v) for k, v in sd.items()}
    oracle = orc.OracleVideoMamba(cfg, p)
    st0 = [(torch.zeros(1, 128, 4), torch.zeros(1, 128, 16)) for _ in range(cfg["depth"])]
    xo = _leaf(x)
    _, _, s1 = oracle.forward(xo[:, :, :2], ssm_state=st0, temporal_pos_offset=0)
    v2, p2, _ = oracle.forward(xo[:, :, 2:], ssm_state=s1, temporal_pos_offset=2)
    # random cotangents: sum(vis^2) is (nearly) invariant under the final RMSNorm, its gradient is round-off
    gv, gp = torch.randn(v2.shape, generator=_gen(14)), torch.randn(p2.shape, generator=_gen(15))
    ((v2 * gv).sum() + (p2 * gp).sum()).backward()

    xg = _leaf(x, DEV)
    st = m.allocate_state(1, device=DEV)
    _, _, g1 = m(xg[:, :, :2], ssm_state=st, temporal_pos_offset=0)
    gv2, gp2, _ = m(xg[:, :, 2:], ssm_state=g1, temporal_pos_offset=2)
    ((gv2 * gv.to(DEV)).sum() + (gp2 * gp.to(DEV)).sum()).backward()
    assert float(xg.grad[:, :, :2].abs().max()) > 0        # reached the first chunk through the state
    _check(xg.grad[:, :, :2], xo.grad[:, :, :2], 5e-4, "dx of the first chunk (through the state only)")
    _check(xg.grad, xo.grad, 5e-4, "dx")
    for name, prm in m.named_parameters():
        if p[name].grad is not None:
            _check(prm.grad, p[name].grad, 5e-4, name)


def test_refiner_backward():
    torch.manual_seed(2)
    blk = BiMambaRefinerBlock(64).to(DEV)
    x = torch.randn(2, 3, 5, 64, device=DEV, requires_grad=True)     # 4-D input: frame-axis flip
    out, state = blk(x)
    go = torch.randn(out.shape, generator=_gen(21))
    ((out * go.to(DEV)).sum() + state[1].sum()).backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    sd = {k: v.detach().cpu() for k, v in blk.state_dict().items()}
    p = {k: _leaf(v) for k, v in sd.items()}
    xo = _leaf(x.detach().cpu())
    want, st = orc.refiner_ref(p, xo)
    ((want * go).sum() + st[1].sum()).backward()
    assert rel_err(out, want) <= 1e-5
    _check(x.grad, xo.grad, 5e-4, "dx")
    for name, prm in blk.named_parameters():
        if p[name].grad is not None:
            _check(prm.grad, p[name].grad, 5e-4, name)


@pytest.mark.parametrize("shape", [(3137 * 2, 768, 384), (1000, 64, 768), (777, 768, 24), (40, 1536, 384), (5000, 56, 768),
                                   (20001, 1536, 384), (9000, 384, 768), (4097, 576, 1152), (6000, 2304, 576),
                                   (300, 128, 512), (8191, 1152, 40), (2000, 256, 128), (1500, 48, 136)])
def test_linear_wgrad_kernel(shape):
    """vmb_linear_wgrad against fp32 matmul, including strided operand views and ragged token / tile edges:
    the tcgen05 kernel (csrc/wgrad_tc.cu: N >= 128, M >= 256; MN-major operands read in place) and the mma.sync
    kernel (csrc/wgrad.cu) for the rest."""
    M, N, K = shape
    g = _gen(M + N)
    dy = torch.randn(M, N + 8, generator=g).to(torch.bfloat16).to(DEV)[:, :N]      # row pitch N + 8
    x = torch.randn(M, 2 * K, generator=g).to(torch.bfloat16).to(DEV)[:, K:]      # the upper half of a wider buffer
    want = dy.float().t() @ x.float()
    for out_dtype in (torch.float32, torch.bfloat16):
        got = ag.linear_wgrad(dy, x, out_dtype)
        assert got.dtype == out_dtype and got.shape == (N, K)
        assert rel_err(got, want) <= (1e-5 if out_dtype == torch.float32 else 5e-3)
    again = ag.linear_wgrad(dy, x, torch.float32)
    assert torch.equal(again, ag.linear_wgrad(dy, x, torch.float32))               # deterministic
