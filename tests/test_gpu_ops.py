"""GPU: operator-level parity of the libvmb200 kernels (called through the C ABI via
videomamba_b200.ops) against the CPU oracle on the same seeded inputs.

Bars (north star): fp32 1e-5, bf16 2e-2, relative = max|a-b| / max|b|."""
import pytest
import torch

from oracle import videomamba_oracle as orc
from oracle.videomamba_oracle import rel_err
from videomamba_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
DTYPES = [torch.float32, torch.bfloat16]


def _tol(dt):
    return 2e-2 if dt == torch.bfloat16 else 1e-5


def _rand(gen, *shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).to(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rms", [True, False])
@pytest.mark.parametrize("dim", [16, 192, 384, 576, 1152])
def test_add_norm(dtype, rms, dim):
    gen = torch.Generator().manual_seed(dim + rms)
    x = _rand(gen, 3, 37, dim, dtype=dtype)
    res = _rand(gen, 3, 37, dim)
    w = (1 + 0.1 * torch.randn(dim, generator=gen)).to(dtype)
    b = None if rms else (0.1 * torch.randn(dim, generator=gen)).to(dtype)
    for residual in (None, res):
        for prenorm in (True, False):
            want = orc.add_norm_ref(x, w, b, residual, 1e-5, prenorm, True, rms)
            got = ops.add_norm(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV),
                               None if residual is None else residual.to(DEV), 1e-5, rms, prenorm,
                               True)
            if prenorm:
                assert got[1].dtype == torch.float32
                assert rel_err(got[0], want[0]) <= _tol(dtype)
                assert rel_err(got[1], want[1]) <= 1e-6
            else:
                assert got.dtype == dtype and rel_err(got, want) <= _tol(dtype)


def test_add_norm_empty_and_dtype_rules():
    x = torch.randn(0, 16, device=DEV)
    assert ops.add_norm(x, torch.ones(16, device=DEV), None, None, 1e-5, True, False, True).shape == (0, 16)
    xb = torch.randn(2, 5, 16, device=DEV, dtype=torch.bfloat16)
    y, r = ops.add_norm(xb, torch.ones(16, device=DEV, dtype=torch.bfloat16), None, None, 1e-5, True,
                        True, False)
    assert r.dtype == torch.bfloat16 and torch.equal(r, xb)   # no residual, same dtype: sum is x
    y, r = ops.add_norm(xb, torch.ones(16, device=DEV, dtype=torch.bfloat16), None,
                        torch.zeros(2, 5, 16, device=DEV, dtype=torch.bfloat16), 1e-5, True, True, True)
    assert r.dtype == torch.bfloat16                          # keeps the incoming residual dtype


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mnk", [(5, 3, 1), (37, 44, 192), (300, 1536, 384), (4500, 56, 768),
                                 (260, 384, 24), (129, 768, 12), (64, 16, 16)])
def test_linear(dtype, mnk):
    M, N, K = mnk
    gen = torch.Generator().manual_seed(M * N + K)
    a = _rand(gen, M, K, dtype=dtype)
    w = _rand(gen, N, K, dtype=dtype, scale=K ** -0.5)
    bias = _rand(gen, N, dtype=dtype)
    got = ops.linear(a.to(DEV), w.to(DEV), bias.to(DEV))
    want = orc._linear(a, w, bias)
    assert got.dtype == dtype and rel_err(got, want) <= _tol(dtype)
    # strided input view (dt_low slice of x_dbl) and no bias
    wide = _rand(gen, M, K + 9, dtype=dtype)
    got = ops.linear(wide.to(DEV)[:, :K], w.to(DEV))
    assert rel_err(got, orc._linear(wide[:, :K], w)) <= _tol(dtype)


@pytest.mark.parametrize("mnk", [(12837, 1536, 384), (8197, 64, 768), (5000, 384, 768),
                                 (4096, 192, 384), (3000, 576, 1152), (2000, 80, 1152),
                                 (999, 2304, 576), (128, 768, 192), (70, 48, 384), (4100, 768, 24)])
def test_tensor_core_projection_exact_integers(mnk):
    """bf16 tcgen05 projection at the production widths (Tiny / Small / Middle in_proj, x_proj,
    out_proj, patch embed, dt_proj): operands are small integers, so every fp32 partial sum is
    exact and the result must EQUAL the fp32 product rounded once to bf16 -- any mis-indexed tile,
    swizzle or k-block shows up as a mismatch, not as noise."""
    M, N, K = mnk
    gen = torch.Generator().manual_seed(M + N + K)
    a = torch.randint(-2, 3, (M, K), generator=gen).to(torch.bfloat16).to(DEV)
    w = torch.randint(-2, 3, (N, K), generator=gen).to(torch.bfloat16).to(DEV)
    bias = torch.randint(-4, 5, (N,), generator=gen).to(torch.bfloat16).to(DEV)
    for b in (None, bias):
        got = ops.linear(a, w, b)
        acc = a.double() @ w.double().t()            # exact (products and sums of small integers)
        if b is not None:
            acc = acc + b.double()
        want = acc.to(torch.float32).to(torch.bfloat16)
        assert torch.equal(got, want), (mnk, (got.float() - want.float()).abs().max().item())
    # output written into a wider buffer (x_dbl pitch) must not touch the padding columns
    from videomamba_b200 import _lib
    lib = _lib.load()
    out = torch.full((M, N + 16), 7.0, dtype=torch.bfloat16, device=DEV)
    rc = lib.vmb_linear_fwd(ops._p(a), K, ops._p(w), K, None, ops._p(out), N + 16, M, N, K,
                            _lib.VMB_BF16, ops._stream(a))
    _lib.check(rc, "vmb_linear_fwd")
    want = (a.double() @ w.double().t()).to(torch.float32).to(torch.bfloat16)
    assert torch.equal(out[:, :N], want) and bool((out[:, N:] == 7.0).all())


@pytest.mark.parametrize("m", [512, 513, 640, 767, 1025, 2049])
@pytest.mark.parametrize("nk", [(192, 16), (256, 72), (384, 384), (512, 40), (768, 136)])
def test_projection_cta_pair_tile_edges(m, nk):
    """The CTA-pair kernel (cta_group::2, 256-row tiles) at its edges: the second CTA of the last pair
    owns 0, 1, 127 or 128 valid rows, K is not a multiple of the 64-wide k-block, N takes one to four
    tiles of 192 / 256.  Integer operands: the result must equal the exact product rounded once."""
    n, k = nk
    gen = torch.Generator().manual_seed(m * 7 + n + k)
    a = torch.randint(-2, 3, (m, k), generator=gen).to(torch.bfloat16).to(DEV)
    w = torch.randint(-2, 3, (n, k), generator=gen).to(torch.bfloat16).to(DEV)
    bias = torch.randint(-4, 5, (n,), generator=gen).to(torch.bfloat16).to(DEV)
    guard = torch.full((m + 300, n), 5.0, dtype=torch.bfloat16, device=DEV)   # rows beyond M stay untouched
    from videomamba_b200 import _lib
    lib = _lib.load()
    rc = lib.vmb_linear_fwd(ops._p(a), k, ops._p(w), k, ops._p(bias), ops._p(guard), n, m, n, k,
                            _lib.VMB_BF16, ops._stream(a))
    _lib.check(rc, "vmb_linear_fwd")
    want = (a.double() @ w.double().t() + bias.double()).to(torch.float32).to(torch.bfloat16)
    assert torch.equal(guard[:m], want)
    assert bool((guard[m:] == 5.0).all())


def test_tensor_core_projection_random_values():
    """Random bf16 operands against a float64 product of the same bf16 values: the only error left
    is the fp32 accumulation order and the final rounding (<= 1 bf16 ulp of the result)."""
    gen = torch.Generator().manual_seed(5)
    for M, N, K in [(6000, 1536, 384), (6000, 384, 768), (6000, 64, 768)]:
        a = torch.randn(M, K, generator=gen).to(torch.bfloat16).to(DEV)
        w = (torch.randn(N, K, generator=gen) * K ** -0.5).to(torch.bfloat16).to(DEV)
        got = ops.linear(a, w).double()
        want = a.double() @ w.double().t()
        err = (got - want).abs()
        assert float((err / (want.abs() + 1e-2)).max()) <= 2.0 ** -8, (M, N, K)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("geom", [(2, 3, 4, 32, 32, 1, 16, 16), (1, 3, 5, 30, 50, 2, 14, 12),
                                  (2, 1, 2, 8, 9, 1, 4, 3), (0, 3, 2, 16, 16, 1, 16, 16)])
def test_patchify_is_the_conv3d_im2col(dtype, geom):
    """Pure data movement: bit exact against the reshape/permute of the strided Conv3d
    (reference PatchEmbed, videomamba.py:359-368), including cropped remainders."""
    B, C, T, H, W, k, ph, pw = geom
    gen = torch.Generator().manual_seed(sum(geom))
    x = _rand(gen, B, C, T, H, W, dtype=dtype)
    t, h, w = T // k, H // ph, W // pw
    want = x[:, :, :t * k, :h * ph, :w * pw].reshape(B, C, t, k, h, ph, w, pw) \
        .permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B * t * h * w, C * k * ph * pw)
    got = ops.patchify(x.to(DEV), k, ph, pw)
    assert got.shape == want.shape
    assert torch.equal(got.cpu(), want)
    if B:       # and the projection over it equals the Conv3d
        wgt = _rand(gen, 8, C, k, ph, pw, dtype=dtype, scale=0.1)
        conv = torch.nn.functional.conv3d(x.float(), wgt.float(), stride=(k, ph, pw))
        proj = (got.float().cpu() @ wgt.float().reshape(8, -1).T).reshape(B, t, h, w, 8).permute(0, 4, 1, 2, 3)
        assert rel_err(proj, conv) <= 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("geom", [(2, 3, 5, 16), (1, 1, 196, 384), (3, 2, 4, 6), (2, 0, 4, 8)])
@pytest.mark.parametrize("with_cls", [True, False])
def test_embed_tokens_matches_two_adds_and_cat(dtype, geom, with_cls):
    """Position embeddings + CLS placement (reference videomamba.py:806-823): same rounding points as
    the two adds in the model dtype, bit exact; CLS only when given (continuation chunks omit it)."""
    B, t, hw, D = geom
    gen = torch.Generator().manual_seed(sum(geom) + with_cls)
    patches = _rand(gen, B, t, hw, D, dtype=dtype)
    spatial, temporal = _rand(gen, hw, D, dtype=dtype), _rand(gen, t, D, dtype=dtype)
    cls_row = _rand(gen, D, dtype=dtype) if with_cls else None
    want = (patches + spatial.reshape(1, 1, hw, D)) + temporal.reshape(1, t, 1, D)
    want = want.reshape(B, t * hw, D)
    if with_cls:
        want = torch.cat((cls_row.reshape(1, 1, D).expand(B, -1, -1), want), dim=1)
    got = ops.embed_tokens(patches.to(DEV), spatial.to(DEV), temporal.to(DEV),
                           None if cls_row is None else cls_row.to(DEV))
    assert got.shape == want.shape
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("geom", [(2, 70, 768, 4), (1, 200, 1152, 4), (3, 5, 16, 2), (1, 2, 24, 4), (2, 130, 384, 3),
                                  (1, 1, 8, 4)])
@pytest.mark.parametrize("reverse", [False, True])
def test_causal_conv1d(dtype, geom, reverse):
    B, L, Di, W = geom
    gen = torch.Generator().manual_seed(L * Di + W)
    x = _rand(gen, B, L, Di, dtype=dtype)
    w = _rand(gen, Di, W, dtype=dtype, scale=0.5)
    b = _rand(gen, Di, dtype=dtype, scale=0.5)
    cs = _rand(gen, B, Di, W, dtype=dtype)
    for state in (None, cs):
        xl = torch.flip(x, dims=[1]) if reverse else x          # logical order
        xcm = xl.transpose(1, 2)
        if state is not None:
            cat = torch.cat([state, xcm], dim=-1)
            want = orc.causal_conv1d_ref(cat, w, b, "silu")[..., -L:]
            want_state = cat[..., -W:]
        else:
            want = orc.causal_conv1d_ref(xcm, w, b, "silu")
            want_state = torch.nn.functional.pad(xcm, (W - L, 0))
        want = want.transpose(1, 2)
        if reverse:
            want = torch.flip(want, dims=[1])
        got, got_state = ops.causal_conv1d_tokens(
            x.to(DEV), w.to(DEV), b.to(DEV), None if state is None else state.to(DEV),
            want_state=True, reverse=reverse)
        assert rel_err(got, want) <= _tol(dtype)
        assert got_state.shape == (B, Di, W)
        assert torch.equal(got_state.cpu(), want_state.contiguous())   # pre-conv inputs: bit exact


@pytest.mark.parametrize("L", [16, 17, 39, 40, 41, 79, 81, 163])
@pytest.mark.parametrize("reverse", [False, True])
def test_causal_conv1d_ring_kernel_chunk_edges(L, reverse):
    """bf16 production kernel (ring-staged, 40 tokens per thread, groups of 4 rows): lengths around the
    chunk and group boundaries, x as the strided first half of xz, with and without carried state.
    Chunked execution (state carried across two calls) must equal the single pass bit for bit."""
    B, Di, W = 2, 768, 4
    bf = torch.bfloat16
    gen = torch.Generator().manual_seed(L)
    xz = _rand(gen, B, L, 2 * Di, dtype=bf)
    w = _rand(gen, Di, W, dtype=bf, scale=0.5)
    b = _rand(gen, Di, dtype=bf, scale=0.5)
    cs = _rand(gen, B, Di, W, dtype=bf)
    x = xz[..., :Di]
    for state in (None, cs):
        xl = torch.flip(x, dims=[1]) if reverse else x
        xcm = xl.transpose(1, 2)
        if state is not None:
            cat = torch.cat([state, xcm], dim=-1)
            want = orc.causal_conv1d_ref(cat, w, b, "silu")[..., -L:]
            want_state = cat[..., -W:]
        else:
            want = orc.causal_conv1d_ref(xcm, w, b, "silu")
            want_state = torch.nn.functional.pad(xcm, (W - L, 0))[..., -W:]
        want = want.transpose(1, 2)
        if reverse:
            want = torch.flip(want, dims=[1])
        got, got_state = ops.causal_conv1d_tokens(
            xz.to(DEV)[..., :Di], w.to(DEV), b.to(DEV), None if state is None else state.to(DEV),
            want_state=True, reverse=reverse)
        assert rel_err(got, want) <= _tol(bf)
        assert torch.equal(got_state.cpu(), want_state.contiguous())
    if not reverse and L >= 32:
        xd = xz.to(DEV)[..., :Di]
        full, st_full = ops.causal_conv1d_tokens(xd, w.to(DEV), b.to(DEV), None, want_state=True)
        cut = L // 2 + 1
        y0, s0 = ops.causal_conv1d_tokens(xd[:, :cut], w.to(DEV), b.to(DEV), None, want_state=True)
        y1, s1 = ops.causal_conv1d_tokens(xd[:, cut:], w.to(DEV), b.to(DEV), s0, want_state=True)
        assert torch.equal(torch.cat([y0, y1], dim=1), full) and torch.equal(s1, st_full)


@pytest.mark.parametrize("dtype", DTYPES)
def test_causal_conv1d_strided_views_and_dropin_signature(dtype):
    gen = torch.Generator().manual_seed(3)
    xz = _rand(gen, 2, 33, 2 * 64, dtype=dtype)
    w = _rand(gen, 64, 4, dtype=dtype, scale=0.5)
    got = ops.causal_conv1d_tokens(xz.to(DEV)[..., :64], w.to(DEV), None)
    want = orc.causal_conv1d_ref(xz[..., :64].transpose(1, 2), w, None, "silu").transpose(1, 2)
    assert rel_err(got, want) <= _tol(dtype)
    xcm = _rand(gen, 2, 24, 19, dtype=dtype)        # (B, D, L) as the reference passes it
    w2 = _rand(gen, 24, 4, dtype=dtype, scale=0.5)
    got = ops.causal_conv1d_fn(xcm.to(DEV), w2.to(DEV), None, activation=None)
    assert got.shape == xcm.shape
    assert rel_err(got, orc.causal_conv1d_ref(xcm, w2, None, None)) <= _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("geom", [(2, 24, 37, 16), (1, 768, 200, 16), (3, 16, 5, 4), (2, 40, 64, 8),
                                  (1, 130, 33, 1)])
def test_selective_scan_fn_dropin(dtype, geom):
    B, D, L, N = geom
    gen = torch.Generator().manual_seed(D + L)
    u = _rand(gen, B, D, L, dtype=dtype)
    delta = _rand(gen, B, D, L, dtype=dtype)
    A = -torch.exp(torch.log(torch.arange(1, N + 1).float()).repeat(D, 1)
                   + 0.1 * torch.randn(D, N, generator=gen))
    Bm = _rand(gen, B, N, L, dtype=dtype)
    Cm = _rand(gen, B, N, L, dtype=dtype)
    Dp = torch.randn(D, generator=gen)
    z = _rand(gen, B, D, L, dtype=dtype)
    bias = torch.randn(D, generator=gen) - 2.0
    h0 = torch.randn(B, D, N, generator=gen)
    for init in (None, h0):
        want, want_last = orc.selective_scan_ref(u, delta, A, Bm, Cm, Dp, z, bias, True, init, True)
        got, last = ops.selective_scan_fn(
            u.to(DEV), delta.to(DEV), A.to(DEV), Bm.to(DEV), Cm.to(DEV), Dp.to(DEV), z=z.to(DEV),
            delta_bias=bias.to(DEV), delta_softplus=True, return_last_state=True,
            initial_state=None if init is None else init.to(DEV))
        assert got.shape == (B, D, L) and last.dtype == torch.float32
        assert rel_err(got, want) <= _tol(dtype)
        assert rel_err(last, want_last) <= _tol(dtype)
    # no z / no D / no softplus / no bias (delta is then a step size: keep it positive, a
    # negative one turns the decay into exponential growth and overflows in both)
    dpos = (delta.float().abs() * 0.1).to(dtype)
    want = orc.selective_scan_ref(u, dpos, A, Bm, Cm)
    got = ops.selective_scan_fn(u.to(DEV), dpos.to(DEV), A.to(DEV), Bm.to(DEV), Cm.to(DEV))
    assert rel_err(got, want) <= _tol(dtype)


def _fused_scan_ref64(u, z, xdbl, w_dt, A, Dp, bias, h0, reverse, R, N, z_gate=False):
    """float64 evaluation of the fused op on the SAME bf16 inputs (mamba_simple.py:413-414 without
    the intermediate rounding of delta, then :30-106).  ``z_gate``: ``z`` is the gate itself."""
    if reverse:
        u, z, xdbl = u.flip(1), z.flip(1), xdbl.flip(1)
    u, z, xd = u.double(), z.double(), xdbl.double()
    dt = torch.nn.functional.softplus(xd[..., :R] @ w_dt[:, :R].double().t() + bias.double())
    Bm, Cm = xd[..., R:R + N], xd[..., R + N:R + 2 * N]
    Bsz, L, Di = u.shape
    h = torch.zeros(Bsz, Di, N, dtype=torch.float64, device=u.device) if h0 is None else h0.double().clone()
    ys = torch.empty(Bsz, L, Di, dtype=torch.float64, device=u.device)
    Ad = A.double()
    for t in range(L):
        h = torch.exp(dt[:, t, :, None] * Ad) * h + (dt[:, t] * u[:, t])[:, :, None] * Bm[:, t, None, :]
        ys[:, t] = (h * Cm[:, t, None, :]).sum(-1)
    y = (ys + u * Dp.double()) * (z if z_gate else z * torch.sigmoid(z))
    return (y.flip(1) if reverse else y), h


SCAN_GEOMS = [(2, 777, 768, 24, 64), (1, 100, 384, 12, 48), (3, 33, 1152, 36, 80), (2, 1, 768, 24, 64),
              # 300-1400 (batch, 16-channel) units: the two-warp kernel (scan2w)
              (8, 100, 768, 24, 64), (5, 333, 1152, 36, 80), (16, 50, 384, 12, 48),
              (7, 15, 768, 24, 64), (24, 17, 768, 24, 64),
              # >= 1406 units: the one-warp kernel (scan1w) -- what bench.py times at batch 32
              (32, 17, 768, 24, 64), (32, 64, 768, 24, 64), (32, 777, 768, 24, 64),
              (20, 100, 1152, 36, 80), (64, 40, 384, 12, 48)]


def _fused_scan_case(geom, geometric):
    Bsz, L, Di, R, Xp = geom
    N = 16
    gen = torch.Generator().manual_seed(L + Di + (7 if geometric else 0))
    bf = torch.bfloat16
    u = _rand(gen, Bsz, L, Di, dtype=bf).to(DEV)
    z = _rand(gen, Bsz, L, Di, dtype=bf).to(DEV)
    xdbl = _rand(gen, Bsz, L, Xp, dtype=bf).to(DEV)
    w_dt = _rand(gen, Di, R, dtype=bf, scale=R ** -0.5).to(DEV)
    if geometric:       # A[d, n] = (n + 1) * A[d, 0] with a per-channel base (S4D-real has base -1)
        base = -torch.exp(0.5 * torch.randn(Di, 1, generator=gen))
        A = (base * torch.arange(1, N + 1).float()).to(DEV)
    else:
        A = -torch.exp(torch.log(torch.arange(1, N + 1).float()).repeat(Di, 1)
                       + 0.1 * torch.randn(Di, N, generator=gen)).to(DEV)
    Dp = torch.randn(Di, generator=gen).to(DEV)
    bias = (torch.randn(Di, generator=gen) - 3.0).to(DEV)
    h0 = torch.randn(Bsz, Di, N, generator=gen).to(DEV)
    return u, z, xdbl, w_dt, A, Dp, bias, h0, gen


@pytest.mark.parametrize("geom", SCAN_GEOMS)
@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("geometric", [False, True])
def test_fused_scan_against_float64(geom, reverse, geometric):
    """The fused dt_proj + scan kernels (bf16) against float64 on identical inputs, with and without an
    initial state, ragged lengths (tile tails), all three production dt ranks, unit counts on both
    sides of the one-warp / two-warp switch, general A and the geometric-A evaluator."""
    Bsz, L, Di, R, Xp = geom
    N = 16
    bf = torch.bfloat16
    u, z, xdbl, w_dt, A, Dp, bias, h0, gen = _fused_scan_case(geom, geometric)
    A2 = (A * ops.LOG2E).contiguous()
    assert ops.is_geometric(A2) == geometric
    for init in (None, h0):
        want, want_h = _fused_scan_ref64(u, z, xdbl, w_dt, A, Dp, bias, init, reverse, R, N)
        got, got_h = ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, init,
                                                     want_last=True, reverse=reverse, a_geometric=geometric)
        assert got.dtype == bf and got_h.dtype == torch.float32
        assert rel_err(got, want) <= 6e-3, rel_err(got, want)
        assert rel_err(got_h, want_h) <= 2e-3, rel_err(got_h, want_h)
    if geometric:       # the general evaluator on the same structured A agrees too
        plain = ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, reverse=reverse)
        want, _ = _fused_scan_ref64(u, z, xdbl, w_dt, A, Dp, bias, None, reverse, R, N)
        assert rel_err(plain, want) <= 6e-3
    # strided views: z living in the second half of an xz buffer, u / y with a wider pitch
    xz = _rand(gen, Bsz, L, 2 * Di, dtype=bf).to(DEV)
    want, _ = _fused_scan_ref64(xz[..., :Di], xz[..., Di:], xdbl, w_dt, A, Dp, bias, None, reverse, R, N)
    got = ops.selective_scan_fused_tokens(xz[..., :Di], xz[..., Di:], xdbl, w_dt, A2, R, N, Dp, bias,
                                          reverse=reverse, a_geometric=geometric)
    assert rel_err(got, want) <= 6e-3


@pytest.mark.parametrize("geom", [(2, 777, 768, 24, 64), (3, 33, 1152, 36, 80), (8, 100, 768, 24, 64),
                                  (16, 50, 384, 12, 48), (32, 64, 768, 24, 64), (20, 100, 1152, 36, 80)])
@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("geometric", [False, True])
def test_fused_scan_with_stored_gate(geom, reverse, geometric):
    """``z_gate``: z holds SiLU(z) already (the in_proj epilogue of ``vmb_linear_fwd_act`` wrote it) and the
    kernels -- one warp and two warps per unit, both evaluators, both directions, the sequence split -- multiply
    by it as it is: against float64 on identical inputs, and bit-identical to the raw-z kernels wherever the
    gate is not involved (the last state)."""
    Bsz, L, Di, R, Xp = geom
    N = 16
    u, z, xdbl, w_dt, A, Dp, bias, h0, gen = _fused_scan_case(geom, geometric)
    A2 = (A * ops.LOG2E).contiguous()
    gate = torch.nn.functional.silu(z.float()).to(torch.bfloat16)
    for init in (None, h0):
        want, want_h = _fused_scan_ref64(u, gate, xdbl, w_dt, A, Dp, bias, init, reverse, R, N, z_gate=True)
        got, got_h = ops.selective_scan_fused_tokens(u, gate, xdbl, w_dt, A2, R, N, Dp, bias, init, want_last=True,
                                                     reverse=reverse, a_geometric=geometric, z_gate=True)
        assert rel_err(got, want) <= 6e-3, rel_err(got, want)
        raw, raw_h = ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, init, want_last=True,
                                                     reverse=reverse, a_geometric=geometric)
        assert torch.equal(got_h, raw_h)
        assert rel_err(got, raw) <= 8e-3, rel_err(got, raw)      # the gate's own rounding to bf16
    ones = torch.ones_like(z)                                     # gate 1: the ungated sum, whatever z would do
    a = ops.selective_scan_fused_tokens(u, ones, xdbl, w_dt, A2, R, N, Dp, bias, reverse=reverse,
                                        a_geometric=geometric, z_gate=True)
    b = ops.selective_scan_fused_tokens(u, 2 * ones, xdbl, w_dt, A2, R, N, Dp, bias, reverse=reverse,
                                        a_geometric=geometric, z_gate=True)
    assert torch.equal((2 * a.float()).to(torch.bfloat16), b)    # exact: a power of two commutes with the rounding


@pytest.mark.parametrize("shape", [(1000, 1536, 384, 768), (515, 768, 192, 384), (300, 1536, 384, 768),
                                   (4100, 2304, 576, 1152), (777, 1536, 384, 0), (2000, 512, 256, 448)])
def test_projection_activation_epilogue(shape):
    """``vmb_linear_fwd_act``: SiLU on the output columns [silu_from, N) inside the tensor-core projection (CTA
    pairs from 512 rows on, the one-CTA kernel below): the columns before ``silu_from`` keep the bits of the plain
    projection, the others are SiLU of the fp32 accumulator rounded once."""
    M, N, K, frm = shape
    gen = torch.Generator().manual_seed(M + N)
    bf = torch.bfloat16
    x = _rand(gen, M, K, dtype=bf).to(DEV)
    w = _rand(gen, N, K, dtype=bf, scale=K ** -0.5).to(DEV)
    b = _rand(gen, N, dtype=bf).to(DEV)
    for bias in (None, b):
        plain = ops.linear(x, w, bias)
        got = ops.linear_act(x, w, bias, silu_from=frm)
        assert torch.equal(got[:, :frm], plain[:, :frm])
        acc = x.double() @ w.double().t() + (0 if bias is None else bias.double())
        want = torch.nn.functional.silu(acc[:, frm:])
        assert rel_err(got[:, frm:], want) <= 6e-3, rel_err(got[:, frm:], want)
        # rounding after the activation: closer to the exact gate than SiLU of the rounded projection
        late = (torch.nn.functional.silu(plain[:, frm:].double()) - want).abs().mean()
        assert (got[:, frm:].double() - want).abs().mean() <= 1.05 * late
    assert torch.equal(ops.linear_act(x, w, b, silu_from=N), ops.linear(x, w, b))
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.linear_act(x, w, b, silu_from=frm + 8)
    with pytest.raises(RuntimeError, match="tensor-core"):
        ops.linear_act(x.float(), w.float(), b.float(), silu_from=frm)


def test_mixer_gate_in_proj_matches_gate_in_scan():
    """``vmb_mixer_args.gate_in_proj`` (default of the modules): the z half of in_proj leaves the GEMM as SiLU(z) and
    the scan multiplies by it; against the path that applies SiLU inside the scan the outputs differ only by the
    position of one bf16 rounding, the states not at all; shapes the tensor-core projection does not take keep
    the raw z silently."""
    from video_mamba.mamba_simple import Mamba
    torch.manual_seed(5)
    for d_model, Bsz, L in ((384, 4, 300), (192, 2, 77), (576, 2, 130)):
        mx = Mamba(d_model=d_model, use_fast_path=False).to(DEV).to(torch.bfloat16)
        x = torch.randn(Bsz, L, d_model, device=DEV).to(torch.bfloat16)
        w = mx._kernel_weights()
        with torch.no_grad():
            a, ca, sa = ops.mixer_fwd(w, x, None, None, True, True, gate_in_proj=True)
            b, cb, sb = ops.mixer_fwd(w, x, None, None, True, True, gate_in_proj=False)
            ra, _, _ = ops.mixer_fwd(w, x, reverse=True, gate_in_proj=True)
            rb, _, _ = ops.mixer_fwd(w, x, reverse=True, gate_in_proj=False)
        assert torch.equal(ca, cb) and torch.equal(sa, sb)
        assert not torch.equal(a, b)
        assert rel_err(a, b) <= 8e-3 and rel_err(ra, rb) <= 8e-3, (rel_err(a, b), rel_err(ra, rb))
    mx = Mamba(d_model=40, use_fast_path=False).to(DEV).to(torch.bfloat16)      # Di 80: generic kernels
    x = torch.randn(2, 9, 40, device=DEV).to(torch.bfloat16)
    with torch.no_grad():
        a, _, _ = ops.mixer_fwd(mx._kernel_weights(), x, gate_in_proj=True)
        b, _, _ = ops.mixer_fwd(mx._kernel_weights(), x, gate_in_proj=False)
    assert torch.equal(a, b)


def test_fused_scan_layouts_agree_bitwise():
    """The one-warp and two-warp kernels (forced through the tune field) run the same arithmetic in the
    same order: identical bits; the sequence split re-associates the carry and is merely close."""
    run = lambda args, **kw: ops.selective_scan_fused_tokens(*args[:4], args[4], 24, 16, args[5], args[6],
                                                             want_last=True, **kw)
    short = _scan_inputs(6, 250)                     # below the split threshold (3 x 96 tokens at > 1 unit per SM)
    auto, h_auto = run(short)
    for tune in (10, 20, 30):
        got, h = run(short, tune=tune)
        assert torch.equal(got, auto) and torch.equal(h, h_auto), tune
    long = _scan_inputs(2, 3000, seed=13)            # 96 units: split into segments when allowed
    one, h_one = run(long, tune=10)
    two, h_two = run(long, tune=30)
    assert torch.equal(one, two) and torch.equal(h_one, h_two)
    whole, h_whole = run(long, tune=20)              # two warps, never split
    assert rel_err(one, whole) <= 4e-3 and rel_err(h_one, h_whole) <= 1e-4


def test_fused_scan_rejects_other_pitches_and_unbuilt_evaluators():
    """x_dbl rows must have the pitch ops.xdbl_pitch(R, N) (every shared-memory offset of the kernels is
    an immediate); evaluators that only exist in measurement builds are refused, not emulated."""
    u, z, xdbl, w_dt, A2, Dp, bias = _scan_inputs(1, 31, Xp=56)
    with pytest.raises(RuntimeError, match="unsupported"):
        ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, 24, 16, Dp, bias)
    u, z, xdbl, w_dt, A2, Dp, bias = _scan_inputs(1, 31)
    lab = ops.selective_scan_fused_tokens
    outs = []
    for tune in (1, 2, 3, 4, 5):         # general-A evaluators: the built default succeeds, others refuse
        try:
            outs.append(lab(u, z, xdbl, w_dt, A2, 24, 16, Dp, bias, tune=tune))
        except RuntimeError as e:
            assert "not part of this build" in str(e)
    assert len(outs) >= 1


def _scan_inputs(Bsz, L, Di=768, R=24, Xp=64, N=16, seed=9):
    gen = torch.Generator().manual_seed(seed)
    bf = torch.bfloat16
    u = _rand(gen, Bsz, L, Di, dtype=bf).to(DEV)
    z = _rand(gen, Bsz, L, Di, dtype=bf).to(DEV)
    xdbl = _rand(gen, Bsz, L, Xp, dtype=bf).to(DEV)
    w_dt = _rand(gen, Di, R, dtype=bf, scale=R ** -0.5).to(DEV)
    A2 = (-torch.exp(torch.log(torch.arange(1, N + 1).float()).repeat(Di, 1)
                     + 0.1 * torch.randn(Di, N, generator=gen)) * ops.LOG2E).to(DEV)
    Dp = torch.ones(Di, device=DEV)
    bias = torch.full((Di,), -3.0, device=DEV)
    return u, z, xdbl, w_dt, A2, Dp, bias


def test_fused_scan_chunk_carry_is_bitwise():
    """Size-independent property: scanning in three chunks with the state carried equals one pass,
    bit for bit (batch large enough that the kernel walks each sequence with one warp)."""
    Bsz, L, R, N = 26, 1500, 24, 16
    u, z, xdbl, w_dt, A2, Dp, bias = _scan_inputs(Bsz, L)
    full, h_full = ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, want_last=True)
    h, parts = None, []
    for lo, hi in ((0, 700), (700, 701), (701, L)):
        y, h = ops.selective_scan_fused_tokens(u[:, lo:hi], z[:, lo:hi], xdbl[:, lo:hi], w_dt, A2, R,
                                               N, Dp, bias, h, want_last=True)
        parts.append(y)
    assert torch.equal(torch.cat(parts, 1), full) and torch.equal(h, h_full)


@pytest.mark.parametrize("reverse", [False, True])
def test_fused_scan_sequence_split_at_long_clip_size(reverse):
    """Long-clip size (25 089 tokens, batch 1): the batch cannot fill the GPU, so the kernel splits
    the sequence into segments with an exact two-pass carry.  The split result must agree with the
    one-warp-per-sequence walk and with a chunked walk, outputs and final state."""
    Bsz, L, R, N = 1, 25089, 24, 16
    u, z, xdbl, w_dt, A2, Dp, bias = _scan_inputs(Bsz, L, seed=11)
    h0 = torch.randn(Bsz, 768, N, generator=torch.Generator().manual_seed(3)).to(DEV)
    from videomamba_b200 import _lib
    assert _lib.load().vmb_fused_scan_workspace_bytes(Bsz, L, 768, N) > 0        # split is planned
    split, h_split = ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, h0,
                                                     want_last=True, reverse=reverse)
    plain, h_plain = ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, h0,
                                                     want_last=True, reverse=reverse, allow_split=False)
    assert torch.isfinite(split.float()).all()
    assert rel_err(split, plain) <= 4e-3 and rel_err(h_split, h_plain) <= 1e-4
    # three chunks with carried state (each chunk is itself split or not, depending on its length)
    order = [(0, 9000), (9000, 9001), (9001, L)]
    if reverse:
        order = [(L - hi, L - lo) for lo, hi in order]
    h, parts = h0, []
    for lo, hi in order:
        y, h = ops.selective_scan_fused_tokens(u[:, lo:hi], z[:, lo:hi], xdbl[:, lo:hi], w_dt, A2, R,
                                               N, Dp, bias, h, want_last=True, reverse=reverse)
        parts.append(y)
    if reverse:
        parts = parts[::-1]
    assert rel_err(torch.cat(parts, 1), plain) <= 4e-3 and rel_err(h, h_plain) <= 1e-4


def test_fused_scan_split_against_float64():
    Bsz, L, R, N = 2, 4100, 24, 16
    u, z, xdbl, w_dt, A2, Dp, bias = _scan_inputs(Bsz, L, seed=5)
    h0 = torch.randn(Bsz, 768, N, generator=torch.Generator().manual_seed(4)).to(DEV)
    want, want_h = _fused_scan_ref64(u, z, xdbl, w_dt, A2 / ops.LOG2E, Dp, bias, h0, False, R, N)
    got, got_h = ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, h0, want_last=True)
    assert rel_err(got, want) <= 6e-3 and rel_err(got_h, want_h) <= 2e-3


def test_selective_scan_golden_fixture(golden):
    g = golden("scan_fp32.pt")
    c = lambda k: g[k].to(DEV)
    got, last = ops.selective_scan_fn(c("u"), c("delta"), c("A"), c("B"), c("C"), c("D"), z=c("z"),
                                      delta_bias=c("delta_bias"), delta_softplus=True,
                                      return_last_state=True, initial_state=c("h0"))
    assert rel_err(got, g["out"]) <= 1e-5 and rel_err(last, g["last"]) <= 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_scan_state_carry_property_at_long_length(dtype):
    """Size-independent property at the long-clip length (cfg-5: 25 089 tokens): scanning in
    two pieces with the state carried equals one scan, and reverse == flip(forward(flip))."""
    B, D, L, N = 1, 256, 25089, 16
    gen = torch.Generator().manual_seed(9)
    mk = lambda *s: _rand(gen, *s, dtype=dtype).to(DEV)
    u, delta, z = mk(B, L, D), mk(B, L, D), mk(B, L, D)
    bc = mk(B, L, 2 * N)
    A2 = (-torch.arange(1, N + 1).float().repeat(D, 1) * ops.LOG2E).to(DEV)
    Dp = torch.ones(D, device=DEV)
    bias = torch.full((D,), -3.0, device=DEV)
    full, last = ops.selective_scan_tokens(u, delta, A2, bc, 0, N, N, Dp, z, bias, True, None, True)
    k = 12545
    a, ha = ops.selective_scan_tokens(u[:, :k], delta[:, :k], A2, bc[:, :k], 0, N, N, Dp, z[:, :k],
                                      bias, True, None, True)
    b, hb = ops.selective_scan_tokens(u[:, k:], delta[:, k:], A2, bc[:, k:], 0, N, N, Dp, z[:, k:],
                                      bias, True, ha, True)
    assert torch.equal(torch.cat([a, b], 1), full) and torch.equal(hb, last)
    fl = lambda t: torch.flip(t, dims=[1]).contiguous()
    rev = ops.selective_scan_tokens(fl(u), fl(delta), A2, fl(bc), 0, N, N, Dp, fl(z), bias, True,
                                    None, False, reverse=True)
    assert torch.equal(fl(rev), full)
    assert torch.isfinite(full.float()).all()


@pytest.mark.parametrize("dtype", DTYPES)
def test_single_step_ops(dtype):
    gen = torch.Generator().manual_seed(21)
    B, Di, W, N = 3, 48, 4, 16
    x = _rand(gen, B, Di, dtype=dtype)
    cs = _rand(gen, B, Di, W, dtype=dtype)
    w = _rand(gen, Di, W, dtype=dtype, scale=0.5)
    b = _rand(gen, Di, dtype=dtype)
    cs_ref = cs.clone()
    want = orc.causal_conv1d_update_ref(x, cs_ref, w, b, "silu")
    cs_dev = cs.to(DEV)
    got = ops.causal_conv1d_update(x.to(DEV), cs_dev, w.to(DEV), b.to(DEV), "silu")
    assert rel_err(got, want) <= _tol(dtype) and torch.equal(cs_dev.cpu(), cs_ref)
    st = torch.randn(B, Di, N, generator=gen)
    dt = _rand(gen, B, Di, dtype=dtype)
    A = -torch.rand(Di, N, generator=gen) * 4
    Bm, Cm, z = _rand(gen, B, N, dtype=dtype), _rand(gen, B, N, dtype=dtype), _rand(gen, B, Di, dtype=dtype)
    Dp, bias = torch.randn(Di, generator=gen), torch.randn(Di, generator=gen)
    st_ref = st.clone()
    want = orc.selective_state_update_ref(st_ref, x, dt, A, Bm, Cm, Dp, z, bias, True)
    st_dev = st.to(DEV)
    got = ops.selective_state_update(st_dev, x.to(DEV), dt.to(DEV), A.to(DEV), Bm.to(DEV),
                                     Cm.to(DEV), Dp.to(DEV), z.to(DEV), bias.to(DEV), True)
    assert rel_err(got, want) <= _tol(dtype) and rel_err(st_dev, st_ref) <= 1e-5


@pytest.mark.parametrize("geom", [(4, 130, 768), (64, 9, 384), (3, 777, 768), (128, 5, 64), (2, 3137, 768),
                                  (1, 512, 1152), (5, 131, 128)])
def test_conv_xproj_fused_equals_separate_kernels(geom):
    """conv1d + SiLU fused into the x_proj projection (vmb_conv_xproj_fwd) against the separate kernels,
    bit for bit: sequences far shorter than a 128-row tile (several clip starts inside one tile and
    inside one thread's 8 rows), lengths that put clip starts at every row phase, a ragged last tile,
    x as the strided first half of xz."""
    B, L, Di = geom
    bf = torch.bfloat16
    gen = torch.Generator().manual_seed(B * L + Di)
    xz = _rand(gen, B, L, 2 * Di, dtype=bf).to(DEV)
    w = _rand(gen, Di, 4, dtype=bf, scale=0.5).to(DEV)
    b = _rand(gen, Di, dtype=bf, scale=0.5).to(DEV)
    wx = _rand(gen, 64, Di, dtype=bf, scale=Di ** -0.5).to(DEV)
    x = xz[..., :Di]
    for bias in (b, None):
        want_xc = ops.causal_conv1d_tokens(x, w, bias)
        want_xd = ops.linear(want_xc, wx)
        xc, xd = ops.conv_xproj_tokens(x, w, bias, wx)
        assert torch.equal(xc, want_xc), float((xc.float() - want_xc.float()).abs().max())
        assert torch.equal(xd, want_xd), float((xd.float() - want_xd.float()).abs().max())
    # and against the oracle conv (fp32 math) at the bf16 tolerance
    ref = orc.causal_conv1d_ref(x.cpu().transpose(1, 2), w.cpu(), b.cpu(), "silu").transpose(1, 2)
    xc, _ = ops.conv_xproj_tokens(x, w, b, wx)
    assert rel_err(xc, ref) <= _tol(bf)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 37, 384), (1, 5, 16), (3, 200, 576), (0, 4, 8)])
def test_gate_blend(dtype, shape):
    """Refiner fusion gate (refiner_backbone.py:129-134): s * fwd + (1 - s) * bwd, s = sigmoid(g1 + g2),
    against the torch expression in float64; one- and two-logit forms."""
    gen = torch.Generator().manual_seed(sum(shape))
    g1, g2, f, b = (_rand(gen, *shape, dtype=dtype) for _ in range(4))
    for second in (g2, None):
        lg = g1.double() + (second.double() if second is not None else 0.0)
        sg = torch.sigmoid(lg)
        want = sg * f.double() + (1.0 - sg) * b.double()
        got = ops.gate_blend(g1.to(DEV), None if second is None else second.to(DEV), f.to(DEV), b.to(DEV))
        assert got.dtype == dtype and got.shape == f.shape
        if f.numel():
            assert rel_err(got, want) <= (1e-6 if dtype == torch.float32 else 8e-3)


def test_state_gather_scatter_roundtrip():
    pool = torch.randn(9, 24, 16, device=DEV)
    idx = torch.tensor([7, 0, 3], device=DEV)
    rows = ops.state_gather(pool, idx)
    assert torch.equal(rows, pool[idx])
    new = torch.randn_like(rows)
    want = pool.clone()
    want[idx] = new
    ops.state_scatter(pool, idx, new)
    assert torch.equal(pool, want)
    poolb = torch.randn(5, 7, 3, device=DEV).to(torch.bfloat16)    # odd row size: scalar path
    assert torch.equal(ops.state_gather(poolb, torch.tensor([4, 4, 1], device=DEV)), poolb[[4, 4, 1]])
    assert ops.state_gather(pool, torch.zeros(0, dtype=torch.long, device=DEV)).shape == (0, 24, 16)


def test_unsupported_inputs_raise_python_errors():
    x = torch.randn(2, 5, 16, device=DEV, dtype=torch.float16)
    with pytest.raises(TypeError, match="float32 and bfloat16"):
        ops.linear(x, torch.randn(4, 16, device=DEV, dtype=torch.float16))
    with pytest.raises(RuntimeError, match="requires CUDA tensors"):
        ops.linear(torch.randn(2, 4), torch.randn(3, 4))
    with pytest.raises(RuntimeError, match="d_conv"):
        ops.causal_conv1d_tokens(torch.randn(1, 4, 8, device=DEV), torch.randn(8, 7, device=DEV), None)


def test_add_norm_and_gate_accept_misaligned_views():
    """a last-dim slice big[..., 2:2+dim] is a view whose base is not 16-byte aligned; the
    vector kernels must not fault on it (ops realigns; the C entry refuses instead of faulting)."""
    from videomamba_b200 import _lib
    dim = 384
    gen = torch.Generator().manual_seed(3)
    for dtype in DTYPES:
        big = _rand(gen, 6, 10, dim + 8, dtype=dtype).to(DEV)
        x = big[..., 2:2 + dim]
        assert x.data_ptr() % 16 != 0
        w = torch.ones(dim, dtype=dtype, device=DEV)
        res = _rand(gen, 6, 10, dim, dtype=torch.float32).to(DEV)
        y, r = ops.add_norm(x, w, None, res, 1e-5, True, True, True)
        y2, r2 = ops.add_norm(x.contiguous(), w, None, res, 1e-5, True, True, True)
        assert torch.equal(y, y2) and torch.equal(r, r2)
        g = ops.gate_blend(x, None, x, x)
        assert torch.equal(g, ops.gate_blend(x.contiguous(), None, x.contiguous(), x.contiguous()))
    # the C entry itself: misaligned base -> VMB_ERR_UNSUPPORTED, never a fault
    lib = _lib.load()
    buf = torch.zeros(4 * dim + 8, dtype=torch.float32, device=DEV)
    import ctypes as C
    p = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
    w32 = torch.ones(dim, device=DEV)
    out = torch.zeros(4, dim, device=DEV)
    rc = lib.vmb_add_norm_fwd(p(buf, 4), 0, dim, None, 0, p(w32), None, 0, p(out), None, 0, 4, dim, 1e-5, 1,
                              None)
    assert rc == -2 and b"aligned" in lib.vmb_last_error()
    torch.cuda.synchronize()


def test_forward_only_entry_points_refuse_backward_and_stale_weights_are_detected():
    """entry points without a backward kernel (the fused inference mixer) raise in backward
    instead of dropping gradients -- the modules themselves train through autograd.py
    (tests/test_gpu_backward.py); writes through .data are caught by verify_weights and cured by
    refresh_weights()."""
    from video_mamba.mamba_simple import Mamba
    torch.manual_seed(2)
    mx = Mamba(d_model=64, use_fast_path=False).to(DEV)
    x = torch.randn(2, 20, 64, device=DEV)
    out = mx(x)                                   # grad enabled, parameters require grad: differentiable
    assert out.requires_grad
    out.sum().backward()
    assert mx.in_proj.weight.grad is not None and mx.A_log.grad is not None
    fused, _, _ = ops.mixer_fwd(mx._kernel_weights(), x)    # the fused inference entry point is not
    assert fused.requires_grad
    with pytest.raises(NotImplementedError, match="forward-only"):
        fused.sum().backward()
    with torch.no_grad():
        base = mx(x)
        assert not base.requires_grad
        mx.A_log.data.add_(0.5)                   # no version bump: the derived A2 copy is now stale
        stale = mx(x)
        assert torch.equal(stale, base)
        mx.verify_weights = True
        with pytest.raises(RuntimeError, match="stale"):
            mx(x)
        mx.verify_weights = False
        mx.refresh_weights()
        fresh = mx(x)
        assert not torch.equal(fresh, base)


def _flip_frames(t, n):
    b, l, c = t.shape
    return torch.flip(t.reshape(b, l // n, n, c), dims=[1]).reshape(b, l, c)


@pytest.mark.parametrize("geom", [(2, 5, 49, 768, 24), (3, 4, 16, 384, 12), (1, 7, 33, 1152, 36), (40, 3, 20, 768, 24)])
def test_frame_axis_reversal_equals_flip_copies(geom):
    """reverse + frame_len: frames walked back to front, tokens of a frame front to back -- the 4-D flip
    of BiMambaRefinerBlock (models/refiner_backbone.py:61-68) without the two gather copies.  Must equal
    flip -> forward kernel -> flip: bit for bit for the scans (same arithmetic per token), to rounding for
    the conv (a different kernel carries the frame walk)."""
    Bsz, T, n, Di, R = geom
    L, N, bf = T * n, 16, torch.bfloat16
    Xp = ops.xdbl_pitch(R, N)
    gen = torch.Generator().manual_seed(Di + n)
    u = _rand(gen, Bsz, L, Di, dtype=bf).to(DEV)
    z = _rand(gen, Bsz, L, Di, dtype=bf).to(DEV)
    xdbl = _rand(gen, Bsz, L, Xp, dtype=bf).to(DEV)
    w_dt = _rand(gen, Di, R, dtype=bf, scale=R ** -0.5).to(DEV)
    A2 = (-torch.exp(torch.log(torch.arange(1, N + 1).float()).repeat(Di, 1)
                     + 0.1 * torch.randn(Di, N, generator=gen)) * ops.LOG2E).to(DEV)
    Dp = torch.randn(Di, generator=gen).to(DEV)
    bias = (torch.randn(Di, generator=gen) - 3.0).to(DEV)
    h0 = torch.randn(Bsz, Di, N, generator=gen).to(DEV)
    # fused scan (one-warp and two-warp kernels by unit count; the frame walk never splits the sequence, so the
    # flipped copy is walked whole as well -- the split re-associates the carry)
    want, want_h = ops.selective_scan_fused_tokens(_flip_frames(u, n), _flip_frames(z, n), _flip_frames(xdbl, n),
                                                   w_dt, A2, R, N, Dp, bias, h0, want_last=True, allow_split=False)
    got, got_h = ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias, h0, want_last=True,
                                                 reverse=True, frame_len=n)
    assert torch.equal(got, _flip_frames(want, n)) and torch.equal(got_h, want_h)
    # op-level scan (generic kernel), bf16 and fp32
    for dt in (bf, torch.float32):
        delta = _rand(gen, Bsz, L, Di, dtype=dt).to(DEV)
        uu, zz, bc = u.to(dt), z.to(dt), xdbl[..., :2 * N].contiguous().to(dt)
        want, want_h = ops.selective_scan_tokens(_flip_frames(uu, n), _flip_frames(delta, n), A2, _flip_frames(bc, n),
                                                 0, N, N, Dp, _flip_frames(zz, n), bias, True, h0, True)
        got, got_h = ops.selective_scan_tokens(uu, delta, A2, bc, 0, N, N, Dp, zz, bias, True, h0, True,
                                               reverse=True, frame_len=n)
        assert torch.equal(got, _flip_frames(want, n)) and torch.equal(got_h, want_h)
    # conv with history in, next state out
    w = _rand(gen, Di, 4, dtype=bf, scale=0.5).to(DEV)
    b = _rand(gen, Di, dtype=bf).to(DEV)
    cs = _rand(gen, Bsz, Di, 4, dtype=bf).to(DEV)
    want, want_cs = ops.causal_conv1d_tokens(_flip_frames(u, n), w, b, cs, True)
    got, got_cs = ops.causal_conv1d_tokens(u, w, b, cs, True, reverse=True, frame_len=n)
    assert rel_err(got, _flip_frames(want, n)) <= 1e-2 and torch.equal(got_cs, want_cs)
    with pytest.raises(RuntimeError, match="whole number of frames"):
        ops.causal_conv1d_tokens(u, w, b, reverse=True, frame_len=n + 1 if L % (n + 1) else n + 2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("pool_type", ["cls", "cls+avg", "cls_cat_avg", "avg"])
@pytest.mark.parametrize("geom", [(3, 1, 196, 384), (2, 4, 49, 192), (5, 8, 196, 576), (1, 1, 3136, 384), (2, 3, 7, 20)])
def test_pool_norm_against_torch(dtype, pool_type, geom):
    """Pooling over the patch tokens + pool_norm (reference videomamba.py:983-1063): all four pool types,
    whole-clip mean (groups = 1) and keep_temporal (per-frame mean), with and without a CLS row."""
    Bsz, G, per, Cdim = geom
    gen = torch.Generator().manual_seed(G * per + Cdim)
    F = torch.nn.functional
    for has_cls in ((True,) if pool_type != "avg" else (True, False)):
        x = _rand(gen, Bsz, (1 if has_cls else 0) + G * per, Cdim, dtype=dtype).to(DEV)
        w = (1 + 0.1 * torch.randn(Cdim, generator=gen)).to(dtype).to(DEV)
        b = (0.1 * torch.randn(Cdim, generator=gen)).to(dtype).to(DEV)
        groups = 1 if pool_type == "cls" else G
        per_g = (G * per) // groups
        got = ops.pool_norm(x, has_cls, groups, per_g, pool_type, w, b, 1e-5)
        cls = x[:, :1]
        patches = x[:, 1:] if has_cls else x
        ln = lambda t: F.layer_norm(t, (Cdim,), w, b, 1e-5)
        avg = patches.reshape(Bsz, groups, per_g, Cdim).mean(2)
        want = {"cls": lambda: ln(cls), "cls+avg": lambda: ln(cls + avg),
                "cls_cat_avg": lambda: ln(torch.cat([cls, avg], dim=1)), "avg": lambda: ln(avg)}[pool_type]()
        assert got.shape == want.shape and got.dtype == dtype
        # LayerNorm of a mean amplifies the last-bit differences of the bf16 mean by 1 / std(mean)
        assert rel_err(got, want) <= (2e-2 if dtype == torch.bfloat16 else 2e-5), rel_err(got, want)
    with pytest.raises(RuntimeError, match="CLS"):
        ops.pool_norm(x[:, : G * per] if not has_cls else x[:, 1:], False, G, per, "cls+avg", w, b, 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_gather_rows_equals_torch_gather(dtype):
    gen = torch.Generator().manual_seed(1)
    for Bsz, L, n, Cdim in ((4, 197, 50, 384), (2, 65, 65, 20), (3, 33, 1, 7)):
        src = _rand(gen, Bsz, L, Cdim, dtype=dtype).to(DEV)
        idx = torch.stack([torch.randperm(L, generator=gen)[:n].sort().values for _ in range(Bsz)]).to(DEV)
        got = ops.gather_rows(src, idx)
        assert torch.equal(got, src.gather(1, idx.unsqueeze(-1).expand(-1, -1, Cdim)))
        view = src[..., : Cdim // 2 * 2][:, :, : max(1, Cdim // 2)]     # strided source rows
        got = ops.gather_rows(view, idx)
        assert torch.equal(got, view.gather(1, idx.unsqueeze(-1).expand(-1, -1, view.shape[-1])))
