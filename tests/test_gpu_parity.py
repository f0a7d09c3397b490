"""GPU: module / model level parity of the CUDA path against (a) the fixtures generated from
the live reference (tests/golden) and (b) the CPU oracle on seeded synthetic weights at sizes
the oracle finishes in seconds; plus the reference's own CUDA test-suite cases
(tests/test_videomamba_regressions.py:249-588, tests/test_public_api_contract.py:68-92).

Bars (north star): x_vis / x_pool / next_state within 1e-5 (fp32) and 2e-2 (bf16),
relative = max|a-b| / max|b|; state shapes and token indexing exact."""
from types import SimpleNamespace

import pytest
import torch

import video_mamba
from oracle import videomamba_oracle as orc
from oracle.videomamba_oracle import rel_err
from video_mamba.mamba_simple import Mamba
from videomamba_b200 import _lib
from videomamba_b200.block import create_block

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tol(dt):
    return 2e-2 if dt == torch.bfloat16 else 1e-5


def _close(got, want, tol):
    assert tuple(got.shape) == tuple(want.shape), (got.shape, want.shape)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    err = rel_err(got, want)
    assert err <= tol, err


def _model_from(cfg, sd, dtype, **over):
    kw = dict(img_size=cfg["img_size"], patch_size=cfg["patch_size"], depth=cfg["depth"],
              embed_dim=cfg["embed_dim"], channels=3, ssm_cfg={"use_fast_path": False},
              rms_norm=cfg.get("rms_norm", True), fused_add_norm=cfg.get("fused_add_norm", True),
              residual_in_fp32=cfg.get("residual_in_fp32", True),
              kernel_size=cfg.get("kernel_size", 1), num_frames=cfg["num_frames"],
              pool_type=cfg.get("pool_type", "cls+avg"))
    kw.update(over)
    m = video_mamba.PretrainVideoMamba(**kw).eval().to(dtype)
    if not kw.get("add_pool_norm", True):
        sd = {k: v for k, v in sd.items() if not k.startswith("pool_norm")}
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


def test_extension_is_loaded_and_reports_blackwell():
    lib = _lib.load()
    import ctypes
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.vmb_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi)) == 0
    assert sm.value > 0 and ma.value >= 10


GOLD = ["model_fp32_rms_fused.pt", "model_bf16_rms_fused.pt", "model_fp32_ln_unfused.pt"]


@pytest.mark.parametrize("name", GOLD)
def test_golden_model_forward(golden, name):
    g = golden(name)
    dt = g["x"].dtype
    m = _model_from(g["cfg"], g["sd"], dt)
    x = g["x"].to(DEV)
    with torch.no_grad():
        x_vis, x_pool = m(x)
        _close(x_vis, g["x_vis"], _tol(dt))
        _close(x_pool, g["x_pool"], _tol(dt))
        _close(m.forward_features(x), g["features"], _tol(dt))
        _close(m(x, keep_temporal=True)[1], g["x_pool_keep_temporal"], _tol(dt))
        mv, mp = m(x, mask=g["mask"].to(DEV))
        _close(mv, g["x_vis_masked"], _tol(dt))
        _close(mp, g["x_pool_masked"], _tol(dt))


def test_golden_temporal_table_interpolation(golden):
    """a.9 with a NON-ZERO temporal table walked beyond its length (reference videomamba.py:655-675):
    full clip and two chunked walks against the live-reference fixture."""
    g = golden("model_fp32_temporal_interp.pt")
    m = _model_from(g["cfg"], g["sd"], torch.float32)
    x = g["x"].to(DEV)
    with torch.no_grad():
        vis, pool = m(x)
        _close(vis, g["full_vis"], 1e-5); _close(pool, g["full_pool"], 1e-5)
        for n in (4, 6, 8):
            rows = m._get_temporal_pos_embedding(n - 2, offset=2, dtype=torch.float32, device=x.device)
            _close(rows, g[f"rows_{n}"], 1e-6)
        for tag, cuts in (("c44", (0, 4, 8)), ("c33", (0, 3, 6))):
            state = m.allocate_state(2, dtype=torch.float32, device=x.device)
            for i in range(len(cuts) - 1):
                lo, hi = cuts[i], cuts[i + 1]
                m.pool_type = "cls+avg" if lo == 0 else "avg"
                v, p_, state = m(x[:, :, lo:hi], ssm_state=state, temporal_pos_offset=lo)
                _close(v, g[f"{tag}_vis{i}"], 1e-5); _close(p_, g[f"{tag}_pool{i}"], 1e-5)
            m.pool_type = "cls+avg"
            for (c, s), (rc, rs) in zip(state, g[f"{tag}_state"]):
                _close(c, rc, 1e-5); _close(s, rs, 1e-5)


@pytest.mark.parametrize("name", GOLD)
@pytest.mark.parametrize("container", ["list", "tuple", "dict"])
def test_golden_streaming(golden, name, container):
    g = golden(name)
    dt = g["x"].dtype
    m = _model_from(g["cfg"], g["sd"], dt)
    x = g["x"].to(DEV)
    state = video_mamba.allocate_state(m, 2, dtype=dt, device=x.device, as_dict=container == "dict")
    if container == "tuple":
        state = tuple(state)
    with torch.no_grad():
        a_vis, a_pool, s1 = m(x[:, :, :2], ssm_state=state, temporal_pos_offset=0)
        m.pool_type = "avg"
        b_vis, b_pool, s2 = m(x[:, :, 2:], ssm_state=s1, temporal_pos_offset=2)
    _close(a_vis, g["chunk0_vis"], _tol(dt)); _close(a_pool, g["chunk0_pool"], _tol(dt))
    _close(b_vis, g["chunk1_vis"], _tol(dt)); _close(b_pool, g["chunk1_pool"], _tol(dt))
    assert type(s2) is {"list": list, "tuple": tuple, "dict": dict}[container]
    video_mamba.validate_state(m, s2, 2)
    for i, (rc, rs) in enumerate(g["state2"]):
        _close(s2[i][0], rc, _tol(dt))
        _close(s2[i][1], rs, _tol(dt))
        assert s2[i][1].dtype == torch.float32


def test_golden_legacy_state_mixer_refiner(golden):
    g = golden("model_fp32_rms_fused.pt")
    m = _model_from(g["cfg"], g["sd"], torch.float32)
    legacy = m.init_ssm_state(2, dtype=torch.float32, device=DEV)
    with torch.no_grad():
        vis, _p, out_state = m(g["x"].to(DEV)[:, :, :2], ssm_state=legacy, temporal_pos_offset=0)
    _close(vis, g["legacy_vis"], 1e-5)
    assert out_state is legacy
    for s, r in zip(legacy, g["legacy_state"]):
        _close(s, r, 1e-5)

    g = golden("mixer_fp32.pt")
    mx = Mamba(d_model=16, d_state=8, d_conv=4, expand=2, use_fast_path=False).eval()
    mx.load_state_dict(g["sd"]); mx.to(DEV)
    x = g["x"].to(DEV)
    with torch.no_grad():
        _close(mx(x), g["full"], 1e-5)
        o1, st1 = mx(x[:, :5], return_state=True)
        o2, st2 = mx(x[:, 5:], state=st1, return_state=True)
    _close(o1, g["out1"], 1e-5); _close(o2, g["out2"], 1e-5)
    for a, b in zip(st1 + st2, g["state1"] + g["state2"]):
        _close(a, b, 1e-5)
    # scripts/check_streaming_state.py:47-55 acceptance: chunked == full at 1e-4
    torch.testing.assert_close(torch.cat([o1, o2], 1).cpu(), g["full"], rtol=1e-4, atol=1e-4)

    s = g["small"]
    sm = Mamba(d_model=8, d_state=4, d_conv=2, expand=2, use_fast_path=False, layer_idx=0).eval()
    sm.load_state_dict(s["sd"]); sm.to(DEV)
    cache = SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    xs = s["x"].to(DEV)
    with torch.no_grad():
        _close(sm(xs[:, :3], inference_params=cache), s["prefill"], 1e-5)
        cache.seqlen_offset = 3
        _close(sm(xs[:, 3:4], inference_params=cache), s["step1"], 1e-5)
        cache.seqlen_offset = 4
        _close(sm(xs[:, 4:5], inference_params=cache), s["step2"], 1e-5)
    _close(cache.key_value_memory_dict[0][0], s["cache_conv"], 1e-5)
    _close(cache.key_value_memory_dict[0][1], s["cache_ssm"], 1e-5)

    g = golden("refiner_fp32.pt")
    blk = video_mamba.BiMambaRefinerBlock(dim=16, ssm_cfg={"use_fast_path": False}, layer_idx=0).eval()
    blk.load_state_dict(g["sd"]); blk.to(DEV)
    with torch.no_grad():
        y3, s3 = blk(g["x3"].to(DEV))
        y4, s4 = blk(g["x4"].to(DEV))
    _close(y3, g["y3"], 1e-5); _close(y4, g["y4"], 1e-5)
    for a, b in zip(s3 + s4, g["s3"] + g["s4"]):
        _close(a, b, 1e-5)


# ---- oracle parity on synthetic weights at the production widths ---------------------------------
def _synthetic(cfg, dtype, perturbed, seed=0):
    sd = orc.synthetic_state_dict(cfg, seed=seed, dtype=torch.float32, perturbed=perturbed)
    return {k: v.to(dtype) for k, v in sd.items()}


WIDTHS = [
    # (embed_dim, depth, frames, batch)  -- Tiny / Small / Middle widths, short clips
    (192, 4, 2, 2),
    (384, 3, 2, 2),
    (576, 2, 1, 1),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("width", WIDTHS)
@pytest.mark.parametrize("perturbed", [False, True])
def test_oracle_parity_production_widths(dtype, width, perturbed):
    dim, depth, frames, batch = width
    cfg = dict(img_size=64, patch_size=16, depth=depth, embed_dim=dim, kernel_size=1,
               num_frames=4, norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True,
               residual_in_fp32=True, pool_type="cls+avg", add_pool_norm=True)
    sd = _synthetic(cfg, dtype, perturbed, seed=dim)
    x = torch.rand(batch, 3, 2 * frames, 64, 64, generator=torch.Generator().manual_seed(1)).to(dtype)
    oracle = orc.OracleVideoMamba(cfg, sd)
    m = _model_from(cfg, sd, dtype)
    state_o = [(torch.zeros(batch, 2 * dim, 4, dtype=dtype), torch.zeros(batch, 2 * dim, 16, dtype=dtype))
               for _ in range(depth)]
    with torch.no_grad():
        want_vis, want_pool = oracle.forward(x)
        got_vis, got_pool = m(x.to(DEV))
        _close(got_vis, want_vis, _tol(dtype)); _close(got_pool, want_pool, _tol(dtype))
        # streaming: two chunks with state carry; continuation chunk has no CLS
        w1 = oracle.forward(x[:, :, :frames], ssm_state=state_o, temporal_pos_offset=0)
        oracle.pool_type = "avg"
        w2 = oracle.forward(x[:, :, frames:], ssm_state=w1[2], temporal_pos_offset=frames)
        st = m.allocate_state(batch, dtype=dtype, device=DEV)
        g1 = m(x.to(DEV)[:, :, :frames], ssm_state=st, temporal_pos_offset=0)
        m.pool_type = "avg"
        g2 = m(x.to(DEV)[:, :, frames:], ssm_state=g1[2], temporal_pos_offset=frames)
    _close(g1[0], w1[0], _tol(dtype)); _close(g1[1], w1[1], _tol(dtype))
    _close(g2[0], w2[0], _tol(dtype)); _close(g2[1], w2[1], _tol(dtype))
    assert g2[0].shape[1] == frames * 16
    for (gc, gs), (wc, ws) in zip(g2[2], w2[2]):
        assert gc.shape == wc.shape and gs.shape == ws.shape and gs.dtype == torch.float32
        assert rel_err(gc, wc) <= _tol(dtype) and rel_err(gs, ws) <= _tol(dtype)
    # chunked == full on the CUDA path itself (zero temporal table unless perturbed)
    if not perturbed:
        stitched = torch.cat([g1[0], g2[0]], dim=1)
        assert rel_err(stitched, got_vis) <= (2e-2 if dtype == torch.bfloat16 else 1e-5)


@pytest.mark.parametrize("frames", [1, 2])
@pytest.mark.parametrize("weights", ["general", "geometric"])
def test_oracle_parity_small_width_batch32(frames, weights):
    """VideoMamba-Small width at the BENCH batch (32 clips @224: 1536 (batch, 16-channel) units, the
    one-warp scan kernel and the CTA-pair projections that bench.py times), 1-2 frames, bf16, against
    the CPU oracle.  `geometric`: S4D-real A with A_log kept in fp32 (a bf16 cast of log(n) breaks the
    structure), which selects the geometric-A evaluator of the scan."""
    dtype, dim, depth, batch = torch.bfloat16, 384, 2, 32
    cfg = dict(img_size=224, patch_size=16, depth=depth, embed_dim=dim, kernel_size=1, num_frames=frames,
               norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True, residual_in_fp32=True,
               pool_type="cls+avg", add_pool_norm=True)
    general = weights == "general"
    sd32 = orc.synthetic_state_dict(cfg, seed=11, dtype=torch.float32, perturbed=general)
    sd = {k: v.to(dtype) for k, v in sd32.items()}
    m = _model_from(cfg, sd, dtype)
    if not general:
        for i, layer in enumerate(m.layers):
            exact = sd32[f"layers.{i}.mixer.A_log"]
            sd[f"layers.{i}.mixer.A_log"] = exact
            layer.mixer.A_log = torch.nn.Parameter(exact.to(DEV))
        assert all(layer.mixer._kernel_weights().a_geometric for layer in m.layers)
    else:
        assert not any(layer.mixer._kernel_weights().a_geometric for layer in m.layers)
    x = torch.rand(batch, 3, frames, 224, 224, generator=torch.Generator().manual_seed(3)).to(dtype)
    with torch.no_grad():
        want_vis, want_pool = orc.OracleVideoMamba(cfg, sd).forward(x)
        got_vis, got_pool = m(x.to(DEV))
    _close(got_vis, want_vis, 2e-2)
    _close(got_pool, want_pool, 2e-2)


@pytest.mark.parametrize("case", ["configs0_tiny_8f_fp32_b2", "configs1_small_16f_bf16_b1"])
def test_oracle_parity_at_baseline_sizes(case):
    """BASELINE.json configs[0] exactly (VideoMamba-Tiny, depth 24, 8 frames @224, fp32, batch 2; bar
    1e-5) and configs[1] at full model / clip size (VideoMamba-Small, depth 24, 16 frames @224, bf16,
    3 137 tokens, batch 1; bar 2e-2), general-A weights, against the CPU oracle on x_vis and x_pool."""
    if case.startswith("configs0"):
        dim, frames, batch, dtype = 192, 8, 2, torch.float32
    else:
        dim, frames, batch, dtype = 384, 16, 1, torch.bfloat16
    cfg = dict(img_size=224, patch_size=16, depth=24, embed_dim=dim, kernel_size=1, num_frames=frames,
               norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True, residual_in_fp32=True,
               pool_type="cls+avg", add_pool_norm=True)
    sd = _synthetic(cfg, dtype, True, seed=7)
    x = torch.rand(batch, 3, frames, 224, 224, generator=torch.Generator().manual_seed(2)).to(dtype)
    oracle = orc.OracleVideoMamba(cfg, sd)          # the reference's rounding points in `dtype`
    m = _model_from(cfg, sd, dtype)
    with torch.no_grad():
        want_vis, want_pool = oracle.forward(x)
        got_vis, got_pool = m(x.to(DEV))
    assert got_vis.shape == (batch, frames * 196, dim)
    _close(got_vis, want_vis, _tol(dtype))
    _close(got_pool, want_pool, _tol(dtype))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("geom", [(384, 16, 4, 1), (384, 16, 4, 32), (576, 16, 4, 5), (40, 8, 3, 3)])
def test_fused_decode_step_matches_oracle(dtype, tol, geom):
    """Mamba.step (reference mamba_simple.py:453-497): in_proj, ONE fused kernel for conv update + x_proj +
    dt_proj + state update + gate (vmb_mixer_step_fwd), out_proj.  Five consecutive tokens after a
    prefill through the inference cache, against the oracle: outputs and both in-place states."""
    d_model, d_state, d_conv, batch = geom
    torch.manual_seed(d_model + batch)
    mx = Mamba(d_model=d_model, d_state=d_state, d_conv=d_conv, use_fast_path=False, layer_idx=0).eval()
    with torch.no_grad():
        mx.A_log.add_(0.1 * torch.randn_like(mx.A_log))
    mx = mx.to(dtype)
    p = {k: v.detach().clone() for k, v in mx.state_dict().items()}
    x = torch.randn(batch, 9, d_model).to(dtype)
    # oracle: prefill 4 tokens, then 5 single-token steps
    di = 2 * d_model
    o_conv = torch.zeros(batch, di, d_conv, dtype=dtype)
    o_ssm = torch.zeros(batch, di, d_state, dtype=dtype)
    want = [orc.mixer_prefill_cache_ref(p, x[:, :4], o_conv, o_ssm)]
    for t in range(4, 9):
        want.append(orc.mixer_step_ref(p, x[:, t:t + 1], o_conv, o_ssm))
    mx.to(DEV)
    cache = SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    launches0 = _lib.load().vmb_launch_count()
    with torch.no_grad():
        got = [mx(x[:, :4].to(DEV), inference_params=cache)]
        before = _lib.load().vmb_launch_count()
        for t in range(4, 9):
            cache.seqlen_offset = t
            got.append(mx(x[:, t:t + 1].to(DEV), inference_params=cache))
    assert _lib.load().vmb_launch_count() - before == 5 * 3      # in_proj, fused step, out_proj per token
    for g, w_ in zip(got, want):
        _close(g, w_, tol)
    g_conv, g_ssm = cache.key_value_memory_dict[0]
    _close(g_conv, o_conv, tol)
    _close(g_ssm, o_ssm, tol)


def test_mixer_fused_conv_xproj_is_bit_identical():
    """Mamba.fuse_conv_xproj routes stateless forward walks through the one-kernel conv + x_proj
    (vmb_conv_xproj_fwd inside vmb_mixer_fwd): same bits as the two-kernel path, at the bench width."""
    torch.manual_seed(9)
    mx = Mamba(d_model=384, use_fast_path=False).eval().to(torch.bfloat16).to(DEV)
    x = torch.randn(4, 700, 384, device=DEV).to(torch.bfloat16)
    with torch.no_grad():
        base = mx(x)
        mx.fuse_conv_xproj = True
        fused = mx(x)
        # with state the fused kernel does not apply; the flag must not change anything
        st = mx.allocate_state(4, dtype=torch.bfloat16, device=DEV)
        o1, s1 = mx(x, state=st, return_state=True)
        mx.fuse_conv_xproj = False
        o2, s2 = mx(x, state=st, return_state=True)
    assert torch.equal(base, fused)
    assert torch.equal(o1, o2) and torch.equal(s1[0], s2[0]) and torch.equal(s1[1], s2[1])


def test_oracle_parity_middle_32f_full_size():
    """BASELINE.json configs[2] at full model / clip size: VideoMamba-Middle (embed 576, depth 32),
    32 frames @224 (6 273 tokens), bf16, batch 1, general-A weights, against the CPU oracle."""
    dtype, dim, frames = torch.bfloat16, 576, 32
    cfg = dict(img_size=224, patch_size=16, depth=32, embed_dim=dim, kernel_size=1, num_frames=frames,
               norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True, residual_in_fp32=True,
               pool_type="cls+avg", add_pool_norm=True)
    sd = _synthetic(cfg, dtype, True, seed=17)
    x = torch.rand(1, 3, frames, 224, 224, generator=torch.Generator().manual_seed(4)).to(dtype)
    m = _model_from(cfg, sd, dtype)
    with torch.no_grad():
        want_vis, want_pool = orc.OracleVideoMamba(cfg, sd).forward(x)
        got_vis, got_pool = m(x.to(DEV))
    assert got_vis.shape == (1, frames * 196, dim)
    _close(got_vis, want_vis, 2e-2)
    _close(got_pool, want_pool, 2e-2)


def test_oracle_parity_long_clip_128_frames_full_size():
    """BASELINE.json configs[4] at full size: VideoMamba-Small, 128 frames @224 (25 089 tokens in one
    sequence: the scan runs split into segments with the exact two-pass carry), bf16, batch 1, general-A
    weights, against the CPU oracle."""
    dtype, dim, frames = torch.bfloat16, 384, 128
    cfg = dict(img_size=224, patch_size=16, depth=24, embed_dim=dim, kernel_size=1, num_frames=frames,
               norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True, residual_in_fp32=True,
               pool_type="cls+avg", add_pool_norm=True)
    sd = _synthetic(cfg, dtype, True, seed=23)
    x = torch.rand(1, 3, frames, 224, 224, generator=torch.Generator().manual_seed(6)).to(dtype)
    m = _model_from(cfg, sd, dtype)
    with torch.no_grad():
        want_vis, want_pool = orc.OracleVideoMamba(cfg, sd).forward(x)
        got_vis, got_pool = m(x.to(DEV))
    assert got_vis.shape == (1, frames * 196, dim)
    _close(got_vis, want_vis, 2e-2)
    _close(got_pool, want_pool, 2e-2)


def test_oracle_parity_streaming_64_frame_chunks_full_size():
    """BASELINE.json configs[3] at full chunk size: VideoMamba-Small, two 64-frame chunks @224
    (12 545 tokens with CLS, then 12 544) with (conv_state, ssm_state) carry and temporal_pos_offset,
    bf16, one stream, general-A weights and a non-zero temporal table, against the CPU oracle:
    x_vis, x_pool and the next state of both chunks."""
    dtype, dim, depth, chunk = torch.bfloat16, 384, 24, 64
    cfg = dict(img_size=224, patch_size=16, depth=depth, embed_dim=dim, kernel_size=1, num_frames=2 * chunk,
               norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True, residual_in_fp32=True,
               pool_type="cls+avg", add_pool_norm=True)
    sd = _synthetic(cfg, dtype, True, seed=19)
    x = torch.rand(1, 3, 2 * chunk, 224, 224, generator=torch.Generator().manual_seed(5)).to(dtype)
    oracle = orc.OracleVideoMamba(cfg, sd)
    m = _model_from(cfg, sd, dtype)
    zero = [(torch.zeros(1, 2 * dim, 4, dtype=dtype), torch.zeros(1, 2 * dim, 16, dtype=dtype))
            for _ in range(depth)]
    with torch.no_grad():
        w0 = oracle.forward(x[:, :, :chunk], ssm_state=zero, temporal_pos_offset=0)
        oracle.pool_type = "avg"
        w1 = oracle.forward(x[:, :, chunk:], ssm_state=w0[2], temporal_pos_offset=chunk)
        st = m.allocate_state(1, dtype=dtype, device=DEV)
        g0 = m(x[:, :, :chunk].to(DEV), ssm_state=st, temporal_pos_offset=0)
        m.pool_type = "avg"
        g1 = m(x[:, :, chunk:].to(DEV), ssm_state=g0[2], temporal_pos_offset=chunk)
    assert g0[0].shape == (1, chunk * 196, dim) and g1[0].shape == (1, chunk * 196, dim)
    for got, want in ((g0, w0), (g1, w1)):
        _close(got[0], want[0], 2e-2); _close(got[1], want[1], 2e-2)
        for (gc, gs), (wc, ws) in zip(got[2], want[2]):
            assert gs.dtype == torch.float32 and gc.shape == wc.shape
            assert rel_err(gc, wc) <= 2e-2 and rel_err(gs, ws) <= 2e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mixer_forced_generic_equals_auto_path(dtype):
    """Both kernel selections of vmb_mixer_fwd must agree with the oracle (and so each other)."""
    from videomamba_b200 import ops
    torch.manual_seed(5)
    mx = Mamba(d_model=384, use_fast_path=False).eval().to(dtype)
    p = {k: v.detach().clone() for k, v in mx.state_dict().items()}
    x = torch.randn(2, 300, 384).to(dtype)
    want, (wc, ws) = orc.mixer_ref(p, x, want_state=True)
    mx.to(DEV)
    w = mx._kernel_weights()
    for path in (0, 1) + ((2,) if dtype == torch.bfloat16 else ()):   # 2 = fused kernels forced
        out, cs, ss = ops.mixer_fwd(w, x.to(DEV), None, None, True, True, path=path)
        assert rel_err(out, want) <= _tol(dtype)
        assert torch.equal(cs.cpu(), wc) and rel_err(ss, ws) <= _tol(dtype)
    # reversed walk == flip / forward / flip
    rev, _, _ = ops.mixer_fwd(w, x.to(DEV), reverse=True)
    want_rev = torch.flip(orc.mixer_ref(p, torch.flip(x, dims=[1])), dims=[1])
    assert rel_err(rev, want_rev) <= _tol(dtype)


# ---- the reference's own CUDA test cases ----------------------------------------------------------
def _small_model(**over):
    kw = dict(img_size=8, patch_size=4, depth=2, embed_dim=16, channels=3,
              ssm_cfg={"use_fast_path": False}, fused_add_norm=False, rms_norm=False,
              residual_in_fp32=False, kernel_size=1, num_frames=4)
    kw.update(over)
    return video_mamba.PretrainVideoMamba(**kw)


def test_reference_cuda_suite_shapes_and_errors():
    m = _small_model().cuda().eval()
    x = torch.randn(1, 3, 4, 8, 8, device=DEV)
    with torch.no_grad():
        x_vis, x_pool = m(x, mask=None, use_image=False)
        assert x_vis.shape == (1, 16, 16) and x_pool.shape == (1, 1, 16)
        feats = m.forward_features(x)
        assert isinstance(feats, torch.Tensor) and feats.shape == (1, 17, 16)
        st = m.init_state(batch_size=1, dtype=x.dtype, device=x.device)
        fv, nxt = m.forward_features(x[:, :, :2], ssm_state=st, temporal_pos_offset=0)
        assert isinstance(nxt, list) and len(nxt) == m.depth
        mask = torch.zeros(1, 17, dtype=torch.bool, device=DEV)
        mv, mp = m(x, mask=mask)
        assert mv.shape[0] == 1 and mp.shape[0] == 1
    nopool = _small_model(add_pool_norm=False).cuda().eval()
    with torch.no_grad():
        out = nopool(x)
    assert isinstance(out, torch.Tensor) and out.shape == (1, 17, 16)
    with pytest.raises(ValueError, match="mask token length mismatch"):
        m(x, mask=torch.zeros(1, 16, dtype=torch.bool, device=DEV))
    bad = torch.zeros(1, 17, dtype=torch.bool, device=DEV); bad[:, 0] = True
    with pytest.raises(ValueError, match="CLS token visible"):
        m(x, mask=bad)
    allp = torch.ones(1, 17, dtype=torch.bool, device=DEV); allp[:, 0] = False
    with pytest.raises(ValueError, match="at least one patch token visible"):
        m(x, mask=allp)
    # non-square runtime resolution (regressions.py:463-469)
    with torch.no_grad():
        xv, xp = m(torch.randn(1, 3, 4, 12, 8, device=DEV))
    assert xv.shape == (1, 4 * 3 * 2, 16) and xp.shape == (1, 1, 16)
    # tubelet models (regressions.py:446-459)
    tub = _small_model(kernel_size=2, num_frames=4).cuda().eval()
    with torch.no_grad():
        xv, xp = tub(x, mask=torch.zeros(1, 1 + 2 * 4, dtype=torch.bool, device=DEV), use_image=True)
    assert xv.shape == (1, 8, 16) and xp.shape == (1, 1, 16)
    # temporal offset changes outputs when the table is non-zero (regressions.py:419-428)
    tm = _small_model(num_frames=8, add_pool_norm=False).cuda().eval()
    with torch.no_grad():
        tm.temporal_pos_embedding.copy_(torch.randn_like(tm.temporal_pos_embedding))
        assert not torch.allclose(tm.forward_features(x, temporal_pos_offset=0),
                                  tm.forward_features(x, temporal_pos_offset=2))


def test_reference_cuda_suite_streaming_and_cache():
    m = _small_model(add_pool_norm=False).cuda().eval()
    x = torch.randn(1, 3, 8, 8, 8, device=DEV)
    with torch.no_grad():
        full = m(x)
        st = m.init_state(batch_size=1, dtype=x.dtype, device=x.device)
        a, st = m(x[:, :, :4], ssm_state=st, temporal_pos_offset=0)
        b, nxt = m(x[:, :, 4:], ssm_state=st, temporal_pos_offset=4)
    assert a.shape[1] == 1 + 4 * 4 and b.shape[1] == 4 * 4
    video_mamba.validate_state(m, nxt, batch_size=1)
    torch.testing.assert_close(torch.cat([a, b], 1), full, rtol=1e-2, atol=1e-2)
    assert rel_err(torch.cat([a, b], 1), full) <= 1e-5

    mx = Mamba(d_model=8, d_state=4, d_conv=2, expand=2, use_fast_path=False, layer_idx=0).cuda().eval()
    cache = SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    with torch.no_grad():
        out_a = mx(torch.randn(2, 1, 8, device=DEV), inference_params=cache)
        cache.seqlen_offset = 1
        out_b = mx(torch.randn(1, 1, 8, device=DEV), inference_params=cache)
    conv, ssm = cache.key_value_memory_dict[0]
    assert out_a.shape == (2, 1, 8) and out_b.shape == (1, 1, 8)
    assert conv.shape[0] == 1 and ssm.shape[0] == 1

    blk = create_block(d_model=16, ssm_cfg={"use_fast_path": False}, rms_norm=False,
                       fused_add_norm=False, residual_in_fp32=False, layer_idx=0).cuda()
    xb = torch.randn(2, 3, 16, device=DEV)
    st = blk.mixer.allocate_state(batch_size=2, dtype=xb.dtype, device=xb.device)
    with torch.no_grad():
        assert len(blk(xb, state=st, return_state=False)) == 2
        assert len(blk(xb, state=st, return_state=True)) == 3
    for pool in ("cls+avg", "cls_cat_avg"):
        pm = _small_model(pool_type=pool).cuda().eval()
        with pytest.raises(ValueError, match="requires a CLS token"):
            pm(x[:, :, :2], keep_temporal=True,
               ssm_state=pm.init_state(1, dtype=x.dtype, device=x.device), temporal_pos_offset=1)


def test_mixer_odd_lengths_and_edge_cases():
    """Ragged / tiny inputs: L smaller than d_conv, L = 1, batch 1, empty batch."""
    torch.manual_seed(3)
    mx = Mamba(d_model=16, d_state=8, d_conv=4, use_fast_path=False).eval()
    p = {k: v.detach().clone() for k, v in mx.state_dict().items()}
    mx.to(DEV)
    for L in (1, 2, 3, 4, 5, 31, 33, 129):
        x = torch.randn(2, L, 16)
        st = (torch.randn(2, 32, 4), torch.randn(2, 32, 8))
        want, (wc, ws) = orc.mixer_ref(p, x, st[0], st[1], want_state=True)
        with torch.no_grad():
            got, (gc, gs) = mx(x.to(DEV), state=(st[0].to(DEV), st[1].to(DEV)), return_state=True)
        assert rel_err(got, want) <= 1e-5 and torch.equal(gc.cpu(), wc) and rel_err(gs, ws) <= 1e-5
        want, (wc, ws) = orc.mixer_ref(p, x, want_state=True)
        with torch.no_grad():
            got, (gc, gs) = mx(x.to(DEV), return_state=True)
        assert rel_err(got, want) <= 1e-5 and torch.equal(gc.cpu(), wc) and rel_err(gs, ws) <= 1e-5
    with torch.no_grad():
        assert mx(torch.randn(0, 5, 16, device=DEV)).shape == (0, 5, 16)


def test_run_to_run_determinism():
    torch.manual_seed(0)
    mx = Mamba(d_model=384).eval().to(torch.bfloat16).to(DEV)
    x = torch.randn(4, 777, 384, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        a = mx(x)
        b = mx(x)
    assert torch.equal(a, b)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_refiner_oracle_parity_production_width(dtype, tol):
    """BiMambaRefinerBlock (refiner_backbone.py:92-135) at the Small width against the oracle: the
    backward block walks the tokens back to front (no flip copies for 3-D input), the fusion gate is
    two projections + one sigmoid/blend kernel, out_proj the tensor-core projection; 4-D input flips
    frames only."""
    torch.manual_seed(0)
    dim = 384
    blk = video_mamba.BiMambaRefinerBlock(dim=dim, ssm_cfg={"use_fast_path": False}, layer_idx=0).eval()
    with torch.no_grad():
        for m in (blk.block_fwd.mixer, blk.block_bwd.mixer):
            m.A_log.add_(0.1 * torch.randn_like(m.A_log))
    sd = {k: v.detach().clone().to(dtype) for k, v in blk.state_dict().items()}
    blk.load_state_dict(sd)
    blk = blk.to(dtype).to(DEV)
    gen = torch.Generator().manual_seed(1)
    x3 = torch.randn(2, 300, dim, generator=gen).to(dtype)
    x4 = torch.randn(2, 5, 49, dim, generator=gen).to(dtype)
    for x in (x3, x4):
        want, want_state = orc.refiner_ref(sd, x)
        with torch.no_grad():
            got, got_state = blk(x.to(DEV))
        assert got.shape == want.shape and got.dtype == dtype
        _close(got, want, tol)
        for a, b in zip(got_state, want_state):
            _close(a, b, tol)


def test_graphed_stream_equals_eager_chunks():
    """videomamba_b200.graphed.GraphedStream: continuation chunks replayed from one CUDA graph (state
    fed back inside the graph, temporal rows of the current offset in a fixed buffer) are
    bit-identical to the eager ``model(x, ssm_state=prev, temporal_pos_offset=off)`` chain."""
    from videomamba_b200.graphed import GraphedStream
    torch.manual_seed(0)
    bf = torch.bfloat16
    m = video_mamba.PretrainVideoMamba(img_size=64, patch_size=16, depth=3, embed_dim=192, channels=3,
                                       ssm_cfg={"use_fast_path": False}, num_frames=16,
                                       pool_type="avg").eval()
    with torch.no_grad():
        m.temporal_pos_embedding.normal_(0, 0.02)
    m = m.to(bf).to(DEV)
    B, T = 3, 2
    gen = torch.Generator().manual_seed(3)
    chunks = [torch.rand(B, 3, T, 64, 64, generator=gen).to(bf).to(DEV) for _ in range(5)]
    with torch.no_grad():
        st = m.allocate_state(B, dtype=bf, device=DEV)
        eager = []
        for i, x in enumerate(chunks):
            vis, pool, st = m(x, ssm_state=st, temporal_pos_offset=i * T)
            eager.append((vis.clone(), pool.clone()))
    runner = GraphedStream(m)
    vis, pool = runner.first(chunks[0])
    assert torch.equal(vis, eager[0][0]) and torch.equal(pool, eager[0][1])
    for i in range(1, 5):
        vis, pool = runner.step(chunks[i])
        assert runner.offset == (i + 1) * T
        assert torch.equal(vis, eager[i][0]), i
        assert torch.equal(pool, eager[i][1]), i
    for (c, s), (ec, es) in zip(runner.state, st):
        assert torch.equal(c, ec) and torch.equal(s, es)
    # the streams restart: the first chunk (CLS included, zero state) is a graph replay of its own, and the
    # continuation graph keeps working on the same carried buffers
    for _ in range(2):
        vis, pool = runner.first(chunks[0])
        assert runner.offset == T and runner._first_graph is not None
        assert torch.equal(vis, eager[0][0]) and torch.equal(pool, eager[0][1])
        for i in range(1, 3):
            vis, pool = runner.step(chunks[i])
            assert torch.equal(vis, eager[i][0]) and torch.equal(pool, eager[i][1]), i
    # an explicit initial state runs the first chunk eagerly
    with torch.no_grad():
        st0 = m.allocate_state(B, dtype=bf, device=DEV)
    vis, pool = runner.first(chunks[0], state=st0)
    assert torch.equal(vis, eager[0][0]) and torch.equal(runner.step(chunks[1])[0], eager[1][0])
    with pytest.raises(ValueError, match="pool_type='avg'"):
        GraphedStream(video_mamba.PretrainVideoMamba(img_size=32, patch_size=16, depth=1, embed_dim=32,
                                                     channels=3, num_frames=4, pool_type="cls+avg"))
