"""CPU: the package's host logic (module control flow, state plumbing, CLS / mask / pooling
rules, container types) against the golden fixtures generated from the live reference.

The kernels are replaced by the oracle through tests/oracle_backend.py (test-only monkeypatch;
the product has no CPU path), so what is under test here is everything ABOVE the C ABI.
fp32 bar 1e-5, bf16 bar 2e-2 (relative = max|a-b| / max|b|)."""
from types import SimpleNamespace

import pytest
import torch

import video_mamba
from oracle.videomamba_oracle import rel_err
from tests import oracle_backend
from video_mamba.mamba_simple import Mamba
from videomamba_b200.block import create_block


@pytest.fixture(autouse=True)
def _oracle_kernels(monkeypatch):
    oracle_backend.install(monkeypatch)


def _tol(t):
    return 2e-2 if t.dtype == torch.bfloat16 else 1e-5


def _close(got, want, tol=None):
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    assert rel_err(got, want) <= (tol or _tol(want))


def _model(g, **over):
    cfg = g["cfg"]
    m = video_mamba.PretrainVideoMamba(
        img_size=cfg["img_size"], patch_size=cfg["patch_size"], depth=cfg["depth"],
        embed_dim=cfg["embed_dim"], channels=3, ssm_cfg={"use_fast_path": False},
        rms_norm=cfg["rms_norm"], fused_add_norm=cfg["fused_add_norm"],
        residual_in_fp32=cfg["residual_in_fp32"], kernel_size=cfg["kernel_size"],
        num_frames=cfg["num_frames"], pool_type=cfg["pool_type"], **over).eval()
    m = m.to(g["x"].dtype)
    sd = g["sd"]
    if not over.get("add_pool_norm", True):
        sd = {k: v for k, v in sd.items() if not k.startswith("pool_norm")}
    m.load_state_dict(sd, strict=True)   # reference state_dict loads 1:1
    return m


CASES = ["model_fp32_rms_fused.pt", "model_bf16_rms_fused.pt", "model_fp32_ln_unfused.pt"]


@pytest.mark.parametrize("name", CASES)
def test_forward_contracts_match_reference(golden, name):
    g = golden(name)
    m = _model(g)
    with torch.no_grad():
        x_vis, x_pool = m(g["x"])
        _close(x_vis, g["x_vis"])
        _close(x_pool, g["x_pool"])
        _close(m.forward_features(g["x"]), g["features"])
        _close(m(g["x"], keep_temporal=True)[1], g["x_pool_keep_temporal"])
        mv, mp = m(g["x"], mask=g["mask"])
        _close(mv, g["x_vis_masked"])
        _close(mp, g["x_pool_masked"])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("container", ["list", "tuple", "dict"])
def test_streaming_matches_reference_and_keeps_container_type(golden, name, container):
    g = golden(name)
    m = _model(g)
    dt = g["x"].dtype
    state = video_mamba.allocate_state(m, 2, dtype=dt, as_dict=container == "dict")
    if container == "tuple":
        state = tuple(state)
    with torch.no_grad():
        a_vis, a_pool, s1 = m(g["x"][:, :, :2], ssm_state=state, temporal_pos_offset=0)
        m.pool_type = "avg"
        b_vis, b_pool, s2 = m(g["x"][:, :, 2:], ssm_state=s1, temporal_pos_offset=2)
    _close(a_vis, g["chunk0_vis"])
    _close(a_pool, g["chunk0_pool"])
    _close(b_vis, g["chunk1_vis"])
    _close(b_pool, g["chunk1_pool"])
    assert type(s2) is {"list": list, "tuple": tuple, "dict": dict}[container]
    video_mamba.validate_state(m, s2, 2)
    items = [s2[i] for i in range(m.depth)]
    for (c, s), (rc, rs) in zip(items, g["state2"]):
        _close(c, rc, _tol(g["x"]))
        _close(s, rs, _tol(g["x"]))
    # input state objects are not mutated on the full-state path
    first = state[0] if container != "dict" else state[0]
    assert float(first[0].abs().max()) == 0.0 and float(first[1].abs().max()) == 0.0


def test_continuation_chunk_omits_cls_and_rejects_cls_pooling(golden):
    g = golden("model_fp32_rms_fused.pt")
    m = _model(g, add_pool_norm=False)
    x = g["x"]
    with torch.no_grad():
        st = m.init_state(2, dtype=x.dtype)
        first, st = m(x[:, :, :2], ssm_state=st, temporal_pos_offset=0)
        second, nxt = m(x[:, :, 2:], ssm_state=st, temporal_pos_offset=2)
    assert first.shape[1] == 1 + 2 * 16 and second.shape[1] == 2 * 16
    video_mamba.validate_state(m, nxt, 2)
    pooled = _model(g)
    for pool in ("cls+avg", "cls_cat_avg", "cls"):
        pooled.pool_type = pool
        with pytest.raises(ValueError, match="requires a CLS token"):
            pooled(x[:, :, 2:], keep_temporal=True, ssm_state=pooled.init_state(2),
                   temporal_pos_offset=1)


def test_chunked_equals_full_sequence(golden):
    """The reference's own acceptance test (tests/test_videomamba_regressions.py:563-588) asks
    1e-2; chunked == full to fp32 round-off here (zero temporal table, frames <= num_frames)."""
    g = golden("model_fp32_rms_fused.pt")
    m = _model(g, add_pool_norm=False)
    x = g["x"]
    with torch.no_grad():
        full = m(x)
        st = m.init_state(2, dtype=x.dtype)
        a, st = m(x[:, :, :1], ssm_state=st, temporal_pos_offset=0)
        b, st = m(x[:, :, 1:3], ssm_state=st, temporal_pos_offset=1)
        c, st = m(x[:, :, 3:], ssm_state=st, temporal_pos_offset=3)
    assert rel_err(torch.cat([a, b, c], 1), full) < 1e-5


def test_legacy_ssm_only_state_is_updated_in_place(golden):
    g = golden("model_fp32_rms_fused.pt")
    m = _model(g)
    legacy = m.init_ssm_state(2, dtype=torch.float32)
    with torch.no_grad():
        vis, _pool, out_state = m(g["x"][:, :, :2], ssm_state=legacy, temporal_pos_offset=0)
    _close(vis, g["legacy_vis"])
    assert out_state is legacy
    for s, r in zip(legacy, g["legacy_state"]):
        _close(s, r)
    assert set(m.init_ssm_state(1, as_dict=True)) == {0, 1, 2}


def test_mask_validation_messages(golden):
    g = golden("model_fp32_rms_fused.pt")
    m = _model(g)
    x = g["x"]
    n = 1 + 4 * 16
    with pytest.raises(ValueError, match="mask token length mismatch"):
        m(x, mask=torch.zeros(2, n - 1, dtype=torch.bool))
    bad = torch.zeros(2, n, dtype=torch.bool)
    bad[:, 0] = True
    with pytest.raises(ValueError, match="CLS token visible"):
        m(x, mask=bad)
    uneven = torch.zeros(2, n, dtype=torch.bool)
    uneven[0, 3:7] = True
    uneven[1, 3:11] = True
    with pytest.raises(ValueError, match="same number of visible tokens"):
        m(x, mask=uneven)
    allp = torch.ones(2, n, dtype=torch.bool)
    allp[:, 0] = False
    with pytest.raises(ValueError, match="at least one patch token visible"):
        m(x, mask=allp)
    few = torch.ones(2, n, dtype=torch.bool)
    few[:, [0, 1, 2]] = False
    with pytest.raises(ValueError, match="at least one visible patch token"):
        m(x, mask=few, keep_temporal=True)
    with pytest.raises(ValueError, match="mask batch size mismatch"):
        m(x, mask=torch.zeros(1, n, dtype=torch.bool))
    with pytest.raises(ValueError, match="temporal_pos_offset must be non-negative"):
        m(x, temporal_pos_offset=-1)


def test_pool_variants_shapes(golden):
    g = golden("model_fp32_rms_fused.pt")
    x = g["x"]
    for pool, want_t, want in (("cls", (2, 1, 32), (2, 1, 32)), ("avg", (2, 4, 32), (2, 1, 32)),
                               ("cls_cat_avg", (2, 5, 32), (2, 2, 32)),
                               ("cls+avg", (2, 4, 32), (2, 1, 32))):
        m = _model(g)
        m.pool_type = pool
        with torch.no_grad():
            assert m(x, keep_temporal=True)[1].shape == want_t
            assert m(x)[1].shape == want
    m.pool_type = "bogus"
    with pytest.raises(ValueError, match="Unsupported pool_type"):
        m(x)
    # runtime resolution differs from the build resolution -> bicubic table resize
    m = _model(g)
    with torch.no_grad():
        xv, xp = m(torch.rand(1, 3, 2, 24, 16))
    assert xv.shape == (1, 2 * 3 * 2, 32) and xp.shape == (1, 1, 32)


def test_mixer_state_flows(golden):
    g = golden("mixer_fp32.pt")
    mx = Mamba(d_model=16, d_state=8, d_conv=4, expand=2, use_fast_path=False).eval()
    mx.load_state_dict(g["sd"])
    x = g["x"]
    with torch.no_grad():
        _close(mx(x), g["full"])
        o1, st1 = mx(x[:, :5], return_state=True)
        o2, st2 = mx(x[:, 5:], state=st1, return_state=True)
        _close(o1, g["out1"]); _close(o2, g["out2"])
        for a, b in zip(st1 + st2, g["state1"] + g["state2"]):
            _close(a, b)
        # state given, return_state False -> tensor only, state untouched
        keep = (st1[0].clone(), st1[1].clone())
        o2b = mx(x[:, 5:], state=st1)
        assert isinstance(o2b, torch.Tensor) and torch.equal(st1[0], keep[0])
        _close(o2b, g["out2"])
        # ssm_state only -> in-place update, conv has no history
        ssm = st1[1].clone()
        mx(x[:, 5:], ssm_state=ssm)
        assert not torch.equal(ssm, st1[1])


def test_inference_cache_prefill_step_and_resize(golden):
    s = golden("mixer_fp32.pt")["small"]
    mx = Mamba(d_model=8, d_state=4, d_conv=2, expand=2, use_fast_path=False, layer_idx=0).eval()
    mx.load_state_dict(s["sd"])
    cache = SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    with torch.no_grad():
        _close(mx(s["x"][:, :3], inference_params=cache), s["prefill"])
        cache.seqlen_offset = 3
        _close(mx(s["x"][:, 3:4], inference_params=cache), s["step1"])
        cache.seqlen_offset = 4
        _close(mx(s["x"][:, 4:5], inference_params=cache), s["step2"])
        conv, ssm = cache.key_value_memory_dict[0]
        _close(conv, s["cache_conv"]); _close(ssm, s["cache_ssm"])
        # batch size change re-allocates the cache entry (regressions.py:473-493)
        out = mx(torch.randn(1, 1, 8), inference_params=cache)
    assert out.shape == (1, 1, 8)
    conv, ssm = cache.key_value_memory_dict[0]
    assert conv.shape[0] == 1 and ssm.shape[0] == 1
    assert set(video_mamba.PretrainVideoMamba(img_size=8, patch_size=4, depth=2, embed_dim=16)
               .allocate_inference_cache(3, 1)) == {0, 1}


def test_block_return_arity():
    blk = create_block(d_model=16, ssm_cfg={"use_fast_path": False}, rms_norm=False,
                       fused_add_norm=False, residual_in_fp32=False, layer_idx=0)
    x = torch.randn(2, 3, 16)
    st = blk.mixer.allocate_state(batch_size=2, dtype=x.dtype)
    with torch.no_grad():
        assert len(blk(x, state=st, return_state=False)) == 2
        assert len(blk(x, state=st, return_state=True)) == 3
        assert len(blk(x)) == 2
        assert len(blk(x, state=st, return_state=True, use_checkpoint=True)) == 3


def test_refiner_matches_reference(golden):
    g = golden("refiner_fp32.pt")
    blk = video_mamba.BiMambaRefinerBlock(dim=16, ssm_cfg={"use_fast_path": False},
                                          layer_idx=0).eval()
    blk.load_state_dict(g["sd"])
    with torch.no_grad():
        y3, s3 = blk(g["x3"])
        y4, s4 = blk(g["x4"])
    _close(y3, g["y3"]); _close(y4, g["y4"])
    for a, b in zip(s3 + s4, g["s3"] + g["s4"]):
        _close(a, b)
    with pytest.raises(ValueError, match=r"\[B, L, C\] or \[B, T, N, C\]"):
        blk(torch.randn(3, 4))
    f, b = blk.allocate_state(2)
    assert f[0].shape == (2, 32, 4) and b[1].shape == (2, 32, 16)


def test_forward_only_guard_makes_backward_fail_loudly():
    """every libvmb200 op writes into fresh buffers with no autograd node; outputs computed from
    tensors that require grad must carry a node whose backward raises, instead of silently training only
    the torch glue."""
    import pytest
    import torch
    from videomamba_b200 import ops

    w = torch.randn(4, 4, requires_grad=True)
    out = torch.zeros(2, 4)                       # stands for a kernel output buffer
    tagged = ops.forward_only(out, w)
    assert tagged.requires_grad and torch.equal(tagged, out)
    with pytest.raises(NotImplementedError, match="forward-only"):
        tagged.sum().backward()
    with torch.no_grad():                         # inference: nothing is attached
        assert ops.forward_only(out, w) is out
    assert ops.forward_only(out, w.detach()) is out
    both = ops.forward_only((out, None, out), w)
    assert both[0].requires_grad and both[1] is None and both[2] is out


def test_mixer_weight_cache_invalidation_hooks():
    """the kernel-ready copies of a mixer's parameters are dropped by load_state_dict, by
    .to() / dtype casts and by refresh_weights(); the key follows storage, version and dtype."""
    import torch
    from videomamba_b200.mixer import Mamba

    m = Mamba(d_model=16, d_state=4)
    k0 = m.weights_key()
    m._weights_key, m._weights = k0, object()
    with torch.no_grad():
        m.A_log.add_(1.0)                         # in-place through autograd's counter: key changes
    assert m.weights_key() != k0
    m._weights_key, m._weights = m.weights_key(), object()
    m.load_state_dict(m.state_dict())
    assert m._weights is None and m._weights_key is None
    m._weights_key, m._weights = m.weights_key(), object()
    m.to(torch.bfloat16)
    assert m._weights is None
    m._weights_key, m._weights = m.weights_key(), object()
    m.A_log.data.mul_(2.0)                        # bypasses the version counter: needs refresh_weights()
    assert m.weights_key() == m._weights_key
    m.refresh_weights()
    assert m._weights is None


def test_split_xz_gradient_arena():
    """``autograd.XZGrad``: the backward nodes of the x / z halves write into one d(xz) buffer and ``SplitXZ.backward``
    returns it without copies; anything else (a foreign gradient, a missing half) takes the copying path and
    gives the same values (host logic only: no kernel runs here)."""
    from videomamba_b200 import autograd as ag
    B, L, Di = 2, 5, 4
    leaf = torch.randn(B, L, 2 * Di, requires_grad=True)
    xz = leaf * 1.0                       # a non-leaf, so a hook sees the tensor SplitXZ.backward returned
    seen = []
    xz.register_hook(lambda g: seen.append(g.data_ptr()))
    # (1) both halves written into the arena -> the buffer itself comes back
    arena = ag.XZGrad(Di)
    x, z = ag.SplitXZ.apply(xz, Di, arena)
    gx = arena.half(0, B, L, xz.dtype, xz.device)
    gz = arena.half(1, B, L, xz.dtype, xz.device)
    buf = arena.buf
    gx.copy_(torch.full((B, L, Di), 2.0))
    gz.copy_(torch.full((B, L, Di), 3.0))
    assert arena.owns(0, gx) and arena.owns(1, gz) and not arena.owns(0, gz)
    torch.autograd.backward([x, z], [gx, gz])
    assert seen == [buf.data_ptr()] and arena.buf is None
    assert torch.equal(leaf.grad[..., :Di], torch.full((B, L, Di), 2.0))
    assert torch.equal(leaf.grad[..., Di:], torch.full((B, L, Di), 3.0))
    # (2) a gradient that does not live in the arena, and a missing half: copied / zero-filled
    xz2 = torch.randn(B, L, 2 * Di, requires_grad=True)
    arena2 = ag.XZGrad(Di)
    x2, z2 = ag.SplitXZ.apply(xz2, Di, arena2)
    foreign = torch.full((B, L, Di), 5.0)
    torch.autograd.backward([x2], [foreign])
    assert torch.equal(xz2.grad[..., :Di], foreign) and torch.count_nonzero(xz2.grad[..., Di:]) == 0
    # (3) without an arena the node behaves the same
    xz3 = torch.randn(B, L, 2 * Di, requires_grad=True)
    x3, z3 = ag.SplitXZ.apply(xz3, Di)
    torch.autograd.backward([x3, z3], [foreign, 2 * foreign])
    assert torch.equal(xz3.grad, torch.cat([foreign, 2 * foreign], dim=-1))
