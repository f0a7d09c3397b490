"""CPU: public-API contract, mirroring the reference's CPU tests
(tests/test_public_api_contract.py:29-66, tests/test_videomamba_regressions.py:35-171, :287-289)."""
from types import SimpleNamespace
from typing import Any

import pytest
import torch

import video_mamba
import videomamba_b200
import videomamba_b200.model as model_module
from video_mamba.mamba_simple import Mamba
from video_mamba.videomamba import PretrainVideoMamba, build_videomamba, load_state_dict
from videomamba_b200.block import create_block


def _small_model(**overrides: Any) -> PretrainVideoMamba:
    kwargs: dict = dict(img_size=8, patch_size=4, depth=2, embed_dim=16, channels=3,
                        ssm_cfg={"use_fast_path": False}, fused_add_norm=False, rms_norm=False,
                        residual_in_fp32=False, kernel_size=1, num_frames=4)
    kwargs.update(overrides)
    return PretrainVideoMamba(**kwargs)


def _cfg(**over):
    base = dict(img_size=8, patch_size=4, depth=2, embed_dim=16, channels=3, drop_path_rate=0.0,
                ssm_cfg={"use_fast_path": False}, norm_epsilon=1e-5, fused_add_norm=False,
                rms_norm=False, residual_in_fp32=False, bimamba=True, pool_type="cls+avg",
                kernel_size=1, num_frames=4, use_checkpoint=False, checkpoint_num=0,
                pretrained=None, ckpt_num_frame=4)
    base.update(over)
    return SimpleNamespace(vision_encoder=SimpleNamespace(**base))


def test_alias_package_reexports_the_same_objects():
    assert video_mamba.build_videomamba is videomamba_b200.build_videomamba
    assert video_mamba.PretrainVideoMamba is videomamba_b200.PretrainVideoMamba
    assert video_mamba.BiMambaRefinerBlock is videomamba_b200.BiMambaRefinerBlock
    assert video_mamba.STREAMING_CONTRACT_VERSION == "1.0.0"
    for name in video_mamba.__all__:
        assert hasattr(video_mamba, name), name


def test_streaming_contract_allocate_and_validate_cpu():
    model = _small_model()
    state = video_mamba.allocate_state(model, batch_size=2, dtype=torch.float32)
    video_mamba.validate_state(model, state, batch_size=2)
    shapes = video_mamba.expected_state_shapes(model, batch_size=2)
    assert len(shapes) == model.depth
    assert shapes[0].conv_state == (2, model.layers[0].mixer.d_inner, 4)
    assert shapes[0].ssm_state == (2, model.layers[0].mixer.d_inner, 16)
    assert model.expected_state_shapes(2) == shapes
    as_dict = video_mamba.allocate_state(model, batch_size=2, as_dict=True)
    assert sorted(as_dict) == [0, 1]
    video_mamba.validate_state(model, as_dict, batch_size=2)
    video_mamba.validate_state(model, tuple(state), batch_size=2)


def test_validate_state_error_messages():
    model = _small_model()
    state = video_mamba.allocate_state(model, batch_size=2)
    with pytest.raises(ValueError, match="State length mismatch: expected 2, got 1"):
        video_mamba.validate_state(model, state[:1], 2)
    with pytest.raises(ValueError, match="State dict keys mismatch"):
        video_mamba.validate_state(model, {0: state[0], 5: state[1]}, 2)
    with pytest.raises(TypeError, match="list, tuple, or dict"):
        video_mamba.validate_state(model, "nope", 2)
    with pytest.raises(TypeError, match="2-tuple"):
        video_mamba.validate_state(model, [state[0], state[1][0]], 2)
    with pytest.raises(TypeError, match="must both be tensors"):
        video_mamba.validate_state(model, [state[0], (state[1][0], None)], 2)
    with pytest.raises(ValueError, match="Layer 1 conv_state shape mismatch"):
        video_mamba.validate_state(model, [state[0], (state[1][0][:1], state[1][1])], 2)
    with pytest.raises(ValueError, match="Layer 0 ssm_state shape mismatch"):
        video_mamba.validate_state(model, [(state[0][0], state[0][1][:, :, :3]), state[1]], 2)
    with pytest.raises(ValueError, match="positive integer"):
        video_mamba.expected_state_shapes(model, 0)
    with pytest.raises(TypeError, match="allocate_state"):
        video_mamba.allocate_state(object(), 1)


def test_model_contract_metadata_and_forward_semantics():
    model = _small_model(add_pool_norm=True)
    assert model.streaming_contract_version == video_mamba.STREAMING_CONTRACT_VERSION
    sem = model.forward_return_semantics()
    assert (sem.without_state, sem.with_state) == ("(x_vis, x_pool)", "(x_vis, x_pool, next_state)")
    sem = _small_model(add_pool_norm=False).forward_return_semantics()
    assert (sem.without_state, sem.with_state) == ("x_vis", "(x_vis, next_state)")
    assert video_mamba.model_forward_return_semantics(model).without_state == "(x_vis, x_pool)"


def test_configure_determinism_reseeds_torch_rng():
    video_mamba.configure_determinism(seed=1234, deterministic=True)
    a = torch.randn(8)
    cfg = video_mamba.configure_determinism(seed=1234, deterministic=True)
    b = torch.randn(8)
    torch.testing.assert_close(a, b)
    assert cfg.deterministic and not cfg.cudnn_benchmark and not cfg.allow_tf32
    video_mamba.configure_determinism(seed=0, deterministic=False)


def test_bimamba_false_is_rejected():
    with pytest.raises(NotImplementedError, match="bimamba=True"):
        _small_model(bimamba=False)


def test_build_videomamba_namespace_with_pretrained(tmp_path):
    path = tmp_path / "mini.pt"
    torch.save(_small_model().state_dict(), path)
    model = build_videomamba(_cfg(pretrained=str(path)))
    assert isinstance(model, PretrainVideoMamba)
    assert "url" in model.default_cfg


def test_load_state_dict_rejects_wrapped_checkpoint(tmp_path):
    path = tmp_path / "wrapped.pt"
    torch.save({"model": _small_model().state_dict()}, path)
    with pytest.raises(ValueError, match="plain state_dict checkpoint"):
        build_videomamba(_cfg(pretrained=str(path)))


def test_build_videomamba_requires_channels_attr():
    cfg = _cfg()
    del cfg.vision_encoder.channels
    cfg.vision_encoder.in_chans = 3
    with pytest.raises(AttributeError):
        build_videomamba(cfg)


def test_load_state_dict_uses_weights_only(tmp_path, monkeypatch):
    model = _small_model()
    path = tmp_path / "mini.pt"
    torch.save(model.state_dict(), path)
    seen: dict = {}
    real = model_module.torch.load

    def spy(*a, **k):
        seen.update(k)
        return real(*a, **k)

    monkeypatch.setattr(model_module.torch, "load", spy)
    load_state_dict(str(path), model, ckpt_num_frame=4, num_frames=4)
    assert seen.get("weights_only") is True


def test_load_state_dict_interpolates_pos_embed_for_non_square_target(tmp_path):
    src = _small_model(img_size=8, patch_size=4)
    dst = _small_model(img_size=(8, 12), patch_size=4)
    path = tmp_path / "sq.pt"
    torch.save(src.state_dict(), path)
    load_state_dict(str(path), dst, ckpt_num_frame=4, num_frames=4)
    assert dst.pos_embed.shape == (1, 1 + 2 * 3, 16)


def test_load_state_dict_interpolates_temporal_table(tmp_path):
    src = _small_model(num_frames=4)
    dst = _small_model(num_frames=8)
    path = tmp_path / "t.pt"
    torch.save(src.state_dict(), path)
    load_state_dict(str(path), dst, ckpt_num_frame=4, num_frames=8)
    assert dst.temporal_pos_embedding.shape == (1, 8, 16)
    with pytest.raises(ValueError, match="ckpt_num_frame must be a positive integer"):
        load_state_dict(str(path), src, ckpt_num_frame=0, num_frames=4)


def test_mamba_forward_requires_cuda_tensor_inputs():
    m = Mamba(d_model=8, d_state=4, d_conv=2, expand=2, use_fast_path=False, layer_idx=0).eval()
    with pytest.raises(RuntimeError, match="requires CUDA tensors"):
        m(torch.randn(1, 2, 8))
    with pytest.raises(RuntimeError, match="requires CUDA tensors"):
        m.step(torch.randn(1, 1, 8), *m.allocate_state(1))


def test_mamba_argument_validation_messages():
    m = Mamba(d_model=8, d_state=4, d_conv=2)
    st = m.allocate_state(1)
    with pytest.raises(ValueError, match="Pass either state or ssm_state, not both."):
        m(torch.randn(1, 2, 8), ssm_state=st[1], state=st)
    cache = SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
    with pytest.raises(ValueError, match="state is not supported with inference_params."):
        m(torch.randn(1, 2, 8), inference_params=cache, state=st)
    blk = create_block(8, ssm_cfg={"d_state": 4, "d_conv": 2})
    with pytest.raises(ValueError, match="Pass either state or ssm_state, not both."):
        blk(torch.randn(1, 2, 8), ssm_state=st[1], state=st)


def test_no_weight_decay_includes_temporal_pos_embedding():
    assert "temporal_pos_embedding" in _small_model().no_weight_decay()


def test_state_dict_names_and_flags():
    m = Mamba(d_model=16)
    assert sorted(dict(m.named_parameters())) == sorted([
        "A_log", "D", "in_proj.weight", "conv1d.weight", "conv1d.bias", "x_proj.weight",
        "dt_proj.weight", "dt_proj.bias", "out_proj.weight"])
    assert m.conv1d.weight.shape == (32, 1, 4) and m.A_log.shape == (32, 16)
    assert m.dt_rank == 1 and Mamba(d_model=384).dt_rank == 24
    assert getattr(m.dt_proj.bias, "_no_reinit") and getattr(m.A_log, "_no_weight_decay")
    assert torch.equal(torch.exp(m.A_log[0]).round(), torch.arange(1, 17).float())
    # the backbone's init zeroes dt_proj.bias (segm_init_weights runs before the GPT-2 init)
    model = _small_model()
    assert float(model.layers[0].mixer.dt_proj.bias.detach().abs().max()) == 0.0
    assert float(model.cls_token.detach().abs().max()) == 0.0
    names = set(model.state_dict())
    assert {"cls_token", "pos_embed", "temporal_pos_embedding", "patch_embed.proj.weight",
            "layers.1.mixer.out_proj.weight", "layers.0.norm.weight", "norm.weight",
            "pool_norm.weight"} <= names
    rms = _small_model(rms_norm=True, fused_add_norm=True)
    assert "layers.0.norm.bias" not in rms.state_dict() and rms.layers[0].norm.bias is None


def test_env_switch_disables_fast_path_flag(monkeypatch):
    monkeypatch.setenv("VIDEOMAMBA_DISABLE_FUSED", "Yes")
    assert Mamba(d_model=8).use_fast_path is False
    monkeypatch.setenv("VIDEOMAMBA_DISABLE_FUSED", "0")
    assert Mamba(d_model=8).use_fast_path is True


def test_forward_shape_errors_on_cpu_inputs():
    model = _small_model(kernel_size=2, num_frames=8)
    with pytest.raises(ValueError, match=r"x must have shape \[B, C, T, H, W\]"):
        model(torch.randn(3, 5, 8, 8))
    with pytest.raises(ValueError, match="must be divisible by tubelet size"):
        model(torch.randn(1, 3, 5, 8, 8))
    with pytest.raises(ValueError, match="must be divisible by tubelet size"):
        model.forward_features(torch.randn(1, 3, 5, 8, 8))
    with pytest.raises(ValueError, match="at least one patch"):
        model(torch.randn(1, 3, 4, 2, 8))
