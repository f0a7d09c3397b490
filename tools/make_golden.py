"""Generate tests/golden/*.pt from the LIVE reference (run in the build container only).

    python tools/make_golden.py

Imports the unmodified reference host code from /root/reference through
oracle/ref_loader.py (leaf stand-ins + in-memory is_cuda guard bypass), runs it on CPU with
seeded inputs and stores inputs, weights and outputs.  The fixtures are what pins oracle/
(tests/test_oracle_golden.py) and, on the GPU box, the CUDA path (tests/test_gpu_parity.py).
Every case records the reference call that produced it.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import reference_modules  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _clone_sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def _perturb(model, gen):
    """Break the init-time special structure (zero dt bias, S4D-real A, zero temporal table)."""
    with torch.no_grad():
        for layer in model.layers:
            mx = layer.mixer
            mx.dt_proj.bias.copy_(torch.randn(mx.dt_proj.bias.shape, generator=gen) - 3.0)
            mx.A_log.add_(0.1 * torch.randn(mx.A_log.shape, generator=gen))
            mx.D.add_(0.1 * torch.randn(mx.D.shape, generator=gen))
            mx.conv1d.bias.add_(0.1 * torch.randn(mx.conv1d.bias.shape, generator=gen))
        model.temporal_pos_embedding.copy_(
            0.02 * torch.randn(model.temporal_pos_embedding.shape, generator=gen))
        model.cls_token.copy_(0.02 * torch.randn(model.cls_token.shape, generator=gen))


def model_case(vm, dtype, name, rms, fused):
    cfg = dict(img_size=32, patch_size=8, depth=3, embed_dim=32, channels=3, kernel_size=1,
               num_frames=4, norm_epsilon=1e-5, rms_norm=rms, fused_add_norm=fused,
               residual_in_fp32=True, pool_type="cls+avg", add_pool_norm=True)
    torch.manual_seed(1234)
    model = vm.PretrainVideoMamba(
        img_size=32, patch_size=8, depth=3, embed_dim=32, channels=3,
        ssm_cfg={"use_fast_path": False}, rms_norm=rms, fused_add_norm=fused,
        residual_in_fp32=True, kernel_size=1, num_frames=4, pool_type="cls+avg").eval()
    gen = torch.Generator().manual_seed(99)
    _perturb(model, gen)
    model = model.to(dtype)
    x = torch.rand(2, 3, 4, 32, 32, generator=gen).to(dtype)
    case = {"cfg": cfg, "sd": _clone_sd(model), "x": x,
            "source": "PretrainVideoMamba.forward models/videomamba/videomamba.py:943-1067"}
    with torch.no_grad():
        x_vis, x_pool = model(x)
        case["x_vis"], case["x_pool"] = x_vis, x_pool
        feats = model.forward_features(x)
        case["features"] = feats
        # keep_temporal pooling
        _, pool_t = model(x, keep_temporal=True)
        case["x_pool_keep_temporal"] = pool_t
        # masked
        mask = torch.zeros(2, 1 + 4 * 16, dtype=torch.bool)
        mask[:, 3::4] = True
        mv, mp = model(x, mask=mask)
        case["mask"], case["x_vis_masked"], case["x_pool_masked"] = mask, mv, mp
        # streaming: 2 + 2 frames, list state, CLS only on the first chunk
        state = model.allocate_state(2, dtype=dtype)
        a_vis, a_pool, state1 = model(x[:, :, :2], ssm_state=state, temporal_pos_offset=0)
        model.pool_type = "avg"
        b_vis, b_pool, state2 = model(x[:, :, 2:], ssm_state=state1, temporal_pos_offset=2)
        model.pool_type = "cls+avg"
        case["chunk0_vis"], case["chunk0_pool"] = a_vis, a_pool
        case["chunk1_vis"], case["chunk1_pool"] = b_vis, b_pool
        case["state1"] = [(c.clone(), s.clone()) for c, s in state1]
        case["state2"] = [(c.clone(), s.clone()) for c, s in state2]
        # legacy ssm-only state (in-place update, CLS re-inserted every chunk)
        legacy = model.init_ssm_state(2, dtype=torch.float32)
        l_vis, l_pool, legacy_out = model(x[:, :, :2], ssm_state=legacy, temporal_pos_offset=0)
        case["legacy_vis"] = l_vis
        case["legacy_state"] = [s.clone() for s in legacy_out]
    torch.save(case, os.path.join(OUT, name))
    return case


def temporal_interp_case(vm):
    """Temporal position table with NON-ZERO rows, walked beyond its length: the reference
    interpolates the table to ``offset + T`` rows and slices (models/videomamba/videomamba.py:655-675),
    so the rows a continuation chunk sees depend on where the chunk ends.  Pins that rule: full clip of
    8 frames on a 4-row table, chunks 4 | 4 (second one interpolates to 8), and a ragged 3 | 3 walk
    (second chunk interpolates to 6)."""
    cfg = dict(img_size=16, patch_size=8, depth=2, embed_dim=32, channels=3, kernel_size=1,
               num_frames=4, norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True,
               residual_in_fp32=True, pool_type="cls+avg", add_pool_norm=True)
    torch.manual_seed(4321)
    model = vm.PretrainVideoMamba(
        img_size=16, patch_size=8, depth=2, embed_dim=32, channels=3,
        ssm_cfg={"use_fast_path": False}, rms_norm=True, fused_add_norm=True,
        residual_in_fp32=True, kernel_size=1, num_frames=4, pool_type="cls+avg").eval()
    gen = torch.Generator().manual_seed(77)
    _perturb(model, gen)
    with torch.no_grad():       # rows of visibly different size, so a wrong row is a large error
        model.temporal_pos_embedding.copy_(torch.randn(model.temporal_pos_embedding.shape, generator=gen))
    x = torch.rand(2, 3, 8, 16, 16, generator=gen)
    case = {"cfg": cfg, "sd": _clone_sd(model), "x": x,
            "source": "_get_temporal_pos_embedding models/videomamba/videomamba.py:655-675"}
    with torch.no_grad():
        case["full_vis"], case["full_pool"] = model(x)
        for tag, cuts in (("c44", (0, 4, 8)), ("c33", (0, 3, 6))):
            state = model.allocate_state(2, dtype=torch.float32)
            for i in range(len(cuts) - 1):
                lo, hi = cuts[i], cuts[i + 1]
                model.pool_type = "cls+avg" if lo == 0 else "avg"
                vis, pool, state = model(x[:, :, lo:hi], ssm_state=state, temporal_pos_offset=lo)
                case[f"{tag}_vis{i}"], case[f"{tag}_pool{i}"] = vis, pool
            model.pool_type = "cls+avg"
            case[f"{tag}_state"] = [(c.clone(), s.clone()) for c, s in state]
        for n in (4, 6, 8):
            case[f"rows_{n}"] = model._get_temporal_pos_embedding(
                n - 2, offset=2, dtype=torch.float32, device=x.device).clone()
    torch.save(case, os.path.join(OUT, "model_fp32_temporal_interp.pt"))


def mixer_case(ms):
    """scripts/check_streaming_state.py:33-55 configuration (d_model 16, d_state 8, 12 = 5 | 7)."""
    torch.manual_seed(7)
    mixer = ms.Mamba(d_model=16, d_state=8, d_conv=4, expand=2, use_fast_path=False).eval()
    x = torch.randn(2, 12, 16)
    with torch.no_grad():
        full = mixer(x)
        o1, st1 = mixer(x[:, :5], return_state=True)
        o2, st2 = mixer(x[:, 5:], state=st1, return_state=True)
    case = {"sd": _clone_sd(mixer), "x": x, "full": full, "out1": o1, "out2": o2,
            "state1": tuple(t.clone() for t in st1), "state2": tuple(t.clone() for t in st2),
            "source": "Mamba.forward models/videomamba/mamba_simple.py:283-451"}
    # odd geometry the reference tests use (tests/test_videomamba_regressions.py:473-493)
    torch.manual_seed(8)
    small = ms.Mamba(d_model=8, d_state=4, d_conv=2, expand=2, use_fast_path=False,
                     layer_idx=0).eval()
    xs = torch.randn(3, 5, 8)
    with torch.no_grad():
        so, sst = small(xs, return_state=True)
        # single-token decode through the inference cache (mamba_simple.py:316-330, :453-497)
        from types import SimpleNamespace
        cache = SimpleNamespace(seqlen_offset=0, key_value_memory_dict={})
        p0 = small(xs[:, :3], inference_params=cache)
        cache.seqlen_offset = 3
        p1 = small(xs[:, 3:4], inference_params=cache)
        cache.seqlen_offset = 4
        p2 = small(xs[:, 4:5], inference_params=cache)
        cconv, cssm = cache.key_value_memory_dict[0]
    case["small"] = {"sd": _clone_sd(small), "x": xs, "out": so,
                     "state": tuple(t.clone() for t in sst),
                     "prefill": p0, "step1": p1, "step2": p2,
                     "cache_conv": cconv.clone(), "cache_ssm": cssm.clone()}
    torch.save(case, os.path.join(OUT, "mixer_fp32.pt"))


def scan_case(ms):
    """Direct call of the reference's in-tree scan (mamba_simple.py:30-106)."""
    gen = torch.Generator().manual_seed(5)
    Bsz, D, L, N = 2, 24, 37, 16
    u = torch.randn(Bsz, D, L, generator=gen)
    delta = torch.randn(Bsz, D, L, generator=gen)
    A = -torch.exp(torch.log(torch.arange(1, N + 1).float()).repeat(D, 1)
                   + 0.1 * torch.randn(D, N, generator=gen))
    Bm = torch.randn(Bsz, N, L, generator=gen)
    Cm = torch.randn(Bsz, N, L, generator=gen)
    Dp = torch.randn(D, generator=gen)
    z = torch.randn(Bsz, D, L, generator=gen)
    bias = torch.randn(D, generator=gen) - 2.0
    h0 = torch.randn(Bsz, D, N, generator=gen)
    out, last = ms._selective_scan_ref(u, delta, A, Bm, Cm, Dp, z, bias, True, h0, True)
    out0, last0 = ms._selective_scan_ref(u, delta, A, Bm, Cm, Dp, z, bias, True, None, True)
    torch.save({"u": u, "delta": delta, "A": A, "B": Bm, "C": Cm, "D": Dp, "z": z,
                "delta_bias": bias, "h0": h0, "out": out, "last": last, "out_no_h0": out0,
                "last_no_h0": last0,
                "source": "_selective_scan_ref models/videomamba/mamba_simple.py:30-106"},
               os.path.join(OUT, "scan_fp32.pt"))


def refiner_case(rb):
    """BiMambaRefinerBlock (models/refiner_backbone.py:92-135), 3-D and 4-D inputs."""
    torch.manual_seed(11)
    blk = rb.BiMambaRefinerBlock(dim=16, ssm_cfg={"use_fast_path": False}, layer_idx=0).eval()
    x3 = torch.randn(2, 10, 16)
    x4 = torch.randn(2, 3, 4, 16)
    with torch.no_grad():
        y3, s3 = blk(x3)
        y4, s4 = blk(x4)
    torch.save({"sd": _clone_sd(blk), "x3": x3, "y3": y3, "s3": tuple(t.clone() for t in s3),
                "x4": x4, "y4": y4, "s4": tuple(t.clone() for t in s4),
                "source": "BiMambaRefinerBlock.forward models/refiner_backbone.py:92-135"},
               os.path.join(OUT, "refiner_fp32.pt"))


def main():
    os.makedirs(OUT, exist_ok=True)
    ms, vm, _st, rb = reference_modules()
    model_case(vm, torch.float32, "model_fp32_rms_fused.pt", rms=True, fused=True)
    model_case(vm, torch.bfloat16, "model_bf16_rms_fused.pt", rms=True, fused=True)
    model_case(vm, torch.float32, "model_fp32_ln_unfused.pt", rms=False, fused=False)
    temporal_interp_case(vm)
    mixer_case(ms)
    scan_case(ms)
    refiner_case(rb)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
