"""TEST INFRASTRUCTURE -- runs in its OWN process (tests/test_gpu_leaf_crosscheck.py spawns it) so that importing
vllm cannot touch the state of the pytest process.

The reference's third-party leaves (``causal_conv1d_fn`` / ``causal_conv1d_update`` of the causal-conv1d wheel,
``selective_scan_fn`` / ``selective_state_update`` / ``rms_norm_fn`` / ``layer_norm_fn`` of the mamba_ssm wheel:
models/videomamba/mamba_simple.py:17-27, models/videomamba/videomamba.py:26-29)
are absent from this image, but vllm ships its own adaptations of exactly those upstream kernels
(vllm/model_executor/layers/mamba/ops/{mamba_ssm,causal_conv1d}.py, "Adapted from
https://github.com/state-spaces/mamba/blob/v2.2.4/..." and csrc/mamba/mamba_ssm/selective_scan_fwd.cu).  This worker
evaluates them on the inputs it is given; the test compares oracle/ and libvmb200 with the results.

    python tests/vllm_leaves_worker.py in.pt out.pt
"""
import sys
import traceback

import torch


def main(src, dst):
    cases = torch.load(src, weights_only=True)
    out = {}
    try:
        from vllm.model_executor.layers.mamba.ops.causal_conv1d import causal_conv1d_fn, causal_conv1d_update
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn, selective_state_update
    except Exception as e:      # noqa: BLE001
        torch.save({"unavailable": f"{type(e).__name__}: {e}"[:500]}, dst)
        return
    dev = "cuda"
    for name, c in cases.items():
        try:
            c = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in c.items()}
            if name.startswith("scan"):
                # u, delta, z (B, D, L); Bm, Cm (B, N, L); the kernel writes the output over z and the last
                # state over ssm_states (read as the initial state where has_initial_state is set)
                u, delta, z = c["u"].contiguous(), c["delta"].contiguous(), c["z"].contiguous().clone()
                has = torch.full((u.shape[0],), c.get("h0") is not None, dtype=torch.bool, device=dev)
                err = None
                for sdt in (torch.float32, u.dtype):        # the state cache: fp32 if the kernel takes it
                    state = (c["h0"].to(sdt).clone() if c.get("h0") is not None
                             else torch.zeros(u.shape[0], u.shape[1], c["A"].shape[1], device=dev, dtype=sdt))
                    zz = z.clone()
                    try:
                        y = selective_scan_fn(u, state, delta, c["A"].contiguous(), c["Bm"].contiguous(),
                                              c["Cm"].contiguous(), c["D"], zz, c["delta_bias"], delta_softplus=True,
                                              has_initial_state=has)
                        torch.cuda.synchronize()
                        err = None
                        break
                    except Exception as e:      # noqa: BLE001
                        err = e
                if err is not None:
                    raise err
                out[name] = {"y": y.cpu(), "last_state": state.cpu()}
            elif name.startswith("update"):
                state = c["state"].clone()
                y = torch.empty_like(c["x"])
                selective_state_update(state, c["x"], c["dt"], c["A"], c["Bm"], c["Cm"], c["D"], c["dt_bias"],
                                       z=c["z"], dt_softplus=True, out=y)
                out[name] = {"y": y.cpu(), "state": state.cpu()}
            elif name.startswith("stepmixer"):
                # single-token decode (mamba_simple.py:453-497): in_proj, causal_conv1d_update, x_proj, dt_proj,
                # selective_state_update, out_proj -- three tokens in a row so the rolled states are used again
                import torch.nn.functional as F
                Di, _, W = c["conv1d.weight"].shape
                N, R = c["A_log"].shape[1], c["dt_proj.weight"].shape[1]
                Bsz = c["hidden"].shape[0]
                cstate = torch.cat([torch.zeros_like(c["conv_state"][:1, :, 1:]), c["conv_state"][..., 1:]], 0).contiguous()
                idx = torch.arange(1, Bsz + 1, device=dev, dtype=torch.int32)
                state = c["ssm_state"].clone()
                A = -torch.exp(c["A_log"].float())
                outs = []
                for t in range(c["hidden"].shape[1]):
                    xz = F.linear(c["hidden"][:, t], c["in_proj.weight"])
                    x, z = xz[:, :Di].contiguous(), xz[:, Di:].contiguous()
                    x = causal_conv1d_update(x, cstate, c["conv1d.weight"].reshape(Di, W), c["conv1d.bias"],
                                             activation="silu", conv_state_indices=idx)
                    dt_low, Bm, Cm = torch.split(F.linear(x, c["x_proj.weight"]), [R, N, N], dim=-1)
                    dt = F.linear(dt_low, c["dt_proj.weight"])
                    y = torch.empty_like(x)
                    selective_state_update(state, x, dt, A, Bm.contiguous(), Cm.contiguous(), c["D"].float(),
                                           c["dt_proj.bias"].float(), z=z, dt_softplus=True, out=y)
                    outs.append(F.linear(y, c["out_proj.weight"]))
                out[name] = {"out": torch.stack(outs, 1).cpu(), "ssm_state": state.cpu(), "conv_tail": cstate[1:].cpu()}
            elif name.startswith("mixer"):
                # the reference's use_fast_path=False forward (models/videomamba/mamba_simple.py:333-339, :369,
                # :381-416, :423-446) with the upstream kernels in the places of its two wheels: projections by
                # torch, conv and scan by vllm's copies; optional streaming state in, next state out
                import torch.nn.functional as F
                h = c["hidden"]
                Bsz, L, _ = h.shape
                Di, _, W = c["conv1d.weight"].shape
                N, R = c["A_log"].shape[1], c["dt_proj.weight"].shape[1]
                xz = F.linear(h, c["in_proj.weight"])                                   # (B, L, 2 Di)
                x_tok, z = xz[..., :Di], xz[..., Di:]
                cstate = torch.zeros(Bsz + 1, Di, W - 1, device=dev, dtype=h.dtype)
                has_c = c.get("conv_state") is not None
                if has_c:
                    cstate[1:] = c["conv_state"][..., 1:]                                # the last W - 1 cached inputs
                cstate = cstate.permute(0, 2, 1).contiguous().permute(0, 2, 1)
                qsl = torch.arange(0, Bsz + 1, device=dev, dtype=torch.int32) * L
                idx = torch.arange(1, Bsz + 1, device=dev, dtype=torch.int32)
                xc = causal_conv1d_fn(x_tok.reshape(Bsz * L, Di).contiguous().t(), c["conv1d.weight"].reshape(Di, W),
                                      c["conv1d.bias"], cstate, qsl, cache_indices=idx,
                                      has_initial_state=torch.full((Bsz,), has_c, dtype=torch.bool, device=dev),
                                      activation="silu").t().reshape(Bsz, L, Di)
                x_dbl = F.linear(xc, c["x_proj.weight"])
                dt_low, Bm, Cm = torch.split(x_dbl, [R, N, N], dim=-1)
                delta = F.linear(dt_low, c["dt_proj.weight"])
                has_s = c.get("ssm_state") is not None
                state = c["ssm_state"].clone() if has_s else torch.zeros(Bsz, Di, N, device=dev, dtype=h.dtype)
                gate = z.transpose(1, 2).contiguous().clone()
                y = selective_scan_fn(xc.transpose(1, 2).contiguous(), state, delta.transpose(1, 2).contiguous(),
                                      -torch.exp(c["A_log"].float()), Bm.transpose(1, 2).contiguous(),
                                      Cm.transpose(1, 2).contiguous(), c["D"].float(), gate, c["dt_proj.bias"].float(),
                                      delta_softplus=True,
                                      has_initial_state=torch.full((Bsz,), has_s, dtype=torch.bool, device=dev))
                out[name] = {"out": F.linear(y.transpose(1, 2), c["out_proj.weight"]).cpu(), "ssm_state": state.cpu(),
                             "conv_tail": cstate[1:].contiguous().cpu()}
            elif name.startswith("addnorm"):
                # (a) the upstream Triton norm kernel family (vllm's copy of mamba_ssm/ops/triton/layernorm_gated.py,
                # ungated) on the fp32 sum x + residual; (b) vllm's own fused add + RMSNorm CUDA kernel
                from vllm.model_executor.layers.mamba.ops.layernorm_gated import _layer_norm_fwd
                acc = (c["x"] + c["residual"]).contiguous()
                rms, _, _ = _layer_norm_fwd(acc, c["weight"], None, c["eps"], is_rms_norm=True)
                ln, _, _ = _layer_norm_fwd(acc, c["weight"], c["bias"], c["eps"], is_rms_norm=False)
                res = {"rms": rms.cpu(), "ln": ln.cpu(), "sum": acc.cpu()}
                try:
                    from vllm import _custom_ops as vops
                    xi, ri = c["x"].clone(), c["residual"].clone()
                    vops.fused_add_rms_norm(xi, ri, c["weight"], c["eps"])
                    res["fused_rms"], res["fused_sum"] = xi.cpu(), ri.cpu()
                except Exception:       # noqa: BLE001
                    res["fused_error"] = traceback.format_exc()[-800:]
                out[name] = res
            elif name.startswith("convstep"):
                # x (B, D); conv_state (B, D, W - 1) rolled in place; the kernel overwrites x with the output
                # (vllm's cache line 0 is its "null block": sequences mapped to it are skipped -> lines 1..B)
                state = torch.cat([torch.zeros_like(c["conv_state"][:1]), c["conv_state"]], dim=0).contiguous()
                x = c["x"].clone()
                idx = torch.arange(1, x.shape[0] + 1, device=dev, dtype=torch.int32)
                y = causal_conv1d_update(x, state, c["weight"], c["bias"], activation="silu", conv_state_indices=idx)
                out[name] = {"y": y.cpu(), "conv_state": state[1:].cpu()}
            elif name.startswith("conv"):
                # x (B, D, L) -> the kernel's varlen layout: (D, B * L) with the channel stride 1
                x = c["x"]
                Bsz, Dm, L = x.shape
                xt = x.permute(0, 2, 1).reshape(Bsz * L, Dm).contiguous().t()
                W = c["weight"].shape[1]
                state = (c["init"].clone() if c.get("init") is not None
                         else torch.zeros(Bsz, Dm, W - 1, device=dev, dtype=x.dtype))
                # cache line 0 is vllm's "null block" (sequences mapped to it are skipped): lines 1..B; the kernel
                # wants the state's channel stride to be 1 as well
                state = torch.cat([torch.zeros_like(state[:1]), state], dim=0)
                state = state.permute(0, 2, 1).contiguous().permute(0, 2, 1)
                qsl = torch.arange(0, Bsz + 1, device=dev, dtype=torch.int32) * L
                idx = torch.arange(1, Bsz + 1, device=dev, dtype=torch.int32)
                has = torch.full((Bsz,), c.get("init") is not None, dtype=torch.bool, device=dev)
                y = causal_conv1d_fn(xt, c["weight"], c["bias"], state, qsl, cache_indices=idx,
                                     has_initial_state=has, activation="silu")
                out[name] = {"y": y.t().reshape(Bsz, L, Dm).permute(0, 2, 1).contiguous().cpu(),
                             "final": state[1:].contiguous().cpu()}
            torch.cuda.synchronize()
        except Exception:       # noqa: BLE001
            out[name] = {"error": traceback.format_exc()[-1500:]}
    torch.save(out, dst)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
