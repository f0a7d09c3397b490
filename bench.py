#!/usr/bin/env python
"""Benchmark of the VideoMamba mixer hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME]     # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]               # the CPU restatement, timed alone

Configurations (BASELINE.json `configs`; the default is the one the metric is quoted on):
    small16f   configs[1]  VideoMamba-S (embed 384, depth 24) 16 frames @224, bf16, 32 clips per GPU per step
    middle32f  configs[2]  VideoMamba-M (embed 576, depth 32) 32 frames @224, bf16, a global batch of 32 clips
                           sharded over the GPUs (strong scaling)
    stream64   configs[3]  VideoMamba-S streaming: 64-frame chunks with (conv_state, ssm_state) carry and
                           temporal_pos_offset, 32 resident streams per GPU (256 over 8) whose states live in a
                           slot pool (StreamPlacement + vmb_state_gather / vmb_state_scatter in the loop)
    long128f   configs[4]  VideoMamba-S 128 frames @224 (25 089 tokens per clip), 2 clips per GPU per step
    train16f   (not a BASELINE configuration) VideoMamba-S 16 frames @224, bf16, 32 clips per GPU:
                           one TRAINING step = forward + backward (libvmb200 backward kernels) + SGD update

A step = one forward over one batch of synthetic clips (stream64: one 64-frame chunk for one group of
streams).  For N > 1 the driver launches this file under torch.distributed.run, one rank per GPU; clips /
streams are sharded along batch, there is no collective on the data path (torch.distributed is used for the
start barrier and the max-over-ranks of the step time only).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) out of it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

MODELS = {
    "small": dict(embed_dim=384, depth=24, tag="S"),
    "tiny": dict(embed_dim=192, depth=24, tag="Ti"),
    "middle": dict(embed_dim=576, depth=32, tag="M"),
}
CONFIGS = {
    # name: model, frames per step, batch per GPU (None = global_batch / world), scaling, unit
    "small16f": dict(model="small", frames=16, batch=32, scaling="weak", unit="clips/s",
                     metric="clips/s VideoMamba-S 16f@224 bf16 forward"),
    "middle32f": dict(model="middle", frames=32, global_batch=32, scaling="strong", unit="clips/s",
                      metric="clips/s VideoMamba-M 32f@224 bf16 forward, global batch 32"),
    "stream64": dict(model="small", frames=64, batch=32, scaling="weak", unit="chunks/s",
                     metric="64-frame chunks/s VideoMamba-S streaming with state carry, 32 streams per GPU"),
    "long128f": dict(model="small", frames=128, batch=2, scaling="weak", unit="clips/s",
                     metric="clips/s VideoMamba-S 128f@224 bf16 forward (25 089 tokens)"),
    "train16f": dict(model="small", frames=16, batch=32, scaling="weak", unit="clips/s",
                     metric="clips/s VideoMamba-S 16f@224 bf16 training step (forward + backward + SGD)"),
}
STREAM_TOTAL_FRAMES = 256      # stream64: a stream is reset after 4 chunks (temporal table of 256 rows)

# Instruction model of the one-warp scan kernel (general A), from its SASS (profiles/r02_scan_sass_*.txt):
# the loop body of one 16-token tile is 789 warp instructions, 152 of them MUFU (geometric A: 790 / 88); a
# MUFU occupies the XU pipe for 8 clk and costs ~4.6 issue slots (profiles/r01_mufu_issue_microbench.txt).
# With SiLU(z) applied by the in_proj epilogue (the default; z_gate kernels) the tile loses 26 / 33 instructions, 8 of
# them MUFU (same SASS listing, tools/scan_tile_instr.py).
SCAN_TILE_INSTR = {"general": (789, 152), "geometric": (790, 88),
                   "general+gate": (763, 144), "geometric+gate": (757, 80)}
MUFU_ISSUE_SLOTS, MUFU_XU_CLK = 4.6, 8.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="small16f", choices=sorted(CONFIGS))
    ap.add_argument("--model", default=None, choices=sorted(MODELS), help="override the configuration's model")
    ap.add_argument("--batch", type=int, default=None, help="override: clips (streams) per GPU per step")
    ap.add_argument("--frames", type=int, default=None, help="override: frames per clip (per chunk)")
    ap.add_argument("--img", type=int, default=224)
    ap.add_argument("--weights", default="perturbed", choices=["perturbed", "init"],
                    help="perturbed: general A (trained-checkpoint-like); init: reference random init")
    ap.add_argument("--in-flight", type=int, default=None,
                    help="steps kept in flight on separate CUDA streams (1 = strictly serial steps)")
    ap.add_argument("--gate-in-scan", action="store_true",
                    help="A/B: SiLU(z) inside the scan instead of the in_proj epilogue (Mamba.gate_in_proj = False)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=4, help="clips in the CPU baseline sample")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.model = args.model or cfg["model"]
    args.frames = args.frames or cfg["frames"]
    args.streaming = args.config == "stream64"
    if args.in_flight is None:
        args.in_flight = 2 if args.streaming else 3
    return args


def per_gpu_batch(args, world):
    cfg = CONFIGS[args.config]
    if args.batch is not None:
        return args.batch
    if "global_batch" in cfg:
        if cfg["global_batch"] % world:
            raise SystemExit(f"global batch {cfg['global_batch']} does not divide over {world} GPUs")
        return cfg["global_batch"] // world
    return cfg["batch"]


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def model_cfg(args, num_frames=None):
    m = MODELS[args.model]
    return dict(img_size=args.img, patch_size=16, depth=m["depth"], embed_dim=m["embed_dim"], kernel_size=1,
                num_frames=num_frames or (STREAM_TOTAL_FRAMES if args.streaming else args.frames),
                norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True, residual_in_fp32=True,
                pool_type="cls+avg", add_pool_norm=True)


def build_model(args, dtype, device):
    import torch
    import video_mamba

    torch.manual_seed(0)
    cfg = model_cfg(args)
    model = video_mamba.PretrainVideoMamba(
        img_size=args.img, patch_size=16, depth=cfg["depth"], embed_dim=cfg["embed_dim"], channels=3,
        ssm_cfg={"use_fast_path": False}, norm_epsilon=1e-5, fused_add_norm=True, rms_norm=True,
        residual_in_fp32=True, bimamba=True, pool_type="cls+avg", kernel_size=1,
        num_frames=cfg["num_frames"]).eval()
    if args.weights == "perturbed":
        # break the S4D-real structure of A and the zero dt bias of the reference's model init, so
        # that the general-A kernels are what is measured
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for blk in model.layers:
                mx = blk.mixer
                mx.A_log.add_(0.1 * torch.randn(mx.A_log.shape, generator=g))
                dt = torch.exp(torch.rand(mx.d_inner, generator=g)
                               * (math.log(0.1) - math.log(0.001)) + math.log(0.001)).clamp(min=1e-4)
                mx.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
            model.temporal_pos_embedding.normal_(0, 0.02, generator=g)
    return model.to(dtype).to(device)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    _NAMES = {
        0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self._NAMES.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def _time_oracle(cfg, clips, frames, img, perturbed, dtype=None):
    import torch
    from oracle import videomamba_oracle as orc

    sd = orc.synthetic_state_dict(cfg, seed=0, perturbed=perturbed)
    model = orc.OracleVideoMamba(cfg, sd)
    x = torch.rand(clips, 3, frames, img, img, generator=torch.Generator().manual_seed(1000))
    t0 = time.perf_counter()
    with torch.no_grad():
        model.forward(x)
    return time.perf_counter() - t0


def cpu_baseline(args, clips: int, with_configs0: bool = True):
    """Times the CPU restatement (oracle/, kind "port") on the host cores: `clips` clips of the bench model /
    clip length in fp32 (the reference's own CPU-runnable precision) and BASELINE.json configs[0] exactly
    (VideoMamba-Tiny, 8 frames @224, fp32, batch 2).  The only place bench.py executes oracle/ besides the
    parity check.  /root/reference does not exist on the GPU box and the reference's kernels live in wheels
    that are absent everywhere, so the live reference cannot be timed: kind is "port"."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample: at most 16 frames per clip (the CPU port is linear in tokens; longer units are scaled
    # by their frame count), and fewer clips for the wider / deeper model
    frames = min(args.frames, 16)
    cfg = model_cfg(args, num_frames=frames)
    clips = max(1, round(clips * (384 * 24) / (cfg["embed_dim"] * cfg["depth"])))
    dt = _time_oracle(cfg, clips, frames, args.img, args.weights == "perturbed")
    scale = frames / args.frames          # units of `frames` frames -> units of args.frames frames
    out = {"value": clips / dt * scale, "unit": CONFIGS[args.config]["unit"], "cores": cores, "kind": "port",
           "sample": f"{clips} clip(s) of VideoMamba-{args.model} {frames}f@{args.img} fp32, "
                     f"full forward, torch CPU ops + per-token scan loop, {dt:.1f} s"
                     + (f"; scaled by {frames}/{args.frames} frames to the unit of this configuration"
                        if scale != 1 else ""),
           "clips": clips,
           "seconds": dt,
           "note": "oracle/ restatement of the reference's use_fast_path=False path; the reference itself "
                   "rejects CPU tensors (mamba_simple.py:304-308) and /root/reference is absent on this box"}
    if with_configs0:
        t = MODELS["tiny"]
        c0 = dict(model_cfg(args, num_frames=8), depth=t["depth"], embed_dim=t["embed_dim"])
        dt0 = _time_oracle(c0, 2, 8, 224, False)
        out["configs0"] = {"value": 2 / dt0, "unit": "clips/s", "seconds": round(dt0, 2),
                           "sample": "BASELINE.json configs[0] exactly: VideoMamba-Tiny (embed 192, depth 24) "
                                     "8 frames @224 fp32 batch 2, reference random init"}
    return out


def workload_name(args, batch):
    m = MODELS[args.model]
    base = f"VideoMamba-{args.model.capitalize()} (embed {m['embed_dim']}, depth {m['depth']})"
    if args.streaming:
        return (f"{base} streaming, {args.frames}-frame chunks @{args.img} with (conv_state, ssm_state) carry, "
                f"{batch} resident streams/GPU")
    return f"{base} {args.frames}f@{args.img} forward, batch {batch}/GPU"


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement timed alone
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    if args.config == "train16f":
        # training step of the CPU restatement: forward + torch-autograd backward, a bounded sample per step
        import torch
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        frames = 2

        def sample():
            t = _oracle_train_step(model_cfg(args, num_frames=frames), 1, frames, args.img)
            return {"value": 1 / t * frames / args.frames, "seconds": t, "cores": cores,
                    "sample": f"1 clip of VideoMamba-{args.model} {frames}f@{args.img} fp32, forward + autograd "
                              f"backward of oracle/, {t:.1f} s; scaled by {frames}/{args.frames} frames"}
        for _ in range(min(args.warmup, 1)):
            sample()
        times = [sample() for _ in range(args.steps)]
    else:
        for _ in range(min(args.warmup, 1)):
            cpu_baseline(args, 1, with_configs0=False)
        # a bounded sample per step: about 48 clips over the whole run (4 per step up to 12 steps, 2 at 20), so that
        # --steps 20 still ends within a few minutes on a 16-core host
        clips = max(1, min(args.cpu_clips, 48 // max(1, args.steps)))
        times = []
        for _ in range(args.steps):
            times.append(cpu_baseline(args, clips, with_configs0=False))
    total = sum(t["seconds"] for t in times)
    value = sum(t["value"] * t["seconds"] for t in times) / total
    last = times[-1]
    batch = per_gpu_batch(args, max(1, args.gpus))
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, batch), "weights": args.weights,
                   "note": "reference rejects CPU tensors and its kernels live in absent wheels; "
                           "this is the CPU restatement (oracle/) of its use_fast_path=False path, "
                           f"each step a sample of {last['clips'] if 'clips' in last else args.cpu_clips} clip(s)"},
        "cpu_baseline": {"value": value, "unit": cfg["unit"], "cores": last["cores"], "kind": "port",
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class ClipWorkload:
    """Independent clips: step i is model(x_i) on one of `nbuf` rotating input buffers."""

    def __init__(self, args, model, batch, dev, dtype, rank, nbuf):
        import torch
        self.model, self.batch = model, batch
        gen = torch.Generator().manual_seed(1000 + rank)
        shape = (batch, 3, args.frames, args.img, args.img)
        self.x_host = [torch.rand(shape, generator=gen).to(dtype).pin_memory() for _ in range(2)]
        self.x_dev = [self.x_host[i % 2].to(dev, non_blocking=True) for i in range(max(2, nbuf))]
        self.units = batch
        self.tokens = batch * (1 + args.frames * (args.img // 16) ** 2)

    def lane_of(self, i, lanes):
        return i % lanes

    def host_input(self, i):
        return self.x_host[i % 2]

    def run(self, x, i):
        return self.model(x)

    def resident(self, i):
        return self.x_dev[i % len(self.x_dev)]


class StreamWorkload:
    """Streaming with state carry: the GPU's resident streams are split into `groups` (one per lane in
    flight); step i advances group i % groups by one chunk.  Stream states live in a slot pool; every step
    gathers its streams' rows, runs model(x, ssm_state=..., temporal_pos_offset=...) and scatters the next
    state back (StreamPlacement + vmb_state_gather / vmb_state_scatter)."""

    def __init__(self, args, model, batch, dev, dtype, rank, world, groups):
        import torch
        from videomamba_b200.replica import StreamPlacement

        self.model, self.dev, self.dtype = model, dev, dtype
        self.chunk = args.frames
        self.chunks_per_stream = STREAM_TOTAL_FRAMES // args.frames
        self.groups = max(1, min(groups, batch))
        # sticky placement of the world * batch global stream ids (the same deterministic table on every
        # rank): a stream lives on one rank for its lifetime and owns one row of that rank's state pool
        place = StreamPlacement(world, batch)
        table = {sid: place.place(sid) for sid in range(world * batch)}
        ids = [sid for sid, (r, _slot) in table.items() if r == rank]
        per = (len(ids) + self.groups - 1) // self.groups
        self.group_ids = [ids[g * per:(g + 1) * per] for g in range(self.groups)]
        self.group_ids = [g for g in self.group_ids if g]
        self.groups = len(self.group_ids)
        self.slots = [torch.tensor(place.local_slots(rank, g), dtype=torch.int32, device=dev)
                      for g in self.group_ids]
        mixers = [blk.mixer for blk in model.layers]
        self.pool_conv = [torch.zeros(batch, m.d_inner, m.d_conv, dtype=dtype, device=dev) for m in mixers]
        self.pool_ssm = [torch.zeros(batch, m.d_inner, m.d_state, dtype=torch.float32, device=dev) for m in mixers]
        self.chunk_idx = [0] * self.groups
        gen = torch.Generator().manual_seed(1000 + rank)
        gb = len(self.group_ids[0])
        shape = (gb, 3, args.frames, args.img, args.img)
        self.x_host = [torch.rand(shape, generator=gen).to(dtype).pin_memory() for _ in range(2)]
        self.x_dev = [self.x_host[i % 2].to(dev, non_blocking=True) for i in range(max(2, self.groups + 1))]
        self.units = gb
        self.tokens = gb * args.frames * (args.img // 16) ** 2

    def lane_of(self, i, lanes):
        return (i % self.groups) % lanes

    def host_input(self, i):
        return self.x_host[i % 2]

    def resident(self, i):
        return self.x_dev[i % len(self.x_dev)]

    def run(self, x, i):
        from videomamba_b200 import ops
        g = i % self.groups
        k = self.chunk_idx[g] % self.chunks_per_stream
        self.chunk_idx[g] += 1
        n = len(self.group_ids[g])
        x = x[:n]
        if k == 0:      # a new stream starts: zero state, CLS token on the first chunk
            state = self.model.allocate_state(n, dtype=self.dtype, device=self.dev)
            self.model.pool_type = "cls+avg"
        else:
            idx = self.slots[g]
            state = [(ops.state_gather(c, idx), ops.state_gather(s, idx))
                     for c, s in zip(self.pool_conv, self.pool_ssm)]
            self.model.pool_type = "avg"
        vis, pool, nxt = self.model(x, ssm_state=state, temporal_pos_offset=k * self.chunk)
        idx = self.slots[g]
        for (c, s), pc, ps in zip(nxt, self.pool_conv, self.pool_ssm):
            ops.state_scatter(pc, idx, c)
            ops.state_scatter(ps, idx, s)
        return vis, pool


def parity_check(args, model, wl, dev, dtype):
    """One unit of the timed workload against the CPU oracle (outside every timed region): clip 0 of the
    first host batch through model() and through oracle/ with the same weights.  Streaming: chunk 0 and
    chunk 1 of one stream with the state carried.  Returns the worst relative error (max|a-b| / max|b|)
    over x_vis and x_pool (and the carried state)."""
    import torch
    from oracle import videomamba_oracle as orc

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = model_cfg(args)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    oracle = orc.OracleVideoMamba(cfg, sd)
    x = wl.x_host[0][:1]
    errs = {}
    with torch.no_grad():
        if not args.streaming:
            model.pool_type = "cls+avg"
            vis, pool = model(x.to(dev))
            w_vis, w_pool = oracle.forward(x)
            errs = {"x_vis": orc.rel_err(vis, w_vis), "x_pool": orc.rel_err(pool, w_pool)}
        else:
            frames = min(args.frames, 16)       # two chunks of <= 16 frames keep the CPU side to seconds
            di = 2 * cfg["embed_dim"]
            zero = [(torch.zeros(1, di, 4, dtype=dtype), torch.zeros(1, di, 16, dtype=dtype))
                    for _ in range(cfg["depth"])]
            st = model.allocate_state(1, dtype=dtype, device=dev)
            model.pool_type = oracle.pool_type = "cls+avg"
            v0, p0, st = model(x[:, :, :frames].to(dev), ssm_state=st, temporal_pos_offset=0)
            wv0, wp0, wst = oracle.forward(x[:, :, :frames], ssm_state=zero, temporal_pos_offset=0)
            model.pool_type = oracle.pool_type = "avg"
            v1, p1, st = model(x[:, :, frames:2 * frames].to(dev), ssm_state=st, temporal_pos_offset=frames)
            wv1, wp1, wst = oracle.forward(x[:, :, frames:2 * frames], ssm_state=wst, temporal_pos_offset=frames)
            errs = {"chunk0_x_vis": orc.rel_err(v0, wv0), "chunk0_x_pool": orc.rel_err(p0, wp0),
                    "chunk1_x_vis": orc.rel_err(v1, wv1), "chunk1_x_pool": orc.rel_err(p1, wp1),
                    "next_state": max(max(orc.rel_err(c, wc), orc.rel_err(s, ws))
                                      for (c, s), (wc, ws) in zip(st, wst))}
            model.pool_type = "cls+avg"
    return max(errs.values()), errs


def scan_op_alone(args, B, L, Di, R, N, dev, iters=10):
    """The fused scan as an operator at the bench shape, timed alone (CUDA events on its stream) for both
    decay evaluators: general A and geometric A (S4D-real structure kept exact)."""
    import torch
    from videomamba_b200 import ops

    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    Xp = ops.xdbl_pitch(R, N)
    u = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
    z = torch.randn(B, L, Di, generator=g, device=dev).to(bf)
    xdbl = torch.randn(B, L, Xp, generator=g, device=dev).to(bf)
    w_dt = (torch.randn(Di, R, generator=g, device=dev) * R ** -0.5).to(bf)
    Dp = torch.ones(Di, device=dev)
    dt = torch.exp(torch.rand(Di, generator=g, device=dev) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
    bias = dt + torch.log(-torch.expm1(-dt))
    out = {}
    for name in ("general", "geometric"):
        if name == "geometric":
            A = -torch.arange(1, N + 1, device=dev).float().repeat(Di, 1)
        else:
            A = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(Di, 1)
                           + 0.1 * torch.randn(Di, N, generator=g, device=dev))
        A2 = (A * ops.LOG2E).contiguous()
        fn = lambda: ops.selective_scan_fused_tokens(u, z, xdbl, w_dt, A2, R, N, Dp, bias,
                                                     a_geometric=name == "geometric",
                                                     z_gate=not args.gate_in_scan)    # as the step runs it
        out[name] = {}
        for label, nstreams in (("alone", 1), ("three_in_flight", 3)):
            streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
            main = torch.cuda.current_stream(dev)

            def go(n):
                for s in streams:
                    s.wait_stream(main)
                for i in range(n):
                    with torch.cuda.stream(streams[i % nstreams]):
                        fn()
                for s in streams:
                    main.wait_stream(s)

            go(3 * nstreams)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            go(iters * nstreams)
            e1.record()
            torch.cuda.synchronize(dev)
            out[name][label] = e0.elapsed_time(e1) / (iters * nstreams)     # ms per launch
    return out


def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               os.environ.get("MASTER_PORT", "29541"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    from videomamba_b200 import _lib
    from videomamba_b200.replica import init_replica_group
    grp = init_replica_group(device=dev)      # NCCL: barrier + max-over-ranks of the timing only
    rank = grp.rank
    lib = _lib.load()
    cfg = CONFIGS[args.config]

    dtype = torch.bfloat16
    if args.gate_in_scan:
        from videomamba_b200.mixer import Mamba
        Mamba.gate_in_proj = False
    model = build_model(args, dtype, dev)
    B = per_gpu_batch(args, world)
    nlanes = max(1, args.in_flight)
    if args.streaming:
        wl = StreamWorkload(args, model, B, dev, dtype, rank, world, nlanes)
        nlanes = min(nlanes, wl.groups)
    else:
        wl = ClipWorkload(args, model, B, dev, dtype, rank, nlanes + 1)
    torch.cuda.synchronize()

    def barrier():
        grp.barrier(dev)

    def max_over_ranks(v: float) -> float:
        return grp.max_over_ranks(v, dev)

    def fwd(x, i):
        with torch.no_grad():
            return wl.run(x, i)

    # ---- device-resident throughput ("value") ----------------------------------------------------
    step_no = [0]

    def next_i():
        step_no[0] += 1
        return step_no[0] - 1

    for _ in range(max(args.warmup, 3)):
        i = next_i()
        fwd(wl.resident(i), i)
    if args.streaming:                          # finish the streams' current pass so every loop starts aligned
        while step_no[0] % (wl.groups * wl.chunks_per_stream):
            i = next_i()
            fwd(wl.resident(i), i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = (C.c_double * 8)()
    cnt = (C.c_int64 * 8)()

    # (1) strictly serial steps with the library's per-stage CUDA events on: the per-kernel times the
    #     roofline object is computed from (each kernel has the GPU to itself here).  Input buffers rotate.
    lib.vmb_prof_enable(1)
    lib.vmb_prof_read(ms, cnt, 1)               # clear anything recorded before
    for k in range(8):
        ms[k] = 0.0
        cnt[k] = 0
    barrier()
    e0.record()
    for _ in range(args.steps):
        i = next_i()
        out = fwd(wl.resident(i), i)
    e1.record()
    barrier()
    serial_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    lib.vmb_prof_read(ms, cnt, 1)
    lib.vmb_prof_enable(0)
    stages = {k: {"ms_per_step": ms[i] / args.steps, "launches_per_step": cnt[i] / args.steps}
              for i, k in enumerate(_lib.PROF_KINDS) if cnt[i]}

    # (2) the headline: the same K steps, `in_flight` of them kept in flight on separate streams.
    #     Steps are independent batches (streaming: independent groups of streams; a group's chunks stay in
    #     order on its lane); the scan is bound by instruction issue and runs faster with two launches
    #     sharing the SMs, and the HBM-bound kernels of one step overlap the scan of another.  Every step is
    #     a full forward; nothing is skipped or cached.
    main = torch.cuda.current_stream()
    lanes = [torch.cuda.Stream() for _ in range(nlanes)]

    def run_steps(n):
        start = torch.cuda.Event()
        start.record(main)
        last = None
        seen = set()
        for _ in range(n):
            i = next_i()
            li = wl.lane_of(i, len(lanes))
            s = lanes[li]
            if li not in seen:
                s.wait_event(start)
                seen.add(li)
            with torch.cuda.stream(s):
                last = fwd(wl.resident(i), i)
        for s in lanes:
            main.wait_stream(s)
        return last

    run_steps(2 * len(lanes))
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    barrier()
    launches0 = lib.vmb_launch_count()
    e0.record()
    out = run_steps(args.steps)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    launches = (lib.vmb_launch_count() - launches0) // args.steps
    ms_per_step = ms_total / args.steps
    value = world * wl.units * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e") ----------------------------
    # A double-buffered serving loop: step i+1's clips are copied host->device on a copy stream
    # while step i computes, and step i's features go device->host on a second copy stream.  Every
    # step's input copy and result copy are inside the timed region.
    vis0, pool0 = out[0], out[1]
    vis_host = torch.empty(vis0.shape, dtype=vis0.dtype).pin_memory()
    pool_host = torch.empty(pool0.shape, dtype=pool0.dtype).pin_memory()
    nbuf = len(lanes) + 1
    x_in = [torch.empty_like(wl.resident(0)) for _ in range(nbuf)]
    comp = torch.cuda.current_stream()
    h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()

    def e2e_loop(n):
        ready = [None] * nbuf
        done = [None] * nbuf
        ids = [next_i() for _ in range(n)]

        def stage_in(j):                        # host -> device copy of step j's clips
            k = j % nbuf
            with torch.cuda.stream(h2d):
                if done[k] is not None:
                    h2d.wait_event(done[k])     # the forward that last read this buffer is finished
                x_in[k].copy_(wl.host_input(ids[j]), non_blocking=True)
                ready[k] = torch.cuda.Event()
                ready[k].record(h2d)

        for j in range(min(len(lanes), n)):
            stage_in(j)
        last_out = None
        for j in range(n):
            k = j % nbuf
            if j + len(lanes) < n:
                stage_in(j + len(lanes))
            cs = lanes[wl.lane_of(ids[j], len(lanes))]
            cs.wait_event(ready[k])
            with torch.cuda.stream(cs):
                res = fwd(x_in[k], ids[j])
                vis, pool = res[0], res[1]
            done[k] = torch.cuda.Event()
            done[k].record(cs)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done[k])
                vis.record_stream(d2h)
                pool.record_stream(d2h)
                vis_host[:vis.shape[0], :vis.shape[1]].copy_(vis, non_blocking=True)
                pool_host[:pool.shape[0]].copy_(pool, non_blocking=True)
                last_out = torch.cuda.Event()
                last_out.record(d2h)
        comp.wait_event(last_out)

    e2e_loop(2 * len(lanes))
    if args.streaming:
        while step_no[0] % (wl.groups * wl.chunks_per_stream):
            i = next_i()
            fwd(wl.resident(i), i)
    barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * wl.units * args.steps / (e2e_ms * 1e-3)
    h2d_bytes = x_in[0].numel() * x_in[0].element_size()
    d2h_bytes = vis_host.numel() * vis_host.element_size() + pool_host.numel() * pool_host.element_size()

    # ---- roofline of the dominant kernel (the selective scan) -------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    mixer = model.layers[0].mixer
    Di, N, R = mixer.d_inner, mixer.d_state, mixer.dt_rank
    L = wl.tokens // wl.units
    tokens = wl.tokens
    es = 2
    w = mixer._kernel_weights()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    units = wl.units * (Di // 16)
    two_warp = 2 * units < 19 * sms             # scan_fast(): two warps per unit below 9.5 units per SM
    evaluator = "geometric" if w.a_geometric else "general"
    bytes_per_token = 3 * Di * es + w.Xp * es   # reads u (conv output), z, the x_dbl row; writes y
    kernel = (("scan2w_kernel (helper + consumer warps" if two_warp else "scan1w_kernel (one warp per unit")
              + f"; {evaluator}-A decay evaluator; TMA-staged tiles; dt_proj + softplus + S6 scan + D skip + "
                "SiLU gate, fused)")
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        wlk = tr["workload"]
        if (wlk["B"], wlk["L"], wlk["Di"]) == (wl.units, L, Di):
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]     # per launch, from ncu
    except Exception:
        pass
    sm_hz = 1e6 * float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))

    def issue_bound(ms_per_launch, which):
        """Instruction-issue and XU-pipe floors of one launch: warp-tokens x slots / (4 schedulers x SMs) / clock."""
        instr, mufu = SCAN_TILE_INSTR[which if args.gate_in_scan else which + "+gate"]
        slots = (instr - mufu) / 16.0 + mufu / 16.0 * MUFU_ISSUE_SLOTS
        xu = mufu / 16.0 * MUFU_XU_CLK
        warp_tokens = tokens * (Di // 16)
        us = lambda clk: warp_tokens * clk / (4 * sms) / sm_hz * 1e6
        return {"us_per_launch": us(slots), "frac": us(slots) / (ms_per_launch * 1e3),
                "issue_slots_per_warp_token": slots, "xu_us_per_launch": us(xu),
                "xu_clk_per_warp_token": xu, "sm_mhz": sm_hz / 1e6,
                "model": "(instr - mufu) + 4.6 * mufu issue slots and 8 * mufu XU clk per 16-channel token; "
                         "592 schedulers"}

    roofline = None
    if "scan" in stages and stages["scan"]["launches_per_step"]:
        # kernel time from the serial pass (the kernel alone on the GPU); share of the serial step.
        # With the sequence split a step has more than one scan launch per layer: per layer = sum.
        layers = len(model.layers)
        per_launch_ms = stages["scan"]["ms_per_step"] / layers
        achieved = tokens * bytes_per_token / (per_launch_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": tokens * bytes_per_token,
                    "ms_per_launch": per_launch_ms, "peak_source": peak_src,
                    "share_of_step": stages["scan"]["ms_per_step"] / serial_ms,
                    "issue_bound": issue_bound(per_launch_ms, evaluator) if not two_warp else None,
                    "note": "the kernel is bound by instruction issue / the XU pipe (16 exponentials per "
                            "token-channel in fp32 state), not by HBM: issue_bound is that floor (DESIGN.md 3.2)"}
    roofline_geometric = None
    if args.config == "small16f" and rank == 0:
        measured = scan_op_alone(args, wl.units, L, Di, R, N, dev)
        roofline_geometric = {}
        for name, times in measured.items():
            entry = {}
            for label, t_ms in times.items():
                gbs = tokens * bytes_per_token / (t_ms * 1e-3) / 1e9
                entry[label] = {"achieved": gbs, "unit": "GB/s", "peak": hbm_peak, "frac": gbs / hbm_peak,
                                "ms_per_launch": t_ms, "issue_bound_frac": issue_bound(t_ms, name)["frac"]}
            entry["issue_bound_us_per_launch"] = issue_bound(times["alone"], name)["us_per_launch"]
            roofline_geometric[name] = entry
        roofline_geometric["note"] = (
            "the fused scan as an operator at the bench shape, both decay evaluators -- general A (what the step "
            "above runs with perturbed weights) and geometric A (exact S4D-real A, fp32 A_log: 2 exponentials + "
            "multiplies per channel) -- timed alone and as the effective time per launch with three launches in "
            "flight on three streams (how `value` runs it: more resident warps per scheduler)")
    # the other HBM-bound kernels of the path against the same peak (algorithmic bytes per token:
    # conv reads x and writes xc; add+norm reads hidden (bf16) + residual (fp32), writes both)
    hbm_kernels = {}
    Dm = mixer.d_model
    for k, bpt in (("conv", 2 * Di * es), ("add_norm", Dm * es + 4 * Dm + 4 * Dm + Dm * es)):
        if k in stages and stages[k]["launches_per_step"]:
            per = stages[k]["ms_per_step"] / stages[k]["launches_per_step"]
            gbs = tokens * bpt / (per * 1e-3) / 1e9
            hbm_kernels[k] = {"achieved": gbs, "unit": "GB/s", "frac": gbs / hbm_peak, "ms_per_launch": per}
    # projections against the tensor roofline (reported beside, not the dominant-kernel object)
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flops = {"in_proj": 2 * Dm * 2 * Di, "x_proj": 2 * Di * (R + 2 * N), "out_proj": 2 * Di * Dm}
    tensor = {}
    for k, f in flops.items():
        if k in stages and stages[k]["launches_per_step"]:
            per = stages[k]["ms_per_step"] / stages[k]["launches_per_step"]
            tf = tokens * f / (per * 1e-3) / 1e12
            tensor[k] = {"tflops": tf, "frac_of_sustained_peak": tf / tc_peak}

    if rank != 0:
        grp.close()
        return 0

    parity = None
    if not args.no_parity_check:
        worst, errs = parity_check(args, model, wl, dev, dtype)
        parity = {"rel_err": worst, "bar": 2e-2, "detail": errs,
                  "what": "one unit of the timed workload (same weights, same clip) against oracle/ on the CPU, "
                          "max|a-b| / max|b|, outside the timed regions"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, args.cpu_clips)
        cpu.pop("seconds", None)

    line = {
        "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args, B), "config": args.config, "tokens_per_unit": L,
                   "units_per_step_per_gpu": wl.units,
                   "weights": ("random init, A_log/dt_bias/temporal embedding perturbed (general-A "
                               "kernels)" if args.weights == "perturbed" else "reference random init"),
                   "parallelism": f"batch-sharded replicas x{world}, no collective",
                   "gate": "SiLU(z) inside the scan" if args.gate_in_scan else "SiLU(z) in the in_proj epilogue",
                   "steps_in_flight": len(lanes),
                   "serial_ms_per_step": serial_ms,
                   "serial_value": world * wl.units / (serial_ms * 1e-3),
                   "inputs": "2 host clips batches / >= 2 device buffers, rotated every step",
                   "l2": "per-step working set (>= 150 MB of activations per layer) exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": cfg["unit"], "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms / args.steps,
                "pipeline": "H2D copy stream, compute lanes, D2H copy stream"},
        "gpu_launches": int(launches),
        "parity_check": parity,
        "roofline": roofline,
        "roofline_geometric": roofline_geometric,
        "cpu_baseline": cpu,
        "stages": stages,
        "hbm_kernels": hbm_kernels,
        "tensor": tensor,
    }
    print(json.dumps(line), flush=True)
    grp.close()
    if parity is not None and not (parity["rel_err"] <= parity["bar"]):
        sys.stderr.write(f"bench.py: parity check failed: {parity}\n")
        return 3
    return 0


# ------------------------------------------------------------------------------------------------
# training step: forward + backward + SGD update
# ------------------------------------------------------------------------------------------------
def _oracle_train_step(cfg, clips, frames, img):
    """One forward + backward of the CPU restatement through torch autograd (seconds)."""
    import torch
    from oracle import videomamba_oracle as orc

    sd = {k: v.clone().requires_grad_(True) for k, v in orc.synthetic_state_dict(cfg, seed=0, perturbed=True).items()}
    model = orc.OracleVideoMamba(cfg, sd)
    x = torch.rand(clips, 3, frames, img, img, generator=torch.Generator().manual_seed(1000))
    t0 = time.perf_counter()
    vis, pool = model.forward(x)
    (pool.float().square().mean() + vis.float().mean()).backward()
    return time.perf_counter() - t0


def _gradient_parity(dev):
    """Gradients of a small fp32 model (depth 2, embed 64, 2 frames @32) against torch autograd through the
    CPU oracle: max over parameters of max|a-b| / max|b|."""
    import torch
    import video_mamba
    from oracle import videomamba_oracle as orc

    cfg = dict(img_size=32, patch_size=16, depth=2, embed_dim=64, kernel_size=1, num_frames=2,
               norm_epsilon=1e-5, rms_norm=True, fused_add_norm=True, residual_in_fp32=True,
               pool_type="cls+avg", add_pool_norm=True)
    sd = orc.synthetic_state_dict(cfg, seed=4, dtype=torch.float32, perturbed=True)
    x = torch.rand(2, 3, 2, 32, 32, generator=torch.Generator().manual_seed(9))
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    vis_o, pool_o = orc.OracleVideoMamba(cfg, p).forward(x)
    gv = torch.randn(vis_o.shape, generator=torch.Generator().manual_seed(10))
    gp = torch.randn(pool_o.shape, generator=torch.Generator().manual_seed(11))
    ((vis_o * gv).sum() + (pool_o * gp).sum()).backward()
    m = video_mamba.PretrainVideoMamba(img_size=32, patch_size=16, depth=2, embed_dim=64, channels=3,
                                       ssm_cfg={"use_fast_path": False}, num_frames=2)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    vis, pool = m(x.to(dev))
    ((vis * gv.to(dev)).sum() + (pool * gp.to(dev)).sum()).backward()
    errs = {}
    for name, prm in m.named_parameters():
        if p[name].grad is not None:
            errs[name] = orc.rel_err(prm.grad, p[name].grad)
    worst = max(errs, key=errs.get)
    return errs[worst], {"worst_parameter": worst, "parameters_checked": len(errs),
                         "forward": max(orc.rel_err(vis, vis_o), orc.rel_err(pool, pool_o))}


def run_train(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               os.environ.get("MASTER_PORT", "29541"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from videomamba_b200 import _lib
    from videomamba_b200 import autograd as ag
    from videomamba_b200 import ops
    from videomamba_b200.replica import init_replica_group
    grp = init_replica_group(device=dev)      # barrier + max-over-ranks of the timing only: replicas do not talk
    rank = grp.rank
    lib = _lib.load()
    cfg = CONFIGS[args.config]
    dtype = torch.bfloat16
    model = build_model(args, dtype, dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-4, foreach=True)
    B = per_gpu_batch(args, world)
    gen = torch.Generator().manual_seed(1000 + rank)
    host = [torch.rand(B, 3, args.frames, args.img, args.img, generator=gen).to(dtype).pin_memory() for _ in range(2)]
    resident = [h.to(dev) for h in host]
    x_in = torch.empty_like(resident[0])
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def step(x):
        opt.zero_grad(set_to_none=True)
        vis, pool = model(x)
        loss = pool.float().square().mean() + vis.float().mean()
        loss.backward()
        opt.step()
        return loss

    for i in range(max(args.warmup, 3)):
        step(resident[i % 2])
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    grp.barrier(dev)
    launches0 = lib.vmb_launch_count()
    e[0].record()
    for i in range(args.steps):
        step(resident[i % 2])
    e[1].record()
    grp.barrier(dev)
    ms_total = grp.max_over_ranks(e[0].elapsed_time(e[1]), dev)
    clocks = sampler.stop()
    launches = (lib.vmb_launch_count() - launches0) // args.steps
    # forward / backward split of one step (serial, same process)
    model.zero_grad(set_to_none=True)
    e[0].record()
    vis, pool = model(resident[0])
    loss = pool.float().square().mean() + vis.float().mean()
    e[1].record()
    loss.backward()
    e[2].record()
    torch.cuda.synchronize()
    fwd_ms, bwd_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    del vis, pool, loss
    # end to end: the clip batch comes from pinned host memory, the loss goes back to the host, every step
    grp.barrier(dev)
    e[0].record()
    for i in range(args.steps):
        x_in.copy_(host[i % 2], non_blocking=True)
        loss_host.copy_(step(x_in).detach(), non_blocking=True)
    e[1].record()
    grp.barrier(dev)
    e2e_ms = grp.max_over_ranks(e[0].elapsed_time(e[1]), dev)
    value = world * B * args.steps / (ms_total * 1e-3)

    # roofline of the dominant kernel: the selective-scan backward, timed alone at the step's shape
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    mixer = model.layers[0].mixer
    Di, N, R = mixer.d_inner, mixer.d_state, mixer.dt_rank
    L = 1 + args.frames * (args.img // 16) ** 2
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *sh: torch.randn(*sh, device=dev, generator=g).to(dtype)
    u, z, dout, bc = rn(B, L, Di), rn(B, L, Di), rn(B, L, Di), rn(B, L, 64)
    w_dt = (torch.randn(Di, R, device=dev, generator=g) * R ** -0.5).to(dtype)
    delta = ops.linear_raw(bc[..., :R], w_dt) - 3          # delta_raw as FusedScanFn.backward recomputes it
    A2 = -(torch.rand(Di, N, device=dev, generator=g) * 16 + 0.5) * 1.4427
    ones, zeros = torch.ones(Di, device=dev), torch.zeros(Di, device=dev)
    nck_bytes = lib.vmb_scan_bwd_ckpt_bytes(B, L, Di)
    saved = torch.empty(nck_bytes, dtype=torch.uint8, device=dev)

    def timed(fn, n=5):
        fn()
        torch.cuda.synchronize()
        e[0].record()
        for _ in range(n):
            fn()
        e[1].record()
        torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]) / n

    # (a) as an operator: its own forward pass writes the state records; (b) as the training step runs it: the
    # fused forward has written them (filled here by one recomputing call: same layout, same size)
    scan_ms = timed(lambda: ag._scan_bwd(u, delta, A2, bc, R, R + N, N, ones, z, zeros, True, None, dout, None, False))
    ops.selective_scan_fused_tokens_raw(u, z, bc, w_dt, A2, R, N, ones, zeros - 3, None, False, bwd_ckpt=saved)
    scan_saved_ms = timed(lambda: ag._scan_bwd(u, delta, A2, bc, R, R + N, N, ones, z, zeros, True, None, dout, None,
                                               False, None, saved))
    # the bf16 kernels against the true-fp32 kernels on the same (bf16-representable) inputs
    f32 = lambda t: t.float()
    got = ag._scan_bwd(u[:2], delta[:2], A2, bc[:2], R, R + N, N, ones, z[:2], zeros, True, None, dout[:2], None, False)
    want = ag._scan_bwd(f32(u[:2]), f32(delta[:2]), A2, f32(bc[:2]), R, R + N, N, ones, f32(z[:2]), zeros, True, None,
                        f32(dout[:2]), None, False)
    rel = lambda a_, b_: float((a_.float() - b_.float()).abs().max() / b_.float().abs().max().clamp_min(1e-30))
    scan_parity = {n_: rel(a_, b_) for n_, a_, b_ in zip(("du", "ddelta", "dz", "dbc", "dA", "dD", "dbias"), got, want)
                   if a_ is not None}
    del got, want, saved
    bytes_per_token = (7 * Di + 4 * N) * 2     # reads u, delta, z, dout + B/C rows; writes du, ddelta, dz + dB/dC rows
    achieved = B * L * bytes_per_token / (scan_saved_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "scan_bwd_fast_kernel + scan_bwd_bc_kernel (selective-scan backward, bf16: one "
                "warp per (batch, 16 channels), 2 channels x 4 states per lane in packed fp32, every reduction an HMMA, "
                "4-token sub-chunks recomputed from the state records the fused forward wrote; TMA loads / stores)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                "algorithmic_bytes_per_launch": B * L * bytes_per_token, "ms_per_launch": scan_saved_ms,
                "ms_per_launch_recomputing_the_records": scan_ms,
                "share_of_step": depth_of(model) * scan_saved_ms / (ms_total / args.steps),
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650",
                "bf16_vs_fp32_kernels": scan_parity,
                "note": "bound by the shared-memory data pipe and instruction issue, not by HBM (ncu, "
                        "profiles/r02_scan_bwd_fast_ncu.txt: LSU data pipe 65 %, issue slots 58 % busy with 2.6 warps "
                        "per scheduler; 448 instructions per 4 tokens of 16 channels x 16 states); the state records "
                        "(12 KB per token and layer, written by the forward, read here) are implementation traffic "
                        "on top of the algorithmic bytes"}
    if rank != 0:
        grp.close()
        return 0
    parity = None
    if not args.no_parity_check:
        worst, detail = _gradient_parity(dev)
        parity = {"rel_err": worst, "bar": 1e-3, "detail": detail,
                  "what": "parameter gradients of a small fp32 model (depth 2, embed 64) against torch autograd "
                          "through oracle/ on the CPU, max|a-b| / max|b| per parameter, worst parameter"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        frames = 2
        t = _oracle_train_step(model_cfg(args, num_frames=frames), 1, frames, args.img)
        cpu = {"value": 1 / t * frames / args.frames, "unit": cfg["unit"], "cores": cores, "kind": "port",
               "sample": f"1 clip of VideoMamba-{args.model} {frames}f@{args.img} fp32, forward + torch-autograd "
                         f"backward of the oracle/ restatement, {t:.1f} s; scaled by {frames}/{args.frames} frames"}
    line = {
        "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"VideoMamba-Small (embed 384, depth 24) {args.frames}f@{args.img} training step, "
                               f"batch {B}/GPU: forward (fused inference scan) + backward (libvmb200 kernels) + "
                               "SGD update", "config": args.config,
                   "forward_ms": fwd_ms, "backward_ms": bwd_ms,
                   "weights": "random init, A_log/dt_bias/temporal embedding perturbed (general A)",
                   "parallelism": f"batch-sharded replicas x{world}, gradients are NOT all-reduced (the reference's "
                                  "DDP scaffolding is out of scope, SURVEY 2 row 12)",
                   "inputs": "2 clip batches rotated every step",
                   "l2": "per-step working set exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": cfg["unit"],
                "h2d_bytes_per_step": x_in.numel() * x_in.element_size(), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches), "parity_check": parity, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    grp.close()
    if parity is not None and not (parity["rel_err"] <= parity["bar"]):
        sys.stderr.write(f"bench.py: gradient parity check failed: {parity}\n")
        return 3
    return 0


def depth_of(model):
    return len(model.layers)


def main():
    args = parse_args()
    # a run that has not finished after 20 minutes is stuck (the default run takes well under 2): every thread's
    # Python stack goes to stderr and the process exits non-zero instead of holding the GPU until someone kills it
    import faulthandler
    faulthandler.dump_traceback_later(1200, exit=True)
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "train16f":
        return run_train(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
