#!/usr/bin/env python
"""Benchmark of the VideoMamba mixer hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the CPU restatement, timed alone

A step = one forward of VideoMamba-S (embed 384, depth 24), 16 frames @224^2, bf16, over a batch of
`--batch` synthetic clips per GPU (BASELINE.json configs[1]).  For N > 1 the driver launches this
file under torch.distributed.run, one rank per GPU; clips are sharded along batch, there is no
collective on the data path (torch.distributed is used for the start barrier and the
max-over-ranks of the step time only).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
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

METRIC = "clips/s VideoMamba-S 16f@224 bf16 forward"
UNIT = "clips/s"

MODELS = {
    "small": dict(embed_dim=384, depth=24),
    "tiny": dict(embed_dim=192, depth=24),
    "middle": dict(embed_dim=576, depth=32),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="small", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU per step")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--img", type=int, default=224)
    ap.add_argument("--weights", default="perturbed", choices=["perturbed", "init"],
                    help="perturbed: general A (trained-checkpoint-like); init: reference random init")
    ap.add_argument("--in-flight", type=int, default=3,
                    help="steps kept in flight on separate CUDA streams (1 = strictly serial steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=4, help="clips in the CPU baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def build_model(args, dtype, device):
    import torch
    import video_mamba

    torch.manual_seed(0)
    m = MODELS[args.model]
    model = video_mamba.PretrainVideoMamba(
        img_size=args.img, patch_size=16, depth=m["depth"], embed_dim=m["embed_dim"], channels=3,
        ssm_cfg={"use_fast_path": False}, norm_epsilon=1e-5, fused_add_norm=True, rms_norm=True,
        residual_in_fp32=True, bimamba=True, pool_type="cls+avg", kernel_size=1,
        num_frames=args.frames).eval()
    if args.weights == "perturbed":
        # break the S4D-real structure of A and the zero dt bias of the reference's model init, so
        # that the general-A kernels are what is measured
        import math
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


def cpu_baseline(args, clips: int):
    """Times the CPU restatement (oracle/, kind "port") of the same workload on the host cores:
    `clips` clips of the bench configuration, fp32 (the reference's own CPU-runnable precision,
    BASELINE.json configs[0]).  The only place bench.py executes oracle/."""
    import torch
    from oracle import videomamba_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = MODELS[args.model]
    cfg = dict(img_size=args.img, patch_size=16, depth=m["depth"], embed_dim=m["embed_dim"],
               kernel_size=1, num_frames=args.frames, norm_epsilon=1e-5, rms_norm=True,
               fused_add_norm=True, residual_in_fp32=True, pool_type="cls+avg", add_pool_norm=True)
    sd = orc.synthetic_state_dict(cfg, seed=0, perturbed=args.weights == "perturbed")
    model = orc.OracleVideoMamba(cfg, sd)
    x = torch.rand(clips, 3, args.frames, args.img, args.img, generator=torch.Generator().manual_seed(1000))
    t0 = time.perf_counter()
    with torch.no_grad():
        model.forward(x)
    dt = time.perf_counter() - t0
    return {"value": clips / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{clips} clip(s) of VideoMamba-{args.model} {args.frames}f@{args.img} fp32, "
                      f"full forward, torch CPU ops + per-token scan loop, {dt:.1f} s",
            "seconds": dt}


def workload_name(args):
    return (f"VideoMamba-{args.model.capitalize()} (embed {MODELS[args.model]['embed_dim']}, depth "
            f"{MODELS[args.model]['depth']}) {args.frames}f@{args.img} forward, batch {args.batch}/GPU")


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement timed alone
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(args, 1)
    times = []
    for _ in range(args.steps):
        times.append(cpu_baseline(args, args.cpu_clips))
    total = sum(t["seconds"] for t in times)
    value = args.cpu_clips * len(times) / total
    last = times[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "weights": args.weights,
                   "note": "reference rejects CPU tensors and its kernels live in absent wheels; "
                           "this is the CPU restatement (oracle/) of its use_fast_path=False path"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port",
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
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

    dtype = torch.bfloat16
    model = build_model(args, dtype, dev)
    B, T, S = args.batch, args.frames, args.img
    gen = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.rand(B, 3, T, S, S, generator=gen).to(dtype).pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        grp.barrier(dev)

    def max_over_ranks(v: float) -> float:
        return grp.max_over_ranks(v, dev)

    def fwd(x):
        with torch.no_grad():
            return model(x)

    # ---- device-resident throughput ("value") ----------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        fwd(x_dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = (C.c_double * 8)()
    cnt = (C.c_int64 * 8)()

    # (1) strictly serial steps with the library's per-stage CUDA events on: the per-kernel times the
    #     roofline object is computed from (each kernel has the GPU to itself here)
    lib.vmb_prof_enable(1)
    lib.vmb_prof_read(ms, cnt, 1)               # clear anything recorded before
    for i in range(8):
        ms[i] = 0.0
        cnt[i] = 0
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = fwd(x_dev)
    e1.record()
    barrier()
    serial_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    lib.vmb_prof_read(ms, cnt, 1)
    lib.vmb_prof_enable(0)
    stages = {k: {"ms_per_step": ms[i] / args.steps, "launches_per_step": cnt[i] // args.steps}
              for i, k in enumerate(_lib.PROF_KINDS) if cnt[i]}

    # (2) the headline: the same K steps, `in_flight` of them kept in flight on separate streams.
    #     Steps are independent batches; the scan is bound by instruction issue and runs faster
    #     with two launches sharing the SMs, and the HBM-bound kernels of one step overlap the scan
    #     of another.  Every step is a full forward; nothing is skipped or cached.
    main = torch.cuda.current_stream()
    lanes = [torch.cuda.Stream() for _ in range(max(1, args.in_flight))]

    def run_steps(n):
        start = torch.cuda.Event()
        start.record(main)
        last = None
        for i in range(n):
            s = lanes[i % len(lanes)]
            if i < len(lanes):
                s.wait_event(start)
            with torch.cuda.stream(s):
                last = fwd(x_dev)
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
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e") ----------------------------
    # A double-buffered serving loop: step i+1's clips are copied host->device on a copy stream
    # while step i computes, and step i's features go device->host on a second copy stream.  Every
    # step's input copy and result copy are inside the timed region.
    vis0, pool0 = out
    vis_host = torch.empty(vis0.shape, dtype=vis0.dtype).pin_memory()
    pool_host = torch.empty(pool0.shape, dtype=pool0.dtype).pin_memory()
    nbuf = len(lanes) + 1
    x_in = [torch.empty_like(x_dev) for _ in range(nbuf)]
    comp = torch.cuda.current_stream()
    h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()

    def e2e_loop(n):
        ready = [None] * nbuf
        done = [None] * nbuf

        def stage_in(i):                        # host -> device copy of step i's clips
            k = i % nbuf
            with torch.cuda.stream(h2d):
                if done[k] is not None:
                    h2d.wait_event(done[k])     # the forward that last read this buffer is finished
                x_in[k].copy_(x_host, non_blocking=True)
                ready[k] = torch.cuda.Event()
                ready[k].record(h2d)

        for i in range(min(len(lanes), n)):
            stage_in(i)
        last_out = None
        for i in range(n):
            k = i % nbuf
            if i + len(lanes) < n:
                stage_in(i + len(lanes))
            cs = lanes[i % len(lanes)]
            cs.wait_event(ready[k])
            with torch.cuda.stream(cs):
                vis, pool = fwd(x_in[k])
            done[k] = torch.cuda.Event()
            done[k].record(cs)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done[k])
                vis.record_stream(d2h)
                pool.record_stream(d2h)
                vis_host.copy_(vis, non_blocking=True)
                pool_host.copy_(pool, non_blocking=True)
                last_out = torch.cuda.Event()
                last_out.record(d2h)
        comp.wait_event(last_out)

    e2e_loop(2 * len(lanes))
    barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    h2d_bytes = x_host.numel() * x_host.element_size()
    d2h_bytes = vis_host.numel() * vis_host.element_size() + pool_host.numel() * pool_host.element_size()

    # ---- roofline of the dominant HBM kernel (the selective scan) --------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    mixer = model.layers[0].mixer
    Di, N, R = mixer.d_inner, mixer.d_state, mixer.dt_rank
    L = vis0.shape[1] + 1
    tokens = B * L
    es = 2
    w = mixer._kernel_weights()
    fused = "dt_proj" not in stages      # the fused scan expands dt inside the kernel
    if fused:
        # fused dt_proj+softplus+scan+gate kernel: reads u (conv output), z, the x_dbl row; writes y
        bytes_per_token = 3 * Di * es + w.Xp * es
        # scan_fast() picks the two-warp kernel below 9.5 (batch, 16-channel) units per SM
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        two_warp = 2 * B * (Di // 16) < 19 * sms and os.environ.get("VMB_SCAN_VARIANT", "0") in ("0", "11", "12")
        kernel = ("scan11_kernel (helper + consumer warps" if two_warp else "scan10_kernel (one warp per unit") + \
                 "; TMA-staged tiles; dt_proj + softplus + S6 scan + D skip + SiLU gate, fused)"
    else:
        # op-level selective_scan_fn: reads u, delta, z, B, C; writes y
        bytes_per_token = 4 * Di * es + 2 * N * es
        kernel = "scan_generic_kernel (softplus + scan + D skip + SiLU gate)"
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        wl = tr["workload"]
        if fused and (wl["B"], wl["L"], wl["Di"]) == (B, L, Di):
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]     # per launch, from ncu
    except Exception:
        pass
    roofline = None
    if "scan" in stages and stages["scan"]["launches_per_step"]:
        # kernel time from the serial pass (the kernel alone on the GPU); share of the serial step
        per_launch_ms = stages["scan"]["ms_per_step"] / stages["scan"]["launches_per_step"]
        achieved = tokens * bytes_per_token / (per_launch_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": tokens * bytes_per_token,
                    "ms_per_launch": per_launch_ms, "peak_source": peak_src,
                    "share_of_step": stages["scan"]["ms_per_step"] / serial_ms,
                    "note": "kernel is bound by instruction issue / MUFU ex2 (16 per token-channel), "
                            "not by HBM: see DESIGN.md 3.2 and profiles/r01_mufu_issue_microbench.txt"}
    # the other HBM-bound kernels of the path against the same peak (algorithmic bytes per token:
    # conv reads x and writes xc; add+norm reads hidden (bf16) + residual (fp32), writes both)
    hbm_kernels = {}
    Dm = mixer.d_model
    # conv alone reads x and writes xc; fused with x_proj (no separate x_proj stage) it also writes x_dbl
    conv_fused = "x_proj" not in stages or not stages["x_proj"]["launches_per_step"]
    conv_bpt = (2 * Di + w.Xp) * es if conv_fused else 2 * Di * es
    for k, bpt in (("conv", conv_bpt), ("add_norm", Dm * es + 4 * Dm + 4 * Dm + Dm * es)):
        if k in stages and stages[k]["launches_per_step"]:
            per = stages[k]["ms_per_step"] / stages[k]["launches_per_step"]
            gbs = tokens * bpt / (per * 1e-3) / 1e9
            name = "conv_xproj" if (k == "conv" and conv_fused) else k
            hbm_kernels[name] = {"achieved": gbs, "unit": "GB/s", "frac": gbs / hbm_peak, "ms_per_launch": per}
    # projections against the tensor roofline (reported beside, not the dominant-kernel object)
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    D = mixer.d_model
    flops = {"in_proj": 2 * D * 2 * Di, "x_proj": 2 * Di * (R + 2 * N), "dt_proj": 2 * R * Di,
             "out_proj": 2 * Di * D}
    tensor = {}
    for k, f in flops.items():
        if k in stages and stages[k]["launches_per_step"]:
            per = stages[k]["ms_per_step"] / stages[k]["launches_per_step"]
            tf = tokens * f / (per * 1e-3) / 1e12
            tensor[k] = {"tflops": tf, "frac_of_sustained_peak": tf / tc_peak}

    if rank != 0:
        grp.close()
        return 0

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, args.cpu_clips)
        cpu.pop("seconds", None)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "tokens_per_clip": L,
                   "weights": ("random init, A_log/dt_bias/temporal embedding perturbed (general-A "
                               "kernels)" if args.weights == "perturbed" else "reference random init"),
                   "parallelism": f"batch-sharded replicas x{world}, no collective",
                   "steps_in_flight": len(lanes),
                   "serial_ms_per_step": serial_ms,
                   "l2": "per-step working set (>= 150 MB of activations per layer) exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms / args.steps,
                "pipeline": "H2D copy stream, compute lanes, D2H copy stream"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stages": stages,
        "hbm_kernels": hbm_kernels,
        "tensor": tensor,
    }
    print(json.dumps(line), flush=True)
    grp.close()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
