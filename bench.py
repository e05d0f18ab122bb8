#!/usr/bin/env python
"""bench.py -- frames/sec of the OpenSora-v1.2 denoising loop on the vsb200 sm_100a path (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--pab]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one iteration of the RFLOW sampling loop (schedulers/scheduling_rflow_open_sora.py:238-250 in the
reference): CFG batch of 2 through STDiT3.forward (28 spatial + 28 temporal blocks), guidance combine, Euler update,
on synthetic latents and random-init weights of the named architecture.  metric = output frames / (sampling steps x
seconds per step).  N > 1 shards the sequence with DSP (strong scaling: the same video, split over N GPUs).

JSON line: value (inputs resident in HBM), e2e (host buffers, H2D/D2H inside the timed region), roofline (the GEMM
kernel, timed live with CUDA events on the launching stream during the timed region), cpu_baseline (the oracle port on
the host cores, bounded sample), clocks, gpu_launches.  --impl reference times the reference's CPU path (oracle port).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (height, width, frames, sampling steps, latent T,H,W, text tokens)
    "opensora_720p_68f_50step": dict(h=720, w=1280, frames=68, steps=50, lat=(20, 90, 160), L=300),
    "opensora_240p_51f_30step": dict(h=240, w=426, frames=51, steps=30, lat=(15, 30, 53), L=300),
}
MODEL = dict(hidden_size=1152, num_heads=16, depth=28, caption_channels=4096, model_max_length=300)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1386.7), tflops_burst=d.get("bf16_tflops", 1674.1),
                    hbm=d.get("hbm_gbs", 6572.9), src="measured (MEASURED_PEAKS.json, sustained bf16 GEMM)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


def step_flops(W, depth=28, C=1152, B=2):
    """Algorithmic FLOPs of one denoising step: returns (attention QK^T + PV only, everything dense)."""
    T, Hl, Wl = W["lat"]
    S = ((Hl + 1) // 2) * ((Wl + 1) // 2)  # 2x2 spatial patches, odd sizes padded
    N, L = B * T * S, W["L"]
    spatial = depth * 4.0 * (B * T) * S * S * C
    temporal = depth * 4.0 * (B * S) * T * T * C
    cross = 2 * depth * 4.0 * N * L * C
    linear = 2 * depth * (2.0 * N * C * 3 * C + 2.0 * N * C * C + 2.0 * N * C * C + 2.0 * (B * L) * C * 2 * C + 2.0 * N * C * C
                          + 16.0 * N * C * C)
    attn = spatial + temporal + cross
    return attn, attn + linear


def attention_roofline(W, sec_per_step, peaks, depth=28):
    """north_star: frames/s 'as achieved fraction of the attention-FLOP roofline' = the time the attention FLOPs alone
    need at the measured tensor peak, over the measured step time (and the same for all dense FLOPs of the step)."""
    attn, total = step_flops(W, depth)
    return {"attention_flops_per_step": attn, "dense_flops_per_step": total, "peak_tflops": peaks["tflops"],
            "frac_attention_only": attn / (peaks["tflops"] * 1e12) / sec_per_step,
            "frac_all_dense_flops": total / (peaks["tflops"] * 1e12) / sec_per_step,
            "note": "fraction of the step time that the listed FLOPs would take at the measured sustained bf16 peak"}


def kernel_fractions(shares, peaks, nvlink_gbs=770.0):
    """Adds `frac_of_peak` to every per-kernel entry: achieved / the measured peak that bounds it (tensor
    pipe for GEMM and attention, HBM copy bandwidth for the elementwise passes and the short attention, NVLink peer copy
    for the DSP reshard -- 770 GB/s is what a bulk peer copy reached on this pool's boxes)."""
    out = {}
    for k, v in shares.items():
        v = dict(v)
        if v.get("unit") == "TFLOP/s":
            v["frac_of_peak"], v["peak"] = v["achieved"] / peaks["tflops"], f"{peaks['tflops']:.0f} TFLOP/s sustained bf16 GEMM"
        elif k == "dsp_switch":
            v["frac_of_peak"], v["peak"] = v["achieved"] / nvlink_gbs, f"{nvlink_gbs:.0f} GB/s peer copy"
        else:
            v["frac_of_peak"], v["peak"] = v["achieved"] / peaks["hbm"], f"{peaks['hbm']:.0f} GB/s HBM copy"
        out[k] = v
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port (CPU restatement of the reference's eager path) on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference(workload, budget_s=8.0, reps=1):
    """Times a bounded sample: ONE (spatial + temporal) block pair of the 28, on T_s of the T latent frames
    (CFG batch 2, full S patches, full text length), bf16 eager, all host threads.  Returns seconds per FULL step
    extrapolated x28 x T/T_s, and a description of the sample."""
    from oracle import stdit3_oracle as O

    torch.set_num_threads(os.cpu_count())
    W = WORKLOADS[workload]
    T, Hl, Wl = W["lat"]
    S = -(-Hl // 2) * -(-Wl // 2)
    C, H = MODEL["hidden_size"], MODEL["num_heads"]
    g = torch.Generator().manual_seed(0)
    bf = torch.bfloat16

    def rnd(*shape, std=0.02):
        return (torch.randn(*shape, generator=g) * std).to(bf)

    sd = {}
    for kind in ("spatial", "temporal"):
        p = f"{kind}_blocks.0."
        sd[p + "scale_shift_table"] = rnd(6, C, std=C**-0.5)
        for name, (o, i) in {"attn.qkv": (3 * C, C), "attn.proj": (C, C), "cross_attn.q_linear": (C, C),
                             "cross_attn.kv_linear": (2 * C, C), "cross_attn.proj": (C, C), "mlp.fc1": (4 * C, C),
                             "mlp.fc2": (C, 4 * C)}.items():
            sd[p + name + ".weight"] = rnd(o, i)
            sd[p + name + ".bias"] = rnd(o)
        sd[p + "attn.q_norm.weight"] = torch.ones(C // H, dtype=bf)
        sd[p + "attn.k_norm.weight"] = torch.ones(C // H, dtype=bf)
    freqs = O.rope_freqs(C // H).to(bf)
    L = W["L"]

    def run(Ts):
        B = 2
        x = rnd(B, Ts * S, C, std=1.0)
        y = rnd(1, B * L, C, std=1.0)
        t, t0 = rnd(B, 6 * C, std=0.5), rnd(B, 6 * C, std=0.5)
        xm = torch.ones(B, Ts, dtype=torch.bool)
        t_a = time.perf_counter()
        with torch.no_grad():
            h = O.stdit3_block(sd, "spatial_blocks.0.", x, y, t, [L] * B, xm, t0, Ts, S, H, False)
            h = O.stdit3_block(sd, "temporal_blocks.0.", h, y, t, [L] * B, xm, t0, Ts, S, H, True, freqs)
        return time.perf_counter() - t_a

    run(1)  # warm the thread pool / allocator
    t1 = run(1)
    Ts = max(1, min(T, int(budget_s / max(t1, 1e-3))))
    scale = 28.0 * T / Ts
    desc = (f"{workload}: 1 of 28 (spatial+temporal) block pairs of the oracle port, CFG batch 2 x {Ts} of {T} latent "
            f"frames x {S} patches, {L} text tokens, bf16 eager, extrapolated x28 x {T}/{Ts}")
    times = [t1 * scale] if (Ts == 1 and reps == 1) else [run(Ts) * scale for _ in range(reps)]
    return times, desc, os.cpu_count(), (lambda: run(Ts) * scale)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    W = WORKLOADS[args.workload]
    per_step = []
    desc = cores = None
    budget = max(2.0, min(12.0, 150.0 / (args.steps + args.warmup)))
    _, desc, cores, again = cpu_reference(args.workload, budget_s=budget)
    for i in range(args.warmup + args.steps):
        t = again()
        if i >= args.warmup:
            per_step.append(t)
    sec = statistics.mean(per_step)
    val = W["frames"] / (W["steps"] * sec)
    line = {
        "impl": "reference", "metric": "frames/sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config(args, W),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _config(args, W):
    return {"workload": args.workload, "resolution": f"{W['h']}x{W['w']}", "frames": W["frames"],
            "sampling_steps": W["steps"], "latent": list(W["lat"]), "cfg_batch": 2, "text_tokens": W["L"],
            "architecture": "STDiT3-XL/2 (hidden 1152, 16 heads x 72, 28+28 blocks)", "pab": bool(args.pab),
            "parallelism": f"dsp{args.gpus}" if args.gpus > 1 else "single",
            "l2": "per-step working set (activations 332 MB/tensor at 720p) exceeds the 126 MB L2; no flush needed"}


def make_line(args, W, world, sec, sec_e2e, sec_profiled, launches, roofline, shares, cpu_base, clocks, peaks, depth,
              h2d_bytes, d2h_bytes):
    """The one JSON line of the ours arm (pure: unit-tested on CPU).  sec / sec_e2e / sec_profiled: seconds for
    args.steps steps of the resident arm, the host-buffer arm and the all-kernels-profiled pass."""
    frames, nsteps = W["frames"], W["steps"]
    per = sec / args.steps
    line = {
        "metric": "frames/sec", "value": frames / (nsteps * per), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config(args, W),
        "e2e": {"value": frames / (nsteps * sec_e2e / args.steps), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": sec_e2e / args.steps * 1e3},
        "gpu_launches": launches, "roofline": roofline,
        "attention_flop_roofline": attention_roofline(W, per * world, peaks, depth),
        "kernels": shares,
        "kernels_note": "gemm: events inside the timed region; the other kinds: a second pass of the same steps with "
                        f"events around every launch ({sec_profiled / args.steps * 1e3:.1f} ms/step with that overhead)",
        "cpu_baseline": cpu_base,
        "clocks": clocks,
    }
    if args.opt:
        line["config"]["options"] = args.opt
    if args.depth:
        line["config"]["depth_override"] = args.depth
        line["invalid"] = "reduced depth (debug run): not a bench value"
    return line


# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import videosys_b200  # noqa: F401
    from videosys_b200 import kernels
    from videosys_b200.core.distributed.parallel_mgr import initialize
    from videosys_b200.core.pab import pab_mgr
    from videosys_b200.models.transformers.open_sora_transformer_3d import STDiT3, STDiT3Config
    from videosys_b200.pipelines.open_sora.pipeline_open_sora import OpenSoraPABConfig
    from videosys_b200.schedulers.scheduling_rflow_open_sora import RFLOW

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        initialize(rank=rank, world_size=world)
    W = WORKLOADS[args.workload]
    bf = torch.bfloat16

    for kv in args.opt:
        k, v = kv.split("=")
        kernels.set_option(k, int(v))
    torch.manual_seed(0)
    cfg = dict(MODEL)
    if args.depth:
        cfg["depth"] = args.depth
    net = STDiT3(STDiT3Config(**cfg)).to(bf).to(dev).eval()
    net.enable_parallel(dp_size=1, sp_size=world)
    sched = RFLOW(num_sampling_steps=W["steps"], cfg_scale=7.0, use_timestep_transform=True)
    if args.pab:
        pab_mgr.set_pab_manager(OpenSoraPABConfig())
        pab_mgr.update_steps(W["steps"])

    T, Hl, Wl = W["lat"]
    g = torch.Generator(device="cpu").manual_seed(1)
    z_host = torch.randn(1, 4, T, Hl, Wl, generator=g).pin_memory()
    y = torch.randn(1, 1, W["L"], MODEL["caption_channels"], generator=g).to(dev, bf)
    y_null = net.y_embedder.y_embedding[None, None].to(bf)
    margs = dict(
        y=torch.cat([y, y_null], 0), mask=torch.ones(1, W["L"], dtype=torch.long, device=dev),
        height=torch.tensor([W["h"]], device=dev, dtype=bf), width=torch.tensor([W["w"]], device=dev, dtype=bf),
        num_frames=torch.tensor([W["frames"]], device=dev, dtype=bf), fps=torch.tensor([24], device=dev, dtype=bf),
    )
    # the timestep schedule is host arithmetic (50 tiny transforms): keep it off the GPU launch list
    margs_cpu = {k: v.cpu() for k, v in margs.items() if k in ("height", "width", "num_frames")}
    timesteps = [t.to(dev) for t in sched.prepare_timesteps(1, "cpu", margs_cpu)]
    fwd_args = {k: v for k, v in margs.items() if k != "num_frames"}
    fwd_args["x_mask"] = torch.ones(2, T, dtype=torch.bool, device=dev)  # generate() always passes an all-true mask
    n_ts = len(timesteps)

    def dt_of(i):
        return (timesteps[i] - timesteps[i + 1] if i < n_ts - 1 else timesteps[i]) / 1000.0

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n calls of fn(i) between events; returns seconds (max over ranks)."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        sync_all()
        sec = e0.elapsed_time(e1) / 1e3
        if world > 1:
            t = torch.tensor([sec], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = t.item()
        return sec

    # ---- resident arm: latents stay in HBM ----
    z = z_host.to(dev, bf)
    state = {"z": z}

    def step_resident(i):
        k = i % n_ts
        state["z"] = sched.step(net, state["z"], timesteps[k], dt_of(k), fwd_args, 7.0)

    for i in range(args.warmup):
        step_resident(i)
    net.reset_pab_state()
    sampler = ClockSampler(local)
    sampler.start()
    # timed region: CUDA-event pairs only around the dominant kernel (every Linear layer's GEMM, 392 launches per step)
    kernels.PROFILE, kernels.PROFILE_KINDS = [], {"gemm"}
    l0 = kernels.launch_count()
    sec = timed(step_resident, args.steps)
    launches = kernels.launch_count() - l0
    prof_gemm = kernels.PROFILE
    clocks = sampler.stop()
    # a second pass of the same steps with events around EVERY launch: the per-kernel breakdown ("kernels"), not timed
    kernels.PROFILE, kernels.PROFILE_KINDS = [], None
    net.reset_pab_state()
    sec_profiled = timed(step_resident, args.steps)
    prof = [p_ for p_ in kernels.PROFILE if p_[0] != "gemm"] + prof_gemm
    kernels.PROFILE = None

    # ---- per-kernel shares from the live CUDA-event pairs ----
    by_kind = {}
    for kind, a, b, work in prof:
        ms = a.elapsed_time(b)
        d = by_kind.setdefault(kind, [0.0, 0.0, 0])
        d[0] += ms
        d[1] += work
        d[2] += 1
    peaks = _peaks()
    gm = by_kind.get("gemm", [1e-9, 0.0, 1])
    gemm_tflops = gm[1] / (gm[0] * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and args.workload == "opensora_720p_68f_50step":
        traffic = json.load(open(tpath)).get("gemm_qkv_720p_n1")
    roofline = {"kernel": "gemm2_bf16_tn_kernel / gemm_bf16_tn_kernel (all Linear layers, 392 launches per step)",
                "bound": "tensor", "achieved": gemm_tflops,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": gemm_tflops / peaks["tflops"], "traffic": traffic,
                "peak_source": peaks["src"], "launches_timed": gm[2], "timed_in": "the timed region (CUDA-event pairs)",
                "share_of_step": gm[0] / (sec * 1e3)}
    shares = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] / args.steps,
                  "achieved": (v[1] / (v[0] * 1e-3) / 1e12) if k in ("gemm", "attn_flash") else (v[1] / (v[0] * 1e-3) / 1e9),
                  "unit": "TFLOP/s" if k in ("gemm", "attn_flash") else "GB/s"} for k, v in by_kind.items()}
    shares = kernel_fractions(shares, peaks)

    # ---- e2e arm: host (pinned) latents in, host velocity/latents out, every step ----
    out_host = torch.empty(1, 4, T, Hl, Wl, dtype=torch.float32).pin_memory()
    zdev = torch.empty(1, 4, T, Hl, Wl, device=dev, dtype=torch.float32)

    def step_e2e(i):
        k = i % n_ts
        zdev.copy_(z_host, non_blocking=True)
        znew = sched.step(net, zdev.to(bf), timesteps[k], dt_of(k), fwd_args, 7.0)
        out_host.copy_(znew.float(), non_blocking=True)

    net.reset_pab_state()
    step_e2e(0)
    sec_e2e = timed(step_e2e, args.steps)
    frames, nsteps = W["frames"], W["steps"]

    if rank == 0:
        cpu_base = None
        if world == 1 and not args.no_cpu_baseline:
            ts, desc, cores, _ = cpu_reference(args.workload, budget_s=10.0)
            v = frames / (nsteps * statistics.mean(ts))
            cpu_base = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc}
        line = make_line(args, W, world, sec, sec_e2e, sec_profiled, int(launches), roofline, shares, cpu_base, clocks,
                         peaks, cfg["depth"], z_host.numel() * 4, out_host.numel() * 4)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="opensora_720p_68f_50step", choices=sorted(WORKLOADS))
    ap.add_argument("--pab", action="store_true", help="enable Pyramid Attention Broadcast (config 5)")
    ap.add_argument("--depth", type=int, default=0, help="debug only: fewer block pairs (marks the line invalid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="kernel selection knob name=value (vsb_set_option)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
