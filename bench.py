#!/usr/bin/env python
"""Benchmark of the Pano-NeRF hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|render] [--impl ours|reference]

N=1 workload = BASELINE.json configs[1]: the configs/panonerf.yaml training step (mip-NeRF + HDR irradiance +
surface-rendering branch, ort + chroma losses) on 8192 synthetic equirectangular rays, 64 coarse + 64 fine samples,
10 env directions x 10 env samples, bf16 tensor-core MLP, forward + backward + gradient all-reduce + Adam.
Under torchrun every rank trains on its own 8192 rays (weak scaling) and the flat 2.45 MB gradient is all-reduced
with NCCL.  `--workload render` times the communication-free full-panorama render (configs[2]) instead.

One JSON line is printed by rank 0.  `--impl reference` times the reference algorithm on the host CPU cores
(oracle/ port of the pure-PyTorch reference; the upstream tree itself is not present on the GPU box).
"""
import argparse
import datetime
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 8192
N_SAMPLES = 64
GRID_HW = (256, 512)
MLP_FLOP_PER_SAMPLE = 1222656.0          # SURVEY.md §8d, 5 density channels
JAC_FLOP_PER_SAMPLE = 1016320.0          # one trunk input-gradient pass


def camera():
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.1, 0.2, 0.3]
    return c2w


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------------------------
def host_batch(rank, n_rays):
    """Host-side (pinned) ray batch + HDR ground truth: rays of the 256 x 512 grid from the package's own equirect
    generator (the GPU arm does not touch oracle/), a seeded random subset, copied back to pinned host memory once."""
    from panonerf_b200.datasets.pano_datasets import generate_rays
    h, w = GRID_HW
    rays = generate_rays(h, w, camera(), 0.0, 10.0, torch.device("cuda", torch.cuda.current_device()))
    g = torch.Generator().manual_seed(rank)
    perm = torch.randperm(h * w, generator=g)[:n_rays].to(rays.origins.device)
    packed = torch.cat([x[perm] for x in rays], dim=1).contiguous().cpu()                     # [n, 14]
    gt = torch.rand(n_rays, 3, generator=g) * 2
    return packed.pin_memory(), gt.pin_memory()


def synth_weights(mlp, seed=4):
    """Deterministic Xavier-uniform-like weights for random-init benchmarks, drawn from a CPU generator in state-dict
    order (the same numbers oracle.synth_state_dict hands the CPU arm: same order, shapes, bounds and generator)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in mlp.state_dict().items():
        s = tuple(v.shape)
        bound = math.sqrt(6.0 / (s[0] + s[1])) if k.endswith("weight") else 0.05
        sd[k] = (torch.rand(s, generator=g) * 2 - 1) * bound
    return sd


def unpack_rays(packed):
    from panonerf_b200.datasets.base_datasets import Rays
    widths = (3, 3, 3, 1, 1, 1, 1, 1)
    out, o = [], 0
    for wd in widths:
        out.append(packed[:, o:o + wd].contiguous())
        o += wd
    return Rays(*out)


def make_system(device, precision="bf16", seed=4, num_samples=None):
    from panonerf_b200.systems.base_system import default_hparams
    from panonerf_b200.systems.panonerf_system import PanoNeRFSystem
    from panonerf_b200.datasets.pano_datasets import generate_lit_rays, pixel_radius
    hp = default_hparams("panonerf", precision=precision)
    hp["train.randomized"] = True
    if num_samples:
        hp["nerf.num_samples"] = int(num_samples)
    system = PanoNeRFSystem(hp).to(device)
    system.mip_nerf.mlp.load_state_dict(synth_weights(system.mip_nerf.mlp, seed))
    radius = pixel_radius(GRID_HW[0], GRID_HW[1], camera(), device)
    system.env_rays = generate_lit_rays(radius, num=10, device=device)
    return system


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from panonerf_b200 import field, ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    torch.manual_seed(4 + rank)
    system = make_system(dev)
    opt = system.configure_optimizers()
    packed_h, gt_h = host_batch(rank, RAYS_PER_GPU)
    packed_d, gt_d = packed_h.to(dev), gt_h.to(dev)
    rays_d = unpack_rays(packed_d)

    def step_eager():
        opt.zero_grad()
        loss = system.training_step((rays_d, gt_d))
        loss.backward()
        opt.step()
        return loss

    # The public API offers the step as a CUDA graph (systems.base_system.GraphedTrainStep): same kernels, one graph
    # launch instead of ~240 kernel launches.  Eager execution remains available (--no-graph) and is what the
    # per-kernel CUDA-event profile below uses.
    graphed = None
    if not args.no_graph:
        from panonerf_b200.systems.base_system import GraphedTrainStep
        try:
            graphed = GraphedTrainStep(system, opt, rays_d, gt_d)
        except Exception as e:                   # capture refused: keep the eager step (still the CUDA path)
            sys.stderr.write(f"bench.py: CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly\n")
            graphed = None

    def step_resident():
        return graphed() if graphed is not None else step_eager()

    def step_e2e():
        p = packed_h.to(dev, non_blocking=True)
        g = gt_h.to(dev, non_blocking=True)
        if graphed is not None:
            return float(graphed(unpack_rays(p), g))
        opt.zero_grad()
        loss = system.training_step((unpack_rays(p), g))
        loss.backward()
        opt.step()
        return float(loss)                       # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = ops.launch_count()
        if profile:
            field.PROFILE = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = field.PROFILE
        field.PROFILE = None
        clocks = sampler.result()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), ops.launch_count() - l0, clocks, prof, out

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # Pre-heat: keep stepping (untimed) until the GPU has been busy for `--preheat` seconds, so that the timed steps
    # run at the clocks a long training run settles to (the sustained roofline denominator then applies even when
    # the timed region itself is short).
    preheat_steps = 0
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < args.preheat:
        for _ in range(10):
            step_resident()
        torch.cuda.synchronize()
        preheat_steps += 10
    ms, launches, clocks, _, last = timed(step_resident, args.steps)
    if graphed is not None:
        launches = graphed.launches_per_step * args.steps      # replayed launches are not seen by the host counter
    for _ in range(2):
        step_e2e()
    ms_e2e, _, _, _, _ = timed(step_e2e, args.steps)
    # per-kernel roofline: the same step, eagerly, with CUDA events around every tensor-core launch
    step_eager()
    ms_prof, _, _, prof, _ = timed(step_eager, args.steps, profile=True)

    total_rays = RAYS_PER_GPU * world * args.steps
    value = total_rays / (ms / 1e3)
    e2e = total_rays / (ms_e2e / 1e3)
    pk = peaks()
    roof = roofline_from_profile(prof, args.steps, pk, step_ms=ms / args.steps, step_flop=step_flops(),
                                 ideal_bytes_per_step=ideal_train_bytes(RAYS_PER_GPU))
    line = {
        "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs/panonerf.yaml training step: 8192 rays/GPU, 64+64 samples, 10 env dirs x 10 "
                               "env samples, surface+ort+chroma losses, fwd+bwd+allreduce+Adam",
                   "rays_per_gpu": RAYS_PER_GPU, "num_samples": N_SAMPLES, "parallelism": f"ray-dp{world}",
                   "l2": "activations per step are several GB (>> 126 MB L2): every step streams from HBM",
                   "randomized": True, "cuda_graph": graphed is not None,
                   "preheat_s": args.preheat, "preheat_steps": preheat_steps,
                   "update_in_graph": bool(graphed is not None and graphed.update_in_graph),
                   "eager_ms_per_step_with_profile_events": ms_prof / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": int(packed_h.numel() * 4 + gt_h.numel() * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roof,
        "final_loss": float(last),
        "step_tflops": step_flops() * world / (ms / args.steps / 1e3) / 1e12,
    }
    if not args.no_extras:
        # the other BASELINE.json configurations, as extra keys of the same line (every rank takes part)
        try:
            line["c4"] = measure_c4(system, opt, dev, world, rank, args.c4_steps)
        except Exception as e:                   # noqa: BLE001
            import traceback
            line["c4"] = {"error": f"{type(e).__name__}: {e}", "traceback": traceback.format_exc()[-3000:]}
        graphed = None
        torch.cuda.empty_cache()
        if world == 1:                           # (single-process configuration; with more ranks its optimiser step
            try:                                 # would enter an all-reduce the other ranks never join)
                line["c1"] = measure_c1(dev)
            except Exception as e:               # noqa: BLE001
                line["c1"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
        try:
            r = measure_render(system, dev, world, rank, local, 512, 1024, args.render_chunk, args.render_steps, 1)
            line["render"] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "e2e",
                                                "gpu_launches", "roofline", "step_tflops", "steps")}
        except Exception as e:                   # noqa: BLE001
            line["render"] = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0 and world == 1:
            # the memory-bound stages alone against the HBM roofline (>= 0.7 of the copy bandwidth is the
            # bar per kernel): tools/bench_micro.py, 16.8 M samples, CUDA events over back-to-back launches
            try:
                system = opt = None
                torch.cuda.empty_cache()
                sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
                import bench_micro
                line["micro_kernels"] = {r["kernel"]: {"ms": round(r["ms"], 4), "GBps": round(r["algorithmic_GBps"], 1),
                                                       "frac_of_hbm_roofline": round(r["frac_of_hbm_roofline"], 3)}
                                         for r in bench_micro.run(24, dev, peaks()["hbm"])}
            except Exception as e:               # noqa: BLE001
                line["micro_kernels"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_rays)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_render(system, dev, world, rank, local, H, W, chunk, steps, warmup):
    n_s = int(system.hparams["nerf.num_samples"])
    """configs[2]: full-panorama inference render (1024x512 equirect = 524 288 rays), PanoMipNeRF with normals +
    env irradiance + surface rendering (what the reference's render_image does), rows sharded over the ranks, no
    inter-GPU communication.  `e2e` additionally copies the rendered HDR images back to pinned host memory."""
    import torch.distributed as dist
    from panonerf_b200 import field, ops
    from panonerf_b200.datasets.pano_datasets import generate_rays
    from panonerf_b200.parallel import shard_rows
    row0, nrows = shard_rows(H, rank, world)
    host_out = torch.empty(3 * 4 + 2, nrows * W, pin_memory=True)

    def render(copy_back):
        rays = generate_rays(H, W, camera(), 0.0, 10.0, dev, row0=row0, nrows=nrows)
        rays = type(rays)(*[x.view(1, nrows, W, -1) for x in rays])
        outs = system.render_image((rays, torch.empty(1, nrows, W, 3, device=dev)), chunk_size=chunk or None)
        if copy_back:
            c_rgb, f_rgb, c_dep, f_dep, nor, alb, _, sf, sd = outs
            k = 0
            for x in (f_rgb, nor, alb, sf, c_dep, f_dep):          # contiguous [1,c,H,W] planes -> pinned host memory
                c = x.shape[1]
                host_out[k:k + c].copy_(x.view(c, nrows * W), non_blocking=True)
                k += c
            torch.cuda.synchronize()
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(copy_back, n):
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = ops.launch_count()
        if not copy_back:
            field.PROFILE = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            render(copy_back)
        e1.record()
        barrier()
        prof = field.PROFILE
        field.PROFILE = None
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), ops.launch_count() - l0, sampler.result(), prof

    for _ in range(max(warmup, 1)):
        render(False)
    ms, launches, clocks, prof = timed(False, steps)
    ms_e2e, _, _, _ = timed(True, steps)
    rays_total = H * W * steps
    flop_per_ray = (2 * n_s + 100) * MLP_FLOP_PER_SAMPLE + n_s * JAC_FLOP_PER_SAMPLE
    return {"metric": "render_rays_per_s", "value": rays_total / (ms / 1e3), "unit": "rays/s", "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 1), "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"full-panorama PanoMipNeRF render {W}x{H}, {n_s}+{n_s} samples, normals + 10x10 env "
                                   f"irradiance + surface rendering, rows sharded over ranks, no collective",
                       "rays_per_forward": chunk or system.render_rays_per_launch(nrows * W), "parallelism": f"ray-shard{world}",
                       "l2": "per-chunk activations are several GB (>> 126 MB L2)"},
            "clocks": clocks,
            "e2e": {"value": rays_total / (ms_e2e / 1e3), "unit": "rays/s", "h2d_bytes_per_step": 48,
                    "d2h_bytes_per_step": int(host_out.numel() * 4), "ms_per_step": ms_e2e / steps},
            "gpu_launches": int(launches),
            "roofline": roofline_from_profile(prof, steps, peaks(), step_ms=ms / steps,
                                              step_flop=H * W * flop_per_ray / world),
            "step_tflops": H * W * flop_per_ray / (ms / steps / 1e3) / 1e12}


def run_render(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    system = make_system(dev, num_samples=args.num_samples)
    H, W = args.render_hw
    line = measure_render(system, dev, world, rank, local, H, W, args.render_chunk, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_c4(system, opt, dev, world, rank, steps):
    """BASELINE.json configs[3]: 65 536 rays per GPU and step, data-parallel (one NCCL all-reduce of the flat gradient
    per step).  The step walks the batch in 8 slices of 8192 rays (gradient accumulation, AccumulatedTrainStep) so
    that the activation planes of a slice (~19 GB) rather than of the whole batch (~150 GB) are resident."""
    import torch.distributed as dist
    from panonerf_b200.systems.base_system import AccumulatedTrainStep
    n, k = 65536, 8
    packed_h, gt_h = host_batch(1000 + rank, n)
    packed_d, gt_d = packed_h.to(dev), gt_h.to(dev)
    rays_d = unpack_rays(packed_d)
    step = AccumulatedTrainStep(system, opt, rays_d, gt_d, micro_batches=k)
    step(rays_d, gt_d)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(rays_d, gt_d)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    return {"workload": "configs/panonerf.yaml training step, 65536 rays/GPU (8 accumulated slices of 8192), "
                        "one flat-gradient all-reduce + Adam per step", "rays_per_gpu": n, "micro_batches": k,
            "steps": steps, "ms_per_step": ms, "value": n * world / (ms / 1e3), "unit": "rays/s",
            "step_tflops": step_flops() * (n / RAYS_PER_GPU) * world / (ms / 1e3) / 1e12,
            "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "final_loss": float(loss)}


def measure_c1(dev, steps=10):
    """BASELINE.json configs[0] on the GPU: configs/mipnerf.yaml training step (MipNeRF, C = 1 density channel, no ort
    loss), 4096 rays of a 64 x 128 equirect grid, 128 coarse + 128 fine samples, forward + backward + Adam, eager (no
    CUDA graph).  The reference runs this configuration in fp32 on the CPU (~27 s per step on 8 cores)."""
    from panonerf_b200.datasets.pano_datasets import generate_rays
    from panonerf_b200.systems.base_system import default_hparams
    from panonerf_b200.systems.mipnerf_system import MipNeRFSystem
    n, ns = 4096, 128
    hp = default_hparams("mipnerf", precision="bf16")
    hp.update({"nerf.num_samples": ns, "train.randomized": True})
    system = MipNeRFSystem(hp).to(dev)
    system.mip_nerf.mlp.load_state_dict(synth_weights(system.mip_nerf.mlp, 4))
    opt = system.configure_optimizers()
    rays = generate_rays(64, 128, camera(), 0.0, 10.0, dev)
    g = torch.Generator().manual_seed(0)
    from panonerf_b200.datasets.base_datasets import Rays
    rays_d = Rays(*[x[:n].contiguous() for x in rays])
    gt_d = (torch.rand(n, 3, generator=g) * 2).to(dev)

    def step():
        opt.zero_grad()
        loss = system.training_step((rays_d, gt_d))
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flop = n * 3 * (2 * ns) * (MLP_FLOP_PER_SAMPLE - 2 * 256 * 4)       # C = 1: 1 220 608 FLOP per sample
    return {"workload": "configs/mipnerf.yaml training step: MipNeRF, 4096 rays, 128+128 samples, fwd+bwd+Adam, eager",
            "rays": n, "num_samples": ns, "steps": steps, "ms_per_step": ms, "value": n / (ms / 1e3), "unit": "rays/s",
            "step_tflops": flop / (ms / 1e3) / 1e12, "final_loss": float(loss)}


def step_flops():
    """Algorithmic MLP FLOPs of one 8192-ray Pano training step: 3x forward for every MLP
    evaluation (fwd + dgrad + wgrad) plus the normals Jacobian pass and its adjoint (3x one trunk pass)."""
    main = 2 * N_SAMPLES
    env = 10 * 10
    return RAYS_PER_GPU * (3 * (main + env) * MLP_FLOP_PER_SAMPLE + 3 * N_SAMPLES * JAC_FLOP_PER_SAMPLE)


# Kernel families of the MLP stage.  Every one of them is bounded by the TENSOR roofline (the
# MLP - forward, data gradients, weight gradients, Jacobian sweep and its adjoint - is the one dense contraction of
# the path); `mlp_fused_kernel<P>` is one templated kernel (programs forward / forward+Jacobian / dgrad chain /
# adjoint sweep) and is reported as one family.  The HBM view (GB/s actually moved, ncu DRAM bytes) is kept beside
# it as `traffic`: in training the planes the weight-gradient kernel consumes dominate it.
FAMILY = {"mlp_fused": "mlp_fused_kernel", "mlp_fused_bwd": "mlp_fused_kernel", "mlp_fused_jadj": "mlp_fused_kernel",
          "wgrad_batch": "wgrad_batch_kernel", "linear_tc": "linear_tc_kernel", "wgrad_tc": "wgrad_tc_kernel"}


def roofline_from_profile(prof, steps, pk, step_ms=None, step_flop=None, ideal_bytes_per_step=None):
    """Per kernel family: ALGORITHMIC FLOPs over the CUDA-event time of its launches inside the
    timed region (events on the launching stream) against the measured SUSTAINED bf16 rate (the kernels run inside a
    long step).  Headline = the family with the most device time.  Extra keys:
      mlp_stage   all MLP families together (fused programs + weight gradients): the section-8d MLP figure;
      whole_step  the step's algorithmic MLP FLOPs over the whole step time;
      hbm         what the headline family moves: bytes its launches read/write by design (GB/s, fraction of the
                  measured copy bandwidth) next to `algorithmic_bytes` = the irreducible I/O of a fully fused
                  pipeline (encodings in, raw outputs out, gradients) - the gap is the activation stash.
    `traffic` = DRAM bytes per launch of the headline family from the committed ncu --set full capture of the same
    command (profiles/r0?_step_traffic.json), null if absent."""
    if not prof:
        return None
    traffic = {}
    for name in ("r02_step_traffic.json", "r01_step_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            traffic = json.load(open(tp))
            break
    fam = {}
    for name, rec in prof.items():
        f = fam.setdefault(FAMILY.get(name, name), {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0, "traffic": 0.0,
                                                    "traffic_known": True, "programs": {}})
        ms = sum(a.elapsed_time(b) for a, b in rec["events"])
        n = len(rec["events"])
        f["ms"] += ms
        f["n"] += n
        f["flops"] += rec["flops"]
        f["bytes"] += rec["bytes"]
        t = traffic.get(name, {}).get("dram_bytes_per_launch")
        if t is None:
            f["traffic_known"] = False
        else:
            f["traffic"] += t * n
        f["programs"][name] = {"kernel_ms_per_step": ms / steps, "launches_per_step": n / steps,
                               "tflops": rec["flops"] / (ms / 1e3) / 1e12,
                               "frac": rec["flops"] / (ms / 1e3) / 1e12 / pk["tf_sust"]}

    def entry(k, f):
        tfs = f["flops"] / (f["ms"] / 1e3) / 1e12
        return {"kernel": k, "bound": "tensor", "achieved": tfs, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": tfs / pk["tf_sust"], "frac_of_burst_peak": tfs / pk["tf_burst"],
                "traffic": (f["traffic"] / f["n"]) if f["traffic_known"] else None,
                "algorithmic_flops_per_launch": f["flops"] / f["n"], "launches_per_step": f["n"] / steps,
                "avg_launch_ms": f["ms"] / f["n"], "kernel_ms_per_step": f["ms"] / steps,
                "hbm": {"designed_bytes_per_launch": f["bytes"] / f["n"],
                        "gbps": f["bytes"] / (f["ms"] / 1e3) / 1e9,
                        "frac_of_copy_bandwidth": f["bytes"] / (f["ms"] / 1e3) / 1e9 / pk["hbm"]},
                "programs": f["programs"] if len(f["programs"]) > 1 else None}

    top = max(fam, key=lambda k: fam[k]["ms"])
    out = entry(top, fam[top])
    out["peak_source"] = pk["src"] + " (MEASURED_PEAKS.json bf16_tflops_sustained: kernels timed inside a long step)"
    out["other_kernels"] = {k: entry(k, f) for k, f in fam.items() if k != top}
    ms_all = sum(f["ms"] for f in fam.values())
    fl_all = sum(f["flops"] for f in fam.values())
    out["mlp_stage"] = {"bound": "tensor", "kernel_ms_per_step": ms_all / steps, "achieved": fl_all / ms_all / 1e9,
                        "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": fl_all / ms_all / 1e9 / pk["tf_sust"]}
    if step_ms is not None and step_flop is not None:
        tfs = step_flop / (step_ms / 1e3) / 1e12
        out["whole_step"] = {"bound": "tensor", "ms_per_step": step_ms, "achieved": tfs, "peak": pk["tf_sust"],
                             "unit": "TFLOP/s", "frac": tfs / pk["tf_sust"]}
    if ideal_bytes_per_step is not None:
        out["algorithmic_bytes_per_step"] = ideal_bytes_per_step
        out["designed_bytes_per_step"] = sum(f["bytes"] for f in fam.values()) / steps
    return out


def ideal_train_bytes(rays):
    """Irreducible HBM I/O of a fully fused training step (SURVEY.md section 8d / VERDICT r1): per MLP evaluation the
    Gaussians in (24 B) and the raw outputs out (32 B), the same again for their gradients in the backward pass, the
    rays (56 B), and the 2.45 MB gradient + parameter + Adam-state buffers."""
    samples = rays * (2 * N_SAMPLES + 100)
    return int(samples * 2 * (24 + 32) + rays * 56 + 613768 * 4 * 5)


# ----------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference algorithm)
# ----------------------------------------------------------------------------------------------------------------
def cpu_train_step(n_rays, threads):
    from oracle import panonerf_oracle as O
    torch.set_num_threads(threads)
    h, w = GRID_HW
    rays = O.equirect_rays(h, w, camera(), 0.0, 10.0)
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(h * w, generator=g)[:n_rays]
    rays = O.Rays(*[getattr(rays, k)[perm].contiguous() for k in O.Rays._fields])
    gt = torch.rand(n_rays, 3, generator=g) * 2
    env = O.fibonacci_env_rays(10, float(rays.radii[0, 0]))
    env = O.Rays(*[x.float() for x in env])
    sd = {k: v.clone().requires_grad_() for k, v in O.synth_state_dict(seed=4, width=256, c_density=5).items()}
    # normals exactly as the reference computes them: vmap(jacrev(compute_graph)) over every fine sample
    cfg = dict(num_samples=N_SAMPLES, normals_impl="jacrev")
    params = list(sd.values())
    opt = torch.optim.Adam(params, lr=2e-4)

    def step():
        opt.zero_grad()
        torch.manual_seed(0)
        out, _ = O.panonerf_forward(sd, rays, env, cfg, randomized=True, train=True)
        loss = O.panonerf_loss(out, rays, gt)
        loss.backward()
        opt.step()
        return float(loss)
    return step


def cpu_baseline(n_rays):
    threads = os.cpu_count() or 1
    step = cpu_train_step(n_rays, threads)
    step()
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    return {"value": n_rays / dt, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": f"one fwd+bwd+Adam step of the same panonerf workload on {n_rays} rays (of 8192), fp32, oracle port "
                      f"of the reference incl. its vmap(jacrev) normals, torch CPU ops, {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_rays = args.cpu_rays
    step = cpu_train_step(n_rays, threads)
    for _ in range(min(args.warmup, 1)):
        step()
    steps = min(args.steps, 3)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = n_rays * steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "train_rays_per_s", "value": v, "unit": "rays/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs/panonerf.yaml training step (same as the GPU arm), bounded CPU sample",
                   "rays_per_step": n_rays, "num_samples": N_SAMPLES},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} steps x {n_rays} rays, oracle port of the reference (upstream tree is "
                                   f"not present on the GPU box), {threads} torch threads"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the training step eagerly instead of as a CUDA graph")
    ap.add_argument("--workload", default="train", choices=["train", "render"])
    ap.add_argument("--num-samples", type=int, default=0, help="render workload: samples per level (default: 64, the YAML value)")
    ap.add_argument("--preheat", type=float, default=2.0, help="seconds of untimed steps before the timed region")
    ap.add_argument("--no-extras", action="store_true", help="skip the C3 render and C4 65536-ray figures")
    ap.add_argument("--c4-steps", type=int, default=3)
    ap.add_argument("--render-steps", type=int, default=2)
    ap.add_argument("--render-hw", type=int, nargs=2, default=[512, 1024])
    ap.add_argument("--render-chunk", type=int, default=0, help="rays per forward (0 = the whole image in one)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "render":
        run_render(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
