#!/usr/bin/env python
"""bench.py -- pose projections/s (forward + analytic d(dist)/d(pose) + step) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): a batch of 65 536 synthetic
21x4 quaternion poses PER GPU (weak scaling), random-init configs/amass.yaml weights (lrelu, "sensitised"
init so that d > 0), one projection step x <- x - d * dd/dx per "step"
(/root/reference/experiments/sample_poses.py:71-74).  With N > 1 every step ends with the NCCL all-gather
of the projected poses that the north star puts at the end of a projection run.

One JSON line on stdout (rank 0).  `value` = poses of all ranks / max-over-ranks device time, inputs
resident in HBM, L2 flushed between timed steps.  `e2e` = same step through the host-buffer C-ABI call
(pinned host memory, H2D and D2H inside the timed region).  `roofline` reports the kernel against the
fp32-FMA peak measured in-process (the binding pipe: 8 060 flop per compulsory HBM byte, SURVEY Appx C)
and `roofline_hbm` the north-star-mandated HBM fraction.  `cpu_baseline` / `--impl reference` time the
oracle's torch-CPU port of the reference path on the host cores (the only places that touch oracle/).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

BATCH_PER_GPU = 65536
FLOPS_PER_PROJECTION = 5_450_416      # 2*MAC, forward + input gradient
BYTES_PER_PROJECTION = 676            # 336 in + 336 out + 4 dist
WEIGHT_SEED, POSE_SEED = 1, 1234
METRIC = "pose projections/s (fwd+grad+step)"
UNIT = "poses/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx, pw = [], set(), None, []
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2]); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


def host_threads():
    """threads this process can really use: min(affinity mask, cgroup cpu quota) -- on the GPU box os.cpu_count()
    says 128 while the container's cgroup grants 16 CPUs; oversubscribing them makes torch ~10x slower."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def cpu_reference_rate(batch, budget_s, threads):
    """oracle torch-CPU port of the reference step (forward, autograd gradient, x - d*g): best poses/s over as
    many repetitions of a `batch`-pose step as fit in ~budget_s seconds (at least 2, at most 8)."""
    from oracle import posendf_torch as otorch
    from oracle.posendf_numpy import default_cfg
    from posendf_b200 import synth
    torch.set_num_threads(threads)
    tp = otorch.to_torch_params(synth.make_params(WEIGHT_SEED), torch.float32)
    x = torch.from_numpy(synth.make_poses(POSE_SEED, batch))
    cfg = default_cfg()
    otorch.project_step(tp, x[: min(batch, 2048)], cfg)      # warm-up
    best = float("inf")
    times = []
    t_start = time.perf_counter()
    while len(times) < 2 or (len(times) < 8 and time.perf_counter() - t_start < budget_s):
        t0 = time.perf_counter()
        otorch.project_step(tp, x, cfg)
        dt = time.perf_counter() - t0
        times.append(dt)
        best = min(best, dt)
    return batch / best, times


def run_reference(args, rank, world):
    """reference arm: the reference's CPU implementation of the path (oracle port; the Python reference cannot
    travel to the GPU box) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    threads = host_threads()
    sample = 8192
    from oracle import posendf_torch as otorch
    from oracle.posendf_numpy import default_cfg
    from posendf_b200 import synth
    torch.set_num_threads(threads)
    tp = otorch.to_torch_params(synth.make_params(WEIGHT_SEED), torch.float32)
    x = torch.from_numpy(synth.make_poses(POSE_SEED, sample))
    cfg = default_cfg()
    for _ in range(max(1, min(args.warmup, 2))):
        otorch.project_step(tp, x, cfg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        otorch.project_step(tp, x, cfg)
    dt = time.perf_counter() - t0
    rate = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 65 536-pose forward + d(dist)/d(pose) projection step per GPU, amass.yaml lrelu",
                   "batch_per_gpu": BATCH_PER_GPU, "sample_per_step": sample,
                   "note": "reference path = torch CPU (F.linear chain + autograd.grad), timed on a bounded 8 192-pose sample per step"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x {sample} poses (forward + autograd grad + step), torch {torch.__version__} CPU"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="poses per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from posendf_b200 import synth
    from posendf_b200.engine import Engine, fp32_peak_tflops

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    eng = Engine(device=local_rank)
    eng.set_weights_flat(synth.flatten_params(synth.make_params(WEIGHT_SEED)))
    poses_host = torch.from_numpy(synth.make_poses(POSE_SEED + rank, B)).pin_memory()
    x0 = poses_host.to(dev)
    x = x0.clone()
    gathered = torch.empty(world * B, 21, 4, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step():
        eng.project_(x, steps=1)
        if world > 1:
            dist.all_gather_into_tensor(gathered, x)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        x.copy_(x0)
        step()
    sync_all()

    # ---- timed region: exactly K steps, device time per step, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = eng.launch_count()
    sync_all()
    for i in range(args.steps):
        x.copy_(x0)
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        step()
        ev[i][1].record()
    sync_all()
    launches = eng.launch_count() - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([float(sum(ms))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = total_ms.item()

    # ---- kernel-only time (no all-gather) for the roofline, same flush discipline
    kms = []
    for i in range(min(args.steps, 10)):
        x.copy_(x0)
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.project_(x, steps=1); b.record()
        torch.cuda.synchronize(dev)
        kms.append(a.elapsed_time(b))
    kernel_ms = float(np.mean(kms))

    # ---- e2e: host buffers through the C-ABI host entry point (H2D + kernel + D2H inside the timed region)
    out_host = torch.empty_like(poses_host).pin_memory()
    dist_host = torch.empty(B, 1).pin_memory()
    for _ in range(2):
        eng.project_host(poses_host, steps=1, out=out_host, dist_out=dist_host)
    sync_all()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        eng.project_host(poses_host, steps=1, out=out_host, dist_out=dist_host)     # synchronous call
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_rate = world * B * e2e_steps / e2e_s.item()
    clocks = sampler.stop()      # sampled across the timed steps, the kernel-only loop and the e2e loop

    if rank == 0:
        peaks, peak_src = measured_peaks()
        p_ffma = fp32_peak_tflops(local_rank, 0)
        p_ffma2 = max(fp32_peak_tflops(local_rank, 1), fp32_peak_tflops(local_rank, 5))   # two operand orders, best one
        p_fp32 = max(p_ffma, p_ffma2)
        rate = world * B * args.steps / (total_ms * 1e-3)
        k_rate = B / (kernel_ms * 1e-3)
        ach_tf = k_rate * FLOPS_PER_PROJECTION / 1e12
        ach_gbs = k_rate * BYTES_PER_PROJECTION / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 65 536-pose forward + d(dist)/d(pose) projection step per GPU, amass.yaml lrelu",
                       "batch_per_gpu": B, "global_batch": world * B, "weights": "random-init amass.yaml, sensitised (SURVEY 8d)",
                       "l2": "256 MiB flush write between timed steps", "parallelism": f"pose-sharded x{world}"
                       + (", NCCL all-gather of projected poses per step" if world > 1 else "")},
            "roofline": {"bound": "fp32_fma", "achieved": ach_tf, "peak": p_fp32, "unit": "TFLOP/s", "frac": ach_tf / p_fp32,
                         "traffic": traffic, "kernel": "pndf_fused_kernel<1>", "kernel_ms": kernel_ms,
                         "peak_source": "in-process FFMA micro-benchmark (pndf_fp32_peak: scalar %.1f, packed f32x2 %.1f TFLOP/s); "
                                        "tensor cores unused: fp32 parity bar 1e-5" % (p_ffma, p_ffma2),
                         "algorithmic_flops_per_pose": FLOPS_PER_PROJECTION},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": ach_gbs / peaks["hbm_gbs"], "peak_source": peak_src + " (of measured)",
                             "algorithmic_bytes_per_pose": BYTES_PER_PROJECTION,
                             "note": "north-star figure; the path is 8 060 flop/byte, i.e. FMA-bound not HBM-bound (SURVEY Appx C)"},
            "e2e": {"value": e2e_rate, "unit": UNIT, "h2d_bytes_per_step": B * 336, "d2h_bytes_per_step": B * 336 + B * 4,
                    "api": "pndf_project_host (pinned host buffers, chunked H2D/kernel/D2H overlap)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            cpu_rate, times = cpu_reference_rate(16384, 20.0, threads)
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{len(times)} x 16 384 poses of the same step (forward + autograd grad + x-d*g), "
                                              f"best of {len(times)}, {sum(times):.1f} s CPU, torch {torch.__version__}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
