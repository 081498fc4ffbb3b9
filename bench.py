#!/usr/bin/env python
"""EBEN BWE train-step throughput (audio-seconds per second @16 kHz) - BASELINE.json metric.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                      # the CPU port of the reference path (oracle/)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
METRIC = "EBEN BWE train-step audio-sec/sec @16kHz"


def synthetic_pairs(batch: int, samples: int, seed: int):
    """0.1*randn clamped to +-1, (B,1,samples) x2, host generator (seed 42 + rank)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    air = (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)
    body = (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)
    return body, air


_JSON_OUT = None


def claim_stdout():
    """Multi-rank runs: keep stdout for the ONE JSON line.  Libraries write banners to fd 1 (NCCL prints
    "NCCL version ..." there when the box sets NCCL_DEBUG), so fd 1 is pointed at stderr for the life of the
    process and the original stdout is kept aside for `emit`."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def layer_microbench(device, B, L):
    """Per-kernel roofline evidence, measured live with CUDA events on the launching stream:
    (a) the dominant kernel of the step = the dense MelGAN stage 4 implicit-GEMM conv,
    (b) the HBM-bound generator residual-unit convs (north_star's 60 % target)."""
    import torch
    from vibravox_b200 import ops
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    tens = peaks.get("bf16_tflops_sustained", 1400.0)
    src = "measured" if peaks else "fallback"

    def timeit(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    out = []
    # (a) MelGAN stage 4: 1024 -> 1024, k41, s4, groups 4 on (B,1024,748)
    T3 = ((L - 41 + 40) // 4 + 1 - 41 + 40) // 4 + 1
    T3 = (T3 - 41 + 40) // 4 + 1
    g = ops.ConvGeom(1024, 1024, 41, 4, 1, 20, 0, 4)
    x = torch.randn(B, 1024, T3, device=device)
    w = torch.randn(1024, 256, 41, device=device) * 0.01
    bias = torch.zeros(1024, device=device)
    To = g.tout(T3)
    flops = 2.0 * B * To * 1024 * 256 * 41
    byt = 4.0 * (B * 1024 * T3 + B * 1024 * To + w.numel())
    t = timeit(lambda: ops.conv_fwd(x, w, g, bias=bias, slope=0.2), reps=5, warm=2)
    kname = "tc_slab_kernel fwd (tcgen05, bf16x3)" if ops.use_tc(g, "fwd") else "gemm_conv_kernel<FWD> (fp32 FMA)"
    # DRAM bytes of this launch from the committed `ncu --set full` capture (profiles/README.md), bs = 32 only
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath) and B == 32 and ops.use_tc(g, "fwd"):
        traffic = json.load(open(tpath)).get("melgan4_fwd_dram_bytes")
    out.append({"kernel": kname + " melgan.4 (1024->1024,k41,s4,g4)", "bound": "tensor",
                "achieved": flops / t / 1e12, "peak": tens, "unit": "TFLOP/s", "frac": flops / t / 1e12 / tens,
                "ms": t * 1e3, "algorithmic_bytes": byt, "peak_source": src + " bf16 sustained", "traffic": traffic,
                "note": "3 bf16 MMAs per fp32 product: the useful-flop ceiling is peak / 3"})
    # (b) generator residual unit convs at C=32, T=11968
    Tb = (L + 32) // 4
    for C, T in ((32, Tb), (64, Tb // 2), (128, Tb // 8)):
        xg = torch.randn(B, C, T, device=device)
        w1 = torch.randn(C, C, 3, device=device) * 0.1
        w2 = torch.randn(C, C, 1, device=device) * 0.1
        g1 = ops.ConvGeom(C, C, 3, 1, 3, 3, 3, 1)
        g2 = ops.ConvGeom(C, C, 1, 1, 1, 0, 0, 1)
        byt = 4.0 * 2 * B * C * T
        t1 = timeit(lambda: ops.conv_fwd(xg, w1, g1))
        t2 = timeit(lambda: ops.conv_fwd(xg, w2, g2, res=xg, slope=0.01))
        for nm, tt, bb in (("dilated k3 d3", t1, byt), ("pointwise+lrelu+res", t2, byt * 1.5)):
            out.append({"kernel": f"conv fwd residual {nm} C={C} T={T}", "bound": "hbm",
                        "achieved": bb / tt / 1e9, "peak": hbm, "unit": "GB/s", "frac": bb / tt / 1e9 / hbm,
                        "ms": tt * 1e3, "algorithmic_bytes": bb, "peak_source": src + " copy", "traffic": None})
    return out


def run_ours(args):
    import torch
    from vibravox_b200 import _lib, parallel
    import vibravox_b200

    if parallel.env_world()[2] > 1:
        claim_stdout()
    rank, local_rank, world = parallel.init_from_env("nccl")
    if world != args.gpus:
        assert world == 1 and args.gpus == 1, f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, S = args.batch, int(args.seconds * SR)
    lm = vibravox_b200.build_model(seed=42, device=dev)
    L = S - (S + 32) % lm.generator.multiple
    body_h, air_h = synthetic_pairs(B, S, parallel.rank_seed(42, rank))
    body_h, air_h = body_h.pin_memory(), air_h.pin_memory()
    body_d, air_d = body_h.to(dev), air_h.to(dev)
    batch = {"audio_body_conducted": body_d, "audio_airborne": air_d}

    graphed = (not args.no_graph) and (not args.profile) and lm.graph_capturable()

    def step():
        # the public call: replays the captured step (after 2 eager calls + 1 capture, all inside the warm-up)
        (lm.training_step_graphed if graphed else lm.training_step)(batch)

    def timed(fn, k):
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        parallel.barrier()
        wall = time.perf_counter() - t0
        return parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev), wall

    if args.profile:
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()     # ncu --profile-from-start off sees exactly --steps steps
        t_dev, _ = timed(step, args.steps)
        torch.cuda.cudart().cudaProfilerStop()
        if rank == 0:
            emit({"profile_run": True, "ms_per_step": t_dev / args.steps * 1e3})
        return
    n_warm = max(args.warmup, 3) + (3 if graphed else 0)   # graph mode: 2 eager calls + the capture come first
    for _ in range(n_warm):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    t_dev, _ = timed(step, args.steps)
    launches = _lib.launch_count() - n0
    if graphed:
        assert lm.graph_launches() > 0, "the step was not captured during the warm-up"
        launches = lm.graph_launches() * args.steps     # kernel nodes of the replayed graph, counted at capture
    clocks = sampler.stop() if rank == 0 else None
    audio_s = world * B * L / SR
    value = audio_s * args.steps / t_dev

    # e2e: the public call with HOST buffers; H2D of the step's inputs and D2H of its losses inside the timed region
    stage_body, stage_air = torch.empty_like(body_d), torch.empty_like(air_d)
    loss_h = torch.empty(2, dtype=torch.float32).pin_memory()
    host_batch = {"audio_body_conducted": body_h, "audio_airborne": air_h}

    def e2e_step():
        if graphed:
            lm.training_step_graphed(host_batch)          # H2D straight into the captured step's input buffers
        else:
            stage_body.copy_(body_h, non_blocking=True)
            stage_air.copy_(air_h, non_blocking=True)
            lm.training_step({"audio_body_conducted": stage_body, "audio_airborne": stage_air})
        loss_h[0:1].copy_(lm.logged["train/generator/backprop_loss"].view(1), non_blocking=True)
        loss_h[1:2].copy_(lm.logged["train/discriminator/backprop_loss"].view(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    _, wall = timed(e2e_step, args.steps)
    wall = parallel.max_over_ranks(wall, dev)
    e2e = {"value": audio_s * args.steps / wall, "unit": "audio-s/s",
           "h2d_bytes_per_step": int(2 * body_h.numel() * 4), "d2h_bytes_per_step": 8,
           "losses": [float(loss_h[0]), float(loss_h[1])],
           "losses_finite": bool(torch.isfinite(loss_h).all())}
    if not e2e["losses_finite"]:
        sys.stderr.write("WARNING: non-finite training losses at the end of the e2e loop - this run is not valid\n")

    if rank != 0:
        return
    from vibravox_b200 import ops as _ops
    dtype = ("fp32 storage; contractions on tcgen05 tensor cores with bf16x3 split operands, fp32 accumulate"
             if _ops.TC_ENABLED else "fp32")
    kernels = layer_microbench(dev, B, L) if not args.no_micro else []
    roof = dict(kernels[0]) if kernels else None
    line = {
        "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"EBEN BWE full train step (gen+disc+MR-STFT/FM/hinge, EMA balancing, 2x Adam) "
                               f"bs={B}x{args.seconds:g}s@16kHz per GPU (L={L} after cut_to_valid_length), "
                               f"m=4 n=32 p=2 q=4 min_channels=24, schedule={lm.schedule}",
                   "batch_per_gpu": B, "samples": L, "parallelism": f"dp{world}",
                   "launch": ({"whole": "one CUDA graph replay per step",
                               "segments": "three CUDA graph replays per step, split at the two eager NCCL gradient "
                                           "all-reduces"}[lm.graph_mode()] if graphed else "eager launches"),
                   "l2": "per-step working set (activations ~ GBs) >> 126 MB L2; no explicit flush"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roof, "kernel_rooflines": kernels,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, bounded=True)
    emit(line)


def _shutdown():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, steps: int, warmup: int = 1, budget_s: float = 1e9):
    """The oracle port of the reference path on the host cores (kind 'port'): the SAME workload as the GPU arm
    (batch, length, seeds), all host threads.  `steps` timed steps after `warmup` untimed ones; when the first
    warm-up step shows that the run would not fit `budget_s`, fewer timed steps are taken and the line says so."""
    import torch
    from oracle import eben_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.cpu_batch
    S = int(args.seconds * SR)
    body, air = synthetic_pairs(B, S, 42)
    if args.workload == "noisybwe":
        body, air = noisy_batch_host(B, S, 42)
    step = O.OracleEBENStep(seed=42)
    t0 = time.perf_counter()
    step.step(body, air)                                  # warm-up (allocator, oneDNN primitives)
    first = time.perf_counter() - t0
    for _ in range(max(warmup, 1) - 1):
        if (time.perf_counter() - t0) + first * (steps + 1) > budget_s:
            break
        step.step(body, air)
    k = steps
    left = budget_s - (time.perf_counter() - t0)
    if first * k > left:
        k = max(1, int(left / first))
    t1 = time.perf_counter()
    for _ in range(k):
        step.step(body, air)
    dt = (time.perf_counter() - t1) / k
    L = S - (S + 32) % 256
    return {"value": B * L / SR / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "s_per_step": dt, "steps": k, "batch": B,
            "sample": f"{k} timed steps (+{max(warmup, 1)} warm-up) of the same train step at bs={B}x{args.seconds:g}s "
                      f"({args.workload}), fp32, torch CPU ({cores} threads)"}


def noisy_batch_host(B: int, S: int, seed: int):
    """Config 4 on the host (the CPU arm's input): the same draws and arithmetic as `NoisyFeeder` below - speech +
    noise[start:start+len] (no rescaling), then the joint crop - restated with slicing (vibravox/utils.py:195-254,50-81)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    Ls, Ln = S + S // 4, 4 * S
    air = (0.1 * torch.randn(B, 1, Ls, generator=g)).clamp(-1, 1)
    body = (0.1 * torch.randn(B, 1, Ls, generator=g)).clamp(-1, 1)
    noise = 0.05 * torch.randn(B, 1, Ln, generator=g)
    start = torch.randint(0, Ln - Ls, (B,), generator=g)
    off = torch.randint(0, Ls - S + 1, (B,), generator=g)
    ob = torch.stack([(body[i, :, :] + noise[i, :, start[i]:start[i] + Ls])[:, off[i]:off[i] + S] for i in range(B)])
    oa = torch.stack([air[i, :, off[i]:off[i] + S] for i in range(B)])
    return ob, oa


def gpu_eager_baseline(args, device, steps: int = 5, warmup: int = 2):
    """SURVEY 2.3 / 8(d) bar: the reference path's arithmetic as plain PyTorch eager modules on the SAME B200 through
    ATen / cuDNN (the oracle port moved to the device; the reference enables cudnn.benchmark, run.py:67-71), once with
    PyTorch's default TF32 convolutions (what the reference itself runs on a GPU) and once with
    fp32_precision='ieee'.  Same batch, same step.  A reported baseline - none of this repo's kernels run here."""
    import torch
    from oracle import eben_oracle as O
    B, S = args.batch, int(args.seconds * SR)
    body, air = synthetic_pairs(B, S, 42)
    if args.workload == "noisybwe":
        body, air = noisy_batch_host(B, S, 42)
    body, air = body.to(device), air.to(device)
    L = S - (S + 32) % 256
    out = {}
    prev_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ("tf32", "ieee"):
            try:
                torch.backends.cudnn.conv.fp32_precision = mode
                torch.backends.cuda.matmul.fp32_precision = mode
            except Exception:
                torch.backends.cudnn.allow_tf32 = mode == "tf32"
                torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
            step = O.OracleEBENStep(seed=42, device=device, host_logs=False)
            for _ in range(warmup):
                logs = step.step(body, air)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(steps):
                logs = step.step(body, air)
            e1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / steps
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"step_ms": ms, "wall_ms": wall * 1e3, "value": B * L / SR / (ms * 1e-3), "unit": "audio-s/s",
                         "loss": float(logs["generator/backprop_loss"])}
            del step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark = prev_bench
        try:
            torch.backends.cudnn.conv.fp32_precision = "tf32"
            torch.backends.cuda.matmul.fp32_precision = "ieee"
        except Exception:
            pass
    out["what"] = (f"oracle port of the reference modules + step, PyTorch eager on this GPU (ATen/cuDNN, cudnn.benchmark), "
                   f"bs={B}x{args.seconds:g}s, {warmup} warm-up + {steps} timed steps, CUDA events")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args, bounded=True)
    S = int(args.seconds * SR)
    L = S - (S + 32) % 256
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.cpu_steps, "warmup": 1, "ms_per_step": cb["s_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"EBEN BWE full train step, CPU port of the reference path, bounded sample "
                                   f"bs={args.cpu_batch}x{args.seconds:g}s@16kHz (L={L})",
                       "batch_per_gpu": args.cpu_batch, "samples": L, "parallelism": "host-cpu"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--cpu-batch", type=int, default=4)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="fp32 FMA kernels everywhere (VBX_TC=0)")
    ap.add_argument("--profile", action="store_true", help="1 warm-up + --steps steps only (for ncu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.no_tc:
        os.environ["VBX_TC"] = "0"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        _shutdown()


if __name__ == "__main__":
    main()
