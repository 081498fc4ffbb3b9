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


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(path)) if os.path.exists(path) else {}
    return (peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0),
            "measured" if peaks else "fallback")


def _timeit(fn, reps=10, warm=3):
    """Seconds per call of `fn`, GPU-side: the `reps` calls are captured in ONE CUDA graph and the replay is timed with
    CUDA events on the launching stream, so a 30-50 us kernel is not measured through the host's launch rate (the
    boxes of this pool differ by 30 % in Python / ctypes launch cost).  Falls back to plain launches if capture fails."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(reps):
                    fn()
        torch.cuda.current_stream().wait_stream(side)
        graph = g
    except Exception:
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph is not None:
        graph.replay()
        torch.cuda.synchronize()
        e0.record()
        graph.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def layer_microbench(device, B, L):
    """Per-kernel roofline evidence, measured live with CUDA events on the launching stream (every tensor set is
    larger than L2 or the kernel is timed over several distinct buffers, see `rot`):
    (a) the dominant kernel of the step = the dense MelGAN stage 4 implicit-GEMM conv (tensor bound),
    (b) the generator ResidualUnit - north_star's "fused EBEN Conv1d kernel", 60 % of HBM target - as ONE fused
        kernel (algorithmic bytes = read x once + write out once = 4*2*B*C*T) and, next to it, the two-kernel form,
    (c) the PQMF analysis / synthesis polyphase kernels (HBM bound)."""
    import torch
    from vibravox_b200 import ops
    hbm, tens, src = _peaks()
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
    t = _timeit(lambda: ops.conv_fwd(x, w, g, bias=bias, slope=0.2), reps=5, warm=2)
    kname = "tc_slab_kernel fwd (tcgen05, bf16x3)" if ops.use_tc(g, "fwd") else "gemm_conv_kernel<FWD> (fp32 FMA)"
    traffic = _traffic("melgan4_fwd_dram_bytes") if (B == 32 and ops.use_tc(g, "fwd")) else None
    out.append({"kernel": kname + " melgan.4 (1024->1024,k41,s4,g4)", "bound": "tensor",
                "achieved": flops / t / 1e12, "peak": tens, "unit": "TFLOP/s", "frac": flops / t / 1e12 / tens,
                "ms": t * 1e3, "algorithmic_bytes": byt, "peak_source": src + " bf16 sustained", "traffic": traffic,
                "note": "3 bf16 MMAs per fp32 product: the useful-flop ceiling is peak / 3"})
    del x, w
    # (b) generator residual units at the three widths of the network
    Tb = (L + 32) // 4
    for C, T in ((32, Tb), (64, Tb // 2), (128, Tb // 8)):
        nbuf = max(2, int(300e6 // (4 * B * C * T)) + 1)          # rotate over > 2 x L2 worth of inputs
        xs = [torch.randn(B, C, T, device=device) for _ in range(nbuf)]
        w1 = torch.randn(C, C, 3, device=device) * 0.1
        w2 = torch.randn(C, C, 1, device=device) * 0.1
        g1 = ops.ConvGeom(C, C, 3, 1, 3, 3, 3, 1)
        g2 = ops.ConvGeom(C, C, 1, 1, 1, 0, 0, 1)
        byt = 4.0 * 2 * B * C * T
        it = [0]

        def rot():
            it[0] = (it[0] + 1) % nbuf
            return xs[it[0]]
        if hasattr(ops, "residual_unit_fwd") and ops.use_fused_unit(C, T, 3):
            pk = ops.residual_unit_pack(w1, w2)
            tf = _timeit(lambda: ops.residual_unit_fwd(rot(), pk, 3, 0.01), reps=20)
            out.append({"kernel": f"fused ResidualUnit fwd (one kernel: dilated k3 d3 -> 1x1 -> LeakyReLU -> +x) C={C} T={T}",
                        "bound": "hbm", "achieved": byt / tf / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": byt / tf / 1e9 / hbm, "ms": tf * 1e3, "algorithmic_bytes": byt,
                        "peak_source": src + " copy", "traffic": _traffic(f"fused_unit_c{C}_dram_bytes")})
        t1 = _timeit(lambda: ops.conv_fwd(rot(), w1, g1))
        xg = xs[0]
        t2 = _timeit(lambda: ops.conv_fwd(rot(), w2, g2, res=xg, slope=0.01))
        for nm, tt, bb in (("dilated k3 d3", t1, byt), ("pointwise+lrelu+res", t2, byt * 1.5)):
            out.append({"kernel": f"conv fwd residual {nm} C={C} T={T} (two-kernel form)", "bound": "hbm",
                        "achieved": bb / tt / 1e9, "peak": hbm, "unit": "GB/s", "frac": bb / tt / 1e9 / hbm,
                        "ms": tt * 1e3, "algorithmic_bytes": bb, "peak_source": src + " copy", "traffic": None})
        del xs
    # (c) PQMF polyphase kernels (pqmf.py:194-213): analysis 4*B*(L + bands*T) bytes, synthesis+sum 4*B*(m*T + L)
    wa = torch.randn(4, 1, 32, device=device)
    sig = [torch.randn(B, 1, L, device=device) for _ in range(24)]
    bnd = [torch.randn(B, 4, Tb, device=device) for _ in range(12)]
    it = [0]
    for bands in (2, 4):
        def f():
            it[0] += 1
            return ops.pqmf_analysis(sig[it[0] % len(sig)], wa, bands)
        ta = _timeit(f, reps=24)
        bb = 4.0 * B * (L + bands * Tb)
        out.append({"kernel": f"pqmf_analysis_kernel bands={bands} L={L}", "bound": "hbm", "achieved": bb / ta / 1e9,
                    "peak": hbm, "unit": "GB/s", "frac": bb / ta / 1e9 / hbm, "ms": ta * 1e3, "algorithmic_bytes": bb,
                    "peak_source": src + " copy", "traffic": None})

    def fs():
        it[0] += 1
        return ops.pqmf_synthesis(bnd[it[0] % len(bnd)], wa, True)
    tsyn = _timeit(fs, reps=24)
    bb = 4.0 * B * (4 * Tb + L)
    out.append({"kernel": f"pqmf_synthesis_kernel (+ band sum) L={L}", "bound": "hbm", "achieved": bb / tsyn / 1e9,
                "peak": hbm, "unit": "GB/s", "frac": bb / tsyn / 1e9 / hbm, "ms": tsyn * 1e3, "algorithmic_bytes": bb,
                "peak_source": src + " copy", "traffic": None})
    return out


def _traffic(key):
    """DRAM bytes per launch from the committed `ncu --set full` captures (profiles/README.md)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            v = json.load(open(path)).get(key)
            if v is not None:
                return v
    return None


def workload_name(args, L):
    B = args.batch
    if args.workload == "noisybwe":
        return (f"NoisyBWE EBEN (config 4): on-the-fly noise mix + joint crop (speech + noise[start:start+len], no rescaling) "
                f"then the full train step with the MR-STFT loss, bs={B}x{args.seconds:g}s@16kHz per GPU (L={L})")
    return (f"EBEN BWE full train step (gen+disc+MR-STFT/FM/hinge, EMA balancing, 2x Adam) "
            f"bs={B}x{args.seconds:g}s@16kHz per GPU (L={L} after cut_to_valid_length), m=4 n=32 p=2 q=4 min_channels=24")


class Feeder:
    """The step's inputs.  bwe: one synthetic (body, air) pair per rank, resident on the device (value) or in pinned
    host memory (e2e).  noisybwe (BASELINE config 4): utterances longer than the crop + a 4x longer noise
    recording per item; every step draws `start ~ randint(0, len_noise - len_speech)` and the crop offset on the
    DEVICE generator (Philox) and mixes + crops in one vbx_noise_mix_crop launch (vibravox/utils.py:195-254,50-81)."""

    def __init__(self, args, dev, rank):
        import torch
        from vibravox_b200 import parallel
        self.args, self.dev, self.noisy = args, dev, args.workload == "noisybwe"
        B, S = args.batch, int(args.seconds * SR)
        self.B, self.S = B, S
        seed = parallel.rank_seed(42, rank)
        if not self.noisy:
            body, air = synthetic_pairs(B, S, seed)
            self.host = [body.pin_memory(), air.pin_memory()]
        else:
            g = torch.Generator().manual_seed(seed)
            self.Ls, self.Ln = S + S // 4, 4 * S
            air = (0.1 * torch.randn(B, 1, self.Ls, generator=g)).clamp(-1, 1)
            body = (0.1 * torch.randn(B, 1, self.Ls, generator=g)).clamp(-1, 1)
            noise = 0.05 * torch.randn(B, 1, self.Ln, generator=g)
            self.host = [body.pin_memory(), air.pin_memory(), noise.pin_memory()]
            self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.resident = [t.to(dev) for t in self.host]
        self.stage = [torch.empty_like(t) for t in self.resident]
        self.h2d_bytes = sum(t.numel() * 4 for t in self.host)
        self.launches_per_step = 1 if self.noisy else 0          # vbx_noise_mix_crop

    def _mix(self, body, air, noise):
        import torch
        from vibravox_b200 import ops
        start = torch.randint(0, self.Ln - self.Ls, (self.B,), generator=self.gen, device=self.dev, dtype=torch.int32)
        off = torch.randint(0, self.Ls - self.S + 1, (self.B,), generator=self.gen, device=self.dev, dtype=torch.int32)
        ob, oa = ops.noise_mix_crop(body, air, noise, start, off, self.S)
        return {"audio_body_conducted": ob, "audio_airborne": oa}

    def device_batch(self):
        if self.noisy:
            return self._mix(*self.resident)
        return {"audio_body_conducted": self.resident[0], "audio_airborne": self.resident[1]}

    def host_batch(self):
        """e2e: this step's raw inputs come from pinned host memory."""
        if self.noisy:
            for d, h in zip(self.stage, self.host):
                d.copy_(h, non_blocking=True)
            return self._mix(*self.stage)
        return {"audio_body_conducted": self.host[0], "audio_airborne": self.host[1]}


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
    if world > 1:
        for opt in lm.configure_optimizers():
            opt.broadcast_(0)
    L = S - (S + 32) % lm.generator.multiple
    feed = Feeder(args, dev, rank)

    graphed = (not args.no_graph) and (not args.profile) and lm.graph_capturable()
    run_step = lm.training_step_graphed if graphed else lm.training_step

    def step():
        # the public call: replays the captured step (after 2 eager calls + 1 capture, all inside the warm-up)
        run_step(feed.device_batch())

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
    # graph mode: the first 2 calls run eagerly and the 3rd captures - those 3 come BEFORE the --warmup replays
    n_pre = 3 if graphed else 0
    for _ in range(n_pre + max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    t_dev, _ = timed(step, args.steps)
    launches = _lib.launch_count() - n0
    if graphed:
        assert lm.graph_launches() > 0, "the step was not captured during the warm-up"
        launches = (lm.graph_launches() + feed.launches_per_step) * args.steps   # kernel nodes of the replayed graph
    clocks = sampler.stop() if rank == 0 else None
    audio_s = world * B * L / SR
    value = audio_s * args.steps / t_dev

    # e2e: the public call with HOST buffers; H2D of the step's inputs and D2H of its losses inside the timed region
    loss_h = torch.empty(2, dtype=torch.float32).pin_memory()

    def e2e_step():
        hb = feed.host_batch()
        if graphed or feed.noisy:
            run_step(hb)                                   # H2D straight into the captured step's input buffers
        else:
            lm.training_step({k: v.to(dev, non_blocking=True) for k, v in hb.items()})
        loss_h[0:1].copy_(lm.logged["train/generator/backprop_loss"].view(1), non_blocking=True)
        loss_h[1:2].copy_(lm.logged["train/discriminator/backprop_loss"].view(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    _, wall = timed(e2e_step, args.steps)
    wall = parallel.max_over_ranks(wall, dev)
    e2e = {"value": audio_s * args.steps / wall, "unit": "audio-s/s",
           "h2d_bytes_per_step": int(feed.h2d_bytes), "d2h_bytes_per_step": 8,
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
        "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload_name(args, L), "batch_per_gpu": B, "samples": L,
                   "parallelism": f"dp{world}", "schedule": lm.schedule,
                   "launch": ({"whole": "one CUDA graph replay per step",
                               "segments": "three CUDA graph replays per step, split at the two eager NCCL gradient "
                                           "all-reduces"}[lm.graph_mode()] if graphed else "eager launches"),
                   "pre_warmup": f"{n_pre} calls before the warm-up (2 eager + 1 CUDA-graph capture)" if n_pre else "none",
                   "l2": "per-step working set (activations ~ GBs) >> 126 MB L2; no explicit flush"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roof, "kernel_rooflines": kernels,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, steps=args.cpu_steps, warmup=1)
    if world == 1 and not args.no_eager_baseline:
        del lm
        torch.cuda.empty_cache()
        try:
            line["gpu_eager_baseline"] = gpu_eager_baseline(args, dev)
            line["gpu_eager_baseline"]["speedup_vs_tf32"] = line["gpu_eager_baseline"]["tf32"]["step_ms"] / line["ms_per_step"]
            line["gpu_eager_baseline"]["speedup_vs_ieee"] = line["gpu_eager_baseline"]["ieee"]["step_ms"] / line["ms_per_step"]
        except Exception as exc:                          # a baseline leg must never cost the product's number
            line["gpu_eager_baseline"] = {"unavailable": repr(exc)[:300]}
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
    fixed = synthetic_pairs(B, S, 42)
    draw = [0]

    def inputs():                                         # noisybwe: the mix + crop is part of every step, as on the GPU
        if args.workload != "noisybwe":
            return fixed
        draw[0] += 1
        return noisy_batch_host(B, S, 42, draw[0])
    step = O.OracleEBENStep(seed=42)
    t0 = time.perf_counter()
    step.step(*inputs())                                  # warm-up (allocator, oneDNN primitives)
    first = time.perf_counter() - t0
    done_warm = 1
    for _ in range(max(warmup, 1) - 1):
        if (time.perf_counter() - t0) + first * (steps + 1) > budget_s:
            break
        step.step(*inputs())
        done_warm += 1
    k = steps
    left = budget_s - (time.perf_counter() - t0)
    if first * k > left:
        k = max(1, int(left / first))
    t1 = time.perf_counter()
    for _ in range(k):
        step.step(*inputs())
    dt = (time.perf_counter() - t1) / k
    L = S - (S + 32) % 256
    return {"value": B * L / SR / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "s_per_step": dt, "steps": k, "warmup": done_warm, "batch": B,
            "sample": f"{k} timed steps (+{done_warm} warm-up) of the same train step at bs={B}x{args.seconds:g}s "
                      f"({args.workload}), fp32, torch CPU ({cores} threads)"}


_NOISY_RAW = {}


def noisy_batch_host(B: int, S: int, seed: int, draw: int = 0):
    """Config 4 on the host (the CPU arm's and the eager-GPU leg's input): the same raw material and arithmetic as
    `Feeder` - speech + noise[start:start+len] (no rescaling), then the joint crop - restated with slicing
    (vibravox/utils.py:195-254,50-81); `draw` selects the step's (start, offset) draws."""
    import torch
    if (B, S, seed) not in _NOISY_RAW:
        g = torch.Generator().manual_seed(seed)
        Ls, Ln = S + S // 4, 4 * S
        air = (0.1 * torch.randn(B, 1, Ls, generator=g)).clamp(-1, 1)
        body = (0.1 * torch.randn(B, 1, Ls, generator=g)).clamp(-1, 1)
        noise = 0.05 * torch.randn(B, 1, Ln, generator=g)
        _NOISY_RAW[(B, S, seed)] = (air, body, noise)
    air, body, noise = _NOISY_RAW[(B, S, seed)]
    Ls, Ln = body.shape[-1], noise.shape[-1]
    g = torch.Generator().manual_seed(seed * 1000003 + draw)
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
    """The reference arm: the reference path's own arithmetic (oracle port, kind 'port') on the host cores of the box,
    SAME workload / batch / length as the GPU arm, honouring --steps / --warmup unless the run would exceed
    --cpu-budget-s (then fewer timed steps are taken and reported)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args.cpu_batch = args.batch
    cb = cpu_baseline(args, steps=args.steps, warmup=args.warmup, budget_s=args.cpu_budget_s)
    S = int(args.seconds * SR)
    L = S - (S + 32) % 256
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": cb["steps"], "warmup": cb["warmup"], "ms_per_step": cb["s_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_name(args, L), "batch_per_gpu": args.batch, "samples": L,
                       "parallelism": f"host-cpu: 1 process, {cb['cores']} threads (no GPU)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def conv_sweep(args):
    """BASELINE.json configs[4]: dilated / strided Conv1d micro-bench, this repo's kernels vs cuDNN
    (torch.nn.functional.conv1d / aten::convolution_backward, fp32 'ieee' and PyTorch's default TF32, cudnn.benchmark
    on as the reference sets it), B=32, C_in=C_out=C, reflect halo, fwd / input gradient / weight gradient."""
    import torch
    import torch.nn.functional as F
    from vibravox_b200 import ops
    dev = "cuda"
    B = 32
    hbm, _, src = _peaks()
    shapes = [(3, 1, 1), (3, 3, 1), (3, 9, 1), (1, 1, 1), (7, 1, 1), (4, 1, 2), (8, 1, 4), (16, 1, 8)]   # (k, d, s)
    Cs = [int(c) for c in args.sweep_c.split(",")]
    Ls = [int(l) for l in args.sweep_l.split(",")]
    # (cudnn.benchmark stays off here: auto-tuning 200 shapes x 6 cuDNN problems costs more GPU time than the sweep itself)
    torch.backends.cudnn.benchmark = False
    t_start = time.perf_counter()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "conv_sweep.md")
    out_f = open(path, "w")
    skipped = 0
    rows, wins = [], {"fwd_ieee": 0, "fwd_tf32": 0, "dgrad_ieee": 0, "dgrad_tf32": 0, "wgrad_ieee": 0, "wgrad_tf32": 0}
    n = 0

    def mode(m):
        torch.backends.cudnn.conv.fp32_precision = m

    for C in Cs:
        for L in Ls:
            for (k, d, s) in shapes:
                if B * C * L > args.sweep_max_elems or time.perf_counter() - t_start > args.sweep_budget_s:
                    skipped += 1
                    continue
                pad = d * (k - 1) // 2 if s == 1 else s - 1
                g = ops.ConvGeom(C, C, k, s, d, pad, pad, 1)
                x = torch.randn(B, C, L, device=dev)
                w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
                xp = F.pad(x, (pad, pad), mode="reflect") if pad else x
                reps = 8 if B * C * L < 1e8 else 3
                y = ops.conv_fwd(x, w, g)
                dy = torch.randn_like(y)
                wt = ops.transpose_weight(w, 1)
                dw = torch.zeros_like(w)
                t = {"fwd": _timeit(lambda: ops.conv_fwd(x, w, g), reps),
                     "dgrad": _timeit(lambda: ops.conv_dgrad(dy, w, wt, g, L), reps),
                     "wgrad": _timeit(lambda: ops.conv_wgrad(x, dy, g, dw=dw), reps)}
                cb = torch.ops.aten.convolution_backward
                for m in ("ieee", "tf32"):
                    mode(m)
                    t["fwd_" + m] = _timeit(lambda: F.conv1d(F.pad(x, (pad, pad), mode="reflect") if pad else x, w, None, s, 0, d), reps)
                    t["dgrad_" + m] = _timeit(lambda: cb(dy, xp, w, None, [s], [0], [d], False, [0], 1, [True, False, False]), reps)
                    t["wgrad_" + m] = _timeit(lambda: cb(dy, xp, w, None, [s], [0], [d], False, [0], 1, [False, True, False]), reps)
                y_tf32 = F.conv1d(xp, w, None, s, 0, d)
                ref = F.conv1d(xp[:1].double(), w.double(), None, s, 0, d)
                e_ours = float((y[:1].double() - ref).abs().max() / ref.abs().max())
                e_tf32 = float((y_tf32[:1].double() - ref).abs().max() / ref.abs().max())
                To = y.shape[2]
                byts = 4.0 * B * (C * L + C * To) + 4.0 * C * C * k
                n += 1
                for op in ("fwd", "dgrad", "wgrad"):
                    for m in ("ieee", "tf32"):
                        wins[f"{op}_{m}"] += t[op] <= t[f"{op}_{m}"]
                rows.append(f"| {C} | {L} | {k},{d},{s} | {t['fwd']*1e3:.3f} | {byts / t['fwd'] / 1e9:.0f} ({100 * byts / t['fwd'] / 1e9 / hbm:.0f}%) | "
                            f"{t['fwd_ieee']*1e3:.3f} | {t['fwd_tf32']*1e3:.3f} | {t['dgrad']*1e3:.3f} | {t['dgrad_ieee']*1e3:.3f} | "
                            f"{t['dgrad_tf32']*1e3:.3f} | {t['wgrad']*1e3:.3f} | {t['wgrad_ieee']*1e3:.3f} | {t['wgrad_tf32']*1e3:.3f} | "
                            f"{e_ours:.1e} | {e_tf32:.1e} |")
                if len(rows) == 1:
                    out_f.write(_sweep_head(hbm, src))
                out_f.write(rows[-1] + "\n")
                out_f.flush()
                del x, w, y, dy, xp, y_tf32, wt
            torch.cuda.empty_cache()
    mode("tf32")
    out_f.close()
    emit({"sweep": "BASELINE config 5: Conv1d microbench vs cuDNN (cudnn.benchmark off)", "shapes": n, "skipped": skipped,
          "ours_at_least_as_fast_as": {k: f"{v}/{n}" for k, v in wins.items()}, "table": "gpurun_out/conv_sweep.md",
          "seconds": round(time.perf_counter() - t_start, 1)})


def _sweep_head(hbm, src):
    return ("B=32, C_in=C_out=C, reflect halo; times in ms (CUDA events over a captured graph, " + src +
            f" HBM copy peak {hbm:.0f} GB/s); cuDNN dgrad / wgrad = aten::convolution_backward on the pre-padded input "
            "(the reflect-pad backward is not charged to cuDNN); cudnn.benchmark off\n\n"
            "| C | L | k,d,s | ours fwd | GB/s (% of peak) | cuDNN ieee fwd | cuDNN tf32 fwd | ours dgrad | cuDNN ieee dgrad | "
            "cuDNN tf32 dgrad | ours wgrad | cuDNN ieee wgrad | cuDNN tf32 wgrad | err ours | err tf32 |\n" + "|" + "---|" * 15 + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bwe", choices=["bwe", "noisybwe"],
                    help="bwe = BASELINE configs[1]/[2] (bs=32x3s); noisybwe = configs[3] (bs=16x3s, noise mix in the loop)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--batch", type=int, default=None, help="per GPU; default 32 (bwe) / 16 (noisybwe)")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--cpu-batch", type=int, default=None, help="cpu_baseline batch (default: the GPU arm's)")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed steps of the in-line cpu_baseline leg")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="--impl reference: wall-clock budget")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="fp32 FMA kernels everywhere (VBX_TC=0)")
    ap.add_argument("--profile", action="store_true", help="1 warm-up + --steps steps only (for ncu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="BASELINE config 5: conv microbench sweep vs cuDNN")
    ap.add_argument("--sweep-c", default="32,64,128,256,512")
    ap.add_argument("--sweep-l", default="4096,8192,16384,32768,65536")
    ap.add_argument("--sweep-budget-s", type=float, default=420.0, help="stop adding shapes after this many seconds")
    ap.add_argument("--sweep-max-elems", type=float, default=2.7e8, help="skip shapes with more than B*C*L elements")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 16 if args.workload == "noisybwe" else 32
    if args.cpu_batch is None:
        args.cpu_batch = args.batch
    if args.no_tc:
        os.environ["VBX_TC"] = "0"
    if args.sweep:
        conv_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        _shutdown()


if __name__ == "__main__":
    main()
