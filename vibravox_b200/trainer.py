"""Minimal stand-in for lightning.Trainer on this path: calls `training_step` for `max_steps` batches, one process
per GPU (torchrun env), logs the device scalars the module recorded every `log_every_n_steps`, and writes / resumes
checkpoints in the layout of the reference's Lightning checkpoints (`configs/callbacks/bwe_checkpoint.yaml`:
`last.ckpt` with the module's `state_dict`, both optimizer states and the global step)."""
from __future__ import annotations

import os
import time
from typing import Optional

import torch

from . import parallel


class Trainer:
    def __init__(self, max_steps: int = 100, log_every_n_steps: int = 10, use_cuda_graph: bool = True,
                 default_root_dir: Optional[str] = None, save_every_n_steps: int = 0, accelerator: str = "gpu",
                 **_ignored):
        self.max_steps, self.log_every, self.use_cuda_graph = max_steps, log_every_n_steps, use_cuda_graph
        self.default_root_dir, self.save_every = default_root_dir, save_every_n_steps
        # "cpu" only places the tensors; the ops still refuse CPU tensors (tests drive it with tests/cpu_shim.py)
        self.accelerator = accelerator
        self.global_step = 0

    # ---- checkpoints ---------------------------------------------------------------------------
    def save_checkpoint(self, module, path: str) -> None:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp"
        torch.save(module.checkpoint(global_step=self.global_step), tmp)
        os.replace(tmp, path)                               # never leaves a half-written last.ckpt behind

    def _last(self) -> Optional[str]:
        return os.path.join(self.default_root_dir, "checkpoints", "last.ckpt") if self.default_root_dir else None

    # ---- the loop ------------------------------------------------------------------------------
    def fit(self, module, datamodule, ckpt_path: Optional[str] = None):
        if self.accelerator == "gpu":
            rank, local_rank, world = parallel.init_from_env("nccl")
            torch.cuda.set_device(local_rank)
            dev = torch.device("cuda", local_rank)
            sync = torch.cuda.synchronize
        else:
            rank, local_rank, world = parallel.env_world()
            dev, sync = torch.device("cpu"), (lambda: None)
        module.to(dev)
        if ckpt_path == "last":
            ckpt_path = self._last()
        if ckpt_path:
            # every rank reads the same file: parameters, Adam moments and the balancing EMA continue where they were
            self.global_step = module.load_checkpoint(torch.load(ckpt_path, map_location=dev, weights_only=False))
        if self.accelerator == "gpu" and world > 1:
            # what DDP does at wrap time: every rank starts from rank 0's parameters / optimizer state / balancing EMA
            for opt in module.configure_optimizers():
                if hasattr(opt, "broadcast_"):
                    opt.broadcast_(0)
            if getattr(module, "_bal", None) is not None:
                for t in module._bal.values():
                    torch.distributed.broadcast(t, src=0)
        it = datamodule.batches(dev, rank)
        for _ in range(self.global_step):                   # the synthetic stream is replayed up to the resume point
            next(it)
        graphed = self.use_cuda_graph and self.accelerator == "gpu" and hasattr(module, "training_step_graphed")
        t0, first = time.perf_counter(), self.global_step
        while self.global_step < self.max_steps:
            batch = next(it)
            if graphed:
                module.training_step_graphed(batch)         # falls back to eager launches when not capturable
            else:
                module.training_step(batch)
            self.global_step += 1
            step = self.global_step
            if rank == 0 and step % self.log_every == 0:
                sync()
                logs = {k.replace("train/", ""): round(float(v), 5) for k, v in module.logged.items()}
                ms = (time.perf_counter() - t0) / (step - first) * 1e3
                print(f"step {step:6d}  {ms:7.1f} ms/step  {logs}", flush=True)
            if rank == 0 and self.save_every and self._last() and step % self.save_every == 0:
                sync()
                self.save_checkpoint(module, self._last())
        sync()
        if rank == 0 and self._last():
            self.save_checkpoint(module, self._last())
        return module
