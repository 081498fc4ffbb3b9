"""Minimal stand-in for lightning.Trainer on this path: calls `training_step` for `max_steps` batches,
one process per GPU (torchrun env), logs the device scalars the module recorded every `log_every_n_steps`."""
from __future__ import annotations

import time

import torch

from . import parallel


class Trainer:
    def __init__(self, max_steps: int = 100, log_every_n_steps: int = 10, use_cuda_graph: bool = True, **_ignored):
        self.max_steps, self.log_every, self.use_cuda_graph = max_steps, log_every_n_steps, use_cuda_graph

    def fit(self, module, datamodule):
        rank, local_rank, world = parallel.init_from_env("nccl")
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        module.to(dev)
        it = datamodule.batches(dev, rank)
        t0 = time.perf_counter()
        for step in range(1, self.max_steps + 1):
            if self.use_cuda_graph and hasattr(module, "training_step_graphed"):
                module.training_step_graphed(next(it))      # falls back to eager launches when not capturable
            else:
                module.training_step(next(it))
            if rank == 0 and step % self.log_every == 0:
                torch.cuda.synchronize()
                logs = {k.replace("train/", ""): round(float(v), 5) for k, v in module.logged.items()}
                print(f"step {step:6d}  {(time.perf_counter() - t0) / step * 1e3:7.1f} ms/step  {logs}", flush=True)
        torch.cuda.synchronize()
        return module
