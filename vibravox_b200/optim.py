"""FlatAdam: torch.optim.Adam semantics (configs/lightning_module/optimizer/adam.yaml:1-9 of the
reference) on ONE flat fp32 bucket per network.

On first use every parameter is re-pointed at a slice of `flat` (values preserved), and gets
a `_vbx_grad` view into `grad` that the backward kernels (weight-norm backward, wgrad, bias
reduction) accumulate into directly.  `step()` is then two launches (tick + fused Adam) over
the whole network, `zero_grad()` one fill, and data-parallel training all-reduces `grad`
once per network.  Use as `_target_: vibravox_b200.optim.FlatAdam` with
`_partial_: true`, exactly where the reference config has torch.optim.Adam.
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from . import ops


class FlatAdam(torch.optim.Optimizer):
    TAIL = 16        # floats appended to the gradient bucket (see materialize)

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, amsgrad: bool = False):
        if weight_decay != 0.0 or amsgrad:
            raise NotImplementedError("FlatAdam implements the reference configuration: weight_decay=0, amsgrad=False")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps)
        # Every parameter stays in param_groups[0]["params"] - frozen ones too (generator.parameters() starts with the
        # two requires_grad=False PQMF banks, pqmf.py:51-56) - so that state indices are those torch.optim.Adam uses
        # over the same iterable and optimizer states travel both ways.  Only the trainable ones get a slice of the
        # flat buckets (`params`); frozen ones never receive a gradient, exactly as Adam skips `grad is None`.
        super().__init__(list(params), defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError("FlatAdam supports a single parameter group")
        self._trainable_idx = [i for i, p in enumerate(self.param_groups[0]["params"]) if p.requires_grad]
        self._params = [self.param_groups[0]["params"][i] for i in self._trainable_idx]
        self._indirect: List[int] = []      # ids of params whose grads arrive through autograd (.grad)
        self.flat = self.grad = self.bucket = self.tail = self.exp_avg = self.exp_avg_sq = self.step_count = None
        self.grad_scale = 1.0               # set to 1/world_size after a sum all-reduce of `grad`

    # ---- layout ---------------------------------------------------------------------------------
    def keep_autograd_grad(self, params: Iterable[torch.Tensor]) -> None:
        """Parameters that must keep ordinary `.grad` semantics (targets of torch.autograd.grad,
        e.g. generator.last_conv.weight in dynamically_balance_losses, eben.py:223-228)."""
        assert self.flat is None, "call before the first step / materialize()"
        self._indirect += [id(p) for p in params]

    @property
    def params(self) -> List[torch.Tensor]:
        """The trainable parameters, in order: the ones that own a slice of the flat buckets."""
        return self._params

    @property
    def all_params(self) -> List[torch.Tensor]:
        return self.param_groups[0]["params"]

    def materialize(self) -> None:
        if self.flat is not None:
            return
        ps = self.params
        dev = ps[0].device
        ops.require_cuda(dev)
        if any(p.device != dev for p in ps):
            raise RuntimeError("FlatAdam needs all parameters on one device")
        offs, total = [], 0
        for p in ps:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # keep every slice 16-byte aligned
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        # the gradient bucket carries TAIL extra floats behind the gradients: the step's logged scalars ride in the
        # same all-reduce (lightning_modules/eben.py: sync_dist semantics without extra collectives)
        self.bucket = torch.zeros(total + self.TAIL, device=dev, dtype=torch.float32)
        self.grad, self.tail = self.bucket[:total], self.bucket[total:]
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self._slices = []
        with torch.no_grad():
            for p, off in zip(ps, offs):
                n = p.numel()
                self.flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)
                slot = self.grad[off:off + n].view(p.shape)
                self._slices.append(slot)
                if id(p) not in self._indirect:
                    p._vbx_grad = slot
                p.grad = None

    # ---- torch.optim API ------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True) -> None:
        self.materialize()
        ops.fill(self.bucket, 0.0)
        for p in self.params:
            p.grad = None

    def gather_autograd_grads(self) -> None:
        """Fold `.grad` of the keep_autograd_grad parameters into the bucket (before an all-reduce)."""
        for p, slot in zip(self.params, self._slices):
            if p.grad is not None:
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                ops.axpby(g, slot, 1.0, 1.0)
                p.grad = None

    def broadcast_(self, src: int = 0) -> None:
        """Data-parallel start-up (what DDP does when it wraps a module): every rank takes rank `src`'s parameters,
        moments and step count, so ranks that were seeded differently cannot silently diverge."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        self.materialize()
        for t in (self.flat, self.exp_avg, self.exp_avg_sq, self.step_count):
            dist.broadcast(t, src=src)

    # ---- checkpoints: the layout torch.optim.Adam writes (what a Lightning checkpoint of the reference holds) ----
    def state_dict(self) -> dict:
        """`torch.optim.Adam.state_dict()` layout - per-parameter `step` / `exp_avg` / `exp_avg_sq` indexed in
        parameter order - so optimizer states travel both ways between this class and the reference's Adam
        (`configs/lightning_module/optimizer/adam.yaml`).  Copies, not views of the flat buckets."""
        g = self.param_groups[0]
        group = {k: v for k, v in g.items() if k != "params"}
        group.update(weight_decay=0.0, amsgrad=False, params=list(range(len(self.all_params))))
        state = {}
        if self.flat is not None and int(self.step_count[0]) > 0:
            step = self.step_count[0].to(torch.float32)
            off = 0
            for i, p in zip(self._trainable_idx, self.params):      # indices in the FULL parameter order (as Adam)
                n = p.numel()
                state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
                off += (n + 3) // 4 * 4
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, state_dict: dict) -> None:
        """Accepts its own `state_dict()` or one written by `torch.optim.Adam` over the same parameters (in the same
        order).  Moments are copied INTO the flat buckets, so a captured CUDA graph keeps pointing at live memory."""
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.all_params):
            raise ValueError("FlatAdam.load_state_dict: expected one parameter group over the same parameters")
        if groups[0].get("amsgrad", False) or groups[0].get("weight_decay", 0.0) != 0.0:
            raise NotImplementedError("FlatAdam implements the reference configuration: weight_decay=0, amsgrad=False")
        g = self.param_groups[0]
        for k in ("lr", "betas", "eps"):
            if k in groups[0]:
                g[k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        self.materialize()
        state = state_dict["state"]
        steps = {int(float(v["step"])) for v in state.values()}
        if len(steps) > 1:
            raise ValueError("FlatAdam keeps one step count for all parameters; the checkpoint has several")
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count.fill_(steps.pop() if steps else 0)
        stray = [k for k in state if int(k) not in self._trainable_idx]
        if stray:
            raise ValueError(f"FlatAdam.load_state_dict: the checkpoint holds state for frozen parameters {stray[:4]}")
        off = 0
        for i, p in zip(self._trainable_idx, self.params):
            n = p.numel()
            st = state.get(i, state.get(str(i)))
            if st is not None:
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"FlatAdam.load_state_dict: parameter {i} has shape {tuple(p.shape)}, "
                                     f"the checkpoint {tuple(st['exp_avg'].shape)}")
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            off += (n + 3) // 4 * 4

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        self.materialize()
        self.gather_autograd_grads()
        g = self.param_groups[0]
        ops.adam_tick(self.step_count)
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, g["lr"],
                      g["betas"][0], g["betas"][1], g["eps"], self.grad_scale)
        # the kernel writes the parameters behind autograd's back (no version bump): drop what was derived from
        # the old values (packed tensor-core tiles of convs that use their weight directly)
        for p in self.params:
            p.__dict__.pop("_vbx_packs", None)
        return None
