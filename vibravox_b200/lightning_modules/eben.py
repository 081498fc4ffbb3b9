"""EBENLightningModule without Lightning (reference: vibravox/lightning_modules/eben.py:9-240).

Same constructor arguments and the same `training_step(batch)` schedule - generator phase
(G forward, MR-STFT / feature-matching / hinge losses, gradient-norm loss balancing on
`generator.last_conv.weight`, backward, optimizer step) then discriminator phase on the
detached pre-update generator outputs - with Lightning's services restated per SURVEY App. B:
toggle_optimizer = requires_grad flips, manual_backward = .backward() (+ one all-reduce of
the flat gradient bucket per network when torch.distributed is initialised), self.log =
device scalars kept in `self.logged` (no host sync inside the step).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch

from .. import ops
from ..functional import Flags, WeightedSumFn
from ..optim import FlatAdam
from ..parallel import allreduce_sum_
from ..torch_modules.utils import share_weight_norm

# fused backward of the discriminator chains (functional.Flags); VBX_CHAIN_FUSION=0 restores the separate aten::add /
# L1-pair backward / LeakyReLU backward passes between the conv stages
_CHAIN_FUSION = __import__("os").environ.get("VBX_CHAIN_FUSION", "1") != "0"
# discriminator-phase backward enqueued on a forked stream before the generator's backward (independent work)
_PHASE_OVERLAP = __import__("os").environ.get("VBX_PHASE_OVERLAP", "1") != "0"
# D(reference) forward started before the generator's forward (independent work)
_EARLY_REFERENCE = __import__("os").environ.get("VBX_EARLY_REFERENCE", "1") != "0"
# Priority of the captured step's main stream (generator chain, losses, optimizers) relative to the discriminator side
# streams.  High (-1) by default: the generator's backward is a long chain of SMALL kernels that runs beside the
# discriminator phase; at equal priority its kernels queue behind the discriminators' large grids, the chain finishes
# ~3.5 ms after everything else and that tail runs alone on a mostly idle GPU (profiles/r2_step_timeline.txt).  With
# priority its kernels slot in as SMs free up and both chains end together: measured 35.6 -> 33.9 ms per step.
_MAIN_PRIORITY = int(__import__("os").environ.get("VBX_MAIN_PRIORITY", "-1"))


class _SegmentedStep:
    """A training step captured as consecutive CUDA graphs; after graph i the i-th gradient bucket is sum-all-reduced
    with an ordinary (eager) collective, so NCCL calls are issued in program order on every rank."""

    def __init__(self, graphs, buckets):
        self.graphs, self.buckets = graphs, buckets

    def replay(self) -> None:
        for i, g in enumerate(self.graphs):
            g.replay()
            if i < len(self.buckets):
                allreduce_sum_(self.buckets[i])        # gradients + the logged scalars packed behind them


class EBENLightningModule(torch.nn.Module):
    def __init__(self, sample_rate: int, generator: torch.nn.Module, discriminator: torch.nn.Module,
                 generator_optimizer, discriminator_optimizer,
                 reconstructive_loss_freq_fn: Optional[torch.nn.Module] = None,
                 reconstructive_loss_time_fn: Optional[torch.nn.Module] = None,
                 feature_matching_loss_fn: Optional[torch.nn.Module] = None,
                 adversarial_loss_fn: Optional[torch.nn.Module] = None,
                 dynamic_loss_balancing: Optional[str] = None, beta_ema: float = 0.9,
                 update_discriminator_ratio: float = 1.0, description: Optional[str] = None,
                 push_to_hub_after_testing: bool = False, schedule: str = "shared"):
        super().__init__()
        self.sample_rate, self.description = sample_rate, description
        self.generator, self.discriminator = generator, discriminator
        self.generator_optimizer = generator_optimizer(params=self.generator.parameters())
        self.discriminator_optimizer = discriminator_optimizer(params=self.discriminator.parameters())
        if isinstance(self.generator_optimizer, FlatAdam):
            self.generator_optimizer.keep_autograd_grad([self.generator.last_conv.weight])
        self.reconstructive_loss_temp_fn = reconstructive_loss_time_fn
        self.reconstructive_loss_freq_fn = reconstructive_loss_freq_fn
        self.feature_matching_loss_fn = feature_matching_loss_fn
        self.adversarial_loss_fn = adversarial_loss_fn
        assert dynamic_loss_balancing in {None, "simple", "ema"}, \
            "dynamic_loss_balancing must be in {None, 'simple', 'ema'}"
        self.dynamic_loss_balancing = dynamic_loss_balancing
        self.beta_ema = beta_ema
        assert 0 <= update_discriminator_ratio <= 1, "update_discriminator_ratio must be in [0, 1]"
        self.update_discriminator_ratio = update_discriminator_ratio
        if push_to_hub_after_testing:
            # eben.py:177-182 of the reference pushes the generator in on_test_end; there is no network on this path
            raise NotImplementedError("push_to_hub_after_testing=True is not supported: call "
                                      "generator.push_to_hub(...) (PyTorchModelHubMixin) yourself after testing")
        self.push_to_hub_after_testing = push_to_hub_after_testing
        self.automatic_optimization = False
        assert schedule in {"shared", "reference"}
        # "reference": the literal op sequence of eben.py:82-130 (4 D forwards, 3 D input-gradient passes).
        # "shared": same losses / gradients / updates with the algebraically redundant work removed
        #           (SURVEY 7.3-6): one D graph serves both phases, loss gradients are taken once at the
        #           generator outputs, combined with the lambdas, and the generator is back-propagated once.
        self.schedule = schedule
        # balancing state (eben.py:73,230-235) lives on the device: EMA of the gradient norms
        self._bal = None
        self.logged: Dict[str, torch.Tensor] = {}
        self._graphs: Dict[tuple, dict] = {}      # training_step_graphed: one captured step per batch shape
        self._capture = None                      # state of a segmented capture in progress (see _capture_segments)
        self.max_graph_shapes = 4                 # captured steps kept (LRU by batch shape); more shapes run eagerly
        self.dataloader_names = None              # base_se.py:52: names of the validation / test dataloaders, if several
        self.graph_warmup_steps = 2

    # ---- Lightning services, restated --------------------------------------------------------
    def optimizers(self, use_pl_optimizer: bool = True):
        return self.generator_optimizer, self.discriminator_optimizer

    def configure_optimizers(self):
        return [self.generator_optimizer, self.discriminator_optimizer]

    def log(self, name: str, value: torch.Tensor, **_) -> None:
        self.logged[name] = value.detach() if isinstance(value, torch.Tensor) else value
        self.__dict__.setdefault("_synced", set()).discard(name)      # (re-)logged: not yet rank-averaged

    def toggle_optimizer(self, optimizer) -> None:
        mine = {id(p) for g in optimizer.param_groups for p in g["params"]}
        self._toggled = []
        for opt in self.configure_optimizers():
            for g in opt.param_groups:
                for p in g["params"]:
                    if id(p) not in mine and p.requires_grad:
                        p.requires_grad = False
                        self._toggled.append(p)

    def untoggle_optimizer(self, optimizer) -> None:
        for p in self._toggled:
            p.requires_grad = True
        self._toggled = []

    @staticmethod
    def _join(module) -> None:
        if hasattr(module, "join_streams"):
            module.join_streams()

    def manual_backward(self, loss: torch.Tensor, optimizer) -> None:
        loss.backward()
        self._join(self.discriminator)
        self._sync_grads(optimizer)

    @property
    def atomic_norms_old(self):
        return None if self._bal is None else self._bal["old"]

    # ---- the hot path ------------------------------------------------------------------------
    def training_step(self, batch: Dict[str, torch.Tensor]):
        ok = (self.schedule == "shared" and self.feature_matching_loss_fn is not None
              and self.adversarial_loss_fn is not None and self.reconstructive_loss_temp_fn is None
              and self.dynamic_loss_balancing is not None)
        return self._training_step_shared(batch) if ok else self._training_step_reference(batch)

    def graph_mode(self) -> str:
        """How `training_step_graphed` launches the step: 'whole' (one CUDA graph), 'segments' (three graphs split
        at the two gradient all-reduces, which stay ordinary eager NCCL calls between the replays) or 'eager'.

        world_size == 1 replays the whole step.  world_size > 1 uses segments: NCCL inside a captured multi-stream
        step measured 98 % scaling on 2 GPUs but one verification run hung (opt-in: VBX_GRAPH_DDP=1), while eager
        launches leave the step host-bound (82 %).  Segments keep every collective out of the graphs - issued from
        the host in program order on all ranks - and still cost only three cudaGraphLaunch calls per step.
        VBX_GRAPH_SEGMENTS=1 forces segments on one GPU (tests), =0 forces eager launches for world_size > 1."""
        import os
        flat = all(isinstance(o, FlatAdam) for o in self.configure_optimizers())
        if not flat or self.update_discriminator_ratio < 1 or getattr(self, "_graph_failed", False):
            return "eager"                   # per-step host decisions / foreign optimizers cannot be captured
        multi = (torch.distributed.is_available() and torch.distributed.is_initialized()
                 and torch.distributed.get_world_size() > 1)
        seg = os.environ.get("VBX_GRAPH_SEGMENTS")
        if not multi:
            return "segments" if seg == "1" else "whole"
        if torch.distributed.get_backend() != "nccl":
            return "eager"
        if os.environ.get("VBX_GRAPH_DDP", "0") == "1":
            return "whole"
        return "eager" if seg == "0" else "segments"

    def graph_capturable(self) -> bool:
        """The step replays from CUDA graphs when nothing in it is decided on the host per step."""
        return self.graph_mode() != "eager"

    def training_step_graphed(self, batch: Dict[str, torch.Tensor]):
        """`training_step` replayed from CUDA graphs (all streams of the step included): the ~1300 kernel launches
        of a step cost one cudaGraphLaunch (three when split at the gradient all-reduces, see `graph_mode`), so the
        step no longer depends on how fast the host can issue them.
        Every call performs exactly one training step: the first `graph_warmup_steps` calls for a batch shape run
        eagerly (they fill the weight / filter caches that outlive a step), the next call captures the step and
        replays it, later calls copy the batch into the captured input buffers and replay.  `batch` tensors may
        live on the host (pinned) or the device.  The returned tensors and `self.logged` are the graph's static
        outputs: valid until the next call.  Learning rates are baked in at capture time."""
        mode = self.graph_mode()
        if mode == "eager":
            dev = self.generator.last_conv.weight.device
            return self.training_step({k: v.to(dev, non_blocking=True) for k, v in batch.items()
                                       if isinstance(v, torch.Tensor)})
        names = ("audio_body_conducted", "audio_airborne")
        key = tuple(tuple(batch[n].shape) for n in names)
        st = self._graphs.get(key)
        if st is not None:
            self._graphs[key] = self._graphs.pop(key)          # most recently used last
        if st is None:
            dev = self.generator.last_conv.weight.device
            if len(self._graphs) >= self.max_graph_shapes:
                # 'pad' collation yields a new length per batch: a captured graph (and its private memory pool) per
                # shape would grow without bound.  Keep the most recent shapes, run a shape seen for the first time
                # while the cache is full eagerly, and evict the least recently used capture.
                self._graphs.pop(next(iter(self._graphs)))
                return self.training_step({k: v.to(dev, non_blocking=True) for k, v in batch.items()
                                           if isinstance(v, torch.Tensor)})
            st = self._graphs[key] = dict(calls=0, graph=None, out=None, logged=None, launches=0,
                                          stream=torch.cuda.Stream(dev, priority=_MAIN_PRIORITY),
                                          inputs={n: torch.empty(batch[n].shape, device=dev, dtype=torch.float32)
                                                  for n in names})
        for n in names:
            st["inputs"][n].copy_(batch[n], non_blocking=True)
        if st["graph"] is None:
            st["calls"] += 1
            side = st["stream"]
            if st["calls"] <= self.graph_warmup_steps:
                # eager, but on the stream the capture will use: autograd's cached gradient-accumulation nodes
                # remember the stream they were created on, and one created on the legacy default stream
                # cannot take part in a capture
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    out = self.training_step(st["inputs"])
                torch.cuda.current_stream().wait_stream(side)
                return out
            from .. import _lib
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            multi = torch.distributed.is_available() and torch.distributed.is_initialized()
            # (thread_local: NCCL's watchdog thread may poll its events while this thread captures)
            error_mode = "thread_local" if multi else "global"
            try:
                if mode == "whole":
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side, capture_error_mode=error_mode):
                        st["out"] = self.training_step(st["inputs"])
                else:
                    graph = self._capture_segments(st, side, error_mode)
            except Exception as exc:           # nothing of the captured step has run: do it eagerly, stay eager
                import warnings
                warnings.warn(f"CUDA-graph capture of the training step failed ({exc}); using eager launches")
                self._graph_failed = True
                torch.cuda.synchronize()
                return self.training_step(st["inputs"])
            st["launches"] = _lib.launch_count() - n0
            st["graph"], st["logged"] = graph, dict(self.logged)
        self.logged = st["logged"]
        st["graph"].replay()
        return st["out"]

    def _capture_segments(self, st, side, error_mode):
        """Capture the step as consecutive graphs sharing one memory pool; `_sync_grads` closes a segment and opens
        the next one wherever the eager step all-reduces a gradient bucket (all side streams are joined there)."""
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        cap = self._capture = dict(pool=torch.cuda.graph_pool_handle(), graphs=[], buckets=[], error_mode=error_mode,
                                   open=False)
        side.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(side):
                self._segment_begin()
                st["out"] = self.training_step(st["inputs"])
                self._segment_end()
        except BaseException:
            if cap["open"]:
                try:
                    self._segment_end()
                except Exception:
                    pass
            raise
        finally:
            self._capture = None
        torch.cuda.current_stream().wait_stream(side)
        return _SegmentedStep(cap["graphs"], cap["buckets"])

    def _segment_begin(self) -> None:
        cap = self._capture
        g = torch.cuda.CUDAGraph()
        g.capture_begin(pool=cap["pool"], capture_error_mode=cap["error_mode"])
        cap["graphs"].append(g)
        cap["open"] = True

    def _segment_end(self) -> None:
        cap = self._capture
        cap["open"] = False
        cap["graphs"][-1].capture_end()

    def graph_launches(self) -> int:
        """Kernels of this library inside one captured step (0 before capture)."""
        return max([st["launches"] for st in self._graphs.values()], default=0)

    def _training_step_shared(self, batch: Dict[str, torch.Tensor]):
        """eben.py:82-130 with the redundant traversals removed.  Equalities used:
        (a) D(enhanced) and D(reference) of the discriminator phase equal those of the generator phase
            (only G was updated in between, and the phase uses the pre-update, detached G outputs), so the
            two D forwards are done once, with a graph, and serve both phases;
        (b) d(sum_i lambda_i L_i) = sum_i lambda_i dL_i with detached lambdas, so each loss is differentiated
            once down to the generator outputs; the norms at last_conv.weight come from pushing those three
            gradients through G's short tail (synthesis, tanh, last_conv), and G is back-propagated once."""
        G, D = self.generator, self.discriminator
        corrupted_speech = G.cut_to_valid_length(batch["audio_body_conducted"])
        reference_speech = G.cut_to_valid_length(batch["audio_airborne"])
        g_opt, d_opt = self.optimizers(use_pl_optimizer=True)
        for opt in (g_opt, d_opt):
            if isinstance(opt, FlatAdam):
                opt.materialize()
        d_params = [p for grp in d_opt.param_groups for p in grp["params"]]
        with share_weight_norm():
            # D(reference) does not depend on the generator: on a GPU it is started first, on the discriminator's side
            # streams, and the generator's forward - a serial chain of small kernels - runs underneath it
            early = _EARLY_REFERENCE and hasattr(D, "forward_multi") and reference_speech.is_cuda
            reference_bands = G.pqmf.forward(reference_speech, "analysis")
            reference_embeddings = None
            if early:
                Flags.gated_chain = _CHAIN_FUSION
                try:
                    (reference_embeddings,) = D.forward_multi([(reference_bands, reference_speech)], join=False)
                finally:
                    Flags.gated_chain = False
            enhanced, enhanced_bands = G(corrupted_speech)
            enh = enhanced.detach().requires_grad_(True)
            bands = enhanced_bands.detach().requires_grad_(True)
            Flags.param_grads = False                      # generator phase: no D parameter gradients
            try:
                losses = OrderedDict()
                if self.reconstructive_loss_freq_fn:
                    losses["reconstructive_loss_freq"] = self.reconstructive_loss_freq_fn(enh, reference_speech)
                # the stage outputs below are consumed by the next stage, the feature-matching loss and (last one) the hinge
                # loss only: the chain contract of functional.Flags, which lets every input-gradient kernel finish the
                # LeakyReLU / feature-matching backward of the stage before it
                Flags.gated_chain = _CHAIN_FUSION
                try:
                    if early:
                        (enhanced_embeddings,) = D.forward_multi([(bands, enh)])        # (joins the reference pass too)
                    elif hasattr(D, "forward_multi"):
                        enhanced_embeddings, reference_embeddings = D.forward_multi(
                            [(bands, enh), (reference_bands, reference_speech)])
                    else:
                        enhanced_embeddings = D(bands=bands, audio=enh)
                        reference_embeddings = D(bands=reference_bands, audio=reference_speech)
                finally:
                    Flags.gated_chain = False
                losses["feature_matching_loss"] = self.feature_matching_loss_fn(enhanced_embeddings,
                                                                                reference_embeddings)
                losses["adv_loss_gen"] = self.adversarial_loss_fn(embeddings=enhanced_embeddings, target=1)
                for key, value in losses.items():
                    self.log(f"train/generator/{key}", value, sync_dist=True)
                # each loss once, down to the generator outputs
                grads = []
                for l in losses.values():
                    Flags.chain_reset()
                    grads.append(torch.autograd.grad(l, (enh, bands), retain_graph=True, allow_unused=True))
                    Flags.chain_check()
                self._join(D)
                lambdas = self._balance_from_output_grads(enhanced, enhanced_bands, grads)
                total_e = torch.empty_like(enh)
                total_b = torch.empty_like(bands)
                first_e = first_b = True
                for i, (ge, gb) in enumerate(grads):
                    if ge is not None:
                        ops.axpby_dev(ge.contiguous(), total_e, lambdas[i:i + 1], 0.0 if first_e else 1.0)
                        first_e = False
                    if gb is not None:
                        ops.axpby_dev(gb.contiguous(), total_b, lambdas[i:i + 1], 0.0 if first_b else 1.0)
                        first_b = False
                with torch.no_grad():
                    backprop_loss_generator = WeightedSumFn.apply(lambdas, *[l.detach() for l in losses.values()])
                self.log("train/generator/backprop_loss", backprop_loss_generator, sync_dist=True)
            finally:
                Flags.param_grads = True
            # The discriminator phase reads only what exists by now (the detached generator outputs, the stored D
            # activations, D's parameters), and the generator's backward / update touches none of that: the two are
            # independent, so the D-phase backward is ENQUEUED FIRST, on a forked stream, and the generator's backward - a
            # serial chain of small kernels that leaves most SMs idle - runs underneath it.  Same arithmetic, same results
            # (the reference runs them back to back, eben.py:107-130); the coin flip of update_discriminator_ratio keeps its
            # place in the RNG sequence (nothing between here and the reference's draw consumes random numbers).
            update = True
            if self.update_discriminator_ratio < 1:
                update = bool(torch.rand(1) < self.update_discriminator_ratio)
            overlap = update and _PHASE_OVERLAP and enh.is_cuda
            if overlap:
                cur = torch.cuda.current_stream()
                side = self._phase_stream(enh.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    self._discriminator_backward(D, reference_embeddings, enhanced_embeddings, d_params)
            torch.autograd.backward((enhanced, enhanced_bands), (total_e, total_b))
            if overlap:
                cur.wait_stream(side)
            self._sync_grads(g_opt)
            g_opt.step()
            g_opt.zero_grad()
            if update:
                if not overlap:
                    self._discriminator_backward(D, reference_embeddings, enhanced_embeddings, d_params)
                self._sync_grads(d_opt)
                d_opt.step()
                d_opt.zero_grad()
        return {"corrupted": corrupted_speech, "enhanced": enhanced.detach(), "reference": reference_speech}

    def _discriminator_backward(self, D, reference_embeddings, enhanced_embeddings, d_params) -> None:
        """Discriminator phase on the graph of the generator phase (eben.py:116-124): the two hinge losses and D's
        parameter gradients, accumulated into the discriminator's flat bucket; joined onto the current stream."""
        real_loss = self.adversarial_loss_fn(embeddings=reference_embeddings, target=1)
        fake_loss = self.adversarial_loss_fn(embeddings=enhanced_embeddings, target=-1)
        self.log("train/discriminator/real_loss", real_loss, sync_dist=True)
        self.log("train/discriminator/fake_loss", fake_loss, sync_dist=True)
        backprop_loss_discriminator = WeightedSumFn.apply(None, real_loss, fake_loss)
        self.log("train/discriminator/backprop_loss", backprop_loss_discriminator, sync_dist=True)
        Flags.skip_leaf_input_grad = True          # nothing upstream of the detached G outputs
        Flags.chain_reset()
        try:
            torch.autograd.backward(backprop_loss_discriminator, inputs=d_params)
        finally:
            Flags.skip_leaf_input_grad = False
        Flags.chain_check()
        self._join(D)

    def _phase_stream(self, device):
        cache = self.__dict__.setdefault("_phase_streams", {})
        if device not in cache:
            cache[device] = torch.cuda.Stream(device=device)
        return cache[device]

        return {"corrupted": corrupted_speech, "enhanced": enhanced.detach(), "reference": reference_speech}

    def _balance_from_output_grads(self, enhanced, enhanced_bands, grads) -> torch.Tensor:
        """dynamically_balance_losses (eben.py:222-240) from the loss gradients at the generator outputs."""
        layer = self.generator.last_conv.weight
        n, dev = len(grads), layer.device
        if self._bal is None or self._bal["old"].numel() != n:
            self._bal = dict(old=torch.zeros(n, device=dev), init=torch.zeros(1, device=dev, dtype=torch.int32))
        sumsq = torch.zeros(n, device=dev, dtype=torch.float64)
        for i, (ge, gb) in enumerate(grads):
            outs, gos = [], []
            if ge is not None:
                outs.append(enhanced); gos.append(ge)
            if gb is not None:
                outs.append(enhanced_bands); gos.append(gb)
            saved, Flags.param_grads = Flags.param_grads, True      # G's own tail: last_conv.weight is the target
            try:
                gw = torch.autograd.grad(outs, layer, grad_outputs=gos, retain_graph=True)[0]
            finally:
                Flags.param_grads = saved
            ops.sumsq(gw.contiguous(), sumsq[i:i + 1])
        lambdas = torch.empty(n, device=dev)
        norms = torch.empty(n, device=dev)
        ops.balance(sumsq, self._bal["old"], self._bal["init"], lambdas, norms, self.beta_ema,
                    1 if self.dynamic_loss_balancing == "ema" else 0)
        self.last_norms, self.last_lambdas = norms, lambdas
        return lambdas

    def _sync_grads(self, optimizer) -> None:
        """Gradient exchange of one network (DDP's all-reduce, mean semantics) - and, in the same collective, the
        scalars logged since the previous exchange: the reference logs every loss with `sync_dist=True`
        (eben.py:103-124), i.e. the rank MEAN; here they are packed behind the gradients in the flat bucket."""
        if getattr(self, "_capture", None) is not None:
            # segmented capture: the all-reduce of this bucket happens between two graph replays
            optimizer.gather_autograd_grads()
            keys = self._pack_logs(optimizer)
            self._segment_end()
            self._capture["buckets"].append(optimizer.bucket)
            from ..parallel import world_size
            optimizer.grad_scale = 1.0 / world_size()
            self._segment_begin()
            self._unpack_logs(optimizer, keys, optimizer.grad_scale)
            return
        if torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            if isinstance(optimizer, FlatAdam):
                optimizer.gather_autograd_grads()
                keys = self._pack_logs(optimizer)
                optimizer.grad_scale = allreduce_sum_(optimizer.bucket)
                self._unpack_logs(optimizer, keys, optimizer.grad_scale)
            else:
                world = torch.distributed.get_world_size()
                for g in optimizer.param_groups:
                    for p in g["params"]:
                        if p.grad is not None:
                            torch.distributed.all_reduce(p.grad)
                            p.grad.div_(world)
                for k in self._unsynced_logs():
                    v = self.logged[k].clone()
                    torch.distributed.all_reduce(v)
                    self.logged[k] = v / world
                    self._synced.add(k)

    def _unsynced_logs(self):
        synced = self.__dict__.setdefault("_synced", set())
        return [k for k, v in self.logged.items() if k not in synced and isinstance(v, torch.Tensor)]

    def _pack_logs(self, optimizer):
        keys = self._unsynced_logs()[:FlatAdam.TAIL]
        for i in range(0, len(keys), 8):
            ops.gather_scalars([self.logged[k].reshape(1) for k in keys[i:i + 8]], optimizer.tail[i:i + 8], 1.0)
        return keys

    def _unpack_logs(self, optimizer, keys, scale: float) -> None:
        if not keys:
            return
        mean = torch.empty(len(keys), device=optimizer.tail.device, dtype=torch.float32)
        ops.axpby(optimizer.tail[:len(keys)], mean, scale, 0.0)
        for i, k in enumerate(keys):
            self.logged[k] = mean[i]
            self._synced.add(k)

    def _training_step_reference(self, batch: Dict[str, torch.Tensor]):
        corrupted_speech = self.generator.cut_to_valid_length(batch["audio_body_conducted"])
        reference_speech = self.generator.cut_to_valid_length(batch["audio_airborne"])
        generator_optimizer, discriminator_optimizer = self.optimizers(use_pl_optimizer=True)
        for opt in (generator_optimizer, discriminator_optimizer):
            if isinstance(opt, FlatAdam):
                opt.materialize()

        # Train Generator
        self.toggle_optimizer(generator_optimizer)
        with share_weight_norm():
            enhanced_speech, decomposed_enhanced_speech = self.generator(corrupted_speech)
            decomposed_reference_speech = self.generator.pqmf.forward(reference_speech, "analysis")
            atomic_losses_generator = self.compute_atomic_losses(
                "generator", enhanced_speech, reference_speech, decomposed_enhanced_speech,
                decomposed_reference_speech)
            for key, value in atomic_losses_generator.items():
                self.log(f"train/generator/{key}", value, sync_dist=True)
            lambdas = None
            if self.dynamic_loss_balancing is not None:
                lambdas = self.dynamically_balance_losses(atomic_losses_generator)
            backprop_loss_generator = WeightedSumFn.apply(lambdas, *atomic_losses_generator.values())
            self.log("train/generator/backprop_loss", backprop_loss_generator, sync_dist=True)
            self.manual_backward(backprop_loss_generator, generator_optimizer)
        generator_optimizer.step()
        generator_optimizer.zero_grad()
        self.untoggle_optimizer(generator_optimizer)

        # Train Discriminator
        self.toggle_optimizer(discriminator_optimizer)
        with share_weight_norm():
            atomic_losses_discriminator = self.compute_atomic_losses(
                "discriminator", enhanced_speech, reference_speech, decomposed_enhanced_speech,
                decomposed_reference_speech)
            update = bool(atomic_losses_discriminator)
            if update and self.update_discriminator_ratio < 1:
                update = bool(torch.rand(1) < self.update_discriminator_ratio)
            if update:
                for key, value in atomic_losses_discriminator.items():
                    self.log(f"train/discriminator/{key}", value, sync_dist=True)
                backprop_loss_discriminator = WeightedSumFn.apply(
                    None, atomic_losses_discriminator["real_loss"], atomic_losses_discriminator["fake_loss"])
                self.log("train/discriminator/backprop_loss", backprop_loss_discriminator, sync_dist=True)
                self.manual_backward(backprop_loss_discriminator, discriminator_optimizer)
                discriminator_optimizer.step()
                discriminator_optimizer.zero_grad()
        self.untoggle_optimizer(discriminator_optimizer)

        return {"corrupted": corrupted_speech, "enhanced": enhanced_speech.detach(), "reference": reference_speech}

    def compute_atomic_losses(self, network: str, enhanced_speech, reference_speech, decomposed_enhanced_speech,
                              decomposed_reference_speech) -> "OrderedDict[str, torch.Tensor]":
        atomic_losses = OrderedDict()
        assert network in {"generator", "discriminator"}
        if network == "generator":
            if self.reconstructive_loss_freq_fn:
                atomic_losses["reconstructive_loss_freq"] = self.reconstructive_loss_freq_fn(
                    enhanced_speech, reference_speech)
            if self.reconstructive_loss_temp_fn:
                atomic_losses["reconstructive_loss_temp"] = self.reconstructive_loss_temp_fn(
                    enhanced_speech, reference_speech)
            if self.feature_matching_loss_fn or self.adversarial_loss_fn:
                enhanced_embeddings = self.discriminator(bands=decomposed_enhanced_speech, audio=enhanced_speech)
                if self.feature_matching_loss_fn:
                    reference_embeddings = self.discriminator(bands=decomposed_reference_speech,
                                                              audio=reference_speech)
                    atomic_losses["feature_matching_loss"] = self.feature_matching_loss_fn(
                        enhanced_embeddings, reference_embeddings)
                if self.adversarial_loss_fn:
                    atomic_losses["adv_loss_gen"] = self.adversarial_loss_fn(embeddings=enhanced_embeddings, target=1)
        else:
            if self.adversarial_loss_fn:
                enhanced_embeddings = self.discriminator(bands=decomposed_enhanced_speech.detach(),
                                                         audio=enhanced_speech.detach())
                reference_embeddings = self.discriminator(bands=decomposed_reference_speech, audio=reference_speech)
                atomic_losses["real_loss"] = self.adversarial_loss_fn(embeddings=reference_embeddings, target=1)
                atomic_losses["fake_loss"] = self.adversarial_loss_fn(embeddings=enhanced_embeddings, target=-1)
        return atomic_losses

    def dynamically_balance_losses(self, atomic_losses) -> torch.Tensor:
        """eben.py:222-240.  Returns the detached lambdas (device tensor); the scaling itself is
        applied inside WeightedSumFn so the un-scaled losses stay available for logging."""
        layer = self.generator.last_conv.weight
        n = len(atomic_losses)
        dev = layer.device
        if self._bal is None or self._bal["old"].numel() != n:
            self._bal = dict(old=torch.zeros(n, device=dev), init=torch.zeros(1, device=dev, dtype=torch.int32))
        sumsq = torch.zeros(n, device=dev, dtype=torch.float64)
        for i, loss in enumerate(atomic_losses.values()):
            grad = torch.autograd.grad(loss, layer, retain_graph=True)[0]
            self._join(self.discriminator)
            ops.sumsq(grad.contiguous(), sumsq[i:i + 1])
        lambdas = torch.empty(n, device=dev)
        norms = torch.empty(n, device=dev)
        ops.balance(sumsq, self._bal["old"], self._bal["init"], lambdas, norms, self.beta_ema,
                    1 if self.dynamic_loss_balancing == "ema" else 0)
        self.last_norms, self.last_lambdas = norms, lambdas
        return lambdas

    # ---- checkpoint / resume ----------------------------------------------------
    def checkpoint(self, global_step: int = 0, epoch: int = 0) -> dict:
        """The keys a Lightning checkpoint of the reference module carries (`state_dict` with `generator.` /
        `discriminator.` prefixes, `optimizer_states` in `configure_optimizers()` order, `global_step`, `epoch`), plus
        the EMA of the balancing norms, which the reference keeps as a plain attribute and therefore loses on resume."""
        ckpt = {"state_dict": {k: v.detach().clone() for k, v in self.state_dict().items()},
                "optimizer_states": [o.state_dict() for o in self.configure_optimizers()],
                "global_step": int(global_step), "epoch": int(epoch)}
        if self._bal is not None:
            ckpt["vbx_balancing"] = {k: v.detach().clone() for k, v in self._bal.items()}
        return ckpt

    def load_checkpoint(self, ckpt: dict, strict: bool = True) -> int:
        """Restore from `checkpoint()` or from a Lightning checkpoint of the reference's EBENLightningModule (same
        parameter names; its `torch.optim.Adam` states load into FlatAdam).  Parameters are copied in place, so
        captured graphs and the flat buckets stay valid.  Returns the stored global step."""
        with torch.no_grad():
            own = self.state_dict()
            nets = ("generator.", "discriminator.")            # loss modules only hold constants (windows, filters)
            missing = [k for k in own if k.startswith(nets) and k not in ckpt["state_dict"]]
            unexpected = [k for k in ckpt["state_dict"] if k.startswith(nets) and k not in own]
            if strict and (missing or unexpected):
                raise KeyError(f"load_checkpoint: missing {missing[:4]}, unexpected {unexpected[:4]}")
            for k, v in ckpt["state_dict"].items():
                if k in own:
                    own[k].copy_(v)
        for opt, sd in zip(self.configure_optimizers(), ckpt.get("optimizer_states", [])):
            opt.load_state_dict(sd)
        for mod in (self.generator, self.discriminator):      # cached effective weights / packed tiles are stale
            for p in mod.parameters():
                p.__dict__.pop("_vbx_packs", None)
        bal = ckpt.get("vbx_balancing")
        dev = self.generator.last_conv.weight.device
        if bal is None:
            if self._bal is not None:                          # a captured graph reads these buffers: reset in place
                self._bal["old"].zero_(); self._bal["init"].zero_()
        elif self._bal is not None and self._bal["old"].numel() == bal["old"].numel():
            for k, v in bal.items():                           # in place, so a captured step keeps reading live memory
                self._bal[k].copy_(v)
        else:
            self._bal = {k: v.to(dev).clone() for k, v in bal.items()}
            self._graphs.clear()                               # (buffers were re-created: captured steps are stale)
        return int(ckpt.get("global_step", 0))

    @torch.no_grad()
    def common_eval_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0, stage: str = "validation",
                         dataloader_idx: int = 0):
        """eben.py:132-165: generator forward on the (cut) body-conducted signal and, when the airborne reference is
        in the batch, the atomic losses of both phases logged as `{stage}/{network}/{loss}[/{dataloader name}]`."""
        corrupted_speech = self.generator.cut_to_valid_length(batch["audio_body_conducted"])
        enhanced_speech, decomposed_enhanced_speech = self.generator(corrupted_speech)
        outputs = {"corrupted": corrupted_speech, "enhanced": enhanced_speech}
        if "audio_airborne" in batch:
            reference_speech = self.generator.cut_to_valid_length(batch["audio_airborne"])
            decomposed_reference_speech = self.generator.pqmf.forward(reference_speech, "analysis")
            outputs["reference"] = reference_speech
            names = getattr(self, "dataloader_names", None)
            dl_name = f"/{names[dataloader_idx]}" if names else ""
            for net_type in ["generator", "discriminator"]:
                for key, value in self.compute_atomic_losses(
                        net_type, enhanced_speech, reference_speech, decomposed_enhanced_speech,
                        decomposed_reference_speech).items():
                    self.log(f"{stage}/{net_type}/{key}{dl_name}", value, sync_dist=True, add_dataloader_idx=False)
        return outputs

    # base_se.py:132-136 (the metric / audio logging of common_eval_logging is torchmetrics territory: out of scope)
    def validation_step(self, batch, batch_idx: int = 0, dataloader_idx: int = 0):
        return self.common_eval_step(batch, batch_idx, "validation", dataloader_idx)

    def test_step(self, batch, batch_idx: int = 0, dataloader_idx: int = 0):
        return self.common_eval_step(batch, batch_idx, "test", dataloader_idx)
