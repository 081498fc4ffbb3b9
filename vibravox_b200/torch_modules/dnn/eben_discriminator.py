"""DiscriminatorEBENMultiScales drop-in (reference: vibravox/torch_modules/dnn/eben_discriminator.py:10-163).
`DiscriminatorEBENMultiScales(q=3, min_channels=24).forward(bands, audio) -> List[List[Tensor]]`:
three PQMF-band discriminators (dilation 1, 2, 3; grouped convs) on `bands[:, -q:, :]` and one
MelGAN discriminator on the waveform; element 0 of every list is the input, intermediate
entries are post-LeakyReLU(0.2), the last is the raw certainty map."""
from __future__ import annotations

import os
from typing import List

import torch
from torch import nn

try:
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass

from ..utils import normalized_conv1d
from .melgan_discriminator import DiscriminatorMelGAN, prepare_stage, run_chain


class DiscriminatorEBEN(nn.Module):
    def __init__(self, dilation: int = 1, q: int = 3, min_channels: int = 24):
        super().__init__()
        self.dilation = dilation
        assert min_channels % q == 0, "min_channels must be a multiple of q"
        c = min_channels

        def act():
            return nn.LeakyReLU(0.2, inplace=True)

        stages = [nn.Sequential(nn.ReflectionPad1d(1),
                                normalized_conv1d(q, c, kernel_size=(3,), stride=(1,), padding=(1,),
                                                  dilation=dilation, groups=q), act())]
        for _ in range(5):
            stages.append(nn.Sequential(normalized_conv1d(c, 2 * c, kernel_size=(7,), stride=(2,), padding=(3,),
                                                          dilation=dilation, groups=q), act()))
            c *= 2
        stages.append(nn.Sequential(normalized_conv1d(c, c, kernel_size=(5,), stride=(1,), padding=(2,),
                                                      dilation=dilation, groups=q), act()))
        stages.append(normalized_conv1d(c, 1, kernel_size=(3,), stride=(1,), padding=(1,), groups=1))
        self.discriminator = nn.ModuleList(stages)

    def forward(self, bands: torch.Tensor) -> List[torch.Tensor]:
        return run_chain(self.discriminator, bands)


class DiscriminatorEBENMultiScales(nn.Module, PyTorchModelHubMixin):
    def __init__(self, q: int = 3, min_channels: int = 24):
        super().__init__()
        self.q = q
        self.pqmf_discriminators = nn.ModuleList(
            [DiscriminatorEBEN(dilation=d, q=q, min_channels=min_channels) for d in (1, 2, 3)])
        self.melgan_discriminator = DiscriminatorMelGAN(alpha_leaky_relu=0.2)

    def forward(self, bands: torch.Tensor, audio: torch.Tensor) -> List[List[torch.Tensor]]:
        return self.forward_multi([(bands, audio)])[0]

    def forward_multi(self, pairs, join: bool = True) -> List[List[List[torch.Tensor]]]:
        """forward() for several independent (bands, audio) pairs at once (e.g. enhanced and reference).
        `join=False` leaves the work running on the side streams (the caller's stream does NOT wait for it): the training
        step starts the reference pass this way before the generator's forward, which does not depend on it, and joins
        with the next call / `join_streams()`.
        The four sub-discriminators are independent chains of small / medium kernels: each runs on its own
        stream so that together they fill the 148 SMs; autograd replays each chain's backward on the stream
        its forward ran on and inserts the cross-stream waits itself."""
        nets = list(self.pqmf_discriminators) + [self.melgan_discriminator]
        jobs = []
        for bands, audio in pairs:
            selected = bands[:, -self.q:, :]
            if not selected.is_contiguous():
                selected = selected.contiguous()
            jobs.append([selected, selected, selected, audio])
        if not (jobs[0][0].is_cuda and _SIDE_STREAMS):
            return [[net(x) for net, x in zip(nets, inputs)] for inputs in jobs]
        # One stream per SUB-DISCRIMINATOR, plus one more per additional PASS (enhanced / reference): what the passes of a
        # network share - effective weights, packed tensor-core tiles - is produced first on the network's main stream
        # (`prepare_stage`), the other pass streams wait for that, and then every pass is an independent chain.  The
        # chains of one network launch the same kernels on different data and can fill each other's half-empty waves.
        # (VBX_D_PASS_STREAMS=1; default off, see below.)
        # Deterministic mode keeps all passes of a network on one stream: bias gradients of both passes accumulate into
        # the same bucket slot and their order must not depend on stream timing.
        from ... import ops
        cur = torch.cuda.current_stream()
        per_pass = _PASS_STREAMS and len(jobs) > 1 and not ops.DETERMINISTIC
        npass = len(jobs) if per_pass else 1
        streams = self._streams(jobs[0][0].device, len(nets) * npass)
        out = [[None] * len(nets) for _ in jobs]
        backward = torch.is_grad_enabled()
        for i, net in enumerate(nets):
            main = streams[i * npass]
            main.wait_stream(cur)
            if per_pass:
                with torch.cuda.stream(main):
                    shared = [t for stage in net.discriminator for t in prepare_stage(stage, backward)]
                for k in range(1, npass):
                    streams[i * npass + k].wait_stream(main)
                    streams[i * npass + k].wait_stream(cur)
                    for t in shared:
                        t.record_stream(streams[i * npass + k])
            for j, inputs in enumerate(jobs):
                st = streams[i * npass + (j if per_pass else 0)]
                with torch.cuda.stream(st):
                    x = inputs[i]
                    x.record_stream(st)
                    emb = net(x)
                    for t in emb[1:]:
                        t.record_stream(cur)            # consumed by the losses on the caller's stream
                    out[j][i] = emb
        if join:
            for st in streams:
                cur.wait_stream(st)
        return out

    def join_streams(self) -> None:
        """Make the caller's stream wait for everything queued on the side streams.  Needed after a backward
        pass whose parameter gradients were accumulated straight into a flat bucket (functional.grad_slot):
        those writes are invisible to autograd, so its end-of-backward synchronisation does not cover them."""
        if not torch.cuda.is_available():
            return
        cur = torch.cuda.current_stream()
        for pool in self.__dict__.get("_vbx_streams", {}).values():
            for st in pool:
                cur.wait_stream(st)

    def _streams(self, device, n):
        cache = self.__dict__.setdefault("_vbx_streams", {})
        pool = cache.setdefault(device, [])
        while len(pool) < n:
            # experiment knob (off by default): the MelGAN chain carries 75 % of the discriminator MACs and is the
            # last of the four to be enqueued; a high-priority stream lets its kernels take SMs first while the three
            # PQMF-band chains fill the gaps (critical-path-first)
            prio = -1 if (_MELGAN_PRIORITY and len(pool) == n - 1) else 0
            pool.append(torch.cuda.Stream(device=device, priority=prio))
        return pool


_SIDE_STREAMS = os.environ.get("VBX_D_STREAMS", "1") != "0"
# one more stream per PASS of a network (off by default: measured 39.77 vs 39.85 ms per step - the step is bound by the
# kernels' own throughput, not by half-empty waves; kept as a knob)
_PASS_STREAMS = os.environ.get("VBX_D_PASS_STREAMS", "0") == "1"
_MELGAN_PRIORITY = os.environ.get("VBX_D_PRIORITY", "0") == "1"
if _SIDE_STREAMS and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
    # leaves created on the caller's stream (detached generator outputs) receive gradients from side streams
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
