"""PseudoQMFBanks drop-in (reference: vibravox/torch_modules/dsp/pqmf.py:16-233).

Same constructor, properties, parameters (`analysis_weights`, `synthesis_weights`,
requires_grad=False) and `forward(signal, stage, bands=-1)`; analysis / synthesis run as the
polyphase kernels vbx_pqmf_analysis / vbx_pqmf_synthesis.  The filter design (Kaiser-sinc
prototype, 5 L-BFGS steps on the cut-off, cosine modulation) is init-time host arithmetic
and reproduces the reference taps bit for bit (tests/test_pqmf_design.py).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from ...functional import PQMFAnalysisFn, PQMFSynthesisFn


class PseudoQMFBanks(nn.Module):
    def __init__(self, decimation: int = 32, kernel_size: int = 1024, beta: int = 9):
        super().__init__()
        assert kernel_size % (4 * decimation) == 0          # pqmf.py:42
        self._decimation, self._kernel_size, self._beta = decimation, kernel_size, beta
        self._cutoff_ratio = self.initialize_cutoff_ratio()
        analysis, synthesis = self.initialize_pqmf_bank()
        self.analysis_weights = nn.Parameter(analysis, requires_grad=False)
        self.synthesis_weights = nn.Parameter(synthesis, requires_grad=False)

    @property
    def kernel_size(self) -> int:
        return self._kernel_size

    @property
    def decimation(self) -> int:
        return self._decimation

    # ---- design (pqmf.py:66-180): fp32 sinc on a centred grid, Kaiser window held in fp64 ----
    def _grid(self) -> torch.Tensor:
        return torch.arange(self._kernel_size) - (self._kernel_size - 1) / 2

    def compute_prototype(self, cutoff_ratio) -> torch.Tensor:
        n = self._kernel_size
        kaiser = torch.kaiser_window(n, periodic=False, beta=self._beta).to(torch.float64)
        lowpass = cutoff_ratio * torch.special.sinc(cutoff_ratio * self._grid())
        return (lowpass.to(torch.float64) * kaiser).to(torch.float32).view(1, 1, n)

    def _objective(self, cutoff_ratio) -> torch.Tensor:
        n, m = self._kernel_size, self._decimation
        proto = self.compute_prototype(cutoff_ratio)
        autocorr = torch.nn.functional.conv1d(torch.nn.functional.pad(proto, (n // 2, n // 2)), proto)
        keep = torch.ones(n + 1, dtype=autocorr.dtype)
        keep[n // 2] = 0                                      # ignore the zero-lag peak
        worst = (autocorr * keep)[..., :: 2 * m].abs().max()
        outside = abs(float(cutoff_ratio.detach()) - 1 / (2 * m)) > 1 / (4 * m)
        return worst + (1 / (4 * m) if outside else 0)

    def initialize_cutoff_ratio(self) -> float:
        cutoff = (torch.ones(1) / (2 * self._decimation)).requires_grad_(True)
        optimizer = torch.optim.LBFGS([cutoff], line_search_fn="strong_wolfe")
        for _ in range(5):
            optimizer.zero_grad()
            self._objective(cutoff).backward()
            optimizer.step(lambda: self._objective(cutoff))
        return cutoff.item()

    def initialize_pqmf_bank(self):
        n, m = self._kernel_size, self._decimation
        proto = self.compute_prototype(self._cutoff_ratio).view(1, n)
        freq = torch.tensor([(2 * k + 1) * math.pi / 2 / m for k in range(m)]).view(m, 1)
        phase = torch.tensor([(-1) ** k * math.pi / 4 for k in range(m)]).view(m, 1)
        arg = freq * self._grid().view(1, n)
        analysis = 2 * torch.flip(proto * torch.cos(arg + phase), [1])
        synthesis = (2 * m) * proto * torch.cos(arg - phase)
        return analysis.view(m, 1, n).contiguous(), synthesis.view(m, 1, n).contiguous()

    # ---- hot path -----------------------------------------------------------------------------
    def forward(self, signal: torch.Tensor, stage: str, bands: int = -1) -> torch.Tensor:
        if stage == "analysis":
            nb = self._decimation if bands == -1 else bands
            return PQMFAnalysisFn.apply(signal, self.analysis_weights, nb)
        if stage == "synthesis":
            return PQMFSynthesisFn.apply(signal, self.synthesis_weights, False)
        raise ValueError(f"Invalid stage '{stage}'. Expected 'analysis' or 'synthesis'.")

    def synthesis_sum(self, bands: torch.Tensor) -> torch.Tensor:
        """sum over bands of forward(bands, "synthesis"), fused in one kernel (eben_generator.py:209-211)."""
        return PQMFSynthesisFn.apply(bands, self.synthesis_weights, True)

    def cut_tensor(self, tensor: torch.Tensor) -> torch.Tensor:
        old_len = tensor.shape[2]
        return torch.narrow(tensor, 2, 0, old_len - (old_len + self._kernel_size) % self._decimation)
