"""MultiResolutionSTFTLoss drop-in for `auraloss.freq.MultiResolutionSTFTLoss` as the reference
configures it (configs/lightning_module/loss_module/multi_stft.yaml:1-18; restated in SURVEY App. C
because auraloss is a third-party, unpinned dependency that is not installable offline).

Supported argument surface = what the reference uses: fft_sizes / hop_sizes / win_lengths,
sample_rate, perceptual_weighting, hann window, w_sc = w_log_mag = 1, L1 log-magnitude distance,
mean reduction.  Anything else raises.

Each STFT is a strided Conv1d (stride = hop, K = win_length, reflect halo win/2) against a
windowed DFT basis, so forward and backward re-use the implicit-GEMM conv kernels; the
magnitude / spectral-convergence / log-magnitude statistics are one fused reduction kernel.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch

from ...functional import MRSTFTFn
from ...ops import ConvGeom


def a_weighting_fir(fs: int = 16000, ntaps: int = 101) -> torch.Tensor:
    """FIR approximation of the analog A-weighting curve: bilinear transform of the standard
    pole/zero set, sampled at 512 frequencies, least-squares FIR fit (auraloss FIRFilter 'aw')."""
    import numpy as np
    import scipy.signal
    f1, f2, f3, f4, a1000 = 20.598997, 107.65265, 737.86223, 12194.217, 1.9997
    num = [(2 * np.pi * f4) ** 2 * (10 ** (a1000 / 20)), 0, 0, 0, 0]
    den = np.polymul([1, 4 * np.pi * f4, (2 * np.pi * f4) ** 2], [1, 4 * np.pi * f1, (2 * np.pi * f1) ** 2])
    den = np.polymul(np.polymul(den, [1, 2 * np.pi * f3]), [1, 2 * np.pi * f2])
    b, a = scipy.signal.bilinear(num, den, fs=fs)
    freqs, resp = scipy.signal.freqz(b, a, worN=512, fs=fs)
    return torch.tensor(scipy.signal.firls(ntaps, freqs, abs(resp), fs=fs).astype("float32"))


def dft_basis(n_fft: int, win_length: int) -> torch.Tensor:
    """(2*bins, 1, win_length): rows [0,bins) = w[j] cos(2 pi k (j+left)/n_fft), rows [bins,2bins) =
    -w[j] sin(...), with the periodic Hann window centred in the n_fft frame as torch.stft does."""
    bins = n_fft // 2 + 1
    left = (n_fft - win_length) // 2
    window = torch.hann_window(win_length, dtype=torch.float64)
    j = torch.arange(win_length, dtype=torch.float64) + left
    k = torch.arange(bins, dtype=torch.float64).view(-1, 1)
    ang = 2 * math.pi * ((k * j) % n_fft) / n_fft
    basis = torch.cat([torch.cos(ang) * window, -torch.sin(ang) * window], dim=0)
    return basis.to(torch.float32).view(2 * bins, 1, win_length).contiguous()


class _Spec:
    """Device-resident constants handed to MRSTFTFn."""

    def __init__(self):
        self.taps = None
        self.fir_geom = None
        self.res = []
        self.eps = 1e-8
        self.counts_cache = {}


class MultiResolutionSTFTLoss(torch.nn.Module):
    def __init__(self, fft_sizes: Sequence[int] = (1024, 2048, 512), hop_sizes: Sequence[int] = (120, 240, 50),
                 win_lengths: Sequence[int] = (600, 1200, 240), window: str = "hann_window", w_sc: float = 1.0,
                 w_log_mag: float = 1.0, w_lin_mag: float = 0.0, w_phs: float = 0.0, sample_rate: float = None,
                 scale: str = None, n_bins: int = None, perceptual_weighting: bool = False,
                 scale_invariance: bool = False, **kwargs):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        if (window != "hann_window" or w_sc != 1.0 or w_log_mag != 1.0 or w_lin_mag != 0.0 or w_phs != 0.0
                or scale is not None or scale_invariance or kwargs):
            raise NotImplementedError("only the configuration of the reference's multi_stft.yaml is implemented")
        self.fft_sizes, self.hop_sizes, self.win_lengths = tuple(fft_sizes), tuple(hop_sizes), tuple(win_lengths)
        self.perceptual_weighting = perceptual_weighting
        if perceptual_weighting:
            if sample_rate is None:
                raise ValueError("`sample_rate` must be supplied when `perceptual_weighting = True`.")
            self.register_buffer("fir_taps", a_weighting_fir(int(sample_rate)).view(1, 1, -1), persistent=False)
        for i, (n_fft, win) in enumerate(zip(self.fft_sizes, self.win_lengths)):
            basis = dft_basis(n_fft, win)
            self.register_buffer(f"basis_{i}", basis, persistent=False)
            # Wk[(ci,k)][co] layout for the col2im backward
            self.register_buffer(f"basis_k_{i}", basis.view(basis.shape[0], win).t().contiguous(), persistent=False)
        self._spec = None

    def _get_spec(self, device) -> _Spec:
        if self._spec is None or self._spec.device != device:
            s = _Spec()
            s.device = device
            if self.perceptual_weighting:
                s.taps = self.fir_taps
                nt = s.taps.shape[-1]
                s.fir_geom = ConvGeom(1, 1, nt, 1, 1, nt // 2, 0, 1)
            for i, (n_fft, hop, win) in enumerate(zip(self.fft_sizes, self.hop_sizes, self.win_lengths)):
                pad = n_fft // 2 - (n_fft - win) // 2
                geom = ConvGeom(1, 2 * (n_fft // 2 + 1), win, hop, 1, pad, pad, 1)
                s.res.append((geom, getattr(self, f"basis_{i}"), getattr(self, f"basis_k_{i}")))
            self._spec = s
        return self._spec

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return MRSTFTFn.apply(input, target, self._get_spec(input.device))
