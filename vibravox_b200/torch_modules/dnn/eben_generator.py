"""EBENGenerator drop-in (reference: vibravox/torch_modules/dnn/eben_generator.py:89-316).

Constructor `EBENGenerator(m, n, p)`, `forward(cut_audio) -> (enhanced, enhanced_decomposed)`,
`cut_to_valid_length`, `.pqmf`, `.last_conv.weight`, `.p`, `.multiple` and every
`state_dict()` key are those of the reference; the arithmetic runs in libvbx_b200 kernels:
PQMF polyphase analysis/synthesis, implicit-GEMM convs with the reflect halo folded in,
residual units with activation + skip add in the conv epilogue, ConvTranspose as the
phase-decomposed dgrad kernel with the decoder skip add in front and LeakyReLU behind.
"""
from __future__ import annotations

import torch
from torch import nn

try:  # same mixin as the reference when huggingface_hub is present (it is not on the hot path)
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass

from ...functional import ConvFn, ConvTransposeFn, LeakyReluFn, ResidualUnitFn, TanhRecomposeFn
from ..dsp.pqmf import PseudoQMFBanks
from ..utils import (conv_geom, conv_trans_geom, effective_weight, normalized_conv1d,
                     normalized_conv_trans1d)


def _conv(conv: nn.Conv1d, x: torch.Tensor, slope: float = 1.0) -> torch.Tensor:
    w, wt = effective_weight(conv, need_wt=x.requires_grad)
    return ConvFn.apply(x, w, wt, conv.bias, conv_geom(conv), slope)


class ResidualUnit(nn.Module):
    def __init__(self, channels: int, nl: nn.LeakyReLU, dilation: int, bias: bool = False):
        super().__init__()
        self.dilated_conv = normalized_conv1d(channels, channels, kernel_size=3, dilation=dilation,
                                              padding="same", bias=bias, padding_mode="reflect")
        self.pointwise_conv = normalized_conv1d(channels, channels, kernel_size=1, padding="same",
                                                bias=bias, padding_mode="reflect")
        self.nl = nl
        assert not bias

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w1, wt1 = effective_weight(self.dilated_conv)
        w2, wt2 = effective_weight(self.pointwise_conv)
        return ResidualUnitFn.apply(x, w1, wt1, w2, wt2, conv_geom(self.dilated_conv),
                                    conv_geom(self.pointwise_conv), self.nl.negative_slope)


def _residuals(channels: int, nl: nn.LeakyReLU) -> nn.Sequential:
    return nn.Sequential(*[ResidualUnit(channels=channels, nl=nl, dilation=d) for d in (1, 3, 9)])


class EncBlock(nn.Module):
    def __init__(self, out_channels: int, stride: int, nl: nn.LeakyReLU, bias: bool = False):
        super().__init__()
        self.nl = nl
        self.residuals = _residuals(out_channels // 2, nl)
        self.conv = normalized_conv1d(out_channels // 2, out_channels, kernel_size=2 * stride, stride=stride,
                                      padding=stride - 1, bias=bias, padding_mode="reflect")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _conv(self.conv, self.residuals(x))


class DecBlock(nn.Module):
    def __init__(self, out_channels: int, stride: int, nl: nn.LeakyReLU, bias: bool = False):
        super().__init__()
        self.nl = nl
        self.residuals = _residuals(out_channels, nl)
        self.conv_trans = normalized_conv_trans1d(2 * out_channels, out_channels, kernel_size=2 * stride,
                                                  stride=stride, padding=stride // 2, output_padding=0, bias=bias)
        assert not bias

    def forward(self, x: torch.Tensor, encoder_output: torch.Tensor) -> torch.Tensor:
        w, wt = effective_weight(self.conv_trans)
        geom, out_pad = conv_trans_geom(self.conv_trans)
        y = ConvTransposeFn.apply(x, encoder_output, w, wt, geom, self.nl.negative_slope, out_pad)
        return self.residuals(y)


class EBENGenerator(nn.Module, PyTorchModelHubMixin):
    def __init__(self, m: int, n: int, p: int):
        """m: PQMF bands (= decimation); n: PQMF kernel size; p: informative bands fed to the net."""
        super().__init__()
        self.p = p
        self.pqmf = PseudoQMFBanks(decimation=m, kernel_size=n)
        self.multiple = 2 * 4 * 8 * m
        self.nl = nn.LeakyReLU(negative_slope=0.01)
        self.first_conv = nn.Conv1d(self.p, 32, kernel_size=3, padding="same", bias=False, padding_mode="reflect")
        self.encoder_blocks = nn.ModuleList([EncBlock(64, 2, self.nl), EncBlock(128, 4, self.nl),
                                             EncBlock(256, 8, self.nl)])
        self.latent_conv = nn.Sequential(
            self.nl,
            normalized_conv1d(256, 64, kernel_size=7, padding="same", bias=False, padding_mode="reflect"),
            self.nl,
            normalized_conv1d(64, 256, kernel_size=7, padding="same", bias=False, padding_mode="reflect"),
            self.nl,
        )
        self.decoder_blocks = nn.ModuleList([DecBlock(128, 8, self.nl), DecBlock(64, 4, self.nl),
                                             DecBlock(32, 2, self.nl)])
        self.last_conv = nn.Conv1d(32, 4, kernel_size=3, padding="same", bias=False, padding_mode="reflect")

    def forward(self, cut_audio: torch.Tensor):
        slope = self.nl.negative_slope
        first_bands = self.pqmf(cut_audio, "analysis", bands=self.p)
        x = _conv(self.first_conv, first_bands)
        x1 = self.encoder_blocks[0](LeakyReluFn.apply(x, slope))
        x2 = self.encoder_blocks[1](LeakyReluFn.apply(x1, slope))
        x3 = self.encoder_blocks[2](LeakyReluFn.apply(x2, slope))
        # latent_conv = nl, conv, nl, conv, nl : the 2nd and 3rd activation ride in the conv epilogues
        x = LeakyReluFn.apply(x3, slope)
        x = _conv(self.latent_conv[1], x, slope)
        x = _conv(self.latent_conv[3], x, slope)
        x = self.decoder_blocks[0](x, x3)
        x = self.decoder_blocks[1](x, x2)
        x = self.decoder_blocks[2](x, x1)
        x = _conv(self.last_conv, x)
        enhanced_speech_decomposed = TanhRecomposeFn.apply(x, first_bands, self.p)
        enhanced_speech = self.pqmf.synthesis_sum(enhanced_speech_decomposed)
        return enhanced_speech, enhanced_speech_decomposed

    def cut_to_valid_length(self, tensor: torch.Tensor) -> torch.Tensor:
        old_len = tensor.shape[2]
        new_len = old_len - (old_len + self.pqmf.kernel_size) % self.multiple
        return torch.narrow(tensor, 2, 0, new_len)
