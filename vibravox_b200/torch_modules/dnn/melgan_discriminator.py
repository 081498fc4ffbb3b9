"""DiscriminatorMelGAN drop-in (reference: vibravox/torch_modules/dnn/melgan_discriminator.py:76-169).
Same container tree (hence state_dict keys); each `conv (+ReflectionPad1d) + bias + LeakyReLU`
stage is one fused implicit-GEMM kernel launch."""
from __future__ import annotations

from typing import List

import torch
from torch import nn

from ...functional import ConvFn, Flags
from ..utils import conv_geom, effective_weight, normalized_conv1d


def _parse_stage(stage: nn.Module):
    extra, slope, conv = 0, 1.0, stage
    if isinstance(stage, nn.Sequential):
        conv = None
        for mod in stage:
            if isinstance(mod, nn.ReflectionPad1d):
                extra = int(mod.padding[0])
            elif isinstance(mod, nn.Conv1d):
                conv = mod
            elif isinstance(mod, nn.LeakyReLU):
                slope = mod.negative_slope
    return conv, extra, slope


def run_chain(stages, x: torch.Tensor) -> List[torch.Tensor]:
    """[x, stage_1(x), stage_2(stage_1(x)), ...] - the embeddings list of the reference forwards.  Under
    functional.Flags.gated_chain (the training step's discriminator calls) every stage is told the slope of the stage before
    it, so that its input-gradient kernel finishes that stage's LeakyReLU / feature-matching backward, and the outputs are
    tagged for the feature-matching loss."""
    chain = Flags.gated_chain and torch.is_grad_enabled()
    embeddings, prev_slope, prev_bias = [x], 1.0, None
    for stage in stages:
        conv, extra, slope = _parse_stage(stage)
        w, wt = effective_weight(conv)
        y = ConvFn.apply(embeddings[-1], w, wt, conv.bias, conv_geom(conv, extra), slope, prev_slope if chain else 1.0,
                         prev_bias if chain else None)
        if chain and slope != 1.0:
            y._vbx_chain = True
        embeddings.append(y)
        prev_slope, prev_bias = slope, conv.bias
    return embeddings


def prepare_stage(stage: nn.Module, backward: bool) -> list:
    """Everything of a stage that is created lazily and then SHARED by all passes of a step - the effective (weight-normed)
    weight and the packed tensor-core tiles of the forward and of the input gradient - produced now, on the current
    stream, so that passes running on other streams only ever read them.  Returns the tensors (for record_stream)."""
    from ... import ops
    conv, extra, _ = _parse_stage(stage)
    geom = conv_geom(conv, extra)
    w, wt = effective_weight(conv)
    made = [w] + ([wt] if wt is not None else [])
    # (the pack cache lives on the tensor OBJECT the passes will hand to the conv ops: `w` itself, not a detached alias)
    if ops.use_tc(geom, "fwd"):
        made.append(ops.get_pack(w, geom, ops.TC_FWD))
    if backward and ops.use_tc(geom, "dgrad"):
        made.append(ops.get_pack(w, geom, ops.TC_DGRAD))
    return made


class DiscriminatorMelGAN(nn.Module):
    def __init__(self, alpha_leaky_relu: float):
        super().__init__()
        a = alpha_leaky_relu

        def act():
            return nn.LeakyReLU(a, inplace=True)

        self.discriminator = nn.ModuleList([
            nn.Sequential(nn.ReflectionPad1d(7), normalized_conv1d(1, 16, kernel_size=(15,), stride=(1,)), act()),
            nn.Sequential(normalized_conv1d(16, 64, kernel_size=(41,), stride=(4,), padding=20, groups=4), act()),
            nn.Sequential(normalized_conv1d(64, 256, kernel_size=(41,), stride=(4,), padding=20, groups=4), act()),
            nn.Sequential(normalized_conv1d(256, 1024, kernel_size=(41,), stride=(4,), padding=20, groups=4), act()),
            nn.Sequential(normalized_conv1d(1024, 1024, kernel_size=(41,), stride=(4,), padding=20, groups=4), act()),
            nn.Sequential(normalized_conv1d(1024, 1024, kernel_size=(5,), stride=(1,), padding=2), act()),
            normalized_conv1d(1024, 1, kernel_size=3, stride=1, padding=1),
        ])

    def forward(self, audio: torch.Tensor) -> List[torch.Tensor]:
        return run_chain(self.discriminator, audio)
