"""Feature-matching loss drop-in (reference: vibravox/torch_modules/losses/feature_loss.py:6-50)."""
from __future__ import annotations

from typing import List

import torch

from ...functional import FeatureMatchingFn


class FeatureLossForDiscriminatorMelganMultiScales(torch.nn.Module):
    def forward(self, embeddings_a: List[List[torch.Tensor]], embeddings_b: List[List[torch.Tensor]]) -> torch.Tensor:
        a, b = [], []
        for scale_a, scale_b in zip(embeddings_a, embeddings_b):
            for layer_a, layer_b in zip(scale_a[1:-1], scale_b[1:-1]):   # skip audio and certainties
                a.append(layer_a)
                b.append(layer_b)
        # reference divisor: number of scales x inner layers of the LAST scale (feature_loss.py:48)
        scale = 1.0 / (len(embeddings_a) * len(embeddings_a[-1][1:-1]))
        # stage outputs of a discriminator chain run under functional.Flags.gated_chain carry this tag: their gradient
        # terms are then applied inside the conv stages' input-gradient kernels
        fused = all(getattr(t, "_vbx_chain", False) for t in a)
        return FeatureMatchingFn.apply(scale, len(a), fused, *a, *b)
