"""Hinge loss drop-in (reference: vibravox/torch_modules/losses/hinge_loss.py:6-43)."""
from __future__ import annotations

from typing import List

import torch

from ...functional import HingeFn


class HingeLossForDiscriminatorMelganMultiScales(torch.nn.Module):
    def forward(self, embeddings: List[List[torch.Tensor]], target: float) -> torch.Tensor:
        return HingeFn.apply(float(target), *[scale[-1] for scale in embeddings])
