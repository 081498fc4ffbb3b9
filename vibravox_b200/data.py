"""Synthetic datamodules with the batch contract of the reference's BWE / noisy-BWE datamodules
(`vibravox/lightning_datamodules/bwe.py:232-293`, `noisybwe.py:254-277`): dicts with
"audio_body_conducted" and "audio_airborne" of shape (B, 1, samples).  The real loaders need the HF hub
(no network here); the noisy variant performs the reference's mix + joint crop arithmetic
(`vibravox/utils.py:195-254,50-81`) on the device with vbx_noise_mix_crop."""
from __future__ import annotations

import re

import torch

from . import ops


def _samples(collate_strategy: str, sample_rate: int) -> int:
    m = re.fullmatch(r"constant_length-(\d+)-ms", collate_strategy)
    if not m:
        raise ValueError("collate_strategy must be 'constant_length-XXX-ms'")
    return int(m.group(1)) * sample_rate // 1000


def synthetic_pairs(batch: int, samples: int, seed: int):
    """(body-conducted, airborne) noise pair of SURVEY 8(d): 0.1*randn clamped to +-1, (B,1,samples) each, drawn
    from a host generator so that every tool / bench / test sees the same batch for a seed."""
    g = torch.Generator().manual_seed(seed)
    air = (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)
    body = (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)
    return body, air


class SyntheticBWEDataModule:
    def __init__(self, sample_rate: int = 16000, batch_size: int = 32,
                 collate_strategy: str = "constant_length-3000-ms", seed: int = 42, id: str = "synthetic_bwe"):
        self.sample_rate, self.batch_size, self.seed, self.id = sample_rate, batch_size, seed, id
        self.samples = _samples(collate_strategy, sample_rate)

    def batches(self, device, rank: int = 0):
        g = torch.Generator().manual_seed(self.seed + rank)
        while True:
            air = (0.1 * torch.randn(self.batch_size, 1, self.samples, generator=g)).clamp(-1, 1)
            body = (0.1 * torch.randn(self.batch_size, 1, self.samples, generator=g)).clamp(-1, 1)
            if torch.device(device).type == "cuda":
                body, air = body.pin_memory(), air.pin_memory()
            yield {"audio_body_conducted": body.to(device, non_blocking=True),
                   "audio_airborne": air.to(device, non_blocking=True)}


class SyntheticNoisyBWEDataModule(SyntheticBWEDataModule):
    def __init__(self, noise_seconds: float = 12.0, **kw):
        super().__init__(**kw)
        self.noise_samples = int(noise_seconds * self.sample_rate)

    def batches(self, device, rank: int = 0):
        g = torch.Generator().manual_seed(self.seed + rank)
        B, S = self.batch_size, self.samples
        Ls = S + S // 4                                    # utterances longer than the crop
        while True:
            air = (0.1 * torch.randn(B, 1, Ls, generator=g)).clamp(-1, 1).to(device)
            body = (0.1 * torch.randn(B, 1, Ls, generator=g)).clamp(-1, 1).to(device)
            noise = (0.05 * torch.randn(B, 1, self.noise_samples, generator=g)).to(device)
            # start ~ randint(0, len_noise - len_speech), offset ~ randint(0, len - target + 1)
            start = torch.randint(0, self.noise_samples - Ls + 1, (B,), generator=g, dtype=torch.int32).to(device)
            off = torch.randint(0, Ls - S + 1, (B,), generator=g, dtype=torch.int32).to(device)
            ob, oa = ops.noise_mix_crop(body, air, noise, start, off, S)
            yield {"audio_body_conducted": ob, "audio_airborne": oa}
