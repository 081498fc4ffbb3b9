"""WaveformDataAugmentation (reference: vibravox/torch_modules/dsp/data_augmentation.py:8-79): the collate-time
augmentation hook of the BWE datamodules.  Same constructor, same decisions from the same draws of torch's global
generator - `rand(1)` for the gate (taken even at the default `p_data_augmentation=0`, so it matters for RNG
parity), then `rand(1)` + `randint` per transform.  Time masking is local (time_masking_waveform.py); speed
perturbation and pitch shift are torchaudio transforms in the reference and are delegated to torchaudio when it is
installed - they are dataloader-side resampling, outside the training-step path this package accelerates."""
from typing import Tuple

import torch

from .time_masking_waveform import TimeMaskingBlockWaveform


class WaveformDataAugmentation(torch.nn.Module):
    def __init__(self, sample_rate, p_data_augmentation=0, p_speed_perturbation=0.3, p_pitch_shift=0.3,
                 p_time_masking=0.3,
                 speed_perturbation_factors=(0.7, 0.8, 0.85, 0.9, 0.95, 1.05, 1.1, 1.15, 1.2, 1.3),
                 pitch_shift_steps=(-4, -3, -2, -1, 1, 2, 3, 4, 5, 6),
                 time_masking_percentage=(1, 2, 3, 4, 5, 6, 7, 8)):
        super().__init__()
        self.sample_rate = sample_rate
        assert 0 <= p_data_augmentation <= 1, "p_data_augmentation must be in [0, 1]"
        assert 0 <= p_speed_perturbation <= 1, "p_speed_perturbation must be in [0, 1]"
        assert 0 <= p_pitch_shift <= 1, "p_pitch_shift must be in [0, 1]"
        assert 0 <= p_time_masking <= 1, "p_time_masking must be in [0, 1]"
        self.apply_data_augmentation = p_data_augmentation
        self.p_speed_perturbation = p_speed_perturbation
        self.p_pitch_shift = p_pitch_shift
        self.p_time_masking = p_time_masking
        self.speed_perturbation_factors = speed_perturbation_factors
        self.pitch_shift_steps = pitch_shift_steps
        self.time_masking_percentage = time_masking_percentage

    @staticmethod
    def _torchaudio():
        try:
            import torchaudio.transforms as T
        except Exception as exc:                 # pragma: no cover - torchaudio is present in this image
            raise RuntimeError("speed perturbation / pitch shift need torchaudio (as in the reference)") from exc
        return T

    def forward(self, waveform_1: torch.Tensor, waveform_2: torch.Tensor = None) -> Tuple[torch.Tensor]:
        if torch.rand(1) < self.apply_data_augmentation:
            if torch.rand(1) < self.p_speed_perturbation:
                factor = self.speed_perturbation_factors[
                    torch.randint(len(self.speed_perturbation_factors), size=(1,)).item()]
                speed = self._torchaudio().SpeedPerturbation(orig_freq=self.sample_rate, factors=[factor])
                waveform_1, _ = speed(waveform_1)
                if waveform_2 is not None:
                    waveform_2, _ = speed(waveform_2)
            if torch.rand(1) < self.p_pitch_shift:
                step = self.pitch_shift_steps[torch.randint(len(self.pitch_shift_steps), size=(1,)).item()]
                pitch = self._torchaudio().PitchShift(self.sample_rate, n_steps=step)
                waveform_1 = pitch(waveform_1)
                if waveform_2 is not None:
                    waveform_2 = pitch(waveform_2)
            if torch.rand(1) < self.p_time_masking:
                pct = self.time_masking_percentage[torch.randint(len(self.time_masking_percentage), size=(1,)).item()]
                mask = TimeMaskingBlockWaveform(masking_percentage=pct)
                waveform_1 = mask(waveform_1)
                if waveform_2 is not None:
                    waveform_2 = mask(waveform_2)
        return waveform_1, waveform_2
