"""TimeMaskingBlockWaveform (reference: vibravox/torch_modules/dsp/time_masking_waveform.py:3-37): zero one
contiguous block of `masking_percentage` % of the samples, in place, at a position drawn with one
`torch.randint(0, T - masked, (1,))` from torch's global generator (same draw as the reference)."""
import torch


class TimeMaskingBlockWaveform(torch.nn.Module):
    def __init__(self, masking_percentage=2):
        super().__init__()
        assert 0 <= masking_percentage <= 100, "masking_percentage should be in [0, 100]"
        self.masking_percentage = masking_percentage

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        time_samples = x.shape[-1]
        masked_samples = int(time_samples * self.masking_percentage / 100)
        first = torch.randint(0, time_samples - masked_samples, (1,)).item()
        x[..., first:first + masked_samples] = 0
        return x
