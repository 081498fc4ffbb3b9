"""Batch assembly for the BWE / noisy-BWE training step.

Mirrors, name for name, the helpers the reference's datamodules call - `pad_audio`, `slice_audio`,
`set_audio_duration`, `mix_speech_and_noise_without_rescaling` (`vibravox/utils.py:7-81,195-254`) and the two
`data_collator`s (`lightning_datamodules/bwe.py:232-293`, `noisybwe.py:225-300`) - with the same arguments, the
same error behaviour and the same consumption of torch's global RNG, so a seeded run draws the same crops and
noise segments as the reference (checked draw for draw in tests/test_collate.py).

Design: the random decisions are taken first (`plan_*`: a few host integers per item, exactly the reference's
`torch.randint` calls in the reference's order), the samples are moved second.  When every item of a noisy-BWE
batch has the same length, lives on the GPU and is at least as long as the target, the move is ONE launch of
`vbx_noise_mix_crop` (mix + joint crop fused, `ops.noise_mix_crop`); otherwise plain slicing / padding on
whatever device the items live on (the reference does this in CPU dataloader workers).  Waveform augmentation
(`torch_modules/dsp/data_augmentation.py` of the reference: SURVEY 2 row 12, out of scope) is a caller-supplied hook:
pass the reference's own `WaveformDataAugmentation` object as `data_augmentation`; without one the collators consume
the one `rand(1)` gate draw the reference's default (probability 0) hook takes, so seeded runs stay in step.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch.nn.utils.rnn import pad_sequence


def pad_audio(audio: torch.Tensor, desired_samples: int) -> torch.Tensor:
    """`vibravox/utils.py:7-31`, including its operator-precedence quirk: the left pad is
    `desired - initial // 2` (not `(desired - initial) // 2`), so the right "pad" is negative and crops - the result
    is `desired - initial // 2` zeros followed by the first `initial // 2` samples."""
    assert audio.shape[-1] <= desired_samples, \
        "The audio signal is longer than the desired duration. Use set_audio_duration instead."
    initial = audio.shape[-1]
    left = desired_samples - initial // 2
    out = audio.new_zeros(audio.shape[:-1] + (desired_samples,))
    out[..., left:] = audio[..., : initial // 2]
    return out


def slice_audio(audio: torch.Tensor, desired_samples: int, offset_samples: int) -> torch.Tensor:
    """`vibravox/utils.py:33-48`."""
    assert audio.shape[-1] >= desired_samples, \
        "The audio signal is shorter than the desired duration. Use pad_audio instead."
    offset_samples = int(offset_samples)
    return audio[..., offset_samples: offset_samples + desired_samples]


def plan_duration(initial_samples: int, desired_samples: int, deterministic: bool) -> Optional[int]:
    """The crop offset `set_audio_duration` would use (None: the item is padded instead).  Draws from torch's
    global generator exactly when the reference does (`utils.py:71-73`)."""
    if initial_samples < desired_samples:
        return None
    if deterministic:
        return (initial_samples - desired_samples) // 2
    return int(torch.randint(low=0, high=initial_samples - desired_samples + 1, size=(1,)))


def set_audio_duration(audio: torch.Tensor, desired_samples: int, audio_bis: Optional[torch.Tensor] = None,
                       deterministic: bool = False):
    """`vibravox/utils.py:50-81`: joint crop (random or centred) or pad of one or two signals."""
    assert audio_bis is None or audio.shape == audio_bis.shape, "The two audio signals must have the same shape."
    offset = plan_duration(audio.shape[-1], desired_samples, deterministic)
    fix = (lambda a: pad_audio(a, desired_samples)) if offset is None else \
        (lambda a: slice_audio(a, desired_samples, offset))
    audio = fix(audio)
    return (audio, fix(audio_bis)) if audio_bis is not None else audio


def _check_lists(speech_batch, noise_batch) -> None:
    if not isinstance(speech_batch, list) or not all(isinstance(t, torch.Tensor) for t in speech_batch):
        raise TypeError("speech_batch must be a list of torch.Tensor")
    if not isinstance(noise_batch, list) or not all(isinstance(t, torch.Tensor) for t in noise_batch):
        raise TypeError("noise_batch must be a list of torch.Tensor")
    if len(speech_batch) != len(noise_batch):
        raise ValueError("speech_batch and noise_batch must have the same length")


def plan_noise_starts(speech_batch: List[torch.Tensor], noise_batch: List[torch.Tensor]) -> List[int]:
    """First sample of the noise segment mixed into each utterance: `randint(0, len_noise - len_speech)` per item,
    in batch order (`utils.py:243-245`; like the reference it raises when the two lengths are equal)."""
    _check_lists(speech_batch, noise_batch)
    starts = []
    for speech, noise in zip(speech_batch, noise_batch):
        if speech.dim() != 1:
            raise ValueError(f"Each speech sample must be a 1D tensor, but got shape {speech.shape}")
        if noise.dim() != 1:
            raise ValueError(f"Each noise sample must be a 1D tensor, but got shape {noise.shape}")
        if noise.size(0) < speech.size(0):
            raise ValueError(f"noise_sample length ({noise.size(0)}) must be >= speech_sample length ({speech.size(0)})")
        starts.append(int(torch.randint(0, noise.size(0) - speech.size(0), (1,)).item()))
    return starts


def mix_speech_and_noise_without_rescaling(speech_batch: List[torch.Tensor], noise_batch: List[torch.Tensor]
                                           ) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """`vibravox/utils.py:195-254`: `speech + noise[start:start+len]`, no SNR scaling."""
    starts = plan_noise_starts(speech_batch, noise_batch)
    sliced = [n[s: s + sp.size(0)] for sp, n, s in zip(speech_batch, noise_batch, starts)]
    return [sp + n for sp, n in zip(speech_batch, sliced)], sliced


def _target_samples(collate_strategy: str, sample_rate: int) -> int:
    assert collate_strategy == "pad" or re.match(r"constant_length-\d+-ms", collate_strategy), \
        "collate_strategy must be 'pad' or match the pattern 'constant_length-XXX-ms'"      # bwe.py:77-79
    return int(sample_rate * int(collate_strategy.split("-")[1]) / 1000)


def _arrays(batch: Sequence[dict], key: str) -> List[torch.Tensor]:
    return [item[key]["array"] if isinstance(item[key], dict) else item[key] for item in batch]


def _on_gpu(t: torch.Tensor) -> bool:
    return t.is_cuda


def _pad_batch(items: List[torch.Tensor]) -> torch.Tensor:
    return pad_sequence(items, batch_first=True, padding_value=0.0).unsqueeze(1)


def _constant_length(body: List[torch.Tensor], air: List[torch.Tensor], samples: int, deterministic: bool):
    outs_b, outs_a = [], []
    for b, a in zip(body, air):
        pb, pa = set_audio_duration(audio=b, desired_samples=samples, audio_bis=a, deterministic=deterministic)
        outs_b.append(pb.unsqueeze(0))
        outs_a.append(pa.unsqueeze(0))
    return torch.stack(outs_b, dim=0), torch.stack(outs_a, dim=0)


def _augment(body: torch.Tensor, air: torch.Tensor, deterministic: bool, data_augmentation, sample_rate: int):
    """`if deterministic is False: with torch.no_grad(): data_augmentation(body, air)` (`bwe.py:282-287`).  The
    reference's default hook (`WaveformDataAugmentation(sample_rate)`, probability 0) is a no-op that still draws one
    `rand(1)` for its gate: without a hook that draw is taken here, so the generator state stays the reference's."""
    if deterministic is not False:
        return body, air
    if data_augmentation is None:
        torch.rand(1)
        return body, air
    with torch.no_grad():
        return data_augmentation(body, air)


def bwe_collate(batch: Sequence[dict], sample_rate: int = 16000, collate_strategy: str = "constant_length-3000-ms",
                deterministic: bool = False, data_augmentation=None) -> Dict[str, torch.Tensor]:
    """`BWELightningDataModule.data_collator` (`bwe.py:232-293`): items carry 'audio_body_conducted' /
    'audio_airborne' as 1-D tensors (or HF `{"array": tensor}` dicts); returns the (B, 1, samples) pair the training
    step consumes.  `data_augmentation` is the datamodule's hook (`bwe.py:32,75-81`), applied when not deterministic."""
    body, air = _arrays(batch, "audio_body_conducted"), _arrays(batch, "audio_airborne")
    if collate_strategy == "pad":
        b, a = _pad_batch(body), _pad_batch(air)
    else:
        b, a = _constant_length(body, air, _target_samples(collate_strategy, sample_rate), deterministic)
    b, a = _augment(b, a, deterministic, data_augmentation, sample_rate)
    return {"audio_body_conducted": b, "audio_airborne": a}


def noisybwe_collate(batch: Sequence[dict], sample_rate: int = 16000,
                     collate_strategy: str = "constant_length-3000-ms", deterministic: bool = False,
                     data_augmentation=None) -> Dict[str, torch.Tensor]:
    """`NoisyBWELightningDataModule.data_collator` (`noisybwe.py:225-300`): real noisy recordings (no airborne
    reference) are padded; otherwise body-conducted speech + a random segment of the speech-free noise recording,
    then the joint crop / pad of (corrupted, airborne)."""
    body = _arrays(batch, "audio_body_conducted")
    if "audio_airborne" not in batch[0]:
        return {"audio_body_conducted": _pad_batch(body)}
    air = _arrays(batch, "audio_airborne")
    noise = _arrays(batch, "audio_body_conducted_speechless_noisy")
    samples = None if collate_strategy == "pad" else _target_samples(collate_strategy, sample_rate)
    fused = (samples is not None and _on_gpu(body[0]) and len({t.shape for t in body}) == 1
             and len({t.shape for t in noise}) == 1 and body[0].size(0) >= samples
             and all(a.shape == body[0].shape for a in air))
    if fused:
        # same draws in the same order as the unfused path: all noise starts first, then one crop offset per item
        from . import ops
        starts = plan_noise_starts(body, noise)
        offs = [plan_duration(body[0].size(0), samples, deterministic) for _ in body]
        dev = body[0].device
        ob, oa = ops.noise_mix_crop(torch.stack(body).unsqueeze(1), torch.stack(air).unsqueeze(1),
                                    torch.stack(noise).unsqueeze(1),
                                    torch.tensor(starts, dtype=torch.int32, device=dev),
                                    torch.tensor(offs, dtype=torch.int32, device=dev), samples)
        ob, oa = _augment(ob, oa, deterministic, data_augmentation, sample_rate)
        return {"audio_body_conducted": ob, "audio_airborne": oa}
    corrupted, _ = mix_speech_and_noise_without_rescaling(body, noise)
    if samples is None:
        b, a = _pad_batch(corrupted), _pad_batch(air)
    else:
        b, a = _constant_length(corrupted, air, samples, deterministic)
    b, a = _augment(b, a, deterministic, data_augmentation, sample_rate)
    return {"audio_body_conducted": b, "audio_airborne": a}
