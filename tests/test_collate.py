"""Batch assembly (vibravox_b200/collate.py) against the reference's own helpers and collators, draw for draw:
both sides are seeded identically and must consume torch's global RNG in the same order.  The reference package is
only present in the build container; without it the property checks still run."""
import os
import sys

import pytest
import torch

from vibravox_b200 import collate as C

REF = "/root/reference"


@pytest.fixture(scope="module")
def ref_utils():
    if not os.path.isdir(os.path.join(REF, "vibravox")):
        pytest.skip("reference package not present")
    sys.path.insert(0, REF)
    try:
        import vibravox.utils as U
    except Exception as exc:
        pytest.skip(f"vibravox.utils not importable: {exc}")
    finally:
        sys.path.remove(REF)
    return U


def items(n, lo, hi, noise=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        L = int(torch.randint(lo, hi, (1,), generator=g))
        it = {"audio_body_conducted": {"array": torch.randn(L, generator=g)},
              "audio_airborne": {"array": torch.randn(L, generator=g)}}
        if noise:
            it["audio_body_conducted_speechless_noisy"] = {"array": torch.randn(noise, generator=g)}
        out.append(it)
    return out


def test_pad_audio_keeps_the_reference_precedence_quirk(ref_utils):
    for L, D in ((10, 10), (7, 12), (1, 5), (8, 9), (0, 4)):
        x = torch.arange(1.0, L + 1).view(1, L)
        want = ref_utils.pad_audio(x, D)
        got = C.pad_audio(x, D)
        assert got.shape == want.shape == (1, D) and torch.equal(got, want), (L, D)
    with pytest.raises(AssertionError):
        C.pad_audio(torch.zeros(9), 8)
    with pytest.raises(AssertionError):
        C.slice_audio(torch.zeros(7), 8, 0)


@pytest.mark.parametrize("deterministic", [True, False])
def test_set_audio_duration_draws_like_the_reference(ref_utils, deterministic):
    torch.manual_seed(3)
    a, b = torch.randn(1, 5000), torch.randn(1, 5000)
    for D in (3000, 5000, 6000):
        torch.manual_seed(11)
        want = ref_utils.set_audio_duration(a, D, audio_bis=b, deterministic=deterministic)
        w1 = ref_utils.set_audio_duration(a, D, deterministic=deterministic)
        torch.manual_seed(11)
        got = C.set_audio_duration(a, D, audio_bis=b, deterministic=deterministic)
        g1 = C.set_audio_duration(a, D, deterministic=deterministic)
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and torch.equal(g1, w1)
    with pytest.raises(AssertionError):
        C.set_audio_duration(a, 100, audio_bis=torch.zeros(1, 4999))


def test_mix_without_rescaling_matches_the_reference(ref_utils):
    g = torch.Generator().manual_seed(1)
    speech = [torch.randn(n, generator=g) for n in (4000, 2500, 3999)]
    noise = [torch.randn(9000, generator=g) for _ in speech]
    torch.manual_seed(5)
    want, wn = ref_utils.mix_speech_and_noise_without_rescaling(speech, noise)
    torch.manual_seed(5)
    got, gn = C.mix_speech_and_noise_without_rescaling(speech, noise)
    assert all(torch.equal(a, b) for a, b in zip(got, want)) and all(torch.equal(a, b) for a, b in zip(gn, wn))
    for bad, err in (((tuple(speech), noise), TypeError), ((speech, noise[:2]), ValueError),
                     (([speech[0].view(1, -1)], [noise[0]]), ValueError), (([noise[0]], [speech[0]]), ValueError)):
        with pytest.raises(err):
            C.mix_speech_and_noise_without_rescaling(*bad)
        with pytest.raises(err):
            ref_utils.mix_speech_and_noise_without_rescaling(*bad)


def _ref_collate(U, batch, sample_rate, strategy, deterministic, noisy):
    """The reference collators' arithmetic (bwe.py:232-293 / noisybwe.py:225-300) spelled with the REFERENCE's own
    helpers; the datamodule classes themselves need `lightning` + `datasets` to import."""
    from torch.nn.utils.rnn import pad_sequence
    body = [it["audio_body_conducted"]["array"] for it in batch]
    air = [it["audio_airborne"]["array"] for it in batch]
    if noisy:
        body, _ = U.mix_speech_and_noise_without_rescaling(
            body, [it["audio_body_conducted_speechless_noisy"]["array"] for it in batch])
    if strategy == "pad":
        return (pad_sequence(body, batch_first=True).unsqueeze(1), pad_sequence(air, batch_first=True).unsqueeze(1))
    samples = int(sample_rate * int(strategy.split("-")[1]) / 1000)
    pairs = [U.set_audio_duration(audio=b, desired_samples=samples, audio_bis=a, deterministic=deterministic)
             for b, a in zip(body, air)]
    return (torch.stack([p[0].unsqueeze(0) for p in pairs]), torch.stack([p[1].unsqueeze(0) for p in pairs]))


def _ref_aug(**kw):
    sys.path.insert(0, REF)
    try:
        from vibravox.torch_modules.dsp.data_augmentation import WaveformDataAugmentation
    finally:
        sys.path.remove(REF)
    return WaveformDataAugmentation(16000, **kw)


@pytest.mark.parametrize("noisy", [False, True])
@pytest.mark.parametrize("strategy", ["pad", "constant_length-250-ms"])
@pytest.mark.parametrize("deterministic", [True, False])
def test_collators_match_the_reference(ref_utils, noisy, strategy, deterministic):
    batch = items(5, 2500, 6000, noise=12000 if noisy else None, seed=7)       # target 4000: some crop, some pad
    for aug in ({}, dict(p_data_augmentation=1, p_speed_perturbation=0, p_pitch_shift=0, p_time_masking=1)):
        torch.manual_seed(21)
        wb, wa = _ref_collate(ref_utils, batch, 16000, strategy, deterministic, noisy)
        if deterministic is False:
            with torch.no_grad():
                wb, wa = _ref_aug(**aug)(wb, wa)
        w_next = torch.rand(1)                               # the generator must be left in the same state
        torch.manual_seed(21)
        got = (C.noisybwe_collate if noisy else C.bwe_collate)(
            batch, 16000, strategy, deterministic, data_augmentation=_ref_aug(**aug) if aug else None)   # the hook = the reference's object
        assert torch.equal(torch.rand(1), w_next)
        assert got["audio_body_conducted"].shape == wb.shape and wb.dim() == 3 and wb.shape[1] == 1
        assert torch.equal(got["audio_body_conducted"], wb) and torch.equal(got["audio_airborne"], wa)
        if aug and deterministic is False:
            assert (wb == 0).all(dim=1).any()                # the masked block is really there


def test_noisy_collate_without_reference_signal_only_pads():
    batch = [{"audio_body_conducted": torch.randn(n)} for n in (30, 50, 40)]
    out = C.noisybwe_collate(batch, 16000, "pad")
    assert list(out) == ["audio_body_conducted"] and out["audio_body_conducted"].shape == (3, 1, 50)
    with pytest.raises(AssertionError):
        C.bwe_collate(items(1, 10, 20), 16000, "constant-250")


def test_fused_device_path_takes_the_same_draws(monkeypatch):
    """Equal-length device batches go through ONE vbx_noise_mix_crop launch; the plan (noise starts, then crop
    offsets) must be the one the item-by-item path draws.  The kernel itself is checked on the GPU
    (tests/test_gpu_kernels.py::test_noise_mix_crop); here a host stand-in with its contract takes its place."""
    from vibravox_b200 import ops
    calls = []

    def stand_in(body, air, noise, start, off, length):
        calls.append((start.tolist(), off.tolist()))
        ob = torch.stack([(body[b, 0] + noise[b, 0, int(s): int(s) + body.shape[-1]])[int(o): int(o) + length]
                          for b, (s, o) in enumerate(zip(start, off))]).unsqueeze(1)
        oa = torch.stack([air[b, 0, int(o): int(o) + length] for b, o in enumerate(off)]).unsqueeze(1)
        return ob, oa

    g = torch.Generator().manual_seed(2)
    batch = [{"audio_body_conducted": torch.randn(6000, generator=g), "audio_airborne": torch.randn(6000, generator=g),
              "audio_body_conducted_speechless_noisy": torch.randn(20000, generator=g)} for _ in range(4)]
    torch.manual_seed(9)
    want = C.noisybwe_collate(batch, 16000, "constant_length-250-ms", deterministic=False)
    monkeypatch.setattr(C, "_on_gpu", lambda t: True)
    monkeypatch.setattr(ops, "noise_mix_crop", stand_in)
    torch.manual_seed(9)
    got = C.noisybwe_collate(batch, 16000, "constant_length-250-ms", deterministic=False)
    assert len(calls) == 1 and len(calls[0][0]) == 4
    assert torch.equal(got["audio_body_conducted"], want["audio_body_conducted"])
    assert torch.equal(got["audio_airborne"], want["audio_airborne"])
