"""The Python side of the product (autograd.Functions, drop-in modules, FlatAdam gradient slots,
the training-step schedule) driven with torch-CPU stand-ins for the kernels (tests/cpu_shim.py)
and checked against the oracle / the reference goldens.  The kernels themselves are checked on the
GPU (-m gpu) and through the CPU emulator (test_emu_conv.py)."""
import os

import pytest
import torch

from cpu_shim import cpu_ops
from oracle import eben_oracle as O


def relerr(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_modules_forward_backward_match_oracle():
    from vibravox_b200.torch_modules.dnn.eben_discriminator import DiscriminatorEBENMultiScales
    from vibravox_b200.torch_modules.dnn.eben_generator import EBENGenerator
    from vibravox_b200.torch_modules.losses.feature_loss import FeatureLossForDiscriminatorMelganMultiScales
    from vibravox_b200.torch_modules.losses.hinge_loss import HingeLossForDiscriminatorMelganMultiScales
    torch.manual_seed(7)
    G, D = EBENGenerator(m=4, n=32, p=1), DiscriminatorEBENMultiScales(q=3, min_channels=24)
    gs = {k: v.detach().clone() for k, v in G.state_dict().items()}
    ds = {k: v.detach().clone() for k, v in D.state_dict().items()}
    body, air = O.synthetic_pairs(1, 4000, seed=3)
    g32 = {k: v.clone().requires_grad_(not k.startswith("pqmf.")) for k, v in gs.items()}
    d32 = {k: v.clone().requires_grad_(True) for k, v in ds.items()}
    x = O.cut_to_valid_length(body, 32, 4)
    a = O.cut_to_valid_length(air, 32, 4)
    y0, b0 = O.generator_forward(g32, x, 1)
    e0 = O.discriminator_forward(d32, b0, y0, 3, 24)
    r0 = O.discriminator_forward(d32, O.pqmf_analysis(a, g32["pqmf.analysis_weights"]), a, 3, 24)
    loss0 = O.feature_matching_loss(e0, r0) + O.hinge_loss(e0, 1) + O.hinge_loss(r0, -1)
    names = [k for k in g32 if g32[k].requires_grad]
    want = torch.autograd.grad(loss0, [g32[k] for k in names] + list(d32.values()))
    with cpu_ops():
        y, bands = G(G.cut_to_valid_length(body))
        assert relerr(y, y0) < 1e-5 and relerr(bands, b0) < 1e-5
        e = D(bands=bands, audio=y)
        r = D(bands=G.pqmf(a, "analysis"), audio=a)
        hinge = HingeLossForDiscriminatorMelganMultiScales()
        loss = FeatureLossForDiscriminatorMelganMultiScales()(e, r) + hinge(e, 1) + hinge(r, -1)
        assert float(loss) == pytest.approx(float(loss0), rel=1e-5)
        gp, dp = dict(G.named_parameters()), dict(D.named_parameters())
        got = torch.autograd.grad(loss, [gp[k] for k in names] + [dp[k] for k in d32])
    for n, u, v in zip(names + list(d32), got, want):
        assert relerr(u, v) < 1e-2, (n, relerr(u, v))   # fp32 vs fp32, different summation order


@pytest.mark.parametrize("schedule", ["shared", "reference"])
def test_training_step_schedule_matches_reference_golden(golden_dir, schedule):
    import vibravox_b200
    gold = torch.load(os.path.join(golden_dir, "train_step.pt"))
    body, air = O.synthetic_pairs(gold["B"], gold["S"], seed=gold["data_seed"])
    with cpu_ops():
        lm = vibravox_b200.build_model(seed=gold["model_seed"], device="cpu")
        lm.schedule = schedule
        for it in range(2):
            out = lm.training_step({"audio_body_conducted": body, "audio_airborne": air})
            want = gold["steps"][it]
            for k, v in want["logs"].items():
                got = float(lm.logged["train/" + k])
                assert got == pytest.approx(v, rel=3e-4, abs=3e-5), (it, k, got, v)
            for a, b in zip(lm.atomic_norms_old.tolist(), want["norms_old"]):
                assert a == pytest.approx(b, rel=5e-4)
        # every trainable parameter moved, gradient bucket is clean, D untouched flags restored
        assert all(p.requires_grad for p in lm.discriminator.parameters())
        assert float(lm.generator_optimizer.grad.abs().sum()) == 0.0
        assert int(lm.generator_optimizer.step_count[0]) == 2 and int(lm.discriminator_optimizer.step_count[0]) == 2
    gsd = lm.generator.state_dict()
    for k, v in gold["g_param_sums_after"].items():
        n = gsd[k].numel()
        assert abs(float(gsd[k].double().sum()) - v) <= 5e-3 * n ** 0.5 + 1e-3, k
