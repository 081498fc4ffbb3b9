"""The CPU oracle (oracle/eben_oracle.py) against the fixtures minted from the reference's own
modules by oracle/make_goldens.py (the reference itself is not available on the GPU box)."""
import os

import pytest
import torch

from oracle import eben_oracle as O


def relerr(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_config1_forward_matches_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "cfg1_forward.pt"))
    torch.manual_seed(gold["seed"])
    gs = O.init_generator_state(gold["m"], gold["n"], gold["p"])
    ds = O.init_discriminator_state(gold["q"], gold["min_channels"])
    x = torch.randn(1, 1, 16000)
    assert torch.equal(x[0, 0, :8], gold["x_head"])
    assert torch.equal(gs["pqmf.analysis_weights"], gold["analysis_weights"])
    with torch.no_grad():
        y, bands = O.generator_forward(gs, O.cut_to_valid_length(x, gold["n"], gold["m"]), gold["p"])
        emb = O.discriminator_forward(ds, bands, y, gold["q"], gold["min_channels"])
    assert relerr(y, gold["enhanced"]) < 2e-6 and relerr(bands, gold["bands"]) < 2e-6
    for s, shapes, cert in zip(emb, gold["emb_shapes"], gold["certainties"]):
        assert [tuple(t.shape) for t in s] == shapes
        assert relerr(s[-1], cert) < 5e-6


def test_training_step_matches_reference_logs(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "train_step.pt"))
    body, air = O.synthetic_pairs(gold["B"], gold["S"], seed=gold["data_seed"])
    step = O.OracleEBENStep(seed=gold["model_seed"])
    logs = step.step(body, air)
    for k, v in gold["steps"][0]["logs"].items():
        assert logs[k] == pytest.approx(v, rel=5e-5, abs=5e-6), k
    assert torch.allclose(step.last["enhanced"][0, 0, :64], gold["steps"][0]["enhanced_head"], atol=2e-6)


def test_pqmf_reconstruction_snr():
    wa, ws, cutoff = O.pqmf_design(4, 32)
    assert cutoff == 0.15886658430099487
    torch.manual_seed(3)
    sig = torch.rand(4, 1, 48008)
    rec = O.pqmf_synthesis(O.pqmf_analysis(sig, wa), ws).sum(1, keepdim=True)
    snr = 10 * torch.log10((rec ** 2).mean() / ((sig - rec) ** 2).mean())
    assert snr > 50
