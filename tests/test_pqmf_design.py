"""PQMF filter design (init-time host arithmetic) against the reference's own design: the cut-off ratios of four
banks recorded from /root/reference by oracle/make_goldens.py (oracle_pin_report.pt), the (4, 32) taps bit for bit
(cfg1_forward.pt), and the oracle's restatement of the same design."""
import os

import pytest
import torch


@pytest.mark.parametrize("m,n", [(4, 32), (4, 64), (8, 64), (2, 16)])
def test_cutoff_ratio_equals_the_reference_design(golden_dir, m, n):
    from oracle import eben_oracle as O
    from vibravox_b200.torch_modules.dsp.pqmf import PseudoQMFBanks
    pin = torch.load(os.path.join(golden_dir, "oracle_pin_report.pt"))
    bank = PseudoQMFBanks(decimation=m, kernel_size=n)
    assert bank._cutoff_ratio == pin[f"pqmf_{m}_{n}_cutoff"]            # bit for bit (a Python float)
    wa, ws, cut = O.pqmf_design(m, n)
    assert cut == bank._cutoff_ratio
    assert torch.equal(wa, bank.analysis_weights.data) and torch.equal(ws, bank.synthesis_weights.data)
    assert not bank.analysis_weights.requires_grad and not bank.synthesis_weights.requires_grad


def test_default_bank_taps_equal_the_reference_taps(golden_dir):
    from vibravox_b200.torch_modules.dsp.pqmf import PseudoQMFBanks
    gold = torch.load(os.path.join(golden_dir, "cfg1_forward.pt"))
    bank = PseudoQMFBanks(decimation=4, kernel_size=32)
    assert torch.equal(bank.analysis_weights.data, gold["analysis_weights"])
    assert torch.equal(bank.synthesis_weights.data, gold["synthesis_weights"])
    assert bank._cutoff_ratio == pytest.approx(0.15886658430099487, abs=0)


def test_kernel_size_must_be_a_multiple_of_four_bands():
    from vibravox_b200.torch_modules.dsp.pqmf import PseudoQMFBanks
    with pytest.raises(AssertionError):                                       # pqmf.py:42
        PseudoQMFBanks(decimation=4, kernel_size=40)
