"""-m gpu: the reference's own unit tests for this path (tests/torch_modules/*_test.py, fixtures of
tests/conftest.py:27-44,186-199 in /root/reference) replayed against the drop-in modules: shapes / types on
randn(4, 1, 15679), parameter counts, 0-d losses for targets +-1 - plus the edge cases of the domain (batch 1,
minimum valid length, p in {1,2,4}, q in {3,4}, non-contiguous input, CPU tensors refused)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def sample():
    torch.manual_seed(0)
    return torch.randn(4, 1, 15679)           # conftest.py: batch_size=4, time_len=15679


def test_eben_generator_output_shape_and_params(sample):
    from vibravox_b200.torch_modules.dnn.eben_generator import EBENGenerator
    G = EBENGenerator(m=4, n=32, p=1).to(DEV)                    # conftest.py:195-199
    cut = G.cut_to_valid_length(sample.to(DEV))
    enhanced, decomposed = G(cut)
    assert enhanced.shape == cut.shape                           # eben_generator_test.py:2-8
    assert decomposed.shape == (4, 4, (cut.shape[2] + 32) // 4)
    assert sum(p.numel() for p in G.parameters()) > 1e3          # eben_generator_test.py:10-14
    assert torch.isfinite(enhanced).all()


def test_discriminator_structure_and_losses(sample):
    from vibravox_b200.torch_modules.dnn.eben_discriminator import DiscriminatorEBENMultiScales
    from vibravox_b200.torch_modules.dnn.eben_generator import EBENGenerator
    from vibravox_b200.torch_modules.losses.feature_loss import FeatureLossForDiscriminatorMelganMultiScales
    from vibravox_b200.torch_modules.losses.hinge_loss import HingeLossForDiscriminatorMelganMultiScales
    G = EBENGenerator(m=4, n=32, p=1).to(DEV)
    D = DiscriminatorEBENMultiScales(q=3, min_channels=24).to(DEV)
    with torch.no_grad():
        cut = G.cut_to_valid_length(sample.to(DEV))
        enhanced, bands = G(cut)
        ref_bands = G.pqmf(cut, "analysis")
        emb_a, emb_b = D(bands=bands, audio=enhanced), D(bands=ref_bands, audio=cut)
    assert isinstance(emb_a, list) and [len(s) for s in emb_a] == [9, 9, 9, 8]
    assert all(isinstance(t, torch.Tensor) for s in emb_a for t in s)
    assert emb_a[0][0].shape == (4, 3, bands.shape[2]) and emb_a[3][0].shape == enhanced.shape
    assert sum(p.numel() for p in D.parameters()) > 1e3
    fm = FeatureLossForDiscriminatorMelganMultiScales()(emb_a, emb_b)
    assert fm.shape == torch.Size([])                            # feature_loss_test.py
    for target in (1, -1):                                       # hinge_loss_test.py
        assert HingeLossForDiscriminatorMelganMultiScales()(emb_a, target).shape == torch.Size([])


@pytest.mark.parametrize("p,q,B,L", [(1, 3, 1, 2600), (2, 4, 1, 2559), (4, 4, 3, 3000), (2, 3, 5, 4001)])
def test_edge_shapes_match_oracle(p, q, B, L):
    """batch 1, the shortest length the reference's discriminator accepts (2528 samples: below that its
    dilation-3 branch runs out of samples and PyTorch raises), odd lengths, every p / q the configs use."""
    from oracle import eben_oracle as O
    from vibravox_b200.torch_modules.dnn.eben_discriminator import DiscriminatorEBENMultiScales
    from vibravox_b200.torch_modules.dnn.eben_generator import EBENGenerator
    torch.manual_seed(p * 10 + q)
    G, D = EBENGenerator(m=4, n=32, p=p), DiscriminatorEBENMultiScales(q=q, min_channels=24)
    gs = {k: v.detach().clone() for k, v in G.state_dict().items()}
    ds = {k: v.detach().clone() for k, v in D.state_dict().items()}
    x = 0.3 * torch.randn(B, 1, L)
    xc = O.cut_to_valid_length(x, 32, 4)
    with torch.no_grad():
        y0, b0 = O.generator_forward(gs, xc, p)
        e0 = O.discriminator_forward(ds, b0, y0, q, 24)
        G, D = G.to(DEV), D.to(DEV)
        big = torch.zeros(B, 1, L + 7, device=DEV)
        big[:, :, :L] = x.to(DEV)
        view = big[:, :, :L]                                      # non-contiguous input
        cut = G.cut_to_valid_length(view)
        assert cut.shape == xc.shape
        y, b = G(cut)
        e = D(bands=b, audio=y)
    assert (y.cpu() - y0).abs().max() < 1e-4 * max(1.0, float(y0.abs().max()))
    assert (b.cpu() - b0).abs().max() < 1e-4
    for sa, sb in zip(e, e0):
        for ta, tb in zip(sa, sb):
            assert ta.shape == tb.shape
            assert (ta.cpu() - tb).abs().max() < 2e-4 * max(1.0, float(tb.abs().max()))


def test_cpu_tensors_are_refused_not_silently_computed():
    from vibravox_b200 import _lib
    from vibravox_b200.torch_modules.dsp.pqmf import PseudoQMFBanks
    with pytest.raises(_lib.VbxError):
        PseudoQMFBanks(4, 32)(torch.randn(1, 1, 1000), "analysis")
    with pytest.raises(ValueError):
        PseudoQMFBanks(4, 32).to(DEV)(torch.randn(1, 1, 1000, device=DEV), "decompose")


def test_too_short_input_raises_like_the_reference():
    """Below 2528 samples the reference raises inside conv1d ("Kernel size can't be greater than actual input
    size"); the drop-in raises VbxError from the descriptor check instead of computing garbage."""
    from vibravox_b200 import _lib
    from vibravox_b200.torch_modules.dnn.eben_discriminator import DiscriminatorEBENMultiScales
    from vibravox_b200.torch_modules.dnn.eben_generator import EBENGenerator
    G, D = EBENGenerator(m=4, n=32, p=2).to(DEV), DiscriminatorEBENMultiScales(q=4, min_channels=24).to(DEV)
    x = G.cut_to_valid_length(torch.randn(1, 1, 1000, device=DEV))
    with torch.no_grad():
        y, b = G(x)
        with pytest.raises((_lib.VbxError, AssertionError)):
            D(bands=b, audio=y)
            torch.cuda.synchronize()
