"""The CUDA conv tile code (gemm_conv.cuh + conv_plan.h) executed on the CPU by tests/emu, checked
against torch's fp64 conv1d / autograd: forward, dgrad (phase-decomposed, reflect images),
wgrad (split reduction) and the col2im scatter form, over the layer geometries of the EBEN
generator / discriminators / STFT-as-conv (reduced lengths)."""
import pytest
import torch
import torch.nn.functional as F

from emu_util import ConvDesc, emu, ref_padded, scatter_weight, tout, transpose_weight

CASES = [
    # B, Cin, Cout, Tin, K, s, d, pad, refl, groups
    (2, 2, 32, 100, 3, 1, 1, 1, 1, 1),        # first_conv
    (2, 32, 32, 300, 3, 1, 9, 9, 9, 1),       # residual dilated conv d=9
    (1, 32, 32, 77, 1, 1, 1, 0, 0, 1),        # residual pointwise conv
    (1, 32, 64, 301, 4, 2, 1, 1, 1, 1),       # EncBlock s=2 (odd length)
    (2, 16, 32, 257, 16, 8, 1, 7, 7, 1),      # EncBlock s=8
    (1, 64, 32, 60, 16, 8, 1, 4, 0, 1),       # conv view of DecBlock ConvTranspose s=8, pad 4
    (2, 8, 24, 130, 3, 1, 2, 2, 1, 4),        # PQMF-disc L0: ReflectionPad1d(1) + zero pad 1, dilation 2
    (2, 24, 48, 131, 7, 2, 2, 3, 0, 4),       # PQMF-disc strided, gcd(dil, stride) = 2
    (2, 24, 48, 131, 7, 2, 3, 3, 0, 4),
    (2, 24, 48, 130, 7, 2, 1, 3, 0, 4),
    (1, 1, 16, 200, 15, 1, 1, 7, 7, 1),       # MelGAN L0
    (2, 16, 64, 403, 41, 4, 1, 20, 0, 4),     # MelGAN L1
    (1, 160, 1, 50, 3, 1, 1, 1, 0, 1),        # certainty conv
    (1, 1, 34, 700, 120, 24, 1, 60, 60, 1),   # STFT-as-conv (lanes walk the taps)
    (2, 200, 136, 37, 5, 1, 1, 2, 0, 1),      # dense, >1 M tile
    (3, 4, 1, 100, 32, 4, 1, 31, 0, 1),       # PQMF-synthesis-shaped
    (1, 1, 1, 300, 101, 1, 1, 50, 0, 1),      # A-weighting FIR
]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_emulated_kernels_match_torch(case):
    B, Cin, Cout, Tin, K, s, d, pad, refl, groups = case
    torch.manual_seed(sum(case))
    To = tout(Tin, K, s, d, pad)
    desc = ConvDesc(B, Cin, Cout, Tin, To, K, s, d, pad, refl, groups)
    x = torch.randn(B, Cin, Tin)
    w = torch.randn(Cout, Cin // groups, K) / (Cin // groups * K) ** 0.5
    bias, res = torch.randn(Cout), torch.randn(B, Cout, To)
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    pre = F.conv1d(ref_padded(x64, pad, refl), w64, bias.double(), s, 0, d, groups)
    want = F.leaky_relu(pre, 0.2) + res.double()
    y = torch.full((B, Cout, To), float("nan"))
    mask = torch.zeros(B, Cout, To, dtype=torch.uint8)
    emu(0, desc, x, w, y, bias=bias, res=res, mask=mask, slope=0.2)
    assert (y.double() - want).abs().max() < 2e-5
    assert (mask.bool() != (pre > 0)).sum() <= 2      # only |pre| ~ 1e-7 can flip

    dy = torch.randn(B, Cout, To)
    plain = F.conv1d(ref_padded(x64, pad, refl), w64, None, s, 0, d, groups)
    gx, gw = torch.autograd.grad(plain, (x64, w64), dy.double())
    dx = torch.full((B, Cin, Tin), float("nan"))
    r2 = torch.randn(B, Cin, Tin)
    emu(1, desc, dy, transpose_weight(w, groups), dx, res=r2)
    assert (dx.double() - (gx + r2.double())).abs().max() < 2e-5
    dw = torch.zeros_like(w)
    emu(2, desc, x, dy, dw)
    assert (dw.double() - gw).abs().max() / gw.abs().max() < 5e-6
    dx2 = torch.zeros(B, Cin, Tin)
    emu(3, desc, dy, scatter_weight(w, groups), dx2)
    assert (dx2.double() - gx).abs().max() < 2e-5


def test_dgrad_gate_stage_matches_the_unfused_passes():
    """vbx_epilogue's gate stage: (dgrad + res + feature-matching gradient of y) * LeakyReLU'(y), bit for bit what the three
    separate passes it replaces compute (aten::add, L1-pair backward, LeakyReLU backward)."""
    torch.manual_seed(3)
    B, Cin, Cout, Tin, K, groups = 2, 12, 24, 50, 7, 4
    To = tout(Tin, K, 2, 1, 3)
    desc = ConvDesc(B, Cin, Cout, Tin, To, K, 2, 1, 3, 0, groups)
    w, dy = torch.randn(Cout, Cin // groups, K), torch.randn(B, Cout, To)
    y, other = torch.randn(B, Cin, Tin), torch.randn(B, Cin, Tin)
    y[0, 0, :5] = other[0, 0, :5]                         # sign(0) terms
    y[0, 1, :5] = 0.0
    coef = torch.tensor([0.37, 0.011])
    plain = torch.empty(B, Cin, Tin)
    emu(1, desc, dy, transpose_weight(w, groups), plain)
    fm = coef[0] * torch.sign(y - other) - coef[1] * torch.sign(y)
    want_fm = (plain + fm) * torch.where(y > 0, 1.0, 0.2)
    want_plain = plain * torch.where(y > 0, 1.0, 0.2)
    got = torch.empty(B, Cin, Tin)
    emu(1, desc, dy, transpose_weight(w, groups), got, gate=(y, 0.2, other, coef))
    assert torch.equal(got, want_fm)
    emu(1, desc, dy, transpose_weight(w, groups), got, gate=(y, 0.2, None, None))
    assert torch.equal(got, want_plain)


def test_dgrad_accumulates_with_beta():
    torch.manual_seed(0)
    B, Cin, Cout, Tin, K = 1, 8, 8, 64, 3
    desc = ConvDesc(B, Cin, Cout, Tin, Tin, K, 1, 1, 1, 1, 1)
    w, dy, old = torch.randn(Cout, Cin, K), torch.randn(B, Cout, Tin), torch.randn(B, Cin, Tin)
    fresh = torch.empty(B, Cin, Tin)
    emu(1, desc, dy, transpose_weight(w, 1), fresh)
    acc = old.clone()
    emu(1, desc, dy, transpose_weight(w, 1), acc, beta=1.0)
    assert torch.allclose(acc, fresh + old, atol=1e-6)


def test_bad_descriptor_is_rejected():
    import ctypes
    from emu_util import Epilogue, lib
    bad = ConvDesc(1, 4, 4, 16, 99, 3, 1, 1, 1, 0, 1)     # wrong Tout
    e = Epilogue(None, None, None, 1.0, 0.0)
    t = torch.zeros(4096)
    rc = lib().emu_conv(0, ctypes.byref(bad), ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(t.data_ptr()),
                        ctypes.byref(e), ctypes.c_void_p(t.data_ptr()))
    assert rc == -1
