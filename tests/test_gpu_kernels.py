"""-m gpu: every kernel family of libvbx_b200.so, called through the C ABI (vibravox_b200.ops),
against plain PyTorch fp64 references of the same op on the host."""
import pytest
import torch
import torch.nn.functional as F

from emu_util import ref_padded, scatter_weight, tout, transpose_weight
from test_emu_conv import CASES

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cuda(*ts):
    return [t.to(DEV) for t in ts]


# one input channel per group -> direct_conv.cu (fwd and input gradient), several tiles per row
DIRECT_CASES = [
    (2, 4, 24, 1500, 3, 1, 2, 2, 1, 4),       # PQMF-disc L0 on all four bands
    (2, 4, 24, 1111, 3, 1, 3, 3, 1, 4),
    (2, 1, 16, 2500, 15, 1, 1, 7, 7, 1),      # MelGAN L0, 3 forward tiles / 5 gradient tiles
    (3, 1, 1, 2100, 101, 1, 1, 50, 0, 1),     # A-weighting FIR
    (2, 3, 12, 1100, 5, 2, 1, 2, 2, 3),       # strided: direct forward, GEMM input gradient
    (2, 1, 16, 2503, 15, 1, 1, 7, 7, 1),      # quad kernels: ragged last quad (scalar stores; old input-gradient kernel)
    (2, 1, 16, 1000, 15, 1, 1, 7, 0, 1),      # quad kernels with a zero halo
    (3, 768, 1, 375, 3, 1, 1, 1, 0, 1),       # certainty conv: skinny_fwd_kernel (8 channel slices) / skinny_wgrad_kernel
    (2, 32, 4, 1001, 3, 1, 1, 1, 1, 1),       # generator last conv: 4 output channels, ragged quads, reflect halo
]


@pytest.mark.parametrize("case", CASES + DIRECT_CASES, ids=[str(c) for c in CASES + DIRECT_CASES])
def test_conv_family_matches_torch(case):
    from vibravox_b200 import ops
    B, Cin, Cout, Tin, K, s, d, pad, refl, groups = case
    geom = ops.ConvGeom(Cin, Cout, K, s, d, pad, refl, groups)
    torch.manual_seed(sum(case))
    To = tout(Tin, K, s, d, pad)
    x = torch.randn(B, Cin, Tin)
    w = torch.randn(Cout, Cin // groups, K) / (Cin // groups * K) ** 0.5
    bias, res = torch.randn(Cout), torch.randn(B, Cout, To)
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    pre = F.conv1d(ref_padded(x64, pad, refl), w64, bias.double(), s, 0, d, groups)
    want = F.leaky_relu(pre, 0.2) + res.double()
    xc, wc, bc, rc = cuda(x, w, bias, res)
    y, mask = ops.conv1d_fwd(xc, wc, geom, bias=bc, res=rc, slope=0.2, want_mask=True)
    assert (y.cpu().double() - want).abs().max() < 2e-5
    assert (mask.cpu().bool() != (pre > 0)).sum() <= 2
    dy = torch.randn(B, Cout, To)
    plain = F.conv1d(ref_padded(x64, pad, refl), w64, None, s, 0, d, groups)
    gx, gw = torch.autograd.grad(plain, (x64, w64), dy.double())
    dyc = dy.to(DEV)
    wt = ops.transpose_weight(wc, groups)
    assert torch.equal(wt.cpu().view(-1), transpose_weight(w, groups).view(-1))
    r2 = torch.randn(B, Cin, Tin)
    dx = ops.conv1d_dgrad(dyc, wt, geom, Tin, res=r2.to(DEV))
    assert (dx.cpu().double() - (gx + r2.double())).abs().max() < 2e-5
    dw = ops.conv1d_wgrad(xc, dyc, geom)
    assert (dw.cpu().double() - gw).abs().max() / gw.abs().max() < 5e-6
    # accumulates onto what the bucket holds; one writer per output in deterministic mode (bit-reproducible)
    prev = ops.set_deterministic(True)
    try:
        base = torch.randn_like(dw)
        d1 = ops.conv1d_wgrad(xc, dyc, geom, dw=base.clone())
        d2 = ops.conv1d_wgrad(xc, dyc, geom, dw=base.clone())
    finally:
        ops.set_deterministic(prev)
    assert torch.equal(d1, d2)
    assert ((d1 - base).cpu().double() - gw).abs().max() / gw.abs().max() < 5e-6
    if ops.use_tc(geom, "wgrad"):
        dwt = ops.tc_conv1d_wgrad(xc, dyc, geom)
        assert (dwt.cpu().double() - gw).abs().max() / gw.abs().max() < 2e-4
    dx2 = ops.conv1d_dgrad_scatter(dyc, scatter_weight(w, groups).to(DEV), geom, Tin)
    assert (dx2.cpu().double() - gx).abs().max() < 2e-5


def test_weight_norm_fwd_bwd():
    from vibravox_b200 import ops
    torch.manual_seed(1)
    for shape, groups in (((64, 32, 4), 1), ((48, 6, 7), 4), ((256, 128, 16), 1), ((1, 768, 3), 1)):
        v = torch.randn(shape)
        g = torch.rand(shape[0], 1, 1) + 0.5
        v64, g64 = v.double().requires_grad_(True), g.double().requires_grad_(True)
        w64 = torch._weight_norm(v64, g64, 0)
        dw = torch.randn(shape)
        dv64, dg64 = torch.autograd.grad(w64, (v64, g64), dw.double())
        gc, vc, dwc = cuda(g, v, dw)
        w, wt, inv = ops.weight_norm_fwd(gc, vc, groups)
        assert (w.cpu().double() - w64).abs().max() < 1e-6
        assert torch.equal(wt.cpu().view(-1), transpose_weight(w.cpu(), groups).view(-1))
        dg, dv = ops.weight_norm_bwd(gc, vc, inv, dwc)
        assert (dg.cpu().double() - dg64).abs().max() < 2e-5 * max(1.0, float(dg64.abs().max()))
        assert (dv.cpu().double() - dv64).abs().max() < 1e-5
        dg2, dv2 = ops.weight_norm_bwd(gc, vc, inv, dwc, dg=dg.clone(), dv=dv.clone(), beta=1.0)
        assert torch.allclose(dv2, 2 * dv, atol=1e-6) and torch.allclose(dg2, 2 * dg, atol=1e-5)


def test_pqmf_elementwise_and_reconstruction():
    """north_star: 'PQMF reconstruction ... checked element-wise'."""
    from oracle import eben_oracle as O
    from vibravox_b200 import ops
    wa, ws, _ = O.pqmf_design(4, 32)
    wac, wsc = cuda(wa, ws)
    torch.manual_seed(2)
    for B, L in ((1, 15840), (3, 47840), (2, 4000)):
        x = torch.randn(B, 1, L)
        for bands in (2, 4):
            want = O.pqmf_analysis(x, wa, bands)
            got = ops.pqmf_analysis(x.to(DEV), wac, bands)
            assert got.shape == want.shape and (got.cpu() - want).abs().max() < 2e-6
        dec = O.pqmf_analysis(x, wa)
        want_full = O.pqmf_synthesis(dec, ws)
        got_full = ops.pqmf_synthesis(dec.to(DEV), wsc, sum_bands=False)
        assert got_full.shape == want_full.shape and (got_full.cpu() - want_full).abs().max() < 5e-6
        got_sum = ops.pqmf_synthesis(dec.to(DEV), wsc, sum_bands=True)
        assert (got_sum.cpu() - want_full.sum(1, keepdim=True)).abs().max() < 1e-5
        rec = ops.pqmf_synthesis(ops.pqmf_analysis(x.to(DEV), wac, 4), wsc, sum_bands=True).cpu()
        ref = want_full.sum(1, keepdim=True)
        snr_ref = 10 * torch.log10((ref ** 2).mean() / ((x - ref) ** 2).mean())
        snr = 10 * torch.log10((rec ** 2).mean() / ((x - rec) ** 2).mean())
        assert abs(float(snr - snr_ref)) < 0.01 and snr > 45
    # other filter banks
    for m, n in ((8, 64), (2, 16)):
        wa2, ws2, _ = O.pqmf_design(m, n)
        x = torch.randn(2, 1, 1000 * m - n)
        want = O.pqmf_analysis(x, wa2)
        got = ops.pqmf_analysis(x.to(DEV), wa2.to(DEV), m)
        assert (got.cpu() - want).abs().max() < 5e-6
        got_s = ops.pqmf_synthesis(want.to(DEV), ws2.to(DEV), sum_bands=False)
        assert (got_s.cpu() - O.pqmf_synthesis(want, ws2)).abs().max() < 1e-5


def test_pqmf_autograd():
    from oracle import eben_oracle as O
    from vibravox_b200.functional import PQMFAnalysisFn, PQMFSynthesisFn
    wa, ws, _ = O.pqmf_design(4, 32)
    x = torch.randn(2, 1, 4064, dtype=torch.float64, requires_grad=True)
    ya = O.pqmf_analysis(x, wa.double(), 2)
    ga = torch.randn_like(ya)
    (gx,) = torch.autograd.grad(ya, x, ga)
    xc = x.detach().float().to(DEV).requires_grad_(True)
    yc = PQMFAnalysisFn.apply(xc, wa.to(DEV), 2)
    (gxc,) = torch.autograd.grad(yc, xc, ga.float().to(DEV))
    assert (gxc.cpu().double() - gx).abs().max() < 1e-5
    b = torch.randn(2, 4, 1024, dtype=torch.float64, requires_grad=True)
    for summed in (True, False):
        ys = O.pqmf_synthesis(b, ws.double())
        if summed:
            ys = ys.sum(1, keepdim=True)
        gs = torch.randn_like(ys)
        (gb,) = torch.autograd.grad(ys, b, gs)
        bc = b.detach().float().to(DEV).requires_grad_(True)
        yc = PQMFSynthesisFn.apply(bc, ws.to(DEV), summed)
        (gbc,) = torch.autograd.grad(yc, bc, gs.float().to(DEV))
        assert (gbc.cpu().double() - gb).abs().max() < 2e-5


def test_elementwise_and_losses():
    from oracle import eben_oracle as O
    from vibravox_b200 import ops
    from vibravox_b200.functional import FeatureMatchingFn, HingeFn, LeakyReluFn, TanhRecomposeFn, WeightedSumFn
    torch.manual_seed(3)
    x = torch.randn(3, 5, 1001, dtype=torch.float64, requires_grad=True)
    xc = x.detach().float().to(DEV).requires_grad_(True)
    g = torch.randn(3, 5, 1001)
    y, yc = F.leaky_relu(x, 0.01), LeakyReluFn.apply(xc, 0.01)
    assert (yc.cpu().double() - y).abs().max() < 1e-6
    (gx,), (gxc,) = torch.autograd.grad(y, x, g.double()), torch.autograd.grad(yc, xc, g.to(DEV))
    assert (gxc.cpu().double() - gx).abs().max() < 1e-6
    first = torch.randn(3, 2, 1001)
    y = torch.tanh(x + torch.cat((first.double(), torch.zeros(3, 3, 1001, dtype=torch.float64)), 1))
    yc = TanhRecomposeFn.apply(xc, first.to(DEV), 2)
    assert (yc.cpu().double() - y).abs().max() < 1e-6
    (gx,), (gxc,) = torch.autograd.grad(y, x, g.double()), torch.autograd.grad(yc, xc, g.to(DEV))
    assert (gxc.cpu().double() - gx).abs().max() < 2e-6
    # bias-gradient reduction
    dy, ref = torch.randn(4, 7, 333), torch.randn(4, 7, 333)
    db = torch.zeros(7, device=DEV)
    dxc = ops.leaky_relu_bwd(dy.to(DEV), ref.to(DEV), 0.2, dbias=db)
    want = dy.double() * torch.where(ref > 0, 1.0, 0.2).double()
    assert (dxc.cpu().double() - want).abs().max() < 1e-6
    assert (db.cpu().double() - want.sum((0, 2))).abs().max() < 1e-4
    # vector (16-byte) and scalar paths, sign from a byte mask or from `ref`, several 1024-position chunks
    for (Bq, Cq, Tq) in ((3, 6, 2500), (2, 5, 2048), (2, 3, 1027), (1, 4, 8)):
        dy, ref = torch.randn(Bq, Cq, Tq), torch.randn(Bq, Cq, Tq)
        mask = (ref > 0).to(torch.uint8)
        want = dy.double() * torch.where(ref > 0, 1.0, 0.2).double()
        for kw in (dict(ref=ref.to(DEV)), dict(ref=None, mask=mask.to(DEV))):
            db = torch.zeros(Cq, device=DEV)
            dxb = ops.leaky_relu_bwd(dy.to(DEV), kw.get("ref"), 0.2, mask=kw.get("mask"), dbias=db)
            dxp = ops.leaky_relu_bwd(dy.to(DEV), kw.get("ref"), 0.2, mask=kw.get("mask"))
            assert (dxb.cpu().double() - want).abs().max() < 1e-6 and (dxp.cpu().double() - want).abs().max() < 1e-6
            assert (db.cpu().double() - want.sum((0, 2))).abs().max() < 2e-4
            db2 = torch.zeros(Cq, device=DEV)
            assert ops.leaky_relu_bwd(dy.to(DEV), kw.get("ref"), 0.2, mask=kw.get("mask"), dbias=db2, want_dx=False) is None
            assert (db2 - db).abs().max() < 1e-6 * float(db.abs().max()) + 1e-5
    # feature matching + hinge against the oracle (fp64)
    shapes = [[(2, 1, 500), (2, 24, 502), (2, 48, 251), (2, 1, 251)], [(2, 1, 2000), (2, 16, 2000), (2, 1, 500)]]
    a = [[torch.randn(s, dtype=torch.float64, requires_grad=True) for s in sc] for sc in shapes]
    b = [[torch.randn(s, dtype=torch.float64) for s in sc] for sc in shapes]
    fm = O.feature_matching_loss(a, b)
    ac = [[t.detach().float().to(DEV).requires_grad_(True) for t in sc] for sc in a]
    bc = [[t.float().to(DEV) for t in sc] for sc in b]
    from vibravox_b200.torch_modules.losses.feature_loss import FeatureLossForDiscriminatorMelganMultiScales
    from vibravox_b200.torch_modules.losses.hinge_loss import HingeLossForDiscriminatorMelganMultiScales
    fmc = FeatureLossForDiscriminatorMelganMultiScales()(ac, bc)
    assert float(fmc) == pytest.approx(float(fm), rel=1e-5)
    inner = [a[0][1], a[0][2], a[1][1]]
    innerc = [ac[0][1], ac[0][2], ac[1][1]]
    for gw, gc in zip(torch.autograd.grad(fm, inner), torch.autograd.grad(fmc, innerc)):
        assert (gc.cpu().double() - gw).abs().max() < 1e-6 * max(1.0, float(gw.abs().max()) * 1e3)
    for target in (1.0, -1.0):
        h = O.hinge_loss(a, target)
        hc = HingeLossForDiscriminatorMelganMultiScales()(embeddings=ac, target=target)
        assert float(hc) == pytest.approx(float(h), rel=1e-5)
        gw = torch.autograd.grad(h, [a[0][-1], a[1][-1]])
        gc = torch.autograd.grad(hc, [ac[0][-1], ac[1][-1]])
        for u, v in zip(gw, gc):
            assert (v.cpu().double() - u).abs().max() < 1e-8 + 1e-5 * float(u.abs().max())
    # weighted sum
    l1 = torch.tensor(2.0, device=DEV, requires_grad=True)
    l2 = torch.tensor(3.0, device=DEV, requires_grad=True)
    lam = torch.tensor([0.5, 4.0], device=DEV)
    tot = WeightedSumFn.apply(lam, l1, l2)
    assert float(tot) == 13.0
    g1, g2 = torch.autograd.grad(tot, (l1, l2))
    assert float(g1) == 0.5 and float(g2) == 4.0


@pytest.mark.parametrize("via_frames", [True, False])
def test_mrstft_loss_and_per_bin_magnitudes(via_frames):
    """north_star: 'per-bin STFT checked element-wise'."""
    from oracle import eben_oracle as O
    from vibravox_b200 import ops
    from vibravox_b200.torch_modules.losses.mrstft_loss import MultiResolutionSTFTLoss
    ops.STFT_VIA_FRAMES = via_frames
    torch.manual_seed(4)
    B, L = 2, 15840
    x = (0.3 * torch.randn(B, 1, L)).double().requires_grad_(True)      # fp32-representable values
    y = (0.3 * torch.randn(B, 1, L)).double()
    taps = O.a_weighting_fir()
    want = O.mrstft_loss(x, y, taps.double())
    (gx,) = torch.autograd.grad(want, x)
    x32 = x.detach().float().requires_grad_(True)
    (g32,) = torch.autograd.grad(O.mrstft_loss(x32, y.float(), taps), x32)
    noise = (g32.double() - gx).norm() / gx.norm()      # the reference arithmetic's own fp32 noise (log of tiny bins)
    mod = MultiResolutionSTFTLoss(fft_sizes=(512, 1024, 2048), hop_sizes=(50, 120, 240),
                                  win_lengths=(240, 600, 1200), sample_rate=16000, perceptual_weighting=True).to(DEV)
    assert torch.allclose(mod.fir_taps.cpu().view(-1), taps, atol=0, rtol=0)
    xc = x.detach().float().to(DEV).requires_grad_(True)
    got = mod(xc, y.float().to(DEV))
    assert float(got) == pytest.approx(float(want), rel=2e-5)
    (gxc,) = torch.autograd.grad(got, xc)
    ops.STFT_VIA_FRAMES = ops.TC_ENABLED
    err = (gxc.cpu().double() - gx).norm() / gx.norm()
    print("mrstft grad rel-L2 vs fp64:", float(err), "fp32 oracle:", float(noise), "via_frames", via_frames)
    # the framed variant runs the DFT on the tensor cores with the 3-way (24-bit) operand split: bf16x3 would
    # give 1.8e-2 here because the log-magnitude term divides by bins far below the frame energy
    assert err < 2 * noise + 1e-4, (float(err), float(noise))
    gn, gn64 = float(gxc.norm()), float(gx.norm())
    assert gn == pytest.approx(gn64, rel=2e-4)              # the norm that drives the loss balancing
    # per-bin magnitudes
    spec = mod._get_spec(torch.device(DEV, torch.cuda.current_device()))
    sig = x.detach().float().view(B, 1, L).to(DEV)
    for (geom, basis, _), (n_fft, hop, win) in zip(spec.res, O.STFT_RESOLUTIONS):
        X = ops.conv1d_fwd(sig, basis, geom).cpu()
        bins = n_fft // 2 + 1
        mag = torch.sqrt(torch.clamp(X[:, :bins] ** 2 + X[:, bins:] ** 2, min=1e-8))
        ref = O.stft_magnitude(x.detach().view(B, L), n_fft, hop, win)
        assert mag.shape == ref.shape
        assert (mag.double() - ref).abs().max() < 1e-4 * float(ref.max())


def test_adam_matches_torch():
    from vibravox_b200.optim import FlatAdam
    torch.manual_seed(5)
    ps = [torch.randn(33, 7), torch.randn(5), torch.randn(4, 3, 2)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    mine = [torch.nn.Parameter(p.clone().to(DEV)) for p in ps]
    o_ref = torch.optim.Adam(ref, lr=3e-4, betas=(0.5, 0.9))
    o_mine = FlatAdam(mine, lr=3e-4, betas=(0.5, 0.9))
    o_mine.keep_autograd_grad([mine[1]])
    o_mine.materialize()
    for it in range(3):
        grads = [torch.randn_like(p) for p in ps]
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        for p, g in zip(mine, grads):
            if hasattr(p, "_vbx_grad"):
                p._vbx_grad.copy_(g.to(DEV))
            else:
                p.grad = g.to(DEV)
        o_ref.step(); o_mine.step(); o_ref.zero_grad(); o_mine.zero_grad()
        assert float(o_mine.grad.abs().sum()) == 0.0
    for p, q in zip(ref, mine):
        assert (p.detach() - q.detach().cpu()).abs().max() < 1e-6


def test_noise_mix_crop():
    from vibravox_b200 import ops
    torch.manual_seed(6)
    B, Ls, Ln, length = 3, 5000, 20000, 4000
    body, air, noise = torch.randn(B, 1, Ls), torch.randn(B, 1, Ls), torch.randn(B, 1, Ln)
    start = torch.tensor([0, 15000, 777], dtype=torch.int32)
    off = torch.tensor([0, 1000, 500], dtype=torch.int32)
    ob, oa = ops.noise_mix_crop(*cuda(body, air, noise, start, off), length)
    for b in range(B):
        s, o = int(start[b]), int(off[b])
        mixed = body[b, 0] + noise[b, 0, s:s + Ls]
        assert torch.allclose(ob[b, 0].cpu(), mixed[o:o + length], atol=1e-6)
        assert torch.equal(oa[b, 0].cpu(), air[b, 0, o:o + length])


TC_CASES = [
    (1, 32, 32, 256, 1, 1, 1, 0, 0, 1), (2, 32, 32, 300, 3, 1, 9, 9, 9, 1), (2, 64, 256, 403, 41, 4, 1, 20, 0, 4),
    (2, 1024, 1024, 60, 5, 1, 1, 2, 0, 1), (3, 768, 768, 50, 5, 1, 2, 2, 0, 4), (2, 32, 64, 301, 4, 2, 1, 1, 1, 1),
    (2, 24, 48, 131, 7, 2, 2, 3, 0, 4), (2, 24, 48, 131, 7, 2, 3, 3, 0, 4), (2, 16, 32, 257, 16, 8, 1, 7, 7, 1),
    (2, 128, 256, 64, 16, 8, 1, 4, 0, 1), (2, 2, 32, 100, 3, 1, 1, 1, 1, 1), (2, 256, 64, 62, 7, 1, 1, 3, 3, 1),
    (5, 40, 24, 97, 5, 3, 2, 4, 2, 2),
    # slab-form kernel: several row tiles straddling batch items, partial 16-channel groups, strided + dilated
    # zero-halo input gradients (merged phases on a unit or dilated tap lattice), stride-1 flipped-tap gradients
    (3, 48, 96, 700, 7, 2, 3, 3, 0, 4), (3, 48, 96, 701, 7, 2, 2, 3, 0, 4), (2, 64, 64, 333, 5, 1, 3, 6, 0, 2),
    (2, 40, 80, 500, 5, 3, 2, 4, 0, 2), (3, 512, 512, 187, 41, 4, 1, 20, 0, 2), (2, 36, 20, 450, 3, 1, 9, 9, 9, 1),
    # narrow groups densified into one block-diagonal conv (MelGAN stage 1, PQMF-discriminator stages 1-2, odd group counts)
    (2, 16, 64, 1203, 41, 4, 1, 20, 0, 4), (2, 24, 48, 533, 7, 2, 1, 3, 0, 4), (3, 12, 36, 222, 5, 1, 1, 2, 0, 3),
    (2, 48, 96, 300, 7, 2, 1, 3, 0, 4),
    # two row tiles per CTA (wide N, long reduction): odd tile count, last pair half empty
    (3, 256, 256, 700, 41, 4, 1, 20, 0, 1), (2, 512, 256, 300, 5, 1, 1, 2, 0, 1),
]


GATE_CASES = CASES[:6] + DIRECT_CASES + TC_CASES


@pytest.mark.parametrize("case", GATE_CASES, ids=[str(c) for c in GATE_CASES])
def test_input_gradient_gate_stage_equals_the_separate_passes(case):
    """vbx_epilogue gate stage on every input-gradient kernel form (fp32 FMA, direct, tensor-core gather / slab / persistent
    / merged phases): dgrad with gate == LeakyReLU'(y) * (dgrad + L1-pair backward of y), BIT FOR BIT the result of the
    passes it replaces (vbx_l1_pair_bwd, aten::add, vbx_leaky_relu_bwd), with and without the feature-matching term."""
    from vibravox_b200 import ops
    B, Cin, Cout, Tin, K, s, d, pad, refl, groups = case
    geom = ops.ConvGeom(Cin, Cout, K, s, d, pad, refl, groups)
    torch.manual_seed(sum(case) + 1)
    To = tout(Tin, K, s, d, pad)
    w = (torch.randn(Cout, Cin // groups, K) / (Cin // groups * K) ** 0.5).to(DEV)
    dy = torch.randn(B, Cout, To, device=DEV)
    y, other = torch.randn(B, Cin, Tin, device=DEV), torch.randn(B, Cin, Tin, device=DEV)
    y[0, 0, :3] = other[0, 0, :3]
    y[0, 0, 3:6] = 0.0
    sums = torch.tensor([float((y - other).abs().sum()), float(y.abs().sum())], device=DEV, dtype=torch.float64)
    go = torch.tensor([0.7], device=DEV)
    coef = ops.fm_coef(sums, 1, go, 0.125)
    da, _ = ops.l1_pair_bwd(y, other, sums, go, 0.125, True, False)
    wt = ops.transpose_weight(w, groups)
    for tc in ([False, True] if ops.use_tc(geom, "dgrad") else [False]):
        def dgrad(gate=None):
            if tc:
                return ops.tc_conv1d_dgrad(dy, ops.tc_pack(w, geom, 1), geom, Tin, gate=gate)
            return ops.conv1d_dgrad(dy, wt, geom, Tin, gate=gate)
        plain = dgrad()
        want_fm = ops.leaky_relu_bwd(plain + da, y, 0.2)
        want = ops.leaky_relu_bwd(plain, y, 0.2)
        assert torch.equal(dgrad((y, 0.2, other, coef)), want_fm), tc
        assert torch.equal(dgrad((y, 0.2, None, None)), want), tc
        # + the bias gradient of the stage that produced y, reduced with the gate (accumulates onto the slot)
        for gate_args, ref in (((y, 0.2, None, None), want), ((y, 0.2, other, coef), want_fm)):
            db = torch.full((Cin,), 0.5, device=DEV)
            got = dgrad(gate_args + (db,))
            assert torch.equal(got, ref), tc
            wdb = ref.double().sum((0, 2)) + 0.5
            assert (db.double() - wdb).abs().max() <= 1e-5 * float(ref.double().abs().sum((0, 2)).max() + 1), tc
    assert torch.equal(ops.fm_gate_bwd(y, other, coef, 0.2, plain), want_fm)
    assert torch.equal(ops.fm_gate_bwd(y, other, coef, 0.2, None), ops.leaky_relu_bwd(da, y, 0.2))
    assert torch.equal(ops.fm_gate_bwd(y, None, None, 0.2, plain), want)


@pytest.mark.parametrize("case", TC_CASES, ids=[str(c) for c in TC_CASES])
def test_tensor_core_conv_family_matches_fp64(case):
    """tcgen05 kernels (bf16x3 split operands, fp32 TMEM accumulate) vs torch fp64: forward with the fused
    epilogue, dgrad (phases + reflect images), wgrad (split reduction).  Tolerance 1e-4 relative (contract);
    measured ~3e-6 + 2.4e-9 per reduction element (accumulator rounding)."""
    from vibravox_b200 import ops
    B, Cin, Cout, Tin, K, s, d, pad, refl, groups = case
    geom = ops.ConvGeom(Cin, Cout, K, s, d, pad, refl, groups)
    torch.manual_seed(sum(case))
    To = tout(Tin, K, s, d, pad)
    x = torch.randn(B, Cin, Tin)
    w = torch.randn(Cout, Cin // groups, K) / (Cin // groups * K) ** 0.5
    bias, res = torch.randn(Cout), torch.randn(B, Cout, To)
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    pre = F.conv1d(ref_padded(x64, pad, refl), w64, bias.double(), s, 0, d, groups)
    want = F.leaky_relu(pre, 0.2) + res.double()
    xc, wc, bc, rc = cuda(x, w, bias, res)
    y, mask = ops.tc_conv1d_fwd(xc, ops.tc_pack(wc, geom, 0), geom, bias=bc, res=rc, slope=0.2, want_mask=True)
    assert (y.cpu().double() - want).norm() / want.norm() < 1e-4
    assert (y.cpu().double() - want).abs().max() < 2e-4 * float(want.abs().max())
    assert (mask.cpu().bool() != (pre > 0)).float().mean() < 1e-3
    dy = torch.randn(B, Cout, To)
    plain = F.conv1d(ref_padded(x64, pad, refl), w64, None, s, 0, d, groups)
    gx, gw = torch.autograd.grad(plain, (x64, w64), dy.double())
    dyc = dy.to(DEV)
    r2 = torch.randn(B, Cin, Tin)
    dx = ops.tc_conv1d_dgrad(dyc, ops.tc_pack(wc, geom, 1), geom, Tin, res=r2.to(DEV))
    assert (dx.cpu().double() - (gx + r2.double())).abs().max() < 2e-4 * float(gx.abs().max())
    dw = ops.tc_conv1d_wgrad(xc, dyc, geom)
    assert (dw.cpu().double() - gw).abs().max() < 2e-4 * float(gw.abs().max())
    assert (dw.cpu().double() - gw).norm() / gw.norm() < 1e-4


def test_stft_framing_unfold_fold():
    from vibravox_b200 import ops
    import cpu_shim
    torch.manual_seed(9)
    for B, L, K, hop in ((2, 4000, 240, 50), (3, 15840, 1200, 240), (1, 700, 600, 120)):
        pad = K // 2
        x = torch.randn(B, 1, L)
        want = cpu_shim.unfold_frames(x, K, hop, pad)
        got = ops.unfold_frames(x.to(DEV), K, hop, pad)
        assert got.shape == want.shape and torch.equal(got.cpu(), want)
        dU = torch.randn_like(want)
        gw = cpu_shim.fold_frames(dU.double(), L, hop, pad)
        gg = ops.fold_frames(dU.to(DEV), L, hop, pad)
        assert (gg.cpu().double() - gw).abs().max() < 1e-5
        acc = torch.ones(B, 1, L, device=DEV)
        ops.fold_frames(dU.to(DEV), L, hop, pad, dx=acc)
        assert (acc.cpu().double() - (gw + 1)).abs().max() < 1e-5


@pytest.mark.parametrize("case", [TC_CASES[2], TC_CASES[3], TC_CASES[5], TC_CASES[7]], ids=str)
def test_tensor_core_three_way_split_is_fp32_grade(case):
    """nsplit=3 ("bf16x6": hi+mid+lo, 6 MMAs): 24 mantissa bits per operand, i.e. no operand rounding left.
    What remains is the TMEM accumulator, which rounds toward zero after every MMA: measured 1.9e-8 relative per
    accumulated MMA, linear in their number (the same term dominates bf16x3 at long reductions)."""
    from vibravox_b200 import ops
    B, Cin, Cout, Tin, K, s, d, pad, refl, groups = case
    geom = ops.ConvGeom(Cin, Cout, K, s, d, pad, refl, groups)
    torch.manual_seed(sum(case) + 1)
    x = torch.randn(B, Cin, Tin)
    w = torch.randn(Cout, Cin // groups, K) / (Cin // groups * K) ** 0.5
    x64, w64 = x.double().requires_grad_(True), w.double()
    want = F.conv1d(ref_padded(x64, pad, refl), w64, None, s, 0, d, groups)
    xc, wc = cuda(x, w)
    y = ops.tc_conv1d_fwd(xc, ops.tc_pack(wc, geom, 0, 3), geom, nsplit=3)
    y2 = ops.tc_conv1d_fwd(xc, ops.tc_pack(wc, geom, 0, 2), geom, nsplit=2)
    e3 = float((y.cpu().double() - want).norm() / want.norm())
    e2 = float((y2.cpu().double() - want).norm() / want.norm())
    print("rel-L2 vs fp64: bf16x6", e3, "bf16x3", e2)
    def n_mma(cred, taps):
        cpad = (cred + 7) // 8 * 8 if cred >= 5 else (4 if cred >= 3 else cred)
        return (cpad * taps + 31) // 32 * 2 * 6
    assert e3 < 3e-7 + 4e-8 * n_mma(Cin // groups, K)
    dy = torch.randn_like(want)
    (gx,) = torch.autograd.grad(want, x64, dy)
    dx = ops.tc_conv1d_dgrad(dy.float().to(DEV), ops.tc_pack(wc, geom, 1, 3), geom, Tin, nsplit=3)
    assert float((dx.cpu().double() - gx).norm() / gx.norm()) < 3e-7 + 4e-8 * n_mma(Cout // groups, K)


def test_flat_adam_update_invalidates_packed_weight_tiles():
    """A conv that uses its parameter directly caches bf16 tiles on the tensor; FlatAdam rewrites the parameter
    through a raw pointer, so the cache has to be dropped by the optimizer (regression: stale tiles after step 1)."""
    from vibravox_b200 import ops
    from vibravox_b200.optim import FlatAdam
    torch.manual_seed(3)
    conv = torch.nn.Conv1d(16, 32, 3).to(DEV)
    opt = FlatAdam(conv.parameters(), lr=0.05, betas=(0.5, 0.9))
    opt.materialize()
    geom = ops.ConvGeom(16, 32, 3, 1, 1, 1, 0, 1)
    x = torch.randn(2, 16, 400, device=DEV)
    assert ops.use_tc(geom, "fwd")
    y0 = ops.conv_fwd(x, conv.weight, geom)
    assert "_vbx_packs" in conv.weight.__dict__
    for p in conv.parameters():
        p._vbx_grad.fill_(1.0)
    opt.step()
    y1 = ops.conv_fwd(x, conv.weight, geom)
    want = F.conv1d(x.double(), conv.weight.detach().double(), None, 1, 1)
    assert (y1.double() - want).abs().max() < 1e-4 * float(want.abs().max())
    assert (y1 - y0).abs().max() > 1e-2


PERSISTENT_CASES = [
    (12, 32, 32, 11968, 3, 1, 3, 3, 3, 1),     # generator residual conv: 1123 row tiles on 444 persistent CTAs
    (20, 24, 48, 11970, 7, 2, 1, 3, 0, 4),     # PQMF-discriminator stage 1, groups densified: 936 tiles on 296 CTAs
]


@pytest.mark.parametrize("case", PERSISTENT_CASES, ids=[str(c) for c in PERSISTENT_CASES])
def test_persistent_slab_many_tiles_per_cta_matches_fma(case):
    """tc_pslab_kernel at sizes where every CTA walks two to three row tiles (both TMEM accumulator buffers reused,
    the slab ring wrapped): forward with bias + LeakyReLU and the input gradient against the fp32 FMA kernels of
    the same library on the same device tensors (the small TC_CASES give each persistent CTA a single tile)."""
    from vibravox_b200 import ops
    B, Cin, Cout, Tin, K, s, d, pad, refl, groups = case
    g = ops.ConvGeom(Cin, Cout, K, s, d, pad, refl, groups)
    To = g.tout(Tin)
    torch.manual_seed(0)
    x = torch.randn(B, Cin, Tin, device=DEV)
    dy = torch.randn(B, Cout, To, device=DEV)
    w = torch.randn(Cout, Cin // groups, K, device=DEV) / (Cin // groups * K) ** 0.5
    bias = torch.randn(Cout, device=DEV)
    y = ops.tc_conv1d_fwd(x, ops.tc_pack(w, g, ops.TC_FWD), g, bias=bias, slope=0.2)
    y0 = ops.conv1d_fwd(x, w, g, bias=bias, slope=0.2)
    assert float((y - y0).norm() / y0.norm()) < 2e-5
    dx = ops.tc_conv1d_dgrad(dy, ops.tc_pack(w, g, ops.TC_DGRAD), g, Tin)
    dx0 = ops.conv1d_dgrad(dy, ops.transpose_weight(w, groups), g, Tin)
    assert float((dx - dx0).norm() / dx0.norm()) < 2e-5


RU_CASES = [  # B, C, T, dilation
    (3, 32, 1000, 1), (3, 32, 1000, 3), (3, 32, 1000, 9), (2, 64, 516, 9), (2, 64, 1280, 3), (5, 16, 132, 1),
    (2, 48, 300, 9), (1, 32, 40, 9), (4, 64, 128, 1), (7, 32, 388, 3), (32, 32, 11968, 3), (32, 64, 5984, 9),
]


@pytest.mark.parametrize("case", RU_CASES, ids=str)
def test_fused_residual_unit_matches_fp64_and_the_two_kernel_form(case):
    """vbx_ru_fwd (one kernel per ResidualUnit: TMA tile load, dilated conv -> TMEM -> re-split -> pointwise conv ->
    LeakyReLU + residual) vs torch fp64 (eben_generator.py:314-316) and vs the two-launch tensor-core form: out, the
    saved h = dilated(x) and the activation mask.  Covers edge tiles (reflect halo on both sides, T < 128, T not a
    multiple of 128), every ring wrap (many tiles per CTA at the full BASELINE sizes) and C = 16..64."""
    from vibravox_b200 import ops
    B, C, T, d = case
    assert ops.use_fused_unit(C, T, d, B)
    torch.manual_seed(sum(case))
    x = torch.randn(B, C, T)
    w1 = torch.randn(C, C, 3) / (3 * C) ** 0.5
    w2 = torch.randn(C, C, 1) / C ** 0.5
    if B * C * T <= 4e6:
        x64 = x.double()
        h64 = F.conv1d(F.pad(x64, (d, d), mode="reflect"), w1.double(), None, 1, 0, d)
        z64 = F.conv1d(h64, w2.double())
        want = x64 + F.leaky_relu(z64, 0.01)
    else:
        x64 = want = None
    xc, w1c, w2c = cuda(x, w1, w2)
    pk = ops.residual_unit_pack(w1c, w2c)
    out, h, mask = ops.residual_unit_fwd(xc, pk, d, 0.01, want_h=True, want_mask=True)
    out_inf, h_none, mask_none = ops.residual_unit_fwd(xc, pk, d, 0.01)
    assert h_none is None and mask_none is None and torch.equal(out, out_inf)
    g1, g2 = ops.ConvGeom(C, C, 3, 1, d, d, d, 1), ops.ConvGeom(C, C, 1, 1, 1, 0, 0, 1)
    h2 = ops.conv_fwd(xc, w1c, g1)
    out2, mask2 = ops.conv_fwd(h2, w2c, g2, res=xc, slope=0.01, want_mask=True)
    # same operand split, same MMA order: the two forms agree to fp32 accumulation-order noise
    assert (h - h2).abs().max() <= 2e-5 * float(h2.abs().max())
    assert (out - out2).abs().max() <= 2e-5 * float(out2.abs().max())
    assert (mask != mask2).float().mean() < 1e-4
    if want is not None:
        assert (h.cpu().double() - h64).norm() / h64.norm() < 1e-4
        assert (out.cpu().double() - want).norm() / want.norm() < 2e-5
        assert (out.cpu().double() - want).abs().max() < 1e-4 * float(want.abs().max())
        assert (mask.cpu().bool() != (z64 > 0)).float().mean() < 1e-3
    # repeated launches on the same buffers are bit-identical (no read of a stale ring slot, no race)
    out3, _, _ = ops.residual_unit_fwd(xc, pk, d, 0.01)
    assert torch.equal(out, out3)


def test_fused_residual_unit_through_autograd_matches_the_unfused_module():
    """ResidualUnitFn with the fused forward (+ saved h / mask) gives the gradients of the two-launch form."""
    from vibravox_b200 import ops
    from vibravox_b200.functional import ResidualUnitFn
    torch.manual_seed(3)
    B, C, T, d = 2, 32, 700, 3
    x = torch.randn(B, C, T, device=DEV, requires_grad=True)
    w1 = (torch.randn(C, C, 3, device=DEV) / (3 * C) ** 0.5).requires_grad_(True)
    w2 = (torch.randn(C, C, 1, device=DEV) / C ** 0.5).requires_grad_(True)
    g1, g2 = ops.ConvGeom(C, C, 3, 1, d, d, d, 1), ops.ConvGeom(C, C, 1, 1, 1, 0, 0, 1)
    go = torch.randn(B, C, T, device=DEV)
    res = {}
    for fused in (True, False):
        old, ops.FUSED_UNIT = ops.FUSED_UNIT, fused
        try:
            y = ResidualUnitFn.apply(x, w1, ops.transpose_weight(w1.detach(), 1), w2, ops.transpose_weight(w2.detach(), 1),
                                     g1, g2, 0.01)
            res[fused] = (y.detach(),) + torch.autograd.grad(y, (x, w1, w2), go)
        finally:
            ops.FUSED_UNIT = old
    for a, b in zip(res[True], res[False]):
        assert (a - b).abs().max() <= 5e-5 * float(b.abs().max())
    with torch.no_grad():                       # inference: no h / mask are produced at all
        y = ResidualUnitFn.apply(x, w1, None, w2, None, g1, g2, 0.01)
        assert (y - res[True][0]).abs().max() == 0


@pytest.mark.parametrize("C,T,d", [(32, 700, 3), (64, 516, 9), (128, 300, 1)])
def test_residual_unit_composed_backward_matches_fp64_and_the_layerwise_backward(C, T, d):
    """Backward of the unit through the composed conv (wf = w2 . w1: one input gradient, one weight gradient, dw1 / dw2 from
    C x C x 3C products) against torch fp64 autograd of the reference formula (eben_generator.py:314-316) and against the
    layer-by-layer backward; the small kernels (vbx_unit_combine / vbx_unit_split_grads) against einsum."""
    from vibravox_b200 import functional, ops
    from vibravox_b200.functional import ResidualUnitFn
    torch.manual_seed(C + d)
    B = 3
    x = torch.randn(B, C, T, device=DEV, requires_grad=True)
    w1 = (torch.randn(C, C, 3, device=DEV) / (3 * C) ** 0.5).requires_grad_(True)
    w2 = (torch.randn(C, C, 1, device=DEV) / C ** 0.5).requires_grad_(True)
    g1, g2 = ops.ConvGeom(C, C, 3, 1, d, d, d, 1), ops.ConvGeom(C, C, 1, 1, 1, 0, 0, 1)
    go = torch.randn(B, C, T, device=DEV)
    wf = ops.unit_combine(w1.detach(), w2.detach())
    want_wf = torch.einsum("om,mik->oik", w2.detach()[:, :, 0].double(), w1.detach().double())
    assert (wf.double() - want_wf).abs().max() < 1e-6
    dwf = torch.randn(C, C, 3, device=DEV)
    dw1, dw2 = ops.unit_split_grads(dwf, w1.detach(), w2.detach(), True, True)
    assert (dw1.double() - torch.einsum("om,oik->mik", w2.detach()[:, :, 0].double(), dwf.double())).abs().max() < 2e-5
    assert (dw2[:, :, 0].double() - torch.einsum("oik,mik->om", dwf.double(), w1.detach().double())).abs().max() < 2e-5
    x64, w164, w264 = (t.detach().double().requires_grad_(True) for t in (x, w1, w2))
    h64 = F.conv1d(F.pad(x64, (d, d), mode="reflect"), w164, None, 1, 0, d)
    y64 = x64 + F.leaky_relu(F.conv1d(h64, w264), 0.01)
    want = torch.autograd.grad(y64, (x64, w164, w264), go.double())
    res = {}
    for mode in ("composed", "layerwise", "zero_halo"):     # zero_halo: composed + slab-form zero-halo dgrad + mirror terms
        old = functional.UNIT_COMPOSED, functional.UNIT_ZERO_HALO_DGRAD
        functional.UNIT_COMPOSED, functional.UNIT_ZERO_HALO_DGRAD = mode != "layerwise", mode == "zero_halo"
        try:
            y = ResidualUnitFn.apply(x, w1, ops.transpose_weight(w1.detach(), 1), w2, ops.transpose_weight(w2.detach(), 1),
                                     g1, g2, 0.01)
            res[mode] = torch.autograd.grad(y, (x, w1, w2), go)
        finally:
            functional.UNIT_COMPOSED, functional.UNIT_ZERO_HALO_DGRAD = old
        assert (y.double() - y64).abs().max() < 2e-4 * float(y64.abs().max())
    for mode, grads in res.items():
        for a, w in zip(grads, want):
            scale = float(w.abs().max())
            assert (a.double() - w).abs().max() < 3e-4 * scale, (mode, C, (a.double() - w).abs().max() / scale)
            assert (a.double() - w).norm() / w.norm() < 1e-4, mode


WG_CASES = [  # B, C, T, dilation, K
    (3, 32, 1000, 1, 3), (3, 32, 1000, 3, 3), (3, 32, 1000, 9, 3), (2, 64, 516, 9, 3), (2, 64, 1280, 3, 3),
    (3, 32, 1000, 1, 1), (2, 64, 516, 1, 1), (1, 32, 40, 9, 3), (7, 32, 388, 3, 3), (32, 32, 11968, 9, 3),
    (32, 64, 5984, 3, 3), (32, 32, 11968, 1, 1), (32, 64, 5984, 1, 1),
]


@pytest.mark.parametrize("case", WG_CASES, ids=str)
def test_unit_weight_gradient_kernel_matches_fp64(case):
    """vbx_ru_wgrad (TMA tile loads, time-reduction MMAs with the accumulators resident in TMEM, fixed-order reduction
    of the per-CTA partials) vs torch fp64 autograd of the reflect-padded conv (eben_generator.py:295-312), incl. edge
    tiles, T < 128, accumulation onto an existing gradient, and bit-identical repeats (no atomics)."""
    from vibravox_b200 import ops
    B, C, T, d, K = case
    assert ops.unit_wgrad_workspace(B, C, T, d, K) > 0
    torch.manual_seed(sum(case))
    x, dy = torch.randn(B, C, T), torch.randn(B, C, T)
    xc, dyc = cuda(x, dy)
    dw = ops.unit_wgrad(xc, dyc, K, d)
    assert torch.equal(dw, ops.unit_wgrad(xc, dyc, K, d))                       # deterministic
    acc = torch.full((C, C, K), 0.5, device=DEV)
    assert (ops.unit_wgrad(xc, dyc, K, d, dw=acc) - (dw + 0.5)).abs().max() <= 1e-5 * float(dw.abs().max())
    pad = d * (K - 1) // 2
    g = ops.ConvGeom(C, C, K, 1, d if K == 3 else 1, pad, pad, 1)
    old = ops.tc_conv1d_wgrad(xc, dyc, g)                                         # the gather-form kernel it replaces
    assert (dw - old).abs().max() <= 1e-4 * float(old.abs().max())
    if B * C * T <= 4e6:
        w64 = torch.zeros(C, C, K, dtype=torch.float64, requires_grad=True)
        xp = F.pad(x.double(), (pad, pad), mode="reflect") if pad else x.double()
        y = F.conv1d(xp, w64, None, 1, 0, d if K == 3 else 1)
        (gw,) = torch.autograd.grad(y, w64, dy.double())
        assert (dw.cpu().double() - gw).abs().max() < 2e-4 * float(gw.abs().max())
        assert (dw.cpu().double() - gw).norm() / gw.norm() < 1e-4
