"""CPU oracle: functional restatement of the EBEN training-step arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain PyTorch on the host, any
float dtype (fp32 = the parity target, fp64 = bracketing).  Everything works on a
flat ``dict[str, Tensor]`` keyed exactly like the reference ``state_dict()``.
Each function cites the reference file:line (relative to /root/reference) whose
arithmetic it follows.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

ENC_STRIDES = (2, 4, 8)          # eben_generator.py:121-127
DEC_STRIDES = (8, 4, 2)          # eben_generator.py:151-157
RES_DILATIONS = (1, 3, 9)        # eben_generator.py:231-233,263-265
G_SLOPE = 0.01                   # eben_generator.py:110
D_SLOPE = 0.2                    # eben_discriminator.py:31,79


# --------------------------------------------------------------------------- PQMF
def pqmf_design(m: int, n: int, beta: float = 9) -> Tuple[Tensor, Tensor, float]:
    """Kaiser/sinc prototype + L-BFGS cut-off search + cosine modulation.

    Follows vibravox/torch_modules/dsp/pqmf.py:66-180 operation by operation so
    that the taps are bit-identical (checked in oracle/make_goldens.py).
    Returns (analysis (m,1,n), synthesis (m,1,n), cutoff).
    """
    assert n % (4 * m) == 0                                   # pqmf.py:42
    grid = torch.arange(n) - (n - 1) / 2                       # fp32
    win64 = torch.kaiser_window(n, periodic=False, beta=beta).to(torch.float64)

    def prototype(cut):                                        # pqmf.py:66-91
        s = cut * torch.special.sinc(cut * grid)               # fp32
        return (s.to(torch.float64) * win64).to(torch.float32).view(1, 1, n)

    def objective(cut):                                        # pqmf.py:103-124
        h = prototype(cut)
        ac = F.conv1d(F.pad(h, (n // 2, n // 2)), h)
        keep = torch.ones(n + 1, dtype=ac.dtype)
        keep[n // 2] = 0
        phi = (ac * keep)[..., :: 2 * m].abs().max()
        off = abs(float(cut.detach()) - 1 / (2 * m)) > 1 / (4 * m)
        return phi + (1 / (4 * m) if off else 0)

    cut = (torch.ones(1) / (2 * m)).requires_grad_(True)       # pqmf.py:126-138
    opt = torch.optim.LBFGS([cut], line_search_fn="strong_wolfe")
    for _ in range(5):
        opt.zero_grad()
        objective(cut).backward()
        opt.step(lambda: objective(cut))
    cutoff = cut.item()

    h = prototype(cutoff).view(1, n)                           # pqmf.py:140-180
    coef = torch.tensor([(2 * k + 1) * math.pi / 2 / m for k in range(m)]).view(m, 1)
    shift = torch.tensor([(-1) ** k * math.pi / 4 for k in range(m)]).view(m, 1)
    arg = coef * grid.view(1, n)
    analysis = 2 * torch.flip(h * torch.cos(arg + shift), [1])
    synthesis = (m * 2) * h * torch.cos(arg - shift)
    return analysis.view(m, 1, n).contiguous(), synthesis.view(m, 1, n).contiguous(), cutoff


def pqmf_analysis(x: Tensor, wa: Tensor, bands: int = -1) -> Tensor:
    """pqmf.py:194-202 : strided FIR bank, zero padding n-1."""
    m, _, n = wa.shape
    w = wa if bands == -1 else wa[:bands]
    return F.conv1d(x, w.to(x.dtype), None, stride=m, padding=n - 1)


def pqmf_synthesis(bands: Tensor, ws: Tensor) -> Tensor:
    """pqmf.py:204-213 : grouped transposed FIR bank (per-band outputs, un-summed)."""
    m, _, n = ws.shape
    return F.conv_transpose1d(bands, ws.to(bands.dtype), None, stride=m,
                              output_padding=m - 2, groups=m, padding=n - 1)


def cut_to_valid_length(x: Tensor, n: int, m: int) -> Tensor:
    """eben_generator.py:215-222 with multiple = 2*4*8*m (:108)."""
    L = x.shape[2]
    return x[:, :, : L - (L + n) % (2 * 4 * 8 * m)]


# ------------------------------------------------------------------ parameter init
def _conv_param(shape, bias: bool):
    """nn.Conv1d / nn.ConvTranspose1d.reset_parameters (kaiming_uniform a=sqrt(5))."""
    w = torch.empty(*shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    b = None
    if bias:
        fan_in = shape[1] * shape[2]
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        b = torch.empty(shape[0]).uniform_(-bound, bound)
    return w, b


def _put_wn(sd: State, prefix: str, shape, bias: bool):
    """weight_norm(dim=0) registration (torch_modules/utils.py:4-9): g=||v||, v=w."""
    w, b = _conv_param(shape, bias)
    if b is not None:
        sd[prefix + ".bias"] = b
    sd[prefix + ".parametrizations.weight.original0"] = torch.norm_except_dim(w, 2, 0)
    sd[prefix + ".parametrizations.weight.original1"] = w


def _put_res(sd: State, prefix: str, c: int):
    _put_wn(sd, prefix + ".dilated_conv", (c, c, 3), False)
    _put_wn(sd, prefix + ".pointwise_conv", (c, c, 1), False)


def init_generator_state(m: int = 4, n: int = 32, p: int = 2) -> State:
    """Parameters in the construction order of EBENGenerator.__init__
    (eben_generator.py:93-166) so a given torch seed yields the same values."""
    sd: State = OrderedDict()
    wa, ws, _ = pqmf_design(m, n)
    sd["pqmf.analysis_weights"], sd["pqmf.synthesis_weights"] = wa, ws
    sd["first_conv.weight"] = _conv_param((32, p, 3), False)[0]
    c = 32
    for i, s in enumerate(ENC_STRIDES):                        # EncBlock :257-284
        for r in range(3):
            _put_res(sd, f"encoder_blocks.{i}.residuals.{r}", c)
        _put_wn(sd, f"encoder_blocks.{i}.conv", (2 * c, c, 2 * s), False)
        c *= 2
    _put_wn(sd, "latent_conv.1", (64, 256, 7), False)          # :129-149
    _put_wn(sd, "latent_conv.3", (256, 64, 7), False)
    for i, s in enumerate(DEC_STRIDES):                        # DecBlock :225-254
        c //= 2
        for r in range(3):
            _put_res(sd, f"decoder_blocks.{i}.residuals.{r}", c)
        _put_wn(sd, f"decoder_blocks.{i}.conv_trans", (2 * c, c, 2 * s), False)
    sd["last_conv.weight"] = _conv_param((4, 32, 3), False)[0]  # :159-166 (out=4 hard-coded)
    return sd


def pqmf_disc_layers(q: int, mc: int, dilation: int):
    """(C_in, C_out, K, stride, pad, dilation, groups) of DiscriminatorEBEN
    (eben_discriminator.py:66-157)."""
    L = [(q, mc, 3, 1, 1, dilation, q)]
    c = mc
    for _ in range(5):
        L.append((c, 2 * c, 7, 2, 3, dilation, q))
        c *= 2
    L.append((c, c, 5, 1, 2, dilation, q))
    L.append((c, 1, 3, 1, 1, 1, 1))
    return L


MELGAN_LAYERS = [                                             # melgan_discriminator.py:89-156
    (1, 16, 15, 1, 0, 1, 1), (16, 64, 41, 4, 20, 1, 4), (64, 256, 41, 4, 20, 1, 4),
    (256, 1024, 41, 4, 20, 1, 4), (1024, 1024, 41, 4, 20, 1, 4),
    (1024, 1024, 5, 1, 2, 1, 1), (1024, 1, 3, 1, 1, 1, 1),
]


def _pqmf_disc_key(k: int, i: int) -> str:
    return f"pqmf_discriminators.{k}.discriminator.{i}" + (".1" if i == 0 else ".0" if i < 7 else "")


def _melgan_key(i: int) -> str:
    return f"melgan_discriminator.discriminator.{i}" + (".1" if i == 0 else ".0" if i < 6 else "")


def init_discriminator_state(q: int = 3, min_channels: int = 24) -> State:
    """Construction order of DiscriminatorEBENMultiScales.__init__ (eben_discriminator.py:18-31)."""
    assert min_channels % q == 0                               # eben_discriminator.py:64
    sd: State = OrderedDict()
    for k, dil in enumerate((1, 2, 3)):
        for i, (ci, co, ks, _, _, _, g) in enumerate(pqmf_disc_layers(q, min_channels, dil)):
            _put_wn(sd, _pqmf_disc_key(k, i), (co, ci // g, ks), True)
    for i, (ci, co, ks, _, _, _, g) in enumerate(MELGAN_LAYERS):
        _put_wn(sd, _melgan_key(i), (co, ci // g, ks), True)
    return sd


# ------------------------------------------------------------------------ forward
def wn_weight(sd: State, prefix: str) -> Tensor:
    """torch._weight_norm(v, g, dim=0): v * (g / ||v||_{dims 1,2})."""
    g = sd[prefix + ".parametrizations.weight.original0"]
    v = sd[prefix + ".parametrizations.weight.original1"]
    return v * (g / torch.norm_except_dim(v, 2, 0))


def _rconv(x: Tensor, w: Tensor, pad: int, stride: int = 1, dilation: int = 1) -> Tensor:
    """Conv1d(padding_mode='reflect'): explicit mirror pad (no edge repeat) + valid conv."""
    if pad:
        x = F.pad(x, (pad, pad), mode="reflect")
    return F.conv1d(x, w, None, stride, 0, dilation)


def _lrelu(x: Tensor, slope: float) -> Tensor:
    return F.leaky_relu(x, slope)


def _res_unit(sd: State, prefix: str, x: Tensor, d: int) -> Tensor:
    """ResidualUnit.forward (eben_generator.py:314-316)."""
    z = _rconv(x, wn_weight(sd, prefix + ".dilated_conv"), d, 1, d)
    z = F.conv1d(z, wn_weight(sd, prefix + ".pointwise_conv"))
    return x + _lrelu(z, G_SLOPE)


def generator_forward(sd: State, cut_audio: Tensor, p: int) -> Tuple[Tensor, Tensor]:
    """EBENGenerator.forward (eben_generator.py:168-213)."""
    wa, ws = sd["pqmf.analysis_weights"], sd["pqmf.synthesis_weights"]
    m = wa.shape[0]
    first = pqmf_analysis(cut_audio, wa, p)
    x = _rconv(first, sd["first_conv.weight"], 1)
    skips = []
    for i, s in enumerate(ENC_STRIDES):
        x = _lrelu(x, G_SLOPE)
        for r, d in enumerate(RES_DILATIONS):
            x = _res_unit(sd, f"encoder_blocks.{i}.residuals.{r}", x, d)
        x = _rconv(x, wn_weight(sd, f"encoder_blocks.{i}.conv"), s - 1, s)
        skips.append(x)
    x = _lrelu(x, G_SLOPE)
    x = _lrelu(_rconv(x, wn_weight(sd, "latent_conv.1"), 3), G_SLOPE)
    x = _lrelu(_rconv(x, wn_weight(sd, "latent_conv.3"), 3), G_SLOPE)
    for i, s in enumerate(DEC_STRIDES):
        x = x + skips[2 - i]
        x = F.conv_transpose1d(x, wn_weight(sd, f"decoder_blocks.{i}.conv_trans"), None,
                               stride=s, padding=s // 2)
        x = _lrelu(x, G_SLOPE)
        for r, d in enumerate(RES_DILATIONS):
            x = _res_unit(sd, f"decoder_blocks.{i}.residuals.{r}", x, d)
    x = _rconv(x, sd["last_conv.weight"], 1)
    b, _, t = first.shape
    filled = torch.cat((first, first.new_zeros(b, m - p, t)), dim=1)
    bands = torch.tanh(x + filled)
    enhanced = pqmf_synthesis(bands, ws).sum(1, keepdim=True)
    return enhanced, bands


def _disc_chain(sd: State, x: Tensor, layers, keyfn, first_reflect: int) -> List[Tensor]:
    out = [x]
    last = len(layers) - 1
    for i, (_, _, _, s, pad, d, g) in enumerate(layers):
        h = out[-1]
        if i == 0:
            h = F.pad(h, (first_reflect, first_reflect), mode="reflect")
        key = keyfn(i)
        h = F.conv1d(h, wn_weight(sd, key), sd[key + ".bias"], s, pad, d, g)
        out.append(h if i == last else _lrelu(h, D_SLOPE))
    return out


def discriminator_forward(sd: State, bands: Tensor, audio: Tensor, q: int,
                          min_channels: int = 24) -> List[List[Tensor]]:
    """DiscriminatorEBENMultiScales.forward (eben_discriminator.py:33-51,159-163;
    melgan_discriminator.py:158-169).  Element 0 of each list is the input."""
    emb = []
    sel = bands[:, -q:, :]
    for k, dil in enumerate((1, 2, 3)):
        emb.append(_disc_chain(sd, sel, pqmf_disc_layers(q, min_channels, dil),
                               lambda i, k=k: _pqmf_disc_key(k, i), 1))
    emb.append(_disc_chain(sd, audio, MELGAN_LAYERS, _melgan_key, 7))
    return emb


# ------------------------------------------------------------------------- losses
def feature_matching_loss(a: List[List[Tensor]], b: List[List[Tensor]]) -> Tensor:
    """feature_loss.py:37-50 (divisor = scales * len(last scale's inner layers))."""
    total = 0.0
    for sa, sb in zip(a, b):
        for la, lb in zip(sa[1:-1], sb[1:-1]):
            total = total + (la - lb).abs().mean() / la.abs().mean()
    return total / (len(a) * len(a[-1][1:-1]))


def hinge_loss(emb: List[List[Tensor]], target: float) -> Tensor:
    """hinge_loss.py:35-43."""
    total = 0.0
    for scale in emb:
        total = total + F.relu(1 - target * scale[-1]).mean()
    return total / len(emb)


def a_weighting_fir(fs: int = 16000, ntaps: int = 101) -> Tensor:
    """auraloss.perceptual.FIRFilter(filter_type='aw') tap design [third-party,
    restated from auraloss 0.4.0; SURVEY App. C].  Returns fp32 (ntaps,)."""
    import numpy as np
    import scipy.signal
    f1, f2, f3, f4, a1000 = 20.598997, 107.65265, 737.86223, 12194.217, 1.9997
    nums = [(2 * np.pi * f4) ** 2 * (10 ** (a1000 / 20)), 0, 0, 0, 0]
    dens = np.polymul([1, 4 * np.pi * f4, (2 * np.pi * f4) ** 2],
                      [1, 4 * np.pi * f1, (2 * np.pi * f1) ** 2])
    dens = np.polymul(np.polymul(dens, [1, 2 * np.pi * f3]), [1, 2 * np.pi * f2])
    b, a = scipy.signal.bilinear(nums, dens, fs=fs)
    w_iir, h_iir = scipy.signal.freqz(b, a, worN=512, fs=fs)
    taps = scipy.signal.firls(ntaps, w_iir, abs(h_iir), fs=fs)
    return torch.tensor(taps.astype("float32"))


STFT_RESOLUTIONS = ((512, 50, 240), (1024, 120, 600), (2048, 240, 1200))   # multi_stft.yaml:3-14


def stft_magnitude(x: Tensor, n_fft: int, hop: int, win: int, eps: float = 1e-8) -> Tensor:
    """auraloss STFTLoss.stft: hann(win) centred in n_fft, center/reflect, onesided;
    sqrt(clamp(re^2+im^2, eps)).  x: (B, L) -> (B, bins, frames)."""
    window = torch.hann_window(win, dtype=x.dtype, device=x.device)
    X = torch.stft(x, n_fft, hop, win, window, return_complex=True)
    return torch.sqrt(torch.clamp(X.real ** 2 + X.imag ** 2, min=eps))


def mrstft_loss(x: Tensor, y: Tensor, taps: Tensor, resolutions=STFT_RESOLUTIONS,
                perceptual: bool = True) -> Tensor:
    """auraloss.freq.MultiResolutionSTFTLoss as configured by
    configs/lightning_module/loss_module/multi_stft.yaml:1-18 (x=input, y=target)."""
    bs, ch, L = x.shape
    if perceptual:
        k = taps.to(x.dtype).view(1, 1, -1)
        pad = k.shape[-1] // 2
        x = F.conv1d(x.reshape(bs * ch, 1, L), k, padding=pad).view(bs, ch, -1)
        y = F.conv1d(y.reshape(bs * ch, 1, L), k, padding=pad).view(bs, ch, -1)
    total = 0.0
    for n_fft, hop, win in resolutions:
        xm = stft_magnitude(x.reshape(-1, x.shape[-1]), n_fft, hop, win)
        ym = stft_magnitude(y.reshape(-1, y.shape[-1]), n_fft, hop, win)
        sc = torch.norm(ym - xm, p="fro") / torch.norm(ym, p="fro")
        lg = (torch.log(xm) - torch.log(ym)).abs().mean()
        total = total + sc + lg
    return total / len(resolutions)


# ------------------------------------------------------------------ training step
class OracleEBENStep:
    """Lightning-free restatement of EBENLightningModule.training_step
    (lightning_modules/eben.py:82-130,184-240) with the semantics of SURVEY App. B:
    toggle_optimizer = requires_grad flips, manual_backward = .backward(),
    optimizer.step(); optimizer.zero_grad(); self.log = recorded into a dict."""

    def __init__(self, m=4, n=32, p=2, q=4, min_channels=24, seed=42, dtype=torch.float32,
                 lr=3e-4, betas=(0.5, 0.9), balancing="ema", beta_ema=0.9,
                 g_state: State = None, d_state: State = None, device="cpu", host_logs: bool = True):
        if g_state is None or d_state is None:
            torch.manual_seed(seed)
            g_state = init_generator_state(m, n, p)
            d_state = init_discriminator_state(q, min_channels)
        self.m, self.n, self.p, self.q, self.mc = m, n, p, q, min_channels
        self.dtype = dtype
        # device != "cpu" is the "same arithmetic through ATen/cuDNN on the GPU" baseline leg of bench.py;
        # host_logs=False keeps the logged scalars on the device (no per-loss host sync, like Lightning's self.log)
        self.device, self.host_logs = torch.device(device), host_logs
        self.g = OrderedDict((k, v.detach().clone().to(device=self.device, dtype=dtype)) for k, v in g_state.items())
        self.d = OrderedDict((k, v.detach().clone().to(device=self.device, dtype=dtype)) for k, v in d_state.items())
        self.g_train = [k for k in self.g if not k.startswith("pqmf.")]   # pqmf.py:51-56
        for k in self.g_train:
            self.g[k].requires_grad_(True)
        for v in self.d.values():
            v.requires_grad_(True)
        self.opt_g = torch.optim.Adam([self.g[k] for k in self.g_train], lr=lr, betas=betas)
        self.opt_d = torch.optim.Adam(list(self.d.values()), lr=lr, betas=betas)
        self.taps = a_weighting_fir().to(device=self.device, dtype=dtype)
        self.balancing, self.beta_ema = balancing, beta_ema
        self.norms_old = None                                   # eben.py:73
        self.last = {}

    def _D(self, bands, audio):
        return discriminator_forward(self.d, bands, audio, self.q, self.mc)

    def generator_losses(self, enhanced, reference, enh_bands, ref_bands):
        """compute_atomic_losses('generator', ...)  eben.py:194-211."""
        out = OrderedDict()
        out["reconstructive_loss_freq"] = mrstft_loss(enhanced, reference, self.taps)
        e = self._D(enh_bands, enhanced)
        r = self._D(ref_bands, reference)
        out["feature_matching_loss"] = feature_matching_loss(e, r)
        out["adv_loss_gen"] = hinge_loss(e, 1)
        return out

    def discriminator_losses(self, enhanced, reference, enh_bands, ref_bands):
        """compute_atomic_losses('discriminator', ...)  eben.py:212-219."""
        e = self._D(enh_bands.detach(), enhanced.detach())
        r = self._D(ref_bands, reference)
        return OrderedDict(real_loss=hinge_loss(r, 1), fake_loss=hinge_loss(e, -1))

    def balance(self, losses):
        """dynamically_balance_losses  eben.py:222-240."""
        w = self.g["last_conv.weight"]
        norms = [torch.autograd.grad(l, w, retain_graph=True)[0].norm().detach()
                 for l in losses.values()]
        if self.norms_old is None or self.balancing == "simple":
            self.norms_old = norms
        if self.balancing == "ema":
            self.norms_old = [self.beta_ema * o + (1 - self.beta_ema) * nw
                              for o, nw in zip(self.norms_old, norms)]
        lambdas = [torch.clamp(1 / (nm + 1e-4), min=0.0, max=1e4) for nm in self.norms_old]
        self.last["norms"] = [self._log(x) for x in norms]
        self.last["lambdas"] = [self._log(x) for x in lambdas]
        return OrderedDict((k, v * lam) for (k, v), lam in zip(losses.items(), lambdas))

    def _log(self, v):
        return float(v.detach()) if self.host_logs else v.detach()

    def step(self, body: Tensor, air: Tensor, keep_grads: bool = False) -> Dict[str, float]:
        logs: Dict[str, float] = {}
        x = cut_to_valid_length(body.to(device=self.device, dtype=self.dtype), self.n, self.m)
        y = cut_to_valid_length(air.to(device=self.device, dtype=self.dtype), self.n, self.m)
        # ---- generator phase (D frozen: toggle_optimizer)
        for v in self.d.values():
            v.requires_grad_(False)
        enhanced, enh_bands = generator_forward(self.g, x, self.p)
        ref_bands = pqmf_analysis(y, self.g["pqmf.analysis_weights"])
        losses = self.generator_losses(enhanced, y, enh_bands, ref_bands)
        for k, v in losses.items():
            logs["generator/" + k] = self._log(v)
        if self.balancing is not None:
            losses = self.balance(losses)
        total = sum(losses.values())
        logs["generator/backprop_loss"] = self._log(total)
        total.backward()
        if keep_grads:
            self.last["g_grads"] = {k: self.g[k].grad.detach().clone() for k in self.g_train}
        self.opt_g.step()
        self.opt_g.zero_grad()
        for v in self.d.values():
            v.requires_grad_(True)
        # ---- discriminator phase (G frozen)
        for k in self.g_train:
            self.g[k].requires_grad_(False)
        dl = self.discriminator_losses(enhanced, y, enh_bands, ref_bands)
        for k, v in dl.items():
            logs["discriminator/" + k] = self._log(v)
        back = dl["real_loss"] + dl["fake_loss"]
        logs["discriminator/backprop_loss"] = self._log(back)
        back.backward()
        if keep_grads:
            self.last["d_grads"] = {k: v.grad.detach().clone() for k, v in self.d.items()}
        self.opt_d.step()
        self.opt_d.zero_grad()
        for k in self.g_train:
            self.g[k].requires_grad_(True)
        self.last["enhanced"] = enhanced.detach()
        self.last["enh_bands"] = enh_bands.detach()
        return logs


def synthetic_pairs(batch: int, samples: int, seed: int, device="cpu") -> Tuple[Tensor, Tensor]:
    """SURVEY 8(d) synthetic inputs: 0.1*randn clamped to +-1, (B,1,samples) x2.
    Generated on the host generator so CPU oracle and GPU path see identical data."""
    g = torch.Generator().manual_seed(seed)
    air = (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)
    body = (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)
    return body.to(device), air.to(device)
