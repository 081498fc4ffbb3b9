"""TEST INFRASTRUCTURE: torch-CPU stand-ins for vibravox_b200.ops so that the *Python* side of the
product (autograd.Functions, drop-in modules, FlatAdam, the training-step schedule) can be
exercised in the GPU-less build container.  Installed only by tests via `with cpu_ops():`;
nothing under vibravox_b200/ imports this and the real ops keep refusing CPU tensors."""
import contextlib

import torch
import torch.nn.functional as F

from vibravox_b200 import ops


def _padded(x, g):
    if g.refl:
        x = F.pad(x, (g.refl, g.refl), mode="reflect")
    if g.pad - g.refl:
        x = F.pad(x, (g.pad - g.refl, g.pad - g.refl))
    return x


def _conv(x, w, g):
    return F.conv1d(_padded(x, g), w, None, g.stride, 0, g.dil, g.groups)


def _epi(v, bias, res, slope, want_mask, out, beta):
    if bias is not None:
        v = v + bias.view(1, -1, 1)
    mask = (v > 0).to(torch.uint8) if want_mask else None
    if slope != 1.0:
        v = F.leaky_relu(v, slope)
    if res is not None:
        v = v + res
    if out is not None:
        v = out.copy_(v + beta * out if beta else v)
    return (v, mask) if want_mask else v


def conv1d_fwd(x, w, g, bias=None, res=None, slope=1.0, want_mask=False, out=None, beta=0.0, nsplit=2):
    return _epi(_conv(x, w, g), bias, res, slope, want_mask, out, beta)


def _untranspose(wt, g):
    Cin_g, Cout_g = g.Cin // g.groups, g.Cout // g.groups
    return wt.reshape(g.groups, Cin_g, Cout_g, g.K).permute(0, 2, 1, 3).reshape(g.Cout, Cin_g, g.K)


def conv1d_dgrad(dy, wt, g, Tin, res=None, slope=1.0, out=None, beta=0.0, bias=None, nsplit=2, gate=None):
    with torch.enable_grad():
        x0 = torch.zeros(dy.shape[0], g.Cin, Tin, requires_grad=True)
        (dx,) = torch.autograd.grad(_conv(x0, _untranspose(wt, g), g), x0, dy)
    if gate is None:
        return _epi(dx, bias, res, slope, False, out, beta)
    assert out is None and not beta
    y, gslope, other, coef = gate[:4]
    got = fm_gate_bwd(y, other, coef, gslope, _epi(dx, bias, res, slope, False, None, 0.0))
    if len(gate) > 4 and gate[4] is not None:
        gate[4].add_(got.sum((0, 2)))
    return got


def fm_coef(sums, n, go, scale):
    s = sums.view(n, 2)
    g = go.view(()).float() * scale
    return torch.stack([(1.0 / s[:, 1]).float() * g, (s[:, 0] / (s[:, 1] * s[:, 1])).float() * g], 1).reshape(-1)


def fm_gate_bwd(y, other, coef, gate_slope, g):
    v = g if g is not None else torch.zeros_like(y)
    if other is not None:
        v = v + (coef[0] * torch.sign(y - other) - coef[1] * torch.sign(y))
    return v * torch.where(y > 0, 1.0, gate_slope)


def unit_combine(w1, w2):
    return torch.einsum("om,mik->oik", w2.reshape(w2.shape[0], w2.shape[1]), w1).contiguous()


def unit_split_grads(dwf, w1, w2, want1, want2):
    w2m = w2.reshape(w2.shape[0], w2.shape[1])
    dw1 = torch.einsum("om,oik->mik", w2m, dwf).contiguous() if want1 else None
    dw2 = torch.einsum("oik,mik->om", dwf, w1).reshape(w2.shape).contiguous() if want2 else None
    return dw1, dw2


def record_event():
    return None


def wait_event(ev):
    return None


def conv1d_wgrad(x, dy, g, dw=None):
    with torch.enable_grad():
        w0 = torch.zeros(g.Cout, g.Cin // g.groups, g.K, requires_grad=True)
        (gw,) = torch.autograd.grad(_conv(x, w0, g), w0, dy)
    if dw is None:
        return gw
    dw += gw
    return dw


def conv1d_dgrad_scatter(dy, wk, g, Tin, dx=None):
    Cin_g, Cout_g = g.Cin // g.groups, g.Cout // g.groups
    w = wk.reshape(g.groups, Cin_g * g.K, Cout_g).permute(0, 2, 1).reshape(g.Cout, Cin_g, g.K)
    wt = w.view(g.groups, Cout_g, Cin_g, g.K).permute(0, 2, 1, 3).contiguous()
    got = conv1d_dgrad(dy, wt, g, Tin)
    if dx is None:
        return got
    dx += got
    return dx


def transpose_weight(w, groups):
    Cout, Cin_g, K = w.shape
    return w.view(groups, Cout // groups, Cin_g, K).permute(0, 2, 1, 3).contiguous().view(Cout, Cin_g, K)


def weight_norm_fwd(g, v, groups, want_wt=True):
    nrm = v.flatten(1).norm(dim=1)
    w = v * (g.view(-1) / nrm).view(-1, 1, 1)
    return w, (transpose_weight(w, groups) if want_wt else None), 1.0 / nrm


def weight_norm_bwd(g, v, inv, dw, dg=None, dv=None, beta=0.0):
    dot = (dw * v).flatten(1).sum(1)
    gg = (dot * inv).view_as(g)
    gv = (g.view(-1) * inv).view(-1, 1, 1) * dw - (g.view(-1) * dot * inv ** 3).view(-1, 1, 1) * v
    if dg is None:
        return gg, gv
    dg.copy_(beta * dg + gg)
    dv.copy_(beta * dv + gv)
    return dg, dv


def pqmf_analysis(x, w, bands, T=None, x_per_band=False):
    m, _, n = w.shape
    L = x.shape[2]
    if T is None:
        T = (L + n - 2) // m + 1
    need = m * (T - 1) + n - (n - 1)          # last index touched + 1
    xp = F.pad(x, (n - 1, max(0, need - L)))
    if x_per_band:
        y = F.conv1d(xp, w[:bands], None, m, 0, 1, bands)
    else:
        y = F.conv1d(xp, w[:bands], None, m)
    return y[:, :, :T].contiguous()


def pqmf_synthesis(x, w, sum_bands, L=None):
    m, _, n = w.shape
    B, bands, T = x.shape
    if L is None:
        L = m * T - n
    full = F.conv_transpose1d(x, w[:bands], None, m, 0, 0, bands)      # length (T-1)m + n, position u + n-1
    full = full[:, :, n - 1:]
    if full.shape[2] < L:
        full = F.pad(full, (0, L - full.shape[2]))
    full = full[:, :, :L]
    return full.sum(1, keepdim=True) if sum_bands else full.contiguous()


def leaky_relu_fwd(x, slope):
    return F.leaky_relu(x, slope)


def leaky_relu_bwd(dy, ref, slope, mask=None, dbias=None, want_dx=True):
    pos = mask.bool() if mask is not None else (ref > 0 if ref is not None else torch.ones_like(dy, dtype=torch.bool))
    v = dy * torch.where(pos, 1.0, slope)
    if dbias is not None:
        dbias += v.sum((0, 2))
    return v if want_dx else None


def tanh_recompose_fwd(x, first, p):
    x = x.clone()
    if p:
        x[:, :p] += first
    return torch.tanh(x)


def tanh_bwd(dy, y):
    return dy * (1 - y * y)


def add(a, b):
    return a + b


def axpby(x, y, alpha, beta):
    y.copy_(alpha * x + beta * y if beta else alpha * x)
    return y


def axpby_dev(x, y, alpha, beta):
    y.copy_(alpha.view(-1)[0] * x + (beta * y if beta else 0))
    return y


def gather_scalars(scalars, out, scale=1.0):
    for i, v in enumerate(scalars):
        out[i] = scale * v.reshape(())
    return out


def use_fused_unit(C, T, dil, B=1):
    return False


def unit_wgrad_workspace(B, C, T, dil, K):
    return -1


def set_deterministic(on):
    return False


def fill(t, value):
    return t.fill_(value)


def l1_pair_sums(a, b, sums):
    sums[0] += (a - b).abs().double().sum()
    sums[1] += a.abs().double().sum()


def fm_finalize(sums, npairs, scale):
    return ((sums[0::2] / sums[1::2]).float().sum() * scale).view(())


def l1_pair_bwd(a, b, sums, go, scale, want_da, want_db):
    s_ab, s_a = sums[0], sums[1]
    g = go.view(()) * scale
    sd = torch.sign(a - b)
    da = (g * (sd / s_a - s_ab / s_a ** 2 * torch.sign(a))).float() if want_da else None
    db = (-g * sd / s_a).float() if want_db else None
    return da, db


def hinge_fwd(c, target, scale, acc):
    acc += F.relu(1 - target * c).double().sum() * scale


def hinge_bwd(c, target, scale, go):
    return torch.where(1 - target * c > 0, -target * go.view(()) * scale, torch.zeros(())).float()


def d2f(src, scale=1.0):
    return (src * scale).float()


def unfold_frames(x, K, hop, pad):
    xp = F.pad(x, (pad, pad), mode="reflect")
    return xp.unfold(2, K, hop)[:, 0].transpose(1, 2).contiguous()


def fold_frames(dU, L, hop, pad, dx=None):
    with torch.enable_grad():
        x0 = torch.zeros(dU.shape[0], 1, L, requires_grad=True)
        (g,) = torch.autograd.grad(unfold_frames(x0, dU.shape[1], hop, pad), x0, dU)
    if dx is None:
        return g
    dx += g
    return dx


def _mags(X, eps):
    bins = X.shape[1] // 2
    return torch.sqrt(torch.clamp(X[:, :bins] ** 2 + X[:, bins:] ** 2, min=eps))


def stft_stats(X, Y, eps, stats):
    xm, ym = _mags(X, eps), _mags(Y, eps)
    stats[0] += ((ym - xm).double() ** 2).sum()
    stats[1] += (ym.double() ** 2).sum()
    stats[2] += (xm.log() - ym.log()).abs().double().sum()


def stft_finalize(stats, counts, nres, w):
    s = stats.view(nres, 3)
    return ((s[:, 0].sqrt() / s[:, 1].sqrt() + s[:, 2] / counts).float().sum() * w).view(())


def stft_bwd(X, Y, eps, stats, count, go, w):
    with torch.enable_grad():
        Xr = X.detach().clone().requires_grad_(True)
        xm, ym = _mags(Xr, eps), _mags(Y, eps)
        loss = (ym - xm).norm() / ym.norm() + (xm.log() - ym.log()).abs().mean()
        (g,) = torch.autograd.grad(loss * w, Xr, go.view(()))
    return g


def weighted_sum(xs, lam):
    terms = torch.stack([x.view(()) * (lam[i] if lam is not None else 1.0) for i, x in enumerate(xs)])
    return terms.sum(), terms


def scalar_mul(go, lam, n):
    return go.view(()) * (lam if lam is not None else torch.ones(n))


def sumsq(x, acc):
    acc += (x.double() ** 2).sum()


def balance(sumsq_t, norms_old, initialised, lambdas, norms_out, beta_ema, mode):
    nm = sumsq_t.sqrt().float()
    norms_out.copy_(nm)
    old = nm.clone() if (int(initialised[0]) == 0 or mode == 0) else norms_old.clone()
    if mode == 1:
        old = beta_ema * old + (1 - beta_ema) * nm
    norms_old.copy_(old)
    lambdas.copy_(torch.clamp(1 / (old + 1e-4), 0.0, 1e4))
    initialised[0] = 1


def adam_tick(step):
    step += 1


def adam_step(p, grad, m, v, step, lr, b1, b2, eps, grad_scale=1.0):
    t = int(step[0])
    g = grad * grad_scale
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = v.sqrt() / (1 - b2 ** t) ** 0.5 + eps
    p.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))


def noise_mix_crop(body, air, noise, start, off, length):
    """(body + noise[start : start + Ls])[off : off + length] and air[off : off + length], per item."""
    Ls = body.shape[-1]
    ob = torch.stack([(body[b, 0] + noise[b, 0, int(s): int(s) + Ls])[int(o): int(o) + length]
                      for b, (s, o) in enumerate(zip(start, off))]).unsqueeze(1)
    oa = torch.stack([air[b, 0, int(o): int(o) + length] for b, o in enumerate(off)]).unsqueeze(1)
    return ob, oa


def require_cuda(device):
    return None


_NAMES = [k for k, v in list(globals().items()) if callable(v) and not k.startswith("_") and hasattr(ops, k)
          and k not in ("contextlib",)]


@contextlib.contextmanager
def cpu_ops():
    saved = {k: getattr(ops, k) for k in _NAMES}
    saved["TC_ENABLED"] = ops.TC_ENABLED
    saved["STFT_VIA_FRAMES"] = ops.STFT_VIA_FRAMES
    ops.TC_ENABLED = False
    ops.STFT_VIA_FRAMES = True
    try:
        for k in _NAMES:
            setattr(ops, k, globals()[k])
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
