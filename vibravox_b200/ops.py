"""Tensor-level wrappers over the C ABI (include/vbx.h): raw device pointers + the current
CUDA stream go down, nothing comes back but a status code.  No autograd here (see
functional.py) and no fallback: every function raises unless its arguments are CUDA fp32
contiguous tensors and libvbx_b200.so is loaded.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, Epilogue, check

Tensor = torch.Tensor


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor], dtype=torch.float32) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.VbxError("vibravox_b200 ops need CUDA tensors (there is no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.VbxError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.VbxError("expected a contiguous tensor")
    return t.data_ptr()


def require_cuda(device) -> None:
    if torch.device(device).type != "cuda":
        raise _lib.VbxError("vibravox_b200 needs CUDA tensors (there is no CPU fallback)")


@dataclass(frozen=True)
class ConvGeom:
    """Static geometry of a Conv1d layer (channels, taps, stride, dilation, halo)."""
    Cin: int
    Cout: int
    K: int
    stride: int = 1
    dil: int = 1
    pad: int = 0      # total halo each side
    refl: int = 0     # of which mirrored (PyTorch 'reflect'); the rest is zeros
    groups: int = 1

    def tout(self, Tin: int) -> int:
        span = Tin + 2 * self.pad - self.dil * (self.K - 1) - 1
        if span < 0:      # same condition (and wording) as ATen's conv shape check
            raise _lib.VbxError(f"Calculated padded input size per channel: ({Tin + 2 * self.pad}). Kernel size: "
                                f"({self.dil * (self.K - 1) + 1}). Kernel size can't be greater than actual input size")
        if self.refl > Tin - 1:
            raise _lib.VbxError(f"Padding size should be less than the corresponding input dimension, but got: "
                                f"padding ({self.refl}, {self.refl}) at dimension 2 of input of length {Tin}")
        return span // self.stride + 1

    def desc(self, B: int, Tin: int) -> ConvDesc:
        return ConvDesc(B, self.Cin, self.Cout, Tin, self.tout(Tin), self.K, self.stride, self.dil,
                        self.pad, self.refl, self.groups)


def _epi(bias=None, res=None, mask=None, slope=1.0, beta=0.0, gate=None) -> Epilogue:
    """`gate` = (y, gate_slope, fm_other, fm_coef[, gate_dbias]): the producer-side LeakyReLU' (+ feature-matching gradient,
    + bias gradient) stage of an input-gradient epilogue (include/vbx.h: vbx_epilogue)."""
    if gate is None:
        return Epilogue(_p(bias), _p(res), _p(mask, torch.uint8), float(slope), float(beta), None, None, None, 1.0, None)
    y, gslope, other, coef = gate[:4]
    dbias = gate[4] if len(gate) > 4 else None           # bias gradient of the stage that produced y (accumulated)
    assert (other is None) == (coef is None)
    if dbias is not None:
        assert dbias.numel() == y.shape[1]
    return Epilogue(_p(bias), _p(res), _p(mask, torch.uint8), float(slope), float(beta), _p(y), _p(other), _p(coef),
                    float(gslope), _p(dbias))


# ------------------------------------------------------------------ conv family
def conv1d_fwd(x: Tensor, w: Tensor, g: ConvGeom, bias: Optional[Tensor] = None,
               res: Optional[Tensor] = None, slope: float = 1.0, want_mask: bool = False,
               out: Optional[Tensor] = None, beta: float = 0.0):
    B, Cin, Tin = x.shape
    assert Cin == g.Cin and tuple(w.shape) == (g.Cout, g.Cin // g.groups, g.K), (x.shape, w.shape, g)
    d = g.desc(B, Tin)
    y = out if out is not None else torch.empty((B, g.Cout, d.Tout), device=x.device, dtype=torch.float32)
    mask = torch.empty(y.shape, device=x.device, dtype=torch.uint8) if want_mask else None
    if res is not None:
        assert res.shape == y.shape
    e = _epi(bias, res, mask, slope, beta)
    check(_lib.load().vbx_conv1d_fwd(ctypes.byref(d), _p(x), _p(w), ctypes.byref(e), _p(y), _stream()),
          "vbx_conv1d_fwd")
    return (y, mask) if want_mask else y


def conv1d_dgrad(dy: Tensor, wt: Tensor, g: ConvGeom, Tin: int, res: Optional[Tensor] = None,
                 slope: float = 1.0, out: Optional[Tensor] = None, beta: float = 0.0,
                 bias: Optional[Tensor] = None, gate=None) -> Tensor:
    """dx (B,Cin,Tin) from dy (B,Cout,Tout); also the forward of ConvTranspose1d."""
    B, Cout, Tout = dy.shape
    d = g.desc(B, Tin)
    assert Cout == g.Cout and Tout == d.Tout, (dy.shape, g, Tin, d.Tout)
    assert wt.numel() == g.Cout * (g.Cin // g.groups) * g.K
    dx = out if out is not None else torch.empty((B, g.Cin, Tin), device=dy.device, dtype=torch.float32)
    if res is not None:
        assert res.shape == dx.shape
    if gate is not None:
        assert gate[0].shape == dx.shape and gate[0].is_contiguous() and (gate[2] is None or gate[2].shape == dx.shape)
    e = _epi(bias, res, None, slope, beta, gate)
    check(_lib.load().vbx_conv1d_dgrad(ctypes.byref(d), _p(dy), _p(wt), ctypes.byref(e), _p(dx), _stream()),
          "vbx_conv1d_dgrad")
    return dx


def conv1d_wgrad(x: Tensor, dy: Tensor, g: ConvGeom, dw: Optional[Tensor] = None) -> Tensor:
    """dw (Cout, Cin/groups, K) += sum_{b,t} dy * x.  Allocates a zeroed dw when none is given."""
    B, Cin, Tin = x.shape
    d = g.desc(B, Tin)
    assert tuple(dy.shape) == (B, g.Cout, d.Tout), (dy.shape, g, Tin)
    if dw is None:
        dw = torch.zeros((g.Cout, g.Cin // g.groups, g.K), device=x.device, dtype=torch.float32)
    check(_lib.load().vbx_conv1d_wgrad(ctypes.byref(d), _p(x), _p(dy), _p(dw), _stream()), "vbx_conv1d_wgrad")
    return dw


def conv1d_dgrad_scatter(dy: Tensor, wk: Tensor, g: ConvGeom, Tin: int, dx: Optional[Tensor] = None) -> Tensor:
    B = dy.shape[0]
    d = g.desc(B, Tin)
    assert tuple(dy.shape) == (B, g.Cout, d.Tout)
    if dx is None:
        dx = torch.zeros((B, g.Cin, Tin), device=dy.device, dtype=torch.float32)
    check(_lib.load().vbx_conv1d_dgrad_scatter(ctypes.byref(d), _p(dy), _p(wk), _p(dx), _stream()),
          "vbx_conv1d_dgrad_scatter")
    return dx


def _geom_only_desc(g: ConvGeom) -> ConvDesc:
    """A valid descriptor for calls that depend on the geometry only (weight packing)."""
    tin = max(g.dil * (g.K - 1) + 1 - 2 * g.pad, g.refl + 1, 1)
    return g.desc(1, tin + (-(tin + 2 * g.pad - g.dil * (g.K - 1) - 1)) % g.stride)


def tc_pack(w: Tensor, g: ConvGeom, mode: int, nsplit: int = 2) -> Tensor:
    """bf16 hi/lo(/lo2) K-major weight tiles for the tensor-core kernels (mode 0 = fwd, 1 = dgrad)."""
    d = _geom_only_desc(g)
    nbytes = _lib.load().vbx_tc_pack_bytes(ctypes.byref(d), mode, nsplit)
    if nbytes <= 0:
        raise _lib.VbxError("vbx_tc_pack_bytes: " + _lib.load().vbx_last_error().decode())
    packed = torch.empty((nbytes,), device=w.device, dtype=torch.uint8)
    check(_lib.load().vbx_tc_pack(ctypes.byref(d), mode, nsplit, _p(w), packed.data_ptr(), _stream()), "vbx_tc_pack")
    return packed


def tc_conv1d_fwd(x: Tensor, packed: Tensor, g: ConvGeom, bias: Optional[Tensor] = None,
                  res: Optional[Tensor] = None, slope: float = 1.0, want_mask: bool = False, nsplit: int = 2):
    B, Cin, Tin = x.shape
    assert Cin == g.Cin
    d = g.desc(B, Tin)
    y = torch.empty((B, g.Cout, d.Tout), device=x.device, dtype=torch.float32)
    mask = torch.empty(y.shape, device=x.device, dtype=torch.uint8) if want_mask else None
    if res is not None:
        assert res.shape == y.shape
    e = _epi(bias, res, mask, slope, 0.0)
    check(_lib.load().vbx_tc_conv1d_fwd(ctypes.byref(d), _p(x), packed.data_ptr(), ctypes.byref(e), _p(y), nsplit,
                                        _stream()), "vbx_tc_conv1d_fwd")
    return (y, mask) if want_mask else y


def tc_conv1d_dgrad(dy: Tensor, packed: Tensor, g: ConvGeom, Tin: int, res: Optional[Tensor] = None,
                    slope: float = 1.0, nsplit: int = 2, gate=None) -> Tensor:
    B, Cout, Tout = dy.shape
    d = g.desc(B, Tin)
    assert Cout == g.Cout and Tout == d.Tout, (dy.shape, g, Tin, d.Tout)
    dx = torch.empty((B, g.Cin, Tin), device=dy.device, dtype=torch.float32)
    if res is not None:
        assert res.shape == dx.shape
    if gate is not None:
        assert gate[0].shape == dx.shape and gate[0].is_contiguous() and (gate[2] is None or gate[2].shape == dx.shape)
    e = _epi(None, res, None, slope, 0.0, gate)
    check(_lib.load().vbx_tc_conv1d_dgrad(ctypes.byref(d), _p(dy), packed.data_ptr(), ctypes.byref(e), _p(dx),
                                          nsplit, _stream()), "vbx_tc_conv1d_dgrad")
    return dx


def tc_conv1d_wgrad(x: Tensor, dy: Tensor, g: ConvGeom, dw: Optional[Tensor] = None) -> Tensor:
    B, Cin, Tin = x.shape
    d = g.desc(B, Tin)
    assert tuple(dy.shape) == (B, g.Cout, d.Tout), (dy.shape, g, Tin)
    if dw is None:
        dw = torch.zeros((g.Cout, g.Cin // g.groups, g.K), device=x.device, dtype=torch.float32)
    check(_lib.load().vbx_tc_conv1d_wgrad(ctypes.byref(d), _p(x), _p(dy), _p(dw), _stream()), "vbx_tc_conv1d_wgrad")
    return dw


# ------------------------------------------------------------------ fused residual unit
FUSED_UNIT = os.environ.get("VBX_FUSED_UNIT", "1") != "0"


def use_fused_unit(C: int, T: int, dil: int, B: int = 1) -> bool:
    """Whether ResidualUnit(C) on a (B, C, T) input runs as the single fused kernel (include/vbx.h: vbx_ru_fwd)."""
    if not (TC_ENABLED and FUSED_UNIT):
        return False
    return bool(_lib.load().vbx_ru_supported(B, C, T, dil))


def residual_unit_pack(w_dil: Tensor, w_pw: Tensor) -> Tensor:
    C = w_dil.shape[0]
    assert tuple(w_dil.shape) == (C, C, 3) and tuple(w_pw.shape) == (C, C, 1), (w_dil.shape, w_pw.shape)
    nbytes = _lib.load().vbx_ru_pack_bytes(C)
    if nbytes <= 0:
        raise _lib.VbxError("vbx_ru_pack_bytes: unsupported channel count")
    packed = torch.empty((nbytes,), device=w_dil.device, dtype=torch.uint8)
    check(_lib.load().vbx_ru_pack(C, _p(w_dil), _p(w_pw), packed.data_ptr(), _stream()), "vbx_ru_pack")
    return packed


def residual_unit_fwd(x: Tensor, packed: Tensor, dil: int, slope: float, want_h: bool = False,
                      want_mask: bool = False):
    """out = x + LeakyReLU(pointwise(dilated(x))) in one launch; optionally also h = dilated(x) and the 1-byte
    activation mask (what the backward kernels read)."""
    B, C, T = x.shape
    out = torch.empty_like(x)
    h = torch.empty_like(x) if want_h else None
    mask = torch.empty(x.shape, device=x.device, dtype=torch.uint8) if want_mask else None
    check(_lib.load().vbx_ru_fwd(B, C, T, dil, float(slope), _p(x), packed.data_ptr(), _p(out), _p(h),
                                 _p(mask, torch.uint8), _stream()), "vbx_ru_fwd")
    return out, h, mask


def unit_wgrad_workspace(B: int, C: int, T: int, dil: int, K: int) -> int:
    """Workspace bytes of the TMA weight-gradient kernel for a residual-unit conv, or -1 when the shape is not taken."""
    if not (TC_ENABLED and FUSED_UNIT):
        return -1
    return int(_lib.load().vbx_ru_wgrad_workspace(B, C, T, dil, K))


def unit_wgrad(x: Tensor, dy: Tensor, K: int, dil: int, dw: Optional[Tensor] = None) -> Tensor:
    """dw (C, C, K) (+)= sum_{b,t} dy * x for a stride-1 C -> C conv (K = 3 dilated / reflect, or K = 1): vbx_ru_wgrad.
    Accumulates into `dw` when given, else returns a fresh tensor."""
    B, C, T = x.shape
    assert dy.shape == x.shape
    nbytes = _lib.load().vbx_ru_wgrad_workspace(B, C, T, dil, K)
    if nbytes <= 0:
        raise _lib.VbxError("vbx_ru_wgrad: unsupported shape")
    ws = torch.empty((nbytes // 4,), device=x.device, dtype=torch.float32)
    beta = 1.0 if dw is not None else 0.0
    if dw is None:
        dw = torch.empty((C, C, K), device=x.device, dtype=torch.float32)
    check(_lib.load().vbx_ru_wgrad(B, C, T, dil, K, _p(x), _p(dy), _p(dw), beta, _p(ws), _stream()), "vbx_ru_wgrad")
    return dw


def transpose_weight(w: Tensor, groups: int) -> Tensor:
    Cout, Cin_g, K = w.shape
    wt = torch.empty_like(w)
    check(_lib.load().vbx_transpose_weight(_p(w), _p(wt), Cout, Cin_g, K, groups, _stream()),
          "vbx_transpose_weight")
    return wt


# ------------------------------------------------------------------ weight norm
def weight_norm_fwd(g: Tensor, v: Tensor, groups: int, want_wt: bool = True) -> Tuple[Tensor, Optional[Tensor], Tensor]:
    R, Cin_g, K = v.shape
    assert g.numel() == R
    w = torch.empty_like(v)
    wt = torch.empty_like(v) if want_wt else None
    inv = torch.empty((R,), device=v.device, dtype=torch.float32)
    check(_lib.load().vbx_weight_norm_fwd(_p(g), _p(v), _p(w), _p(wt), _p(inv), R, Cin_g, K, groups, _stream()),
          "vbx_weight_norm_fwd")
    return w, wt, inv


def weight_norm_bwd(g: Tensor, v: Tensor, inv: Tensor, dw: Tensor, dg: Optional[Tensor] = None,
                    dv: Optional[Tensor] = None, beta: float = 0.0) -> Tuple[Tensor, Tensor]:
    R = v.shape[0]
    row = v.numel() // R
    dg = dg if dg is not None else torch.empty_like(g)
    dv = dv if dv is not None else torch.empty_like(v)
    check(_lib.load().vbx_weight_norm_bwd(_p(g), _p(v), _p(inv), _p(dw), _p(dg), _p(dv), R, row, beta, _stream()),
          "vbx_weight_norm_bwd")
    return dg, dv


# ------------------------------------------------------------------ PQMF
def pqmf_analysis(x: Tensor, w: Tensor, bands: int, T: Optional[int] = None, x_per_band: bool = False) -> Tensor:
    m, _, n = w.shape
    B, C, L = x.shape
    assert C == (bands if x_per_band else 1), "PQMF analysis expects a mono signal"
    if T is None:
        T = (L + n - 2) // m + 1
    y = torch.empty((B, bands, T), device=x.device, dtype=torch.float32)
    check(_lib.load().vbx_pqmf_analysis(_p(x), _p(w), _p(y), B, L, T, m, n, bands, int(x_per_band), _stream()),
          "vbx_pqmf_analysis")
    return y


def pqmf_synthesis(x: Tensor, w: Tensor, sum_bands: bool, L: Optional[int] = None) -> Tensor:
    m, _, n = w.shape
    B, bands, T = x.shape
    assert bands <= m
    if L is None:
        L = m * T - n
    y = torch.empty((B, 1 if sum_bands else bands, L), device=x.device, dtype=torch.float32)
    check(_lib.load().vbx_pqmf_synthesis(_p(x), _p(w), _p(y), B, T, L, m, n, bands, int(sum_bands), _stream()),
          "vbx_pqmf_synthesis")
    return y


# ------------------------------------------------------------------ element-wise
def leaky_relu_fwd(x: Tensor, slope: float) -> Tensor:
    y = torch.empty_like(x)
    check(_lib.load().vbx_leaky_relu_fwd(_p(x), _p(y), x.numel(), slope, _stream()), "vbx_leaky_relu_fwd")
    return y


def leaky_relu_bwd(dy: Tensor, ref: Optional[Tensor], slope: float, mask: Optional[Tensor] = None,
                   dbias: Optional[Tensor] = None, want_dx: bool = True) -> Optional[Tensor]:
    B, C, T = dy.shape
    dx = torch.empty_like(dy) if want_dx else None
    check(_lib.load().vbx_leaky_relu_bwd(_p(dy), _p(ref), _p(mask, torch.uint8), _p(dx), _p(dbias), B, C, T,
                                         slope, 0.0, _stream()), "vbx_leaky_relu_bwd")
    return dx


def tanh_recompose_fwd(x: Tensor, first: Optional[Tensor], p: int) -> Tensor:
    B, m, T = x.shape
    y = torch.empty_like(x)
    check(_lib.load().vbx_tanh_recompose_fwd(_p(x), _p(first), _p(y), B, m, p, T, _stream()),
          "vbx_tanh_recompose_fwd")
    return y


def tanh_bwd(dy: Tensor, y: Tensor) -> Tensor:
    dx = torch.empty_like(dy)
    check(_lib.load().vbx_tanh_bwd(_p(dy), _p(y), _p(dx), dy.numel(), _stream()), "vbx_tanh_bwd")
    return dx


def add(a: Tensor, b: Tensor) -> Tensor:
    assert a.shape == b.shape
    y = torch.empty_like(a)
    check(_lib.load().vbx_add(_p(a), _p(b), _p(y), a.numel(), _stream()), "vbx_add")
    return y


def axpby(x: Tensor, y: Tensor, alpha: float, beta: float) -> Tensor:
    check(_lib.load().vbx_axpby(_p(x), _p(y), x.numel(), alpha, beta, _stream()), "vbx_axpby")
    return y


def axpby_dev(x: Tensor, y: Tensor, alpha: Tensor, beta: float) -> Tensor:
    """y = alpha[0]*x + beta*y with a device scalar alpha (a 1-element view)."""
    check(_lib.load().vbx_axpby_dev(_p(x), _p(y), x.numel(), _p(alpha), beta, _stream()), "vbx_axpby_dev")
    return y


DETERMINISTIC = False


def set_deterministic(on: bool) -> bool:
    """Fixed-order reductions (include/vbx.h: vbx_set_deterministic); returns the previous setting."""
    global DETERMINISTIC
    DETERMINISTIC = bool(on)
    return bool(_lib.load().vbx_set_deterministic(1 if on else 0))


def gather_scalars(scalars, out: Tensor, scale: float = 1.0) -> Tensor:
    """out[i] = scale * scalars[i] for up to 8 one-element device tensors, one launch."""
    n = len(scalars)
    assert 1 <= n <= 8 and out.numel() >= n
    ptrs = [_p(s) for s in scalars] + [None] * (8 - n)
    check(_lib.load().vbx_gather_scalars(*ptrs, n, float(scale), _p(out), _stream()), "vbx_gather_scalars")
    return out


def fill(t: Tensor, value: float) -> Tensor:
    check(_lib.load().vbx_fill(_p(t), t.numel(), value, _stream()), "vbx_fill")
    return t


# ------------------------------------------------------------------ losses / reductions
def _zeros_f64(n: int, device) -> Tensor:
    return torch.zeros((n,), device=device, dtype=torch.float64)


def l1_pair_sums(a: Tensor, b: Tensor, sums: Tensor) -> None:
    assert a.shape == b.shape
    check(_lib.load().vbx_l1_pair_sums(_p(a), _p(b), a.numel(), sums.data_ptr(), _stream()), "vbx_l1_pair_sums")


def fm_finalize(sums: Tensor, npairs: int, scale: float) -> Tensor:
    loss = torch.empty((), device=sums.device, dtype=torch.float32)
    check(_lib.load().vbx_fm_finalize(_p(sums, torch.float64), npairs, scale, _p(loss), _stream()), "vbx_fm_finalize")
    return loss


def l1_pair_bwd(a: Tensor, b: Tensor, sums: Tensor, go: Tensor, scale: float, want_da: bool, want_db: bool):
    da = torch.empty_like(a) if want_da else None
    db = torch.empty_like(b) if want_db else None
    check(_lib.load().vbx_l1_pair_bwd(_p(a), _p(b), a.numel(), sums.data_ptr(), _p(go), scale, _p(da), _p(db),
                                      _stream()), "vbx_l1_pair_bwd")
    return da, db


def reflect_fold_k3(dy: Tensor, w: Tensor, dx: Tensor, dil: int) -> Tensor:
    """dx += mirror terms: turns the zero-halo input gradient of a k3 / stride-1 / reflect-halo = dilation conv into the
    reflect-halo one (vbx_reflect_fold_k3)."""
    B, C, T = dx.shape
    assert dy.shape == dx.shape and tuple(w.shape) == (C, C, 3)
    check(_lib.load().vbx_reflect_fold_k3(_p(dy), _p(w), _p(dx), B, C, T, dil, _stream()), "vbx_reflect_fold_k3")
    return dx


def unit_combine(w1: Tensor, w2: Tensor) -> Tensor:
    """wf (C, C, K) = w2 (C, C[, 1]) composed with w1 (C, C, K): the residual unit's two convs as one (vbx_unit_combine)."""
    C, _, K = w1.shape
    wf = torch.empty_like(w1)
    check(_lib.load().vbx_unit_combine(_p(w1), _p(w2), C, K, _p(wf), _stream()), "vbx_unit_combine")
    return wf


def unit_split_grads(dwf: Tensor, w1: Tensor, w2: Tensor, want1: bool, want2: bool):
    """(dw1, dw2) of the unit's two convs from the gradient of the composed weight (vbx_unit_split_grads)."""
    C, _, K = w1.shape
    dw1 = torch.empty_like(w1) if want1 else None
    dw2 = torch.empty_like(w2) if want2 else None
    check(_lib.load().vbx_unit_split_grads(_p(dwf), _p(w1), _p(w2), C, K, _p(dw1), _p(dw2), 0.0, _stream()),
          "vbx_unit_split_grads")
    return dw1, dw2


def record_event():
    """An event on the current stream (hand-over of side-channel tensors between backward nodes on different streams)."""
    ev = torch.cuda.Event()
    ev.record()
    return ev


def wait_event(ev) -> None:
    torch.cuda.current_stream().wait_event(ev)


def fm_coef(sums: Tensor, n: int, go: Tensor, scale: float) -> Tensor:
    """(2n,) floats: per layer the two scalars of the feature-matching gradient (vbx_fm_coef)."""
    coef = torch.empty((2 * n,), device=sums.device, dtype=torch.float32)
    check(_lib.load().vbx_fm_coef(sums.data_ptr(), n, _p(go), scale, _p(coef), _stream()), "vbx_fm_coef")
    return coef


def fm_gate_bwd(y: Tensor, other: Optional[Tensor], coef: Optional[Tensor], gate_slope: float,
                g: Optional[Tensor]) -> Tensor:
    """((g or 0) + feature-matching gradient of y) * LeakyReLU'(y) as one pass (vbx_fm_gate_bwd)."""
    out = torch.empty_like(y)
    check(_lib.load().vbx_fm_gate_bwd(_p(y), _p(other), _p(coef), gate_slope, _p(g), y.numel(), _p(out), _stream()),
          "vbx_fm_gate_bwd")
    return out


def hinge_fwd(c: Tensor, target: float, scale: float, acc: Tensor) -> None:
    check(_lib.load().vbx_hinge_fwd(_p(c), c.numel(), target, scale, _p(acc, torch.float64), _stream()),
          "vbx_hinge_fwd")


def hinge_bwd(c: Tensor, target: float, scale: float, go: Tensor) -> Tensor:
    dc = torch.empty_like(c)
    check(_lib.load().vbx_hinge_bwd(_p(c), c.numel(), target, scale, _p(go), _p(dc), _stream()), "vbx_hinge_bwd")
    return dc


def d2f(src: Tensor, scale: float = 1.0) -> Tensor:
    dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    check(_lib.load().vbx_d2f(_p(src, torch.float64), _p(dst), src.numel(), scale, _stream()), "vbx_d2f")
    return dst


def unfold_frames(x: Tensor, K: int, hop: int, pad: int) -> Tensor:
    """(B,1,L) -> (B,K,F) frame matrix, reflect halo `pad` (STFT framing)."""
    B, C, L = x.shape
    assert C == 1
    F = (L + 2 * pad - K) // hop + 1
    U = torch.empty((B, K, F), device=x.device, dtype=torch.float32)
    check(_lib.load().vbx_unfold_frames(_p(x), _p(U), B, L, K, hop, pad, _stream()), "vbx_unfold_frames")
    return U


def fold_frames(dU: Tensor, L: int, hop: int, pad: int, dx: Optional[Tensor] = None) -> Tensor:
    """Adjoint of unfold_frames; accumulates into dx when given."""
    B, K, F = dU.shape
    beta = 1.0 if dx is not None else 0.0
    if dx is None:
        dx = torch.empty((B, 1, L), device=dU.device, dtype=torch.float32)
    check(_lib.load().vbx_fold_frames(_p(dU), _p(dx), B, L, K, hop, pad, beta, _stream()), "vbx_fold_frames")
    return dx


def stft_stats(X: Tensor, Y: Tensor, eps: float, stats: Tensor) -> None:
    B, C2, F = X.shape
    assert X.shape == Y.shape and C2 % 2 == 0
    check(_lib.load().vbx_stft_stats(_p(X), _p(Y), B, C2 // 2, F, eps, stats.data_ptr(), _stream()), "vbx_stft_stats")


def stft_finalize(stats: Tensor, counts: Tensor, nres: int, w: float) -> Tensor:
    loss = torch.empty((), device=stats.device, dtype=torch.float32)
    check(_lib.load().vbx_stft_finalize(_p(stats, torch.float64), _p(counts, torch.float64), nres, w, _p(loss),
                                        _stream()), "vbx_stft_finalize")
    return loss


def stft_bwd(X: Tensor, Y: Tensor, eps: float, stats: Tensor, count: float, go: Tensor, w: float) -> Tensor:
    B, C2, F = X.shape
    dX = torch.empty_like(X)
    check(_lib.load().vbx_stft_bwd(_p(X), _p(Y), B, C2 // 2, F, eps, stats.data_ptr(), float(count), _p(go), w,
                                   _p(dX), _stream()), "vbx_stft_bwd")
    return dX


def weighted_sum(xs, lam: Optional[Tensor]):
    n = len(xs)
    ptrs = [_p(x) for x in xs] + [None] * (4 - n)
    terms = torch.empty((n,), device=xs[0].device, dtype=torch.float32)
    total = torch.empty((), device=xs[0].device, dtype=torch.float32)
    check(_lib.load().vbx_weighted_sum(*ptrs, n, _p(lam), _p(terms), _p(total), _stream()), "vbx_weighted_sum")
    return total, terms


def scalar_mul(go: Tensor, lam: Optional[Tensor], n: int) -> Tensor:
    out = torch.empty((n,), device=go.device, dtype=torch.float32)
    check(_lib.load().vbx_scalar_mul(_p(go), _p(lam), _p(out), n, _stream()), "vbx_scalar_mul")
    return out


def sumsq(x: Tensor, acc: Tensor) -> None:
    check(_lib.load().vbx_sumsq(_p(x), x.numel(), acc.data_ptr(), _stream()), "vbx_sumsq")


def balance(sumsq_t: Tensor, norms_old: Tensor, initialised: Tensor, lambdas: Tensor, norms_out: Tensor,
            beta_ema: float, mode: int) -> None:
    n = lambdas.numel()
    check(_lib.load().vbx_balance(_p(sumsq_t, torch.float64), _p(norms_old), _p(initialised, torch.int32),
                                  _p(lambdas), _p(norms_out), n, beta_ema, mode, _stream()), "vbx_balance")


def adam_tick(step: Tensor) -> None:
    check(_lib.load().vbx_adam_tick(_p(step, torch.int32), _stream()), "vbx_adam_tick")


def adam_step(p: Tensor, grad: Tensor, m: Tensor, v: Tensor, step: Tensor, lr: float, b1: float, b2: float,
              eps: float, grad_scale: float = 1.0) -> None:
    assert p.numel() == grad.numel() == m.numel() == v.numel()
    check(_lib.load().vbx_adam_step(_p(p), _p(grad), _p(m), _p(v), p.numel(), _p(step, torch.int32), lr, b1, b2,
                                    eps, grad_scale, _stream()), "vbx_adam_step")


def noise_mix_crop(body: Tensor, air: Tensor, noise: Tensor, start: Tensor, off: Tensor, length: int):
    B, Ls = body.shape[0], body.shape[-1]
    Ln = noise.shape[-1]
    out_body = torch.empty((B, 1, length), device=body.device, dtype=torch.float32)
    out_air = torch.empty((B, 1, length), device=body.device, dtype=torch.float32)
    check(_lib.load().vbx_noise_mix_crop(_p(body), _p(air), _p(noise), _p(start, torch.int32), _p(off, torch.int32),
                                         _p(out_body), _p(out_air), B, Ls, Ln, length, _stream()),
          "vbx_noise_mix_crop")
    return out_body, out_air


# ------------------------------------------------------------------ kernel selection
# Dense-enough layers run on the tcgen05 tensor-core kernels (bf16x3 split operands, fp32 accumulate);
# the rest (a handful of tiny-channel layers, the STFT-as-conv with stride >= 50) on the fp32 FMA
# kernels.  VBX_TC=0 forces the fp32 FMA kernels everywhere.
TC_ENABLED = os.environ.get("VBX_TC", "1") != "0"
TC_FWD, TC_DGRAD = 0, 1
# STFT as framing (unfold) + a pointwise conv over the frame axis (so that it rides the tensor-core conv
# kernels) instead of one strided conv with a 240..1200-tap kernel on the FMA path.  ON whenever the tensor-core
# path is on, and run with the 3-way operand split (nsplit = 3, "bf16x6", 24 mantissa bits): the log-magnitude
# term divides by bins that sit 60-100 dB under the frame energy (A-weighting), which turned the 2^-17 operand
# rounding of the 2-way split into a 1.8e-2 gradient error / +1 % gradient-norm bias (measured); with the 3-way
# split the gradient error is the fp32 oracle's own 4e-3 vs fp64.  VBX_STFT_VIA_FRAMES=0 (or VBX_TC=0)
# runs the STFT as one strided conv on the fp32 FMA kernel instead.
STFT_VIA_FRAMES = os.environ.get("VBX_STFT_VIA_FRAMES", "1" if TC_ENABLED else "0") == "1"


# Experiment knob (default 0 = off): weight gradients whose (input channels per group x taps) is at most this run on
# the fp32 FMA split-K kernel instead of the gather-form tensor-core one.  The narrow pointwise weight gradients
# sit at 4-8x their HBM roofline on the tensor-core kernel (M = Cout padded to 128 rows) while their FMA time is
# within reach of it (profiles/r1_layer_roofline_table.md).
WGRAD_FMA_MAX_CK = int(os.environ.get("VBX_WGRAD_FMA_MAX_CK", "0"))


def use_tc(g: ConvGeom, kind: str) -> bool:
    if not TC_ENABLED or g.stride > 8:
        return False
    cin_g, cout_g = g.Cin // g.groups, g.Cout // g.groups
    if kind == "fwd" and cin_g == 1 and cout_g <= 16 and g.K <= 128 and g.stride <= 2:
        return False                                  # one input channel per group: direct kernel (direct_conv.cu)
    if kind == "fwd" and g.groups == 1 and g.Cout in (1, 4) and g.K == 3 and g.stride == 1 and g.Cin >= 8:
        return False                                  # certainty convs / last conv: streaming kernel (skinny_fwd_kernel)
    if kind == "fwd":
        return cout_g >= 8 or cin_g * g.K >= 512
    if kind == "dgrad":
        return cin_g >= 4
    if cin_g * g.K <= WGRAD_FMA_MAX_CK:
        return False
    return cout_g >= 4 and cin_g * g.K >= 3           # wgrad


def get_pack(w: Tensor, g: ConvGeom, mode: int, nsplit: int = 2) -> Tensor:
    """Packed bf16 hi/lo tiles of `w`, cached on the tensor object (an effective weight lives for one
    phase of one step; parameters used directly are keyed by their version counter)."""
    cache = w.__dict__.setdefault("_vbx_packs", {})
    key = (g, mode, nsplit, w._version)
    pk = cache.get(key)
    if pk is None:
        pk = tc_pack(w if w.is_contiguous() else w.contiguous(), g, mode, nsplit)
        for k in [k for k in cache if k[3] != w._version]:
            del cache[k]
        cache[key] = pk
    return pk


def conv_fwd(x: Tensor, w: Tensor, g: ConvGeom, bias=None, res=None, slope: float = 1.0, want_mask: bool = False,
             nsplit: int = 2):
    if use_tc(g, "fwd"):
        return tc_conv1d_fwd(x, get_pack(w, g, TC_FWD, nsplit), g, bias=bias, res=res, slope=slope,
                             want_mask=want_mask, nsplit=nsplit)
    return conv1d_fwd(x, w, g, bias=bias, res=res, slope=slope, want_mask=want_mask)


def conv_dgrad(dy: Tensor, w: Tensor, wt: Optional[Tensor], g: ConvGeom, Tin: int, res=None, slope: float = 1.0,
               nsplit: int = 2, gate=None):
    """`gate` = (y, gate_slope, fm_other | None, fm_coef | None): dx is the gradient of the activation y; the epilogue
    adds y's feature-matching gradient and applies LeakyReLU'(y) (include/vbx.h: vbx_epilogue)."""
    if use_tc(g, "dgrad"):
        return tc_conv1d_dgrad(dy, get_pack(w, g, TC_DGRAD, nsplit), g, Tin, res=res, slope=slope, nsplit=nsplit,
                               gate=gate)
    if wt is None:
        wt = transpose_weight(w, g.groups)
    return conv1d_dgrad(dy, wt, g, Tin, res=res, slope=slope, gate=gate)


def conv_wgrad(x: Tensor, dy: Tensor, g: ConvGeom, dw: Optional[Tensor] = None) -> Tensor:
    if use_tc(g, "wgrad"):
        return tc_conv1d_wgrad(x, dy, g, dw=dw)
    return conv1d_wgrad(x, dy, g, dw=dw)
