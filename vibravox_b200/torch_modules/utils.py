"""Parameter containers and weight handling shared by the drop-in EBEN modules.

The containers are the same torch classes the reference uses
(vibravox/torch_modules/utils.py:4-9: weight_norm-parametrised nn.Conv1d /
nn.ConvTranspose1d), so `state_dict()` keys, shapes and the default initialisation under a
given seed are identical.  Their own `forward` is never called: the drop-in modules read
the raw parameters (`parametrizations.weight.original0/1`, `bias`) and run the
libvbx_b200 kernels through vibravox_b200.functional.
"""
from __future__ import annotations

import contextlib
from typing import Optional, Tuple

import torch
from torch import nn

from ..functional import TransposeWeightFn, WeightNormFn
from ..ops import ConvGeom


def normalized_conv1d(*args, **kwargs) -> nn.Conv1d:
    return nn.utils.parametrizations.weight_norm(nn.Conv1d(*args, **kwargs))


def normalized_conv_trans1d(*args, **kwargs) -> nn.ConvTranspose1d:
    return nn.utils.parametrizations.weight_norm(nn.ConvTranspose1d(*args, **kwargs))


def _one(v) -> int:
    return int(v[0]) if isinstance(v, (tuple, list)) else int(v)


def conv_geom(conv: nn.Conv1d, extra_reflect: int = 0) -> ConvGeom:
    """Geometry of an nn.Conv1d (optionally preceded by nn.ReflectionPad1d(extra_reflect))."""
    K, s, d = _one(conv.kernel_size), _one(conv.stride), _one(conv.dilation)
    if isinstance(conv.padding, str):
        if conv.padding != "same":
            raise ValueError(f"unsupported padding {conv.padding!r}")
        total = d * (K - 1)
        if total % 2:
            raise ValueError("asymmetric 'same' padding is not supported")
        pad = total // 2
    else:
        pad = _one(conv.padding)
    if conv.padding_mode == "reflect":
        refl = pad
    elif conv.padding_mode == "zeros":
        refl = 0
    else:
        raise ValueError(f"unsupported padding_mode {conv.padding_mode!r}")
    if extra_reflect:
        if refl:
            raise ValueError("ReflectionPad1d in front of a reflect-padded conv is not supported")
        pad, refl = pad + extra_reflect, extra_reflect
    return ConvGeom(conv.in_channels, conv.out_channels, K, s, d, pad, refl, conv.groups)


def conv_trans_geom(ct: nn.ConvTranspose1d) -> Tuple[ConvGeom, int]:
    """Geometry of the forward conv whose input-gradient is this ConvTranspose1d, + output_padding."""
    if ct.padding_mode != "zeros":
        raise ValueError("ConvTranspose1d supports zeros padding only")
    return (ConvGeom(ct.out_channels, ct.in_channels, _one(ct.kernel_size), _one(ct.stride),
                     _one(ct.dilation), _one(ct.padding), 0, ct.groups), _one(ct.output_padding))


# ---- effective weights ---------------------------------------------------------------------
# The reference recomputes g*v/||v|| on every forward (168x per training step).
# Inside `share_weight_norm()` a layer's effective weight is computed once and re-used by
# every forward of that scope (e.g. D(enhanced) and D(reference) of one phase); outside, it
# is recomputed per call exactly like the reference.
_WCACHE: Optional[dict] = None


@contextlib.contextmanager
def share_weight_norm():
    global _WCACHE
    prev, _WCACHE = _WCACHE, {}
    try:
        yield
    finally:
        _WCACHE = prev


def is_parametrized(conv: nn.Module) -> bool:
    return hasattr(conv, "parametrizations") and "weight" in conv.parametrizations


def effective_weight(conv: nn.Module, need_wt: bool = True):
    """(w, wt) of a conv container; wt is the group-transposed copy used by the dgrad kernel."""
    groups = conv.groups
    if is_parametrized(conv):
        p = conv.parametrizations.weight
        g, v = p.original0, p.original1
        key = None
        if _WCACHE is not None:
            key = (id(conv), g._version, v._version, g.requires_grad, v.requires_grad, torch.is_grad_enabled())
            hit = _WCACHE.get(key)
            if hit is not None:
                return hit
        out = WeightNormFn.apply(g, v, groups)
        if key is not None:
            _WCACHE[key] = out
        return out
    w = conv.weight
    wt = TransposeWeightFn.apply(w, groups) if need_wt else None
    return w, wt
