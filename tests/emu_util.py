"""ctypes driver for tests/emu/libemu_conv.so (CPU emulation of the CUDA conv tile code)."""
import ctypes
import os
import subprocess

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emu", "libemu_conv.so")
SRC = os.path.join(HERE, "emu", "emu_conv.cpp")


class ConvDesc(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in
                ("B", "Cin", "Cout", "Tin", "Tout", "K", "stride", "dil", "pad", "refl", "groups")]


class Epilogue(ctypes.Structure):
    _fields_ = [("bias", ctypes.c_void_p), ("res", ctypes.c_void_p), ("mask", ctypes.c_void_p),
                ("slope", ctypes.c_float), ("beta", ctypes.c_float),
                ("gate", ctypes.c_void_p), ("fm_other", ctypes.c_void_p), ("fm_coef", ctypes.c_void_p),
                ("gate_slope", ctypes.c_float), ("gate_dbias", ctypes.c_void_p)]


def build():
    deps = [SRC] + [os.path.join(HERE, "..", "vibravox_b200", "csrc", f)
                    for f in ("gemm_conv.cuh", "conv_plan.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    return ctypes.CDLL(SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = build()
    return _lib


def tout(Tin, K, s, d, pad):
    return (Tin + 2 * pad - d * (K - 1) - 1) // s + 1


def transpose_weight(w, groups):
    """W[co][ci_g][k] -> Wt[g][ci_g][co_g][k]"""
    Cout, Cin_g, K = w.shape
    return w.view(groups, Cout // groups, Cin_g, K).permute(0, 2, 1, 3).contiguous()


def scatter_weight(w, groups):
    """W[co][ci_g][k] -> Wk[g][(ci,k)][co_g]"""
    Cout, Cin_g, K = w.shape
    return w.view(groups, Cout // groups, Cin_g * K).permute(0, 2, 1).contiguous()


def ref_padded(x, pad, refl):
    if refl:
        x = F.pad(x, (refl, refl), mode="reflect")
    if pad - refl:
        x = F.pad(x, (pad - refl, pad - refl))
    return x


def emu(mode, desc, a, b, out, bias=None, res=None, mask=None, slope=1.0, beta=0.0, gate=None):
    """gate = (y, gate_slope, fm_other | None, fm_coef | None): the gate stage of vbx_epilogue."""
    def ptr(t):
        return t.data_ptr() if t is not None else None
    y, gslope, other, coef = gate if gate is not None else (None, 1.0, None, None)
    e = Epilogue(ptr(bias), ptr(res), ptr(mask), slope, beta, ptr(y), ptr(other), ptr(coef), gslope, None)
    r = lib().emu_conv(mode, ctypes.byref(desc), ctypes.c_void_p(a.data_ptr()),
                       ctypes.c_void_p(b.data_ptr()), ctypes.byref(e), ctypes.c_void_p(out.data_ptr()))
    assert r == 0, r
    return out
