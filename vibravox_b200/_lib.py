"""ctypes binding of libvbx_b200.so (include/vbx.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or
does not export the ABI, importing an op raises.  torch is imported first so that the
library resolves the same libcudart.so.12 PyTorch already loaded (one runtime, shared
streams / primary context).
"""
from __future__ import annotations

import ctypes
import os

import torch  # noqa: F401  (must precede the CDLL load, see module docstring)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvbx_b200.so")
ABI_VERSION = 7

c_int, c_i64, c_f, c_d, c_p = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p


class ConvDesc(ctypes.Structure):
    """vbx_conv_desc"""
    _fields_ = [(k, c_int) for k in
                ("B", "Cin", "Cout", "Tin", "Tout", "K", "stride", "dil", "pad", "refl", "groups")]


class Epilogue(ctypes.Structure):
    """vbx_epilogue"""
    _fields_ = [("bias", c_p), ("res", c_p), ("mask", c_p), ("slope", c_f), ("beta", c_f),
                ("gate", c_p), ("fm_other", c_p), ("fm_coef", c_p), ("gate_slope", c_f), ("gate_dbias", c_p)]


_PD, _PE = ctypes.POINTER(ConvDesc), ctypes.POINTER(Epilogue)

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/vbx.h 1:1
SIGNATURES = {
    "vbx_abi_version": [],
    "vbx_last_error": [],
    "vbx_launch_count": [],
    "vbx_set_tensor_core_mode": [c_int],
    "vbx_set_deterministic": [c_int],
    "vbx_gather_scalars": [c_p] * 8 + [c_int, c_f, c_p, c_p],
    "vbx_conv1d_fwd": [_PD, c_p, c_p, _PE, c_p, c_p],
    "vbx_conv1d_dgrad": [_PD, c_p, c_p, _PE, c_p, c_p],
    "vbx_conv1d_wgrad": [_PD, c_p, c_p, c_p, c_p],
    "vbx_conv1d_dgrad_scatter": [_PD, c_p, c_p, c_p, c_p],
    "vbx_tc_pack_bytes": [_PD, c_int, c_int],
    "vbx_tc_pack": [_PD, c_int, c_int, c_p, c_p, c_p],
    "vbx_tc_conv1d_fwd": [_PD, c_p, c_p, _PE, c_p, c_int, c_p],
    "vbx_tc_conv1d_dgrad": [_PD, c_p, c_p, _PE, c_p, c_int, c_p],
    "vbx_tc_conv1d_wgrad": [_PD, c_p, c_p, c_p, c_p],
    "vbx_ru_supported": [c_int, c_int, c_int, c_int],
    "vbx_ru_pack_bytes": [c_int],
    "vbx_ru_pack": [c_int, c_p, c_p, c_p, c_p],
    "vbx_ru_set_profile_buffer": [c_p],
    "vbx_ru_fwd": [c_int, c_int, c_int, c_int, c_f, c_p, c_p, c_p, c_p, c_p, c_p],
    "vbx_ru_wgrad_workspace": [c_int, c_int, c_int, c_int, c_int],
    "vbx_ru_wgrad": [c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_f, c_p, c_p],
    "vbx_transpose_weight": [c_p, c_p, c_int, c_int, c_int, c_int, c_p],
    "vbx_weight_norm_fwd": [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p],
    "vbx_weight_norm_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_f, c_p],
    "vbx_pqmf_analysis": [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "vbx_pqmf_synthesis": [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "vbx_leaky_relu_fwd": [c_p, c_p, c_i64, c_f, c_p],
    "vbx_leaky_relu_bwd": [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_f, c_f, c_p],
    "vbx_tanh_recompose_fwd": [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p],
    "vbx_tanh_bwd": [c_p, c_p, c_p, c_i64, c_p],
    "vbx_add": [c_p, c_p, c_p, c_i64, c_p],
    "vbx_axpby": [c_p, c_p, c_i64, c_f, c_f, c_p],
    "vbx_axpby_dev": [c_p, c_p, c_i64, c_p, c_f, c_p],
    "vbx_l1_pair_sums": [c_p, c_p, c_i64, c_p, c_p],
    "vbx_fm_finalize": [c_p, c_int, c_f, c_p, c_p],
    "vbx_l1_pair_bwd": [c_p, c_p, c_i64, c_p, c_p, c_f, c_p, c_p, c_p],
    "vbx_reflect_fold_k3": [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p],
    "vbx_unit_combine": [c_p, c_p, c_int, c_int, c_p, c_p],
    "vbx_unit_split_grads": [c_p, c_p, c_p, c_int, c_int, c_p, c_p, c_f, c_p],
    "vbx_fm_coef": [c_p, c_int, c_p, c_f, c_p, c_p],
    "vbx_fm_gate_bwd": [c_p, c_p, c_p, c_f, c_p, c_i64, c_p, c_p],
    "vbx_hinge_fwd": [c_p, c_i64, c_f, c_f, c_p, c_p],
    "vbx_hinge_bwd": [c_p, c_i64, c_f, c_f, c_p, c_p, c_p],
    "vbx_d2f": [c_p, c_p, c_int, c_f, c_p],
    "vbx_unfold_frames": [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "vbx_fold_frames": [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_f, c_p],
    "vbx_stft_stats": [c_p, c_p, c_int, c_int, c_int, c_f, c_p, c_p],
    "vbx_stft_finalize": [c_p, c_p, c_int, c_f, c_p, c_p],
    "vbx_stft_bwd": [c_p, c_p, c_int, c_int, c_int, c_f, c_p, c_d, c_p, c_f, c_p, c_p],
    "vbx_weighted_sum": [c_p, c_p, c_p, c_p, c_int, c_p, c_p, c_p, c_p],
    "vbx_scalar_mul": [c_p, c_p, c_p, c_int, c_p],
    "vbx_sumsq": [c_p, c_i64, c_p, c_p],
    "vbx_balance": [c_p, c_p, c_p, c_p, c_p, c_int, c_f, c_int, c_p],
    "vbx_adam_tick": [c_p, c_p],
    "vbx_adam_step": [c_p, c_p, c_p, c_p, c_i64, c_p, c_f, c_f, c_f, c_f, c_f, c_p],
    "vbx_fill": [c_p, c_i64, c_f, c_p],
    "vbx_noise_mix_crop": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p],
}
_RESTYPES = {"vbx_last_error": ctypes.c_char_p, "vbx_launch_count": ctypes.c_uint64,
             "vbx_tc_pack_bytes": ctypes.c_int64, "vbx_ru_pack_bytes": ctypes.c_int64,
             "vbx_ru_wgrad_workspace": ctypes.c_int64}


class VbxError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises - never falls back - when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VbxError(f"{LIB_PATH} not built: run `python -m vibravox_b200.build` "
                       "(the EBEN ops have no CPU / PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise VbxError(f"{LIB_PATH} does not export {name}")
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    if lib.vbx_abi_version() != ABI_VERSION:
        raise VbxError(f"ABI mismatch: library {lib.vbx_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, name: str) -> None:
    if rc != 0:
        msg = load().vbx_last_error().decode(errors="replace")
        raise VbxError(f"{name} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().vbx_launch_count())
