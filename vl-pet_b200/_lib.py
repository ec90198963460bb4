"""ctypes binding of libvlpet.so (include/vlpet.h).  The product has NO CPU fallback: if the library cannot be
built or loaded, importing this module raises, and every compute call raises when it returns non-zero."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))

F32, BF16 = 0, 1
GATE_NONE, GATE_LARGE, GATE_MIDDLE_X, GATE_MIDDLE_Y, GATE_SMALL = range(5)
GATE_IDS = {"none": GATE_NONE, "large": GATE_LARGE, "middle_x": GATE_MIDDLE_X, "middle_y": GATE_MIDDLE_Y,
            "small": GATE_SMALL}
IMPL_AUTO, IMPL_GENERIC, IMPL_FUSED = 0, 1, 2
IMPL_IDS = {"auto": IMPL_AUTO, "generic": IMPL_GENERIC, "fused": IMPL_FUSED}

_vp = C.c_void_p
_fp = C.c_void_p  # float* passed as raw address


class K1Desc(C.Structure):
    _fields_ = [("M", C.c_int64), ("L", C.c_int32), ("d", C.c_int32), ("r", C.c_int32), ("rg", C.c_int32),
                ("gate", C.c_int32), ("add_gate", C.c_int32), ("dtype", C.c_int32), ("impl", C.c_int32),
                ("s", C.c_float), ("alpha", C.c_float), ("kappa", C.c_float), ("p_drop", C.c_float),
                ("seed", C.c_uint64), ("seed_dev", C.c_void_p)]


class K1Params(C.Structure):
    _fields_ = [(n, _vp) for n in ("Wd", "bd", "Wu", "bu", "Gd", "gbd", "Gu", "gbu", "gw", "gb", "gz")]


class K1Grads(C.Structure):
    _fields_ = [(n, _fp) for n in ("dWd", "dbd", "dWu", "dbu", "dGd", "dgbd", "dGu", "dgbu", "dgw", "dgb", "dgz")]


class WgradPair(C.Structure):
    _fields_ = [("A", _vp), ("lda", C.c_int64), ("B", _vp), ("ldb", C.c_int64), ("nb_valid", C.c_int32),
                ("transposed", C.c_int32), ("out", _fp), ("bias", _fp), ("scale", C.c_float)]


class K2Desc(C.Structure):
    _fields_ = [("M", C.c_int64), ("d", C.c_int32), ("r", C.c_int32), ("dtype", C.c_int32), ("impl", C.c_int32),
                ("sf", C.c_float)]


class K2Params(C.Structure):
    _fields_ = [(n, _vp) for n in ("Wd", "bd", "Wu", "bu")]


class K2Grads(C.Structure):
    _fields_ = [(n, _fp) for n in ("dWd", "dbd", "dWu", "dbu")]


class K3Desc(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int32), ("F", C.c_int32), ("d", C.c_int32), ("V", C.c_int32),
                ("n_img", C.c_int32), ("rms", C.c_int32), ("dtype", C.c_int32), ("impl", C.c_int32),
                ("eps", C.c_float)]


class K3Params(C.Structure):
    _fields_ = [(n, _vp) for n in ("Wf", "bf", "ln_f_w", "ln_f_b", "Wp", "bp", "ln_p_w", "ln_p_b", "E_img", "E_obj")]


class K3Grads(C.Structure):
    _fields_ = [(n, _fp) for n in ("dWf", "dbf", "dln_f_w", "dln_f_b", "dWp", "dbp", "dln_p_w", "dln_p_b", "dE_img")]


class K3LRDesc(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int32), ("F", C.c_int32), ("d", C.c_int32), ("r", C.c_int32), ("rg", C.c_int32),
                ("V", C.c_int32), ("n_img", C.c_int32), ("gated", C.c_int32), ("residual", C.c_int32), ("dtype", C.c_int32),
                ("impl", C.c_int32), ("eps", C.c_float)]


K3LR_PARAM_NAMES = ("Wd", "bd", "Wu", "bu", "Gd", "gbd", "Gu", "gbu", "ln_f_w", "ln_f_b", "Wp", "bp", "ln_p_w", "ln_p_b",
                    "E_img", "E_obj")
K3LR_GRAD_NAMES = ("dWd", "dbd", "dWu", "dbu", "dGd", "dgbd", "dGu", "dgbu", "dln_f_w", "dln_f_b", "dWp", "dbp", "dln_p_w",
                   "dln_p_b", "dE_img")


class K3LRParams(C.Structure):
    _fields_ = [(n, _vp) for n in K3LR_PARAM_NAMES]


class K3LRGrads(C.Structure):
    _fields_ = [(n, _fp) for n in K3LR_GRAD_NAMES]


# every symbol include/vlpet.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "vlpet_k1_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(K1Desc)]),
    "vlpet_k1_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(K1Desc)]),
    "vlpet_k1_fwd": (C.c_int, [C.POINTER(K1Desc), _vp, _vp, C.POINTER(K1Params), _vp, _vp, C.c_size_t, _vp]),
    "vlpet_k1_bwd": (C.c_int, [C.POINTER(K1Desc), _vp, _vp, _vp, C.POINTER(K1Params), _vp, _vp, C.POINTER(K1Grads),
                               _vp, C.c_size_t, _vp]),
    "vlpet_k1_fwd_is_fused": (C.c_int, [C.POINTER(K1Desc)]),
    "vlpet_k1_bwd_is_fused": (C.c_int, [C.POINTER(K1Desc)]),
    "vlpet_k2_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(K2Desc)]),
    "vlpet_k2_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(K2Desc)]),
    "vlpet_k2_fwd": (C.c_int, [C.POINTER(K2Desc), _vp, _vp, C.POINTER(K2Params), _vp, _vp, C.c_size_t, _vp]),
    "vlpet_k2_bwd": (C.c_int, [C.POINTER(K2Desc), _vp, _vp, C.POINTER(K2Params), _vp, C.POINTER(K2Grads), _vp,
                               C.c_size_t, _vp]),
    "vlpet_k2_is_fused": (C.c_int, [C.POINTER(K2Desc)]),
    "vlpet_k3_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(K3Desc)]),
    "vlpet_k3_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(K3Desc)]),
    "vlpet_k3_save_floats": (C.c_size_t, [C.POINTER(K3Desc)]),
    "vlpet_k3_fwd": (C.c_int, [C.POINTER(K3Desc), _vp, _vp, _vp, _vp, C.POINTER(K3Params), _vp, _vp, _vp, C.c_size_t,
                               _vp]),
    "vlpet_k3_bwd": (C.c_int, [C.POINTER(K3Desc), _vp, _vp, _vp, _vp, C.POINTER(K3Params), _vp, _vp,
                               C.POINTER(K3Grads), _vp, C.c_size_t, _vp]),
    "vlpet_k3lr_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(K3LRDesc)]),
    "vlpet_k3lr_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(K3LRDesc)]),
    "vlpet_k3lr_fwd": (C.c_int, [C.POINTER(K3LRDesc), _vp, _vp, _vp, _vp, C.POINTER(K3LRParams), _vp, _vp, _vp, C.c_size_t, _vp]),
    "vlpet_k3lr_bwd": (C.c_int, [C.POINTER(K3LRDesc), _vp, _vp, _vp, _vp, C.POINTER(K3LRParams), _vp, C.POINTER(K3LRGrads), _vp,
                                 C.c_size_t, _vp]),
    "vlpet_wgrad_bf16": (C.c_int, [C.POINTER(WgradPair), C.c_int32, C.c_int64, C.c_int32, C.c_int32, _vp]),
    "vlpet_layernorm_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_float, C.c_int32, _vp]),
    "vlpet_layernorm_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, _vp]),
    "vlpet_dropout_add_layernorm_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_float, C.c_float,
                                                  C.c_uint64, _vp, _vp]),
    "vlpet_dropout_add_layernorm_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_float,
                                                  C.c_uint64, _vp, _vp]),
    "vlpet_gelu_dropout_fwd": (C.c_int, [_vp, _vp, C.c_int64, C.c_float, C.c_uint64, _vp, _vp]),
    "vlpet_gelu_dropout_bwd": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_float, C.c_uint64, _vp, _vp]),
    "vlpet_attn_fwd": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_float, C.c_uint64, _vp, _vp]),
    "vlpet_attn_bwd": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64,
                                 C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, _vp, _vp]),
    "vlpet_ce_fwd": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int64, _vp]),
    "vlpet_ce_bwd": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int64, _vp]),
    "vlpet_grid_maxpool": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "vlpet_cast_f32_to_bf16": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "vlpet_adamw_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_float, C.c_int32, _vp, _vp, _vp]),
    "vlpet_adamw_step_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, _vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                       _vp, _vp, _vp]),
    "vlpet_sumsq": (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    "vlpet_version": (C.c_int, []),
    "vlpet_last_error": (C.c_char_p, []),
    "vlpet_launch_count": (C.c_uint64, []),
    "vlpet_device_info": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
}


class VlpetError(RuntimeError):
    pass


def _load():
    # Always go through build(): it is a digest comparison when the library is current, and it rebuilds a stale one
    # (the ctypes layouts below must match the loaded code).  VLPET_LIB loads a developer variant as is.
    path = os.environ.get("VLPET_LIB") or _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    return lib, path


lib, LIB_PATH = _load()


def check(code: int, what: str):
    if code != 0:
        raise VlpetError(f"{what} failed (code {code}): {lib.vlpet_last_error().decode(errors='replace')}")


def launch_count() -> int:
    return int(lib.vlpet_launch_count())
