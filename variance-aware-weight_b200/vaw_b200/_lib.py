"""ctypes binding of libvaw_b200.so (the C ABI declared in include/vaw_b200.h).

The library is the product: there is no Python/torch fallback.  `lib()` raises if the shared object is missing,
and every wrapper raises `VawError` when a call returns a negative status.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libvaw_b200.so")

F32, BF16 = 0, 1

# ModelMeanType codes (reference enum values)
MEAN_PREVIOUS_X, MEAN_START_X, MEAN_EPSILON, MEAN_VELOCITY, MEAN_VECTOR, MEAN_SCORE = 1, 2, 3, 4, 5, 6

# reverse-step modes (include/vaw_b200.h VAW_RS_*)
RS_DDPM, RS_DDIM, RS_DDIM_REVERSE, RS_MOMENTS = range(4)

W_CONSTANT, W_LAMBDA, W_MIN_SNR, W_MAX_SNR, W_DEBIAS, W_MIN_DEBIAS, W_MAX_DEBIAS, W_P2, W_TRUNC_SNR, W_SNR, W_INV_SNR = range(11)

(EPI_BF16, EPI_F32, EPI_GELU_TANH, EPI_GELU_ERF, EPI_GATE_RES, EPI_RES, EPI_DGELU_TANH, EPI_DGELU_ERF, EPI_SILU,
 EPI_DSILU, EPI_ALIGN_MSE) = range(11)


class VawError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p),
        ("lda", C.c_longlong), ("ldb", C.c_longlong),
        ("a_mn", C.c_int), ("b_mn", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("epilogue", C.c_int),
        ("out", C.c_void_p), ("out2", C.c_void_p),
        ("bias", C.c_void_p), ("resid", C.c_void_p), ("gate", C.c_void_p), ("aux", C.c_void_p),
        ("ldo", C.c_longlong), ("ldg", C.c_longlong),
        ("rows_per_sample", C.c_int), ("accumulate", C.c_int), ("tile_n", C.c_int), ("resid_mod", C.c_int),
        ("k_splits", C.c_int), ("split_ws", C.c_void_p), ("split_ws_elems", C.c_longlong), ("cta_group", C.c_int),
    ]


_P, _LL, _I, _D, _F = C.c_void_p, C.c_longlong, C.c_int, C.c_double, C.c_float

# name -> argtypes (restype is int for all of these)
_SIGS = {
    "vaw_device_check": [],
    "vaw_qsample_target": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _LL, _LL, _P],
    "vaw_wmse_fwd_bwd": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I, _LL, _LL, _P],
    "vaw_reverse_step": [_P, _I, _LL, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _LL, _LL, _P],
    "vaw_cfg_combine": [_P, _P, _I, _F, _LL, _P],
    "vaw_scale_rows": [_P, _P, _P, _I, _LL, _LL, _P],
    "vaw_loss_weight_lut": [_P, _P, _I, _I, _I, _D, _D, _D, _P],
    "vaw_sampler_sample": [_I, _P, _P, _P, _I, _I, _D, _P, _LL, _P, _P, _P, _P, _P, _P],
    "vaw_sampler_update": [_P, _P, _P, _P, _LL, _I, _I, _P],
    "vaw_pack_tloss": [_P, _P, _P, _P, _LL, _LL, _P],
    "vaw_add_f32": [_P, _P, _LL, _P],
    "vaw_gemm_bf16": [C.POINTER(GemmArgs), _P],
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly: the CUDA extension is mandatory."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VawError(
                f"{LIB_PATH} is missing: build it with `python variance-aware-weight_b200/build.py` "
                "(vaw_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        l.vaw_last_error.restype = C.c_char_p
        l.vaw_last_error.argtypes = []
        l.vaw_version.restype = C.c_int
        for name, args in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = C.c_int
            fn.argtypes = args
        _lib = l
    return _lib


def register(name: str, argtypes) -> None:
    """Late registration of a signature (used by the model engine module)."""
    _SIGS[name] = argtypes
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes


def call(name: str, *args) -> None:
    l = lib()
    rc = getattr(l, name)(*args)
    if rc != 0:
        msg = l.vaw_last_error()
        raise VawError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    """Kernels launched by the library in this process (include/vaw_b200.h: vaw_launch_count)."""
    fn = lib().vaw_launch_count
    fn.restype = C.c_ulonglong
    fn.argtypes = []
    return int(fn())


def ptr(t) -> int | None:
    """Device (or host) pointer of a torch tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VawError("vaw_b200 kernels need CUDA tensors (no CPU fallback)")
