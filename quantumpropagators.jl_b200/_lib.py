"""ctypes binding of ``libqprop_b200.so`` (the C ABI of ``include/qprop.h``).

Stand-in for the ``ccall`` bindings of the Julia host (``julia/QPropB200.jl``): same entry
points, same argument meaning.  There is no fallback -- if the shared library is missing
or a call fails, :class:`QPropLibraryError` / :class:`QPropError` is raised.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libqprop_b200.so")

QP_OK = 0
QP_ERR_INVALID_ARG = -1
QP_ERR_CUDA = -2
QP_ERR_OOM = -3
QP_ERR_NOT_CONVERGED = -4
QP_ERR_NORMALIZATION = -5
QP_ERR_UNSUPPORTED = -6
QP_ERR_INTERNAL = -7

QP_LAYOUT_CSC = 0
QP_LAYOUT_CSR = 1

QP_FORMAT_AUTO = 0
QP_FORMAT_CSR = 1
QP_FORMAT_SELL = 2
QP_FORMAT_DENSE = 3
QP_FORMAT_SELLD = 4
QP_FORMAT_LR = 5
QP_FORMAT_BITFLIP = 6
FORMAT_NAMES = {0: "auto", 1: "csr", 2: "sell", 3: "dense", 4: "selld", 5: "leftright", 6: "bitflip"}


class QPropLibraryError(RuntimeError):
    """The CUDA library is missing or cannot be loaded (no CPU fallback exists)."""


class QPropError(RuntimeError):
    """A library call returned a non-zero status."""

    def __init__(self, status, message):
        super().__init__(f"[{status}] {message}")
        self.status = status
        self.message = message


class c128(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


def to_c128(z) -> c128:
    z = complex(z)
    return c128(z.real, z.imag)


_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double
_P = C.POINTER

# name -> (restype, argtypes); mirrors include/qprop.h declaration by declaration
SIGNATURES = {
    "qp_version": (_i32, []),
    "qp_status_string": (C.c_char_p, [_i32]),
    "qp_ctx_create": (_i32, [_i32, _P(_vp)]),
    "qp_ctx_destroy": (_i32, [_vp]),
    "qp_sync": (_i32, [_vp]),
    "qp_last_error": (C.c_char_p, [_vp]),
    "qp_ctx_stream": (_i32, [_vp, _P(_vp)]),
    "qp_ctx_launch_count": (_i32, [_vp, _P(_i64)]),
    "qp_timer_enable": (_i32, [_vp, _i32]),
    "qp_timer_get": (_i32, [_vp, C.c_char_p, _P(_i64), _P(_f64)]),
    "qp_timer_reset": (_i32, [_vp]),
    "qp_op_upload_sparse": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _P(_vp)]),
    "qp_op_upload_dense": (_i32, [_vp, _i64, _vp, _P(_vp)]),
    "qp_op_create_leftright": (_i32, [_vp, _i64, _i32, _P(_vp), _P(_vp), _vp, _P(_vp)]),
    "qp_op_destroy": (_i32, [_vp]),
    "qp_op_info": (_i32, [_vp, _P(_i64), _P(_i64), _P(_i64), _P(_i32)]),
    "qp_gen_create": (_i32, [_vp, _i32, _P(_vp), _i32, _i32, _P(_vp)]),
    "qp_gen_destroy": (_i32, [_vp]),
    "qp_gen_info": (_i32, [_vp, _P(_i32), _P(_i64), _P(_i64), _P(_i64)]),
    "qp_gen_storage": (_i32, [_vp, _P(_i64), _P(_i32), _P(_i32)]),
    "qp_gen_tile_info": (_i32, [_vp, _P(_i32), _P(_i32), _P(_i32), _P(_i32), _vp]),
    "qp_state_create": (_i32, [_vp, _i64, _i64, _P(_vp)]),
    "qp_state_destroy": (_i32, [_vp]),
    "qp_state_info": (_i32, [_vp, _P(_i64), _P(_i64)]),
    "qp_state_devptr": (_i32, [_vp, _P(_vp)]),
    "qp_state_upload": (_i32, [_vp, _vp, _i64, _i64]),
    "qp_state_download": (_i32, [_vp, _vp, _i64, _i64]),
    "qp_state_upload_async": (_i32, [_vp, _vp, _i64, _i64]),
    "qp_state_download_async": (_i32, [_vp, _vp, _i64, _i64]),
    "qp_copy": (_i32, [_vp, _vp]),
    "qp_fill": (_i32, [_vp, c128]),
    "qp_scal": (_i32, [_vp, c128]),
    "qp_axpy": (_i32, [c128, _vp, _vp]),
    "qp_dot": (_i32, [_vp, _vp, _vp]),
    "qp_norm": (_i32, [_vp, _vp]),
    "qp_gen_mul": (_i32, [_vp, _vp, c128, c128, _vp, _vp]),
    "qp_gen_dot": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "qp_cheby_create": (_i32, [_vp, _vp, _P(_vp)]),
    "qp_cheby_destroy": (_i32, [_vp]),
    "qp_cheby_set_coeffs": (_i32, [_vp, _vp, _i32, _f64, _f64, _f64, _f64]),
    "qp_cheby_step": (_i32, [_vp, _vp, _vp, _i32, _f64, _i32]),
    "qp_cheby_propagate": (_i32, [_vp, _vp, _vp, _i32, _i32, _f64, _i32, _vp, _vp, _vp]),
    "qp_gen_expval": (_i32, [_vp, _vp, _vp, _vp]),
    "qp_cheby_step_bytes": (_i32, [_vp, _P(_i64)]),
    "qp_krylov_create": (_i32, [_vp, _vp, _i32, _P(_vp)]),
    "qp_krylov_destroy": (_i32, [_vp]),
    "qp_arnoldi": (_i32, [_vp, _vp, _vp, _i32, _f64, _i32, _f64, _vp, _i32, _P(_i32)]),
    "qp_arnoldi_extend": (_i32, [_vp, _vp, _i32, _f64, _f64, _vp, _i32, _P(_i32)]),
    "qp_krylov_combine": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32]),
    "qp_newton_step": (_i32, [_vp, _vp, _vp, _vp, _f64, _i32, _vp, _vp, _f64, _f64, _i32, _P(_i32)]),
    "qp_newton_last": (_i32, [_vp, _P(_i32), _P(_i32), _P(_f64), _vp, _vp, _i32]),
    "qp_diagonalize_hessenberg": (_i32, [_vp, _i32, _i32, _i32, _vp, _P(_i32)]),
    "qp_extend_leja": (_i32, [_vp, _i32, _P(_i32), _vp, _i32, _i32]),
    "qp_extend_newton_coeffs": (_i32, [_vp, _i32, _P(_i32), _vp, _i32, _i32, _vp, _vp, _f64]),
    "qp_ens_shard": (_i32, [_i64, _i32, _i32, _P(_i64), _P(_i64)]),
    "qp_ens_create": (_i32, [_vp, _i32, _P(_vp)]),
    "qp_ens_unique_id": (_i32, [_vp]),
    "qp_ens_create_rank": (_i32, [_vp, _i32, _i32, _vp, _P(_vp)]),
    "qp_ens_destroy": (_i32, [_vp]),
    "qp_ens_info": (_i32, [_vp, _P(_i32), _P(_i32), _P(_i32)]),
    "qp_ens_ctx": (_i32, [_vp, _i32, _P(_vp), _P(_i32)]),
    "qp_ens_gather_states": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "qp_ens_gather_expvals": (_i32, [_vp, _vp, _i32, _i64, _vp]),
    "qp_krylov_get": (_i32, [_vp, _i32, _vp]),
    "qp_krylov_set": (_i32, [_vp, _i32, _vp]),
}

QP_FUNC_EXPMI, QP_FUNC_EXP, QP_FUNC_CALLBACK = 0, 1, 2
NEWTON_FUNC = C.CFUNCTYPE(None, _P(c128), _P(c128), _vp)

_lib = None


def load():
    """Load the shared library (once).  Raises QPropLibraryError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QPropLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C quantumpropagators.jl_b200/csrc`.  qprop_b200 has no CPU fallback."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover
        raise QPropLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise QPropLibraryError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, ctx_handle=None):
    if status == QP_OK:
        return
    lib = load()
    msg = lib.qp_last_error(ctx_handle)
    msg = msg.decode("utf-8", "replace") if msg else ""
    if not msg:
        msg = lib.qp_status_string(status).decode()
    raise QPropError(status, msg)


def as_c128_array(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.complex128)
    if shape is not None and out.shape != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {out.shape}")
    return out


def ptr(arr: np.ndarray):
    return arr.ctypes.data_as(C.c_void_p)
