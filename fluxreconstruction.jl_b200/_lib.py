"""ctypes binding of libfrb200.so (include/frb200.h).  No torch types cross this boundary.

There is deliberately no fallback: if the shared library is missing or a compute call
fails (for instance because no B200 is present) an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FRB200_LIB selects another build of the same ABI (kernel experiments, scripts/build_variants.py)
LIB_PATH = os.environ.get("FRB200_LIB") or os.path.join(_HERE, "lib", "libfrb200.so")

c_dp = C.POINTER(C.c_double)


class FRBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfrb200 error {code}: {msg}")
        self.code = code


class Operators(C.Structure):
    _fields_ = [
        ("deg", C.c_int32),
        ("ll", c_dp), ("lr", c_dp), ("lpdm", c_dp), ("dgl", c_dp), ("dgr", c_dp), ("dll", c_dp), ("dlr", c_dp),
    ]


# name -> (restype, argtypes); mirrors include/frb200.h one to one
SIGNATURES = {
    "frb_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "frb_ctx_destroy": (C.c_int32, [C.c_void_p]),
    "frb_last_error": (C.c_char_p, [C.c_void_p]),
    "frb_device_info": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.c_char_p, C.c_int32]),
    "frb_advection1d_create": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(Operators), c_dp, C.c_double, C.c_int32,
                                           C.c_int32, C.POINTER(C.c_void_p)]),
    "frb_euler1d_create": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(Operators), c_dp, C.c_double, C.c_int32,
                                       C.POINTER(C.c_void_p)]),
    "frb_euler2d_create": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Operators), C.c_double,
                                       C.c_double, C.c_double, C.POINTER(C.c_void_p)]),
    "frb_euler2d_curv_create": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Operators), c_dp, c_dp, c_dp,
                                            c_dp, C.c_int32, C.c_double, C.POINTER(C.c_void_p)]),
    "frb_euler2d_curv_set_vertices": (C.c_int32, [C.c_void_p, c_dp, c_dp]),
    "frb_bgk1d_create": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Operators), c_dp, c_dp, c_dp,
                                     C.c_double, C.POINTER(C.c_void_p)]),
    "frb_bgk1d_set_model": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double]),
    "frb_ns2d_create": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Operators)] + [C.c_double] * 9
                        + [C.POINTER(C.c_void_p)]),
    "frb_tri_euler_create": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), c_dp, c_dp,
                                         C.POINTER(C.c_int32), c_dp, c_dp, c_dp, C.c_double, C.POINTER(C.c_void_p)]),
    "frb_prob_destroy": (C.c_int32, [C.c_void_p]),
    "frb_state_len": (C.c_int64, [C.c_void_p]),
    "frb_interior_dofs": (C.c_int64, [C.c_void_p]),
    "frb_state_upload": (C.c_int32, [C.c_void_p, c_dp]),
    "frb_state_download": (C.c_int32, [C.c_void_p, c_dp]),
    "frb_state_device_ptr": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "frb_rhs": (C.c_int32, [C.c_void_p, c_dp, c_dp, C.c_double]),
    "frb_rhs_pipelined": (C.c_int32, [C.c_void_p, c_dp, c_dp, C.c_int32]),
    "frb_step_host": (C.c_int32, [C.c_void_p, c_dp, c_dp, C.c_int32, C.c_double, C.c_int32]),
    "frb_set_step_hooks": (C.c_int32, [C.c_void_p, C.c_int32, c_dp]),
    "frb_step": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double, C.c_int32]),
    "frb_step_tableau": (C.c_int32, [C.c_void_p, C.c_int32, c_dp, c_dp, C.c_double, C.c_int32]),
    "frb_ghost_fill": (C.c_int32, [C.c_void_p, C.c_int32]),
    "frb_limiter_positivity": (C.c_int32, [C.c_void_p, c_dp, C.POINTER(C.c_int32)]),
    "frb_filter_modal": (C.c_int32, [C.c_void_p, c_dp, c_dp, C.c_int32, C.c_double, C.c_double, C.c_double,
                                     C.c_int32, C.POINTER(C.c_int32)]),
    "frb_set_filter_hook": (C.c_int32, [C.c_void_p, C.c_int32, c_dp, c_dp, C.c_int32, C.c_double, C.c_double,
                                        C.c_double, C.c_int32]),
    "frb_time_stage": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "frb_last_timing": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "frb_set_kernel": (C.c_int32, [C.c_void_p, C.c_int32]),
    "frb_set_flux": (C.c_int32, [C.c_void_p, C.c_int32]),
    "frb_set_profiling": (C.c_int32, [C.c_void_p, C.c_int32]),
    "frb_stage_timing": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "frb_host_alloc": (C.c_int32, [C.c_int64, C.POINTER(C.c_void_p)]),
    "frb_host_free": (C.c_int32, [C.c_void_p]),
    "frb_halo_export": (C.c_int32, [C.c_void_p, C.POINTER(C.c_ubyte)]),
    "frb_halo_connect": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte)]),
    "frb_halo_sync": (C.c_int32, [C.c_void_p]),
    "frb_halo_disconnect": (C.c_int32, [C.c_void_p]),
}

_lib = None


def lib():
    """Load libfrb200.so (built by build.py / __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not found: run `python fluxreconstruction.jl_b200/build.py` (nvcc, sm_100a). "
                "There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().frb_last_error(None)
        raise FRBError(rc, msg.decode() if msg else "?")


def dptr(a: np.ndarray):
    assert a.dtype == np.float64
    return a.ctypes.data_as(c_dp)


def fortran_ptr(a: np.ndarray):
    """Pointer to a Julia-layout (column-major) float64 array; refuses silent copies."""
    if a.dtype != np.float64 or not a.flags.f_contiguous:
        raise ValueError("state arrays must be float64 and Fortran-ordered (Julia memory layout)")
    return a.ctypes.data_as(c_dp)


class Context:
    """One CUDA device + stream (frb_ctx_create)."""

    _default = None

    def __init__(self, device: int = -1):
        self.h = C.c_void_p()
        check(lib().frb_ctx_create(device, C.byref(self.h)))

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls(int(os.environ.get("LOCAL_RANK", "-1")) if "LOCAL_RANK" in os.environ else -1)
        return cls._default

    def info(self):
        sm, ma, mi = C.c_int32(), C.c_int32(), C.c_int32()
        buf = C.create_string_buffer(128)
        check(lib().frb_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), buf, 128))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), name=buf.value.decode())

    def close(self):
        if self.h:
            lib().frb_ctx_destroy(self.h)
            self.h = C.c_void_p()


def make_operators(deg, ll, lr, lpdm, dgl, dgr, dll=None, dlr=None):
    """Pack the FRPSpace operator arrays for the ABI.  Returns (struct, keepalive)."""
    keep = [np.ascontiguousarray(x, dtype=np.float64) for x in (ll, lr, dgl, dgr)]
    lp = np.asfortranarray(lpdm, dtype=np.float64)  # ps.dl, column-major [m,k]
    keep.append(lp)
    ops = Operators()
    ops.deg = int(deg)
    ops.ll, ops.lr, ops.dgl, ops.dgr = (dptr(k) for k in keep[:4])
    ops.lpdm = lp.ctypes.data_as(c_dp)
    if dll is not None and dlr is not None:
        a, b = np.ascontiguousarray(dll, dtype=np.float64), np.ascontiguousarray(dlr, dtype=np.float64)
        keep += [a, b]
        ops.dll, ops.dlr = dptr(a), dptr(b)
    return ops, keep


def pinned_empty(shape, order="F"):
    """float64 array in page-locked host memory (cudaMallocHost) for the host-buffer paths."""
    n = int(np.prod(shape))
    p = C.c_void_p()
    check(lib().frb_host_alloc(n * 8, C.byref(p)))
    buf = (C.c_double * n).from_address(p.value)
    a = np.frombuffer(buf, dtype=np.float64).reshape(shape, order=order)
    _PINNED[a.ctypes.data] = p
    return a


def pinned_free(a):
    p = _PINNED.pop(a.ctypes.data, None)
    if p is not None:
        check(lib().frb_host_free(p))


_PINNED = {}
