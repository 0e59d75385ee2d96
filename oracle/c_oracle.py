"""ctypes loader for the C oracle (oracle/fr_oracle.c)  --  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  PARITY UNPINNED (see fr_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libfr_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("fr_oracle.c", "fr_oracle_gks.c", "fr_oracle_curv.c", "fr_arbiter.c", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libfr_oracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.fro_work2d_create.restype = C.c_void_p
        _LIB.fro_work2d_create.argtypes = [C.c_int, C.c_int, C.c_int]
        _LIB.fro_work2d_destroy.argtypes = [C.c_void_p]
    return _LIB


def _p(a):
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.flags.c_contiguous)
    return a.ctypes.data_as(_dp)


def _ops(ps):
    """row-major lpdm[m][k] plus the 1-D vectors."""
    return (
        np.ascontiguousarray(ps.ll),
        np.ascontiguousarray(ps.lr),
        np.ascontiguousarray(ps.dl),
        np.ascontiguousarray(ps.dhl),
        np.ascontiguousarray(ps.dhr),
    )


def num_threads() -> int:
    return lib().fro_num_threads()


def set_num_threads(n: int):
    lib().fro_set_num_threads(int(n))


def operators(deg: int, correction: str = "radau"):
    n = deg + 1
    r, w, ll, lr, dgl, dgr = (np.zeros(n) for _ in range(6))
    lpdm = np.zeros((n, n))
    rc = lib().fro_operators(
        deg, {"radau": 0, "sd": 1, "huynh": 2}[correction], _p(r), _p(w), _p(ll), _p(lr), _p(lpdm), _p(dgl), _p(dgr)
    )
    assert rc == 0
    return dict(r=r, w=w, ll=ll, lr=lr, lpdm=lpdm, dgl=dgl, dgr=dgr)


def rhs_advection1d(u, ps, a, bc="period", variant="packaged"):
    u = np.asfortranarray(u, dtype=np.float64)
    du = np.zeros_like(u, order="F")
    ncell, nsp = u.shape
    J = np.ascontiguousarray(ps.J[ps.ng : ps.ng + ncell])
    ll, lr, dl, dhl, dhr = _ops(ps)
    rc = lib().fro_rhs_adv1d(
        _p(u), _p(du), ncell, nsp, _p(J), _p(ll), _p(lr), _p(dl), _p(dhl), _p(dhr), C.c_double(a),
        1 if bc == "period" else 0, 1 if variant == "lowlevel" else 0,
    )
    assert rc == 0
    return du


def rhs_euler1d(u, ps, gamma, bc="dirichlet"):
    u = np.asfortranarray(u, dtype=np.float64)
    du = np.zeros_like(u, order="F")
    ncell, nsp, _ = u.shape
    J = np.ascontiguousarray(ps.J[ps.ng : ps.ng + ncell])
    ll, lr, dl, dhl, dhr = _ops(ps)
    rc = lib().fro_rhs_euler1d(
        _p(u), _p(du), ncell, nsp, _p(J), _p(ll), _p(lr), _p(dl), _p(dhl), _p(dhr), C.c_double(gamma),
        1 if bc == "period" else 0,
    )
    assert rc == 0
    return du


class Work2D:
    def __init__(self, nx, ny, nsp):
        self.h = lib().fro_work2d_create(nx, ny, nsp)
        if not self.h:
            raise MemoryError("fro_work2d_create")

    def __del__(self):
        if getattr(self, "h", None):
            lib().fro_work2d_destroy(self.h)
            self.h = None


def rhs_euler2d(u, ps, gamma, work=None, du=None):
    u = np.asfortranarray(u, dtype=np.float64)
    if du is None:
        du = np.empty_like(u, order="F")
    nx, ny, nsp = u.shape[0] - 2, u.shape[1] - 2, u.shape[2]
    work = work or Work2D(nx, ny, nsp)
    ll, lr, dl, dhl, dhr = _ops(ps)
    rc = lib().fro_rhs_euler2d(
        _p(u), _p(du), C.c_void_p(work.h), C.c_double(ps.Jx), C.c_double(ps.Jy), _p(ll), _p(lr), _p(dl), _p(dhl),
        _p(dhr), C.c_double(gamma),
    )
    assert rc == 0
    return du


_GHOST = {"wave_x": 0, "wave_y": 1, "copy": 2, None: -1, "none": -1}
_SCHEME = {"euler": 0, "midpoint": 1, "ssprk3": 2}


def ghost_fill_euler2d(u, mode="wave_x"):
    assert u.flags.f_contiguous
    lib().fro_ghost_fill_euler2d(_p(u), u.shape[0] - 2, u.shape[1] - 2, u.shape[2], _GHOST[mode])
    return u


def integrate_euler2d(u, ps, gamma, dt, nsteps, scheme="midpoint", ghost="wave_x", limiter_weights=None,
                      inplace=False):
    """``inplace``: advance the caller's array (Fortran order, float64) instead of a copy -- the timed loops of
    bench.py step a 2 GB state and must not copy it every step."""
    if inplace:
        assert u.dtype == np.float64 and u.flags.f_contiguous
    else:
        u = np.array(u, dtype=np.float64, order="F", copy=True)
    nx, ny, nsp = u.shape[0] - 2, u.shape[1] - 2, u.shape[2]
    ll, lr, dl, dhl, dhr = _ops(ps)
    wts = None if limiter_weights is None else np.asfortranarray(limiter_weights, dtype=np.float64)
    rc = lib().fro_integrate_euler2d(
        _p(u), nx, ny, nsp, C.c_double(ps.Jx), C.c_double(ps.Jy), _p(ll), _p(lr), _p(dl), _p(dhl), _p(dhr),
        C.c_double(gamma), C.c_double(dt), int(nsteps), _SCHEME[scheme], _GHOST[ghost],
        _p(wts) if wts is not None else None,
    )
    assert rc == 0
    return u


def integrate_euler1d(u, ps, gamma, bc, dt, nsteps, scheme="midpoint", limiter_weights=None):
    u = np.array(u, dtype=np.float64, order="F", copy=True)
    ncell, nsp, _ = u.shape
    J = np.ascontiguousarray(ps.J[ps.ng : ps.ng + ncell])
    ll, lr, dl, dhl, dhr = _ops(ps)
    wts = None if limiter_weights is None else np.ascontiguousarray(limiter_weights, dtype=np.float64)
    rc = lib().fro_integrate_euler1d(
        _p(u), ncell, nsp, _p(J), _p(ll), _p(lr), _p(dl), _p(dhl), _p(dhr), C.c_double(gamma),
        1 if bc == "period" else 0, C.c_double(dt), int(nsteps), _SCHEME[scheme],
        _p(wts) if wts is not None else None,
    )
    assert rc == 0
    return u


def integrate_advection1d(u, ps, a, bc, variant, dt, nsteps, scheme="midpoint"):
    u = np.array(u, dtype=np.float64, order="F", copy=True)
    ncell, nsp = u.shape
    J = np.ascontiguousarray(ps.J[ps.ng : ps.ng + ncell])
    ll, lr, dl, dhl, dhr = _ops(ps)
    rc = lib().fro_integrate_adv1d(
        _p(u), ncell, nsp, _p(J), _p(ll), _p(lr), _p(dl), _p(dhl), _p(dhr), C.c_double(a),
        1 if bc == "period" else 0, 1 if variant == "lowlevel" else 0, C.c_double(dt), int(nsteps), _SCHEME[scheme],
    )
    assert rc == 0
    return u


def rhs_bgk1d(u, dx, velo, weights, ll, lr, lpdm, dgl, dgr, tau=1e-2):
    u = np.asfortranarray(u, dtype=np.float64)
    du = np.empty_like(u, order="F")
    ncell, nu, nsp = u.shape
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (dx, velo, weights, ll, lr, lpdm, dgl, dgr)]
    rc = lib().fro_rhs_bgk1d(_p(u), _p(du), ncell, nu, nsp, *[_p(x) for x in a], C.c_double(tau))
    assert rc == 0
    return du


def integrate_bgk1d(u, dx, velo, weights, ll, lr, lpdm, dgl, dgr, tau, dt, nsteps, scheme="midpoint"):
    u = np.array(u, dtype=np.float64, order="F", copy=True)
    ncell, nu, nsp = u.shape
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (dx, velo, weights, ll, lr, lpdm, dgl, dgr)]
    rc = lib().fro_integrate_bgk1d(
        _p(u), ncell, nu, nsp, *[_p(x) for x in a], C.c_double(tau), C.c_double(dt), int(nsteps), _SCHEME[scheme]
    )
    assert rc == 0
    return u


def limiter_euler1d(u, gamma, weights, ll, lr):
    assert u.flags.f_contiguous
    w, l1, l2 = (np.ascontiguousarray(x, dtype=np.float64) for x in (weights, ll, lr))
    return lib().fro_limiter_euler1d(_p(u), u.shape[0], u.shape[1], C.c_double(gamma), _p(w), _p(l1), _p(l2))


def limiter_euler2d(u, gamma, weights, ll, lr):
    assert u.flags.f_contiguous
    w = np.asfortranarray(weights, dtype=np.float64)
    l1, l2 = (np.ascontiguousarray(x, dtype=np.float64) for x in (ll, lr))
    return lib().fro_limiter_euler2d(
        _p(u), u.shape[0] - 2, u.shape[1] - 2, u.shape[2], C.c_double(gamma), _p(w), _p(l1), _p(l2)
    )


def _ns_args(ps, K, gamma, mu, omega, dt, lam0, lid):
    ops = [np.ascontiguousarray(x, dtype=np.float64) for x in (ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr, ps.dll, ps.dlr)]
    scal = [C.c_double(x) for x in (K, gamma, mu, omega, dt, lam0, lid)]
    return ops, scal


def ns_boundary(u, gamma, lam0=1.0, lid=0.15):
    assert u.flags.f_contiguous
    lib().fro_ns_boundary(_p(u), u.shape[4] - 2, u.shape[3] - 2, u.shape[1], C.c_double(gamma), C.c_double(lam0),
                          C.c_double(lid))
    return u


def rhs_ns2d(u, ps, K, gamma, mu, omega, dt, lam0=1.0, lid=0.15):
    """u's ghosts are rewritten in place (boundary!), as in the reference."""
    assert u.flags.f_contiguous and u.dtype == np.float64
    du = np.empty_like(u, order="F")
    ops, scal = _ns_args(ps, K, gamma, mu, omega, dt, lam0, lid)
    rc = lib().fro_rhs_ns2d(_p(u), _p(du), u.shape[4] - 2, u.shape[3] - 2, u.shape[1], C.c_double(ps.Jx),
                            C.c_double(ps.Jy), *[_p(x) for x in ops], *scal)
    assert rc == 0
    return du


def integrate_ns2d(u, ps, K, gamma, mu, omega, dt, nsteps, lam0=1.0, lid=0.15):
    u = np.array(u, dtype=np.float64, order="F", copy=True)
    ops, scal = _ns_args(ps, K, gamma, mu, omega, dt, lam0, lid)
    rc = lib().fro_integrate_ns2d(_p(u), u.shape[4] - 2, u.shape[3] - 2, u.shape[1], C.c_double(ps.Jx),
                                  C.c_double(ps.Jy), *[_p(x) for x in ops], *scal, int(nsteps))
    assert rc == 0
    return u


def rhs_euler2d_curv(u, ps, n1, n2, gamma, corr="sp", fpc=None, fy_index="l", wall_xlo=False):
    """dudt! of dev/parallelogram.jl:80-165 / dev/cylinder2.jl:52-164 (oracle/fr_oracle_curv.c); same arguments
    as fr_oracle_curv.rhs_euler2d_curv (HLL)."""
    u = np.asfortranarray(u, dtype=np.float64)
    du = np.zeros_like(u, order="F")
    nx, ny, nsp = u.shape[0] - 2, u.shape[1] - 2, u.shape[2]
    iJ = np.asfortranarray(ps.iJ, dtype=np.float64)
    n1, n2 = np.asfortranarray(n1, dtype=np.float64), np.asfortranarray(n2, dtype=np.float64)
    assert iJ.shape == u.shape[:4] + (2, 2) and n1.shape == (nx + 1, ny, 2) and n2.shape == (nx, ny + 1, 2)
    fp = None
    if corr == "fp":
        fp = np.asfortranarray(fpc, dtype=np.float64)
        assert fp.shape == (nx, ny, nsp, 4)
    ll, lr, dl, dhl, dhr = _ops(ps)
    fn = lib().fro_rhs_euler2d_curv
    fn.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_double] + [_dp] * 5
    rc = fn(_p(u), _p(du), nx, ny, nsp, _p(iJ), _p(n1), _p(n2), None if fp is None else _p(fp),
            (1 if fy_index == "l" else 0) | (2 if wall_xlo else 0), gamma, _p(ll), _p(lr), _p(dl), _p(dhl), _p(dhr))
    assert rc == 0
    return du


# ---- extended-precision arbiter (oracle/fr_arbiter.c): x87 long double, cell-local --------------------
def arbiter_euler2d_cells(u, ps, gamma, cells):
    """du of the cells ``cells`` [(i, j), ...] (1-based interior indices of the ghosted array) evaluated in
    long double from the double inputs; returns [ncells, nsp, nsp, 4] (k, l, m)."""
    u = np.asfortranarray(u, dtype=np.float64)
    nx, ny, nsp = u.shape[0] - 2, u.shape[1] - 2, u.shape[2]
    cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 2)
    assert cells[:, 0].min() >= 1 and cells[:, 0].max() <= nx and cells[:, 1].min() >= 1 and cells[:, 1].max() <= ny
    out = np.empty((cells.shape[0], 4, nsp, nsp), dtype=np.float64)
    ll, lr, dl, dhl, dhr = _ops(ps)
    rc = lib().fra_rhs_euler2d_cells(
        _p(u), nx, ny, nsp, C.c_double(ps.Jx), C.c_double(ps.Jy), _p(ll), _p(lr), _p(dl), _p(dhl), _p(dhr),
        C.c_double(gamma), cells.ctypes.data_as(C.POINTER(C.c_int32)), int(cells.shape[0]), _p(out))
    assert rc == 0
    return out.transpose(0, 3, 2, 1)  # memory order k fastest -> [c, k, l, m]


def arbiter_bgk1d_cells(u, dx, velo, weights, ll, lr, lpdm, dgl, dgr, tau, cells):
    """du[cell, :, :] of the listed cells (0-based) in long double; returns [ncells, nu, nsp]."""
    u = np.asfortranarray(u, dtype=np.float64)
    ncell, nu, nsp = u.shape
    cells = np.ascontiguousarray(cells, dtype=np.int32).ravel()
    out = np.empty((cells.size, nsp, nu), dtype=np.float64)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (dx, velo, weights, ll, lr, lpdm, dgl, dgr)]
    rc = lib().fra_rhs_bgk1d_cells(_p(u), ncell, nu, nsp, *[_p(x) for x in a], C.c_double(tau),
                                   cells.ctypes.data_as(C.POINTER(C.c_int32)), int(cells.size), _p(out))
    assert rc == 0
    return out.transpose(0, 2, 1)
