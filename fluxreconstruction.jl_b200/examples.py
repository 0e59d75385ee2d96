"""Initial conditions of the reference's example scripts, as the scripts set them up (host side, NumPy):
what a user of ``example/*.jl`` writes before ``ODEProblem(...)``.  Used by the measurement scripts and handy for
trying the package; arrays come back in the reference's memory layout (Fortran order, ghosts included)."""
from __future__ import annotations

import numpy as np

from .tools import maxwellian, prim_conserve

__all__ = ["ic_advection1d", "ic_sod1d", "ic_wave1d", "ic_wave2d", "ic_bgk1d", "ic_kinetic_advection1d", "ic_cavity"]


def _xp(ps):
    return ps.xpg[ps.ng: ps.ng + ps.nx]


def ic_advection1d(ps):
    """example/advection_lowlevel.jl:55-58: u = sin(pi x) at the solution points, [nx, nsp]."""
    return np.asfortranarray(np.sin(np.pi * _xp(ps)))


def ic_sod1d(ps, gamma=5.0 / 3.0):
    """example/euler_lowlevel.jl:18-28: prim = [1, 0, 0.5] left of x = 0.5, [0.3, 0, 0.625] right of it."""
    left = (ps.x[ps.ng: ps.ng + ps.nx] <= 0.5)[:, None]
    w = prim_conserve(np.where(left, np.array([1.0, 0.0, 0.5]), np.array([0.3, 0.0, 0.625])), gamma)
    return np.asfortranarray(np.broadcast_to(w[:, None, :], (ps.nx, ps.deg + 1, 3)))


def ic_wave1d(ps, gamma=5.0 / 3.0, amp=0.1):
    """The smooth alternative commented at example/euler_lowlevel.jl:26: a density wave at unit velocity."""
    x = _xp(ps)
    one = np.ones_like(x)
    return np.asfortranarray(prim_conserve(np.stack([1.0 + amp * np.sin(2.0 * np.pi * x), one, one], axis=-1), gamma))


def ic_wave2d(ps, gamma=5.0 / 3.0, direction="x"):
    """example/euler2d_wave.jl:115-120 (x wave) / :146-151 (y wave): rho = 1 + 0.1 sin(2 pi x), prim = [rho, 1, 0, rho]."""
    along_x = direction == "x"
    rho = 1.0 + 0.1 * np.sin(2.0 * np.pi * ps.xpg[..., 0 if along_x else 1])
    prim = np.stack([rho, np.full_like(rho, 1.0 if along_x else 0.0), np.full_like(rho, 0.0 if along_x else 1.0), rho], -1)
    return np.asfortranarray(prim_conserve(prim, gamma))


def ic_bgk1d(ps, velo):
    """example/bgk_wave.jl:33-40: the Maxwellian of rho = 1 + 0.1 sin(2 pi x), U = 1, T = 1 / rho; [nx, nu, nsp]."""
    rho = 1.0 + 0.1 * np.sin(2.0 * np.pi * _xp(ps))
    prim = np.stack([rho, np.ones_like(rho), rho], axis=-1)  # lambda = 1 / T = rho
    f0 = maxwellian(np.asarray(velo)[None, None, :], prim)  # [nx, nsp, nu]
    return np.asfortranarray(np.transpose(f0, (0, 2, 1)))


def ic_kinetic_advection1d(ps, velo, a=1.0):
    """example/advection_kinetic.jl:37-45: u = 1 - sin(pi x), f = maxwellian(v, [u, a, 1])."""
    rho = 1.0 - np.sin(np.pi * _xp(ps))
    prim = np.stack([rho, np.full_like(rho, a), np.ones_like(rho)], axis=-1)
    f0 = maxwellian(np.asarray(velo)[None, None, :], prim)
    return np.asfortranarray(np.transpose(f0, (0, 2, 1)))


def ic_cavity(ps, gamma=5.0 / 3.0):
    """example/ns_cavity.jl:33-36: the gas at rest, prim = [1, 0, 0, 1]; [4, nsp, nsp, ny+2, nx+2]."""
    nsp = ps.deg + 1
    u = np.empty((4, nsp, nsp, ps.ny + 2, ps.nx + 2), order="F")
    u[...] = prim_conserve(np.array([1.0, 0.0, 0.0, 1.0]), gamma)[:, None, None, None, None]
    return u
