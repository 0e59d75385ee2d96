"""Host-side mirror of the reference's problem constructors and of the SciML driving
interface, bound to libfrb200 through its C ABI.

Reference API                                   here
------------------------------------------------------------------------------------------
FRAdvectionProblem(u, tspan, ps, a, bc)         FRAdvectionProblem(u, tspan, ps, a, bc)
  src/Equation/eq_advection.jl:1-26
FREulerProblem(u, tspan, ps, γ, bc)             FREulerProblem(u, tspan, ps, gamma, bc)
  src/Equation/eq_euler.jl:1-27
ODEProblem(dudt!, u0, tspan, p) of              Euler2DProblem(u0, tspan, ps, gamma)
  example/euler2d_wave.jl:35-122
ODEProblem(mol!, f0, tspan, p) of               BGKProblem(f0, tspan, ps, velo, weights, tau)
  example/bgk_wave.jl:69-132
prob.f(du, u, p, t)                             prob.f(du, u, p, t)      (host arrays in/out)
init(prob, Midpoint(); adaptive=false, dt)      init(prob, Midpoint(), dt=dt)
step!(itg); itg.u                               step_(itg); itg.u
solve(prob, alg; adaptive=false, dt)            solve(prob, alg, dt=dt)
positive_limiter(u, γ, weights, ll, lr)         itg.set_hooks(limiter_weights=...) / prob.limiter(...)

Boundary symbols are the reference's: ``:dirichlet`` and ``:period`` (strings, with or
without the leading colon).  State arrays are float64 in the reference's index order and
Fortran (Julia) memory layout, ghost cells included where the reference has them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Context, check, fortran_ptr, lib, make_operators

BC = {"dirichlet": 0, "period": 1}
GHOST = {None: -1, "none": -1, "wave_x": 0, "wave_y": 1, "copy": 2, "periodic": 3, "cylinder": 4}
KERNEL = {"auto": 0, "generic": 1, "march": 2, "rc": 3, "one_pass": 4, "curv_march": 5}
FLUX = {"hll": 0, "lf": 1, "roe": 2}


class Euler:
    code = 0
    stages = 1


class Midpoint:
    code = 1
    stages = 2


class SSPRK33:
    """Shu-Osher three-stage SSP scheme (OrdinaryDiffEq's SSPRK33); not used by the reference's
    scripts but named by the north star."""

    code = 2
    stages = 3


class ExplicitRK:
    """Any explicit Runge-Kutta scheme by its Butcher tableau (``frb_step_tableau``): ``A`` strictly
    lower triangular [s, s], ``b`` [s]."""

    code = -1

    def __init__(self, A, b):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.b = np.ascontiguousarray(b, dtype=np.float64)
        self.stages = self.b.size
        if self.A.shape != (self.stages, self.stages) or np.triu(self.A).any():
            raise ValueError("A must be a strictly lower triangular [s, s] matrix")

    @property
    def tableau(self):
        return self.A, self.b


class RK4(ExplicitRK):
    """The classical fourth-order scheme (OrdinaryDiffEq's RK4 with adaptive=false)."""

    def __init__(self):
        A = np.zeros((4, 4))
        A[1, 0], A[2, 1], A[3, 2] = 0.5, 0.5, 1.0
        super().__init__(A, [1 / 6, 1 / 3, 1 / 3, 1 / 6])


class Tsit5(ExplicitRK):
    """Tsitouras' 5(4) pair as the reference uses it: fixed step, ``adaptive=false, dt=dt``
    (example/advection_highlevel.jl:26, example/euler1d_convergence.jl:133).  Coefficients:
    Ch. Tsitouras, Comput. Math. Appl. 62 (2011) 770-775; the seventh (FSAL) stage only feeds the error
    estimate and the next step's k1, so six stages are evaluated.  tests/ verify the 17 order
    conditions up to order 5 on this table."""

    def __init__(self):
        A = np.zeros((6, 6))
        A[1, :1] = [0.161]
        A[2, :2] = [-0.008480655492356989, 0.335480655492357]
        A[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
        A[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
        A[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
                    -0.028269050394068383]
        b = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
             2.324710524099774]
        super().__init__(A, b)


def modal_filter_diag(n, lam):
    """[KB] KitBase.modal_filter!(u, lam; filter=:l2) as a diagonal: mode k >= 1 is divided by
    1 + lam k^2 (k+1)^2."""
    k = np.arange(n, dtype=np.float64)
    d = 1.0 / (1.0 + lam * (k + 1.0) ** 2 * k**2)
    d[0] = 1.0
    return d


def _sym(s):
    return str(s).lstrip(":")


class _Problem:
    """Common part: owns the library handle; f!(du,u,p,t); device-resident stepping."""

    def __init__(self, u0, tspan, ctx=None):
        self._ctx = ctx  # the device context is opened on first use, after the arguments have been checked
        self.u0 = np.array(u0, dtype=np.float64, order="F", copy=True)
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.h = C.c_void_p()
        self._keep = []
        self.p = None  # the reference passes a parameter tuple; kept for signature parity

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = Context.default()
        return self._ctx

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if self.h:
            lib().frb_prob_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- sizes ------------------------------------------------------------------------
    @property
    def state_len(self):
        return lib().frb_state_len(self.h)

    @property
    def dofs(self):
        return lib().frb_interior_dofs(self.h)

    # -- f!(du, u, p, t) ----------------------------------------------------------------
    def f(self, du, u, p=None, t=0.0):
        """The SciML in-place RHS with host arrays: uploads u, evaluates on the GPU,
        downloads du.  Returns None like the reference."""
        if u.shape != self.u0.shape or du.shape != self.u0.shape:
            raise ValueError("f!: array shape does not match the problem")
        check(lib().frb_rhs(self.h, fortran_ptr(u), fortran_ptr(du), float(t)))
        return None

    def f_pipelined(self, du, u, p=None, t=0.0, nslab=16):
        """f!(du,u,p,t) with host arrays, streamed through the device in ``nslab`` row slabs so that
        upload, residual and download overlap (2-D Euler; other problems fall back to ``f``)."""
        if u.shape != self.u0.shape or du.shape != self.u0.shape:
            raise ValueError("f!: array shape does not match the problem")
        check(lib().frb_rhs_pipelined(self.h, fortran_ptr(u), fortran_ptr(du), int(nslab)))
        return None

    def step_host(self, u_in, u_out, alg, dt, nslab=32):
        """step!(itg) with the state on the host between steps: ``u_out`` = one step of ``alg`` from ``u_in`` (host
        arrays, may be the same).  2-D Euler problems stream the state through the device in ``nslab`` row slabs
        (upload, stages and download overlapped); the ghost cells are the caller's and stay frozen.  Everything else
        is upload + step + download."""
        if u_in.shape != self.u0.shape or u_out.shape != self.u0.shape:
            raise ValueError("step_host: array shape does not match the problem")
        if hasattr(alg, "tableau"):
            self.upload(u_in)
            self.step(alg, dt, 1)
            self.download(u_out)
            return None
        check(lib().frb_step_host(self.h, fortran_ptr(u_in), fortran_ptr(u_out), alg.code, float(dt), int(nslab)))
        return None

    def rhs_resident(self, du=None):
        """L(u) of the resident state; left on the device (timing / chaining) unless ``du`` (a host
        array of the state's shape) is given."""
        if du is not None and du.shape != self.u0.shape:
            raise ValueError("f!: array shape does not match the problem")
        check(lib().frb_rhs(self.h, None, None if du is None else fortran_ptr(du), 0.0))
        return du

    # -- state --------------------------------------------------------------------------
    def upload(self, u):
        if u.shape != self.u0.shape:
            raise ValueError("upload: array shape does not match the problem")
        check(lib().frb_state_upload(self.h, fortran_ptr(np.asfortranarray(u, dtype=np.float64))))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.u0.shape, dtype=np.float64, order="F")
        check(lib().frb_state_download(self.h, fortran_ptr(out)))
        return out

    def device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().frb_state_device_ptr(self.h, C.byref(p)))
        return p.value

    # -- hooks / stepping -----------------------------------------------------------------
    def set_hooks(self, ghost=None, limiter_weights=None):
        w = None
        if limiter_weights is not None:
            w = np.asfortranarray(limiter_weights, dtype=np.float64)
        check(lib().frb_set_step_hooks(self.h, GHOST[ghost], None if w is None else _lib.dptr(w.ravel(order="F"))))

    def step(self, alg, dt, nsteps=1):
        if hasattr(alg, "tableau"):
            A, b = alg.tableau
            check(lib().frb_step_tableau(self.h, b.size, _lib.dptr(A.ravel()), _lib.dptr(b), float(dt), int(nsteps)))
        else:
            check(lib().frb_step(self.h, alg.code, float(dt), int(nsteps)))

    def ghost_fill(self, mode):
        check(lib().frb_ghost_fill(self.h, GHOST[mode]))

    def limiter(self, weights):
        """positive_limiter on every interior cell of the resident state; raises on the
        reference's @assert condition."""
        w = np.asfortranarray(weights, dtype=np.float64).ravel(order="F")
        nbad = C.c_int32()
        check(lib().frb_limiter_positivity(self.h, _lib.dptr(w), C.byref(nbad)))
        return nbad.value

    def modal_filter(self, ps, lam, eps=None, S0=None, kappa=None, ghosts=None, when=None):
        """Shock sensor + modal l2 filter on every element of the resident state (Euler problems):
        example/euler_highlevel.jl:37-52 (1-D defaults eps = 1e-6, kappa = 4) and
        example/shock-vortex.jl:308-321 (2-D defaults eps = 0, kappa = 9, ghost cells included).
        ``when`` = "before" / "after" registers it as a hook of every step instead (None: run now
        and return the number of filtered elements; "off" removes the hook)."""
        two_d = self.u0.ndim == 5
        eps = (0.0 if two_d else 1e-6) if eps is None else eps
        kappa = (9.0 if two_d else 4.0) if kappa is None else kappa
        S0 = -3.0 * np.log10(ps.deg) if S0 is None else S0
        ghosts = two_d if ghosts is None else ghosts
        iV = np.asfortranarray(ps.iV, dtype=np.float64)
        F = np.asfortranarray(ps.V @ np.diag(modal_filter_diag(iV.shape[0], lam)) @ ps.iV, dtype=np.float64)
        args = (_lib.dptr(iV.ravel(order="F")), _lib.dptr(F.ravel(order="F")), iV.shape[0], float(eps), float(S0),
                float(kappa), int(bool(ghosts)))
        if when is None:
            n = C.c_int32()
            check(lib().frb_filter_modal(self.h, *args, C.byref(n)))
            return n.value
        code = {"off": 0, "before": 1, "after": 2}[when]
        check(lib().frb_set_filter_hook(self.h, code, *args))
        return None

    def set_flux(self, flux="hll"):
        """common flux of the Euler problems: "hll" (the reference's flux_hll!), "lf", "roe" """
        check(lib().frb_set_flux(self.h, FLUX[_sym(flux)]))

    def set_kernel(self, kind="auto"):
        check(lib().frb_set_kernel(self.h, KERNEL[kind]))

    # -- measurement ------------------------------------------------------------------------
    def time_stage(self, stage_kind=1, iters=10) -> float:
        ms = C.c_float()
        check(lib().frb_time_stage(self.h, int(stage_kind), int(iters), C.byref(ms)))
        return ms.value

    def set_profiling(self, on=True):
        check(lib().frb_set_profiling(self.h, 1 if on else 0))

    def stage_timing(self):
        """(summed device ms of the stage launches of the last step/rhs call, their count)"""
        ms, n = C.c_float(), C.c_int64()
        check(lib().frb_stage_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_timing(self):
        ms, n = C.c_float(), C.c_int64()
        check(lib().frb_last_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def _ops_of(ps, slopes=False):
    return make_operators(ps.deg, ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr, ps.dll if slopes else None,
                          ps.dlr if slopes else None)


class FRAdvectionProblem(_Problem):
    """src/Equation/eq_advection.jl:1-26; RHS frode_advection! :55-77.  u[ncell, nsp].
    variant="lowlevel" reproduces example/advection_lowlevel.jl:4-47 (periodic, 1e-8 seam)."""

    def __init__(self, u, tspan, ps, a, bc, variant="packaged", ctx=None):
        super().__init__(u, tspan, ctx)
        ncell, nsp = self.u0.shape
        if nsp != ps.deg + 1:
            raise ValueError("u must be [ncell, deg+1]")
        ops, self._keep = _ops_of(ps)
        J = np.ascontiguousarray(ps.interior(ps.J) if ncell == ps.nx else ps.J, dtype=np.float64)
        self._keep.append(J)
        check(lib().frb_advection1d_create(self.ctx.h, ncell, C.byref(ops), _lib.dptr(J), float(a), BC[_sym(bc)],
                                           1 if variant == "lowlevel" else 0, C.byref(self.h)))
        self.upload(self.u0)


class FREulerProblem(_Problem):
    """src/Equation/eq_euler.jl:1-27; RHS frode_euler! :29-98.  u[ncell, nsp, 3]."""

    def __init__(self, u, tspan, ps, gamma, bc, ctx=None):
        super().__init__(u, tspan, ctx)
        ncell, nsp, nv = self.u0.shape
        if nsp != ps.deg + 1 or nv != 3:
            raise ValueError("u must be [ncell, deg+1, 3]")
        self.gamma = float(gamma)
        ops, self._keep = _ops_of(ps)
        J = np.ascontiguousarray(ps.interior(ps.J) if ncell == ps.nx else ps.J, dtype=np.float64)
        self._keep.append(J)
        check(lib().frb_euler1d_create(self.ctx.h, ncell, C.byref(ops), _lib.dptr(J), self.gamma, BC[_sym(bc)],
                                       C.byref(self.h)))
        self.upload(self.u0)


class Euler2DProblem(_Problem):
    """ODEProblem(dudt!, u0, tspan, p) of example/euler2d_wave.jl:35-122 on an FRPSpace2D with
    one ghost ring.  u0[nx+2, ny+2, nsp, nsp, 4]."""

    def __init__(self, u0, tspan, ps, gamma, ctx=None, kernel="auto"):
        super().__init__(u0, tspan, ctx)
        nsp = ps.deg + 1
        if self.u0.shape != (ps.nx + 2, ps.ny + 2, nsp, nsp, 4):
            raise ValueError("u0 must be [nx+2, ny+2, nsp, nsp, 4] (one ghost ring)")
        self.gamma = float(gamma)
        ops, self._keep = _ops_of(ps)
        check(lib().frb_euler2d_create(self.ctx.h, ps.nx, ps.ny, C.byref(ops), ps.Jx, ps.Jy, self.gamma,
                                       C.byref(self.h)))
        if kernel != "auto":
            self.set_kernel(kernel)
        self.upload(self.u0)


class Euler2DCurvProblem(_Problem):
    """ODEProblem(dudt!, u0, tspan, p) of the curvilinear scratch scripts: dev/parallelogram.jl:80-194
    (``corr="sp"``: correction factors from the solution-point ``ps.iJ``) and dev/cylinder2.jl:52-170
    (``corr="fp"``: factors from the flux-point ``ps.Ji``; ``wall_xlo=True``: mirror wall on x face 1).
    ``ps = FRPSpace2D(base, deg)`` with one ghost ring; ``n1[nx+1, ny, 2]``, ``n2[nx, ny+1, 2]`` are the unit
    face normals the scripts keep as globals (default: ``face_normals(ps.vertices)``).  ``fy_index="l"``
    reproduces the scripts' ``fy_interaction[i, j, l, m]`` (parallelogram.jl:147-148); the default ``"k"`` is
    the index the rectangular scripts use (euler2d_wave.jl:100-103).  u0[nx+2, ny+2, nsp, nsp, 4]."""

    def __init__(self, u0, tspan, ps, gamma, n1=None, n2=None, corr="sp", fy_index="k", wall_xlo=False, ctx=None,
                 metric="stored"):
        super().__init__(u0, tspan, ctx)
        from .spaces import correction_factors_fp, face_normals

        nsp = ps.deg + 1
        if ps.ngx != 1 or ps.ngy != 1:
            raise ValueError("the space must carry one ghost ring (embed ghost-less directions first)")
        if self.u0.shape != (ps.nx + 2, ps.ny + 2, nsp, nsp, 4):
            raise ValueError("u0 must be [nx+2, ny+2, nsp, nsp, 4] (one ghost ring)")
        if n1 is None or n2 is None:
            n1, n2 = face_normals(ps.vertices)
        n1, n2 = np.asfortranarray(n1, dtype=np.float64), np.asfortranarray(n2, dtype=np.float64)
        if n1.shape != (ps.nx + 1, ps.ny, 2) or n2.shape != (ps.nx, ps.ny + 1, 2):
            raise ValueError("n1 must be [nx+1, ny, 2] and n2 [nx, ny+1, 2]")
        self.gamma = float(gamma)
        iJ = np.asfortranarray(ps.iJ, dtype=np.float64)
        fpc = None
        if corr == "fp":
            fpc = correction_factors_fp(ps.Ji, n1, n2)
        elif corr != "sp":
            raise ValueError("corr must be 'sp' or 'fp'")
        flags = (1 if fy_index == "l" else 0) | (2 if wall_xlo else 0)
        ops, self._keep = _ops_of(ps)
        check(lib().frb_euler2d_curv_create(self.ctx.h, ps.nx, ps.ny, C.byref(ops), _lib.dptr(iJ), _lib.dptr(n1),
                                            _lib.dptr(n2), None if fpc is None else _lib.dptr(fpc), flags,
                                            self.gamma, C.byref(self.h)))
        self._ps = ps
        if metric != "stored":
            self.set_metric(metric)
        self.upload(self.u0)

    def set_metric(self, metric):
        """"stored": the kernels read ps.iJ (the reference's array); "vertices": they evaluate it from
        ps.vertices and ps.xpl on the fly (frb_euler2d_curv_set_vertices)."""
        if metric == "vertices":
            v = np.asfortranarray(self._ps.vertices, dtype=np.float64)
            r = np.ascontiguousarray(self._ps.xpl, dtype=np.float64)
            check(lib().frb_euler2d_curv_set_vertices(self.h, _lib.dptr(v), _lib.dptr(r)))
        elif metric == "stored":
            check(lib().frb_euler2d_curv_set_vertices(self.h, None, None))
        else:
            raise ValueError("metric must be 'stored' or 'vertices'")


class BGKProblem(_Problem):
    """ODEProblem(mol!, f0, tspan, p) of example/bgk_wave.jl:69-132.  f0[ncell, nu, nsp].
    ``model="advection"`` is the mol! of example/advection_kinetic.jl:73-132 instead: the same residual relaxing
    towards the Maxwellian of prim = [rho, a, 1] (the script's tau is 2e-3)."""

    def __init__(self, f0, tspan, ps, velo, weights, tau=1e-2, ctx=None, model="bgk", a=1.0, kernel="auto"):
        super().__init__(f0, tspan, ctx)
        ncell, nu, nsp = self.u0.shape
        if nsp != ps.deg + 1 or nu != len(velo):
            raise ValueError("f0 must be [ncell, nu, deg+1]")
        ops, self._keep = _ops_of(ps)
        dx = np.ascontiguousarray(ps.interior(ps.dx), dtype=np.float64)
        v = np.ascontiguousarray(velo, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        self._keep += [dx, v, w]
        check(lib().frb_bgk1d_create(self.ctx.h, ncell, nu, C.byref(ops), _lib.dptr(dx), _lib.dptr(v), _lib.dptr(w),
                                     float(tau), C.byref(self.h)))
        if model != "bgk":
            check(lib().frb_bgk1d_set_model(self.h, {"bgk": 0, "advection": 1}[model], float(a)))
        if kernel != "auto":
            self.set_kernel(kernel)  # "one_pass": the register-tile kernel instead of the two launches
        self.upload(self.u0)


class Integrator:
    """What ``init(prob, alg; adaptive=false, dt)`` returns in the reference's scripts.

    ``itg.u`` hands out a host copy of the state; because the reference's user code
    mutates ``itg.u`` in place between steps (ghost fill, limiter, filters), a handed-out
    array is uploaded again before the next step.  Code that wants the fast path leaves
    ``itg.u`` alone and registers the per-step work with ``set_hooks``."""

    def __init__(self, prob, alg, dt):
        self.prob, self.alg, self.dt = prob, alg, float(dt)
        self.t = prob.tspan[0]
        self.iter = 0
        self._host = None

    @property
    def u(self):
        if self._host is None:
            self._host = self.prob.download()
        return self._host

    def set_u(self, u):
        self.prob.upload(u)
        self._host = None

    def set_hooks(self, ghost=None, limiter_weights=None):
        self.prob.set_hooks(ghost, limiter_weights)

    def step(self, nsteps=1):
        if self._host is not None:
            self.prob.upload(self._host)
            self._host = None
        self.prob.step(self.alg, self.dt, nsteps)
        self.t += nsteps * self.dt
        self.iter += nsteps


def init(prob, alg, dt, adaptive=False, **_):
    if adaptive:
        raise NotImplementedError("only fixed-step integration is accelerated (SURVEY 0.1)")
    return Integrator(prob, alg if not isinstance(alg, type) else alg(), dt)


def step_(itg, nsteps=1):
    """step!(itg)"""
    itg.step(nsteps)


def solve(prob, alg, dt, adaptive=False, **_):
    """solve(prob, alg; adaptive=false, dt): full steps of dt and, like OrdinaryDiffEq, a shortened last step
    that lands exactly on tspan[2] when dt does not divide the span."""
    itg = init(prob, alg, dt, adaptive)
    span = prob.tspan[1] - prob.tspan[0]
    n = int(np.floor(span / dt * (1.0 + 4.0 * np.finfo(float).eps) + 1e-12))
    if n > 0:
        itg.step(n)
    rest = span - n * dt
    if rest > 1e-12 * max(abs(span), abs(dt)):
        itg.dt = rest
        itg.step(1)
        itg.dt = float(dt)
    itg.t = prob.tspan[1]
    return itg


HALO_BLOB_BYTES = 6 * 64 + 8


class DistributedEuler2D(Euler2DProblem):
    """Row-slab parallel 2-D Euler problem: one process per GPU, this rank holds rows
    ``slab.start .. slab.stop`` of the global mesh in a local array [nx+2, ny_local+2, nsp, nsp, 4].

    ``dist`` is an initialised ``torch.distributed`` module (or anything with ``get_rank``,
    ``get_world_size``, ``all_gather_object`` and ``barrier``): it is used once, to hand the
    neighbours' CUDA-IPC blobs around.  Per-stage halo traffic never touches the host: the stage
    kernel stores its boundary rows into the neighbour's memory over NVLink and raises a flag
    (csrc/frb_halo.cu).  The global y seam follows the reference's frozen per-step ghost fill
    (``ghost`` = "wave_x" / "wave_y", example/euler2d_wave.jl:127-132,159-164)."""

    def __init__(self, u0_local, tspan, ps_local, gamma, dist, ctx=None, ghost="wave_x", kernel="auto"):
        super().__init__(u0_local, tspan, ps_local, gamma, ctx=ctx, kernel=kernel)
        self.set_hooks(ghost=ghost)
        self._connect(dist)

    def _connect(self, dist):
        self.dist = dist
        self.rank, self.nranks = dist.get_rank(), dist.get_world_size()
        if self.nranks > 1:
            blob = (C.c_ubyte * HALO_BLOB_BYTES)()
            check(lib().frb_halo_export(self.h, blob))
            blobs = [None] * self.nranks
            dist.all_gather_object(blobs, bytes(blob))  # also orders every rank's upload before any push
            lo = (C.c_ubyte * HALO_BLOB_BYTES).from_buffer_copy(blobs[(self.rank - 1) % self.nranks])
            hi = (C.c_ubyte * HALO_BLOB_BYTES).from_buffer_copy(blobs[(self.rank + 1) % self.nranks])
            check(lib().frb_halo_connect(self.h, self.rank, self.nranks, lo, hi))
            dist.barrier()

    def resync(self):
        """after a new upload on every rank"""
        if self.nranks > 1:
            self.dist.barrier()
            check(lib().frb_halo_sync(self.h))

    def close(self):
        if self.h and self.nranks > 1:
            try:
                self.dist.barrier()  # nobody unmaps while a neighbour may still be storing
            except Exception:
                pass
            lib().frb_halo_disconnect(self.h)
        super().close()


def ref_vhs_vis(Kn, alpha, omega):
    """KitBase.ref_vhs_vis(Kn, alpha, omega) (used by example/ns_cavity.jl:17,30)."""
    import math

    return 5.0 * (alpha + 1.0) * (alpha + 2.0) * math.sqrt(math.pi) / (
        4.0 * alpha * (5.0 - 2.0 * omega) * (7.0 - 2.0 * omega)) * Kn


class _SlabMixin:
    """The halo plumbing shared by the slab-parallel problems (see DistributedEuler2D)."""

    _connect = DistributedEuler2D._connect
    resync = DistributedEuler2D.resync

    def close(self):
        if self.h and self.nranks > 1:
            try:
                self.dist.barrier()
            except Exception:
                pass
            lib().frb_halo_disconnect(self.h)
        super().close()


class NSCavityProblem(_Problem):
    """ODEProblem(dudt!, u0, tspan, p) of example/ns_cavity.jl:147-380: 2-D Navier-Stokes with the
    gas-kinetic flux, isothermal walls and a moving lid.  u0[4, nsp, nsp, ny+2, nx+2] (variable
    fastest).  ``gas`` mirrors KitBase.Gas: K, gamma, mu_ref (μᵣ), omega (ω); ``dt`` is the step that
    enters the time-averaged interface flux; ``lid`` is pb[2] of :337; ``lambda_wall`` the λ0 of
    boundary!(u, p, 1.0)."""

    def __init__(self, u0, tspan, ps, K, gamma, mu_ref, omega, dt, lid=0.15, lambda_wall=1.0, ctx=None):
        super().__init__(u0, tspan, ctx)
        nsp = ps.deg + 1
        if self.u0.shape != (4, nsp, nsp, ps.ny + 2, ps.nx + 2):
            raise ValueError("u0 must be [4, nsp, nsp, ny+2, nx+2]")
        ops, self._keep = _ops_of(ps, slopes=True)
        check(lib().frb_ns2d_create(self.ctx.h, ps.nx, ps.ny, C.byref(ops), ps.Jx, ps.Jy, float(K), float(gamma),
                                    float(mu_ref), float(omega), float(dt), float(lid), float(lambda_wall),
                                    C.byref(self.h)))
        self.upload(self.u0)


class TriEulerProblem(_Problem):
    """ODEProblem(dudt!, u0, tspan, p) of dev/sod.jl:31-130: 2-D Euler on a triangle mesh with
    p = (ps.cellType, ps.J, ps.lf, ps.cellNormals, ps.fpn, ps.dl, ps.phi, gamma) of a TriFRPSpace
    (struct.jl:305-352).  u0[ncell, np, 4].  ``fpn`` holds (cell, face, point) per flux point,
    shape [ncell, 3, deg+1, 3]; ``fpn_base`` = 1 for the reference's 1-based tuples (<= 0: no
    neighbour), 0 for 0-based arrays with -1."""

    def __init__(self, u0, tspan, cell_type, J, lf, cell_normal, fpn, dl, phi, gamma, fpn_base=1, ctx=None):
        super().__init__(u0, tspan, ctx)
        ncell, npts, nv = self.u0.shape
        deg = lf.shape[1] - 1
        if nv != 4 or npts != (deg + 1) * (deg + 2) // 2:
            raise ValueError("u0 must be [ncell, (deg+1)(deg+2)/2, 4]")
        f32 = lambda a: np.asfortranarray(a, dtype=np.int32)  # noqa: E731
        f64 = lambda a: np.asfortranarray(a, dtype=np.float64)  # noqa: E731
        fp = np.asarray(fpn, dtype=np.int64) + (1 - int(fpn_base))  # -> 1-based
        fp = f32(np.moveaxis(fp, -1, 0))  # [3, ncell, 3, deg+1]: the memory image of Julia's array of tuples
        keep = [f32(cell_type), f64(J), f64(cell_normal), fp, f64(lf), f64(dl), f64(phi)]
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
        check(lib().frb_tri_euler_create(self.ctx.h, ncell, deg, ip(keep[0]), fortran_ptr(keep[1]),
                                         fortran_ptr(keep[2]), ip(keep[3]), fortran_ptr(keep[4]),
                                         fortran_ptr(keep[5]), fortran_ptr(keep[6]), float(gamma), C.byref(self.h)))
        self._keep = keep
        self.upload(self.u0)

    @classmethod
    def from_space(cls, ps, u0, tspan, gamma, cell_type=None, ctx=None):
        """ODEProblem(dudt!, u0, tspan, (ps.cellType, ps.J, ps.lf, ps.cellNormals, ps.fpn, ps.∂l, ps.ϕ, γ))
        for a ``TriFRPSpace`` of this package (0-based ``fpn``); ``cell_type`` overrides ``ps.cellType``
        (dev/sod.jl:7-16 retags wall cells as 2 before building the problem)."""
        ct = ps.cellType if cell_type is None else cell_type
        return cls(u0, tspan, ct, ps.J, ps.lf, ps.cellNormals, ps.fpn, ps.dl, ps.phi, gamma, fpn_base=0, ctx=ctx)


class DistributedEuler2DCurv(_SlabMixin, Euler2DCurvProblem):
    """Row-slab parallel curvilinear 2-D Euler problem (SURVEY 8f-2 + 8e): one process per GPU, this rank holds rows
    ``slab.start .. slab.stop`` of the global mesh in a local array [nx+2, ny_local+2, nsp, nsp, 4].  ``ps_local`` is
    the ``FRPSpace2D(base, deg)`` of the rank's rows plus one ghost row on either side (the metric is per element, so
    a space built from ``vertices[:, start-1 : stop+2]`` carries the global mesh's ``iJ`` of those rows bit for bit);
    ``n1 = n1_global[:, start-1 : stop]``, ``n2 = n2_global[:, start-1 : stop+1]``.  Interior slab boundaries are
    exchanged after every stage over NVLink, the global y seam follows the per-step periodic ghost fill of
    dev/parallelogram.jl:201-205 (``ghost="periodic"``)."""

    def __init__(self, u0_local, tspan, ps_local, gamma, dist, n1=None, n2=None, corr="sp", fy_index="k", ctx=None,
                 ghost="periodic", kernel="auto"):
        super().__init__(u0_local, tspan, ps_local, gamma, n1, n2, corr=corr, fy_index=fy_index, ctx=ctx)
        if kernel != "auto":
            self.set_kernel(kernel)
        self.set_hooks(ghost=ghost)
        self._connect(dist)


class DistributedNSCavity(_SlabMixin, NSCavityProblem):
    """Column-slab parallel cavity (cfg5): one process per GPU, this rank holds columns ``slab.start .. slab.stop``
    of the global mesh in a local array [4, nsp, nsp, ny+2, nx_local+2] (``i`` is the slowest index of the
    reference's layout, so a slab and each of its halo columns are contiguous).  The walls of the slab's interior
    boundaries are replaced by the neighbour's boundary column of the current stage, stored into this rank's memory
    over NVLink after every stage (csrc/frb_halo.cu); the x walls stay with the first / last rank, the y walls and
    the lid with everybody."""

    def __init__(self, u0_local, tspan, ps_local, K, gamma, mu_ref, omega, dt, dist, lid=0.15, lambda_wall=1.0,
                 ctx=None):
        super().__init__(u0_local, tspan, ps_local, K, gamma, mu_ref, omega, dt, lid=lid, lambda_wall=lambda_wall,
                         ctx=ctx)
        self._connect(dist)
