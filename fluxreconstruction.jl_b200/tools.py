"""Host-side mirror of the small public helpers of the reference that sit next to the hot path and that
its own test script calls (test/runtests.jl:6-46): error norms (src/tools.jl:43-53), ``shock_detector``
and the single-cell ``positive_limiter`` methods (src/dissipation.jl:13-206), the exponential filters
(src/Polynomial/poly_filter.jl), ``interp_face!`` / ``poly_derivative!`` (src/interpolate.jl,
src/derivative.jl), ``rs_jacobi`` for straight-sided quadrilaterals (src/Geometry/geo_jacobi.jl:77-108)
and the triangle coordinate maps (src/Geometry/geo_transform.jl:65-128).

NumPy, setup / post-processing time only: the per-step versions of the limiter and of the sensor +
filter pass are device kernels behind ``frb_limiter_positivity`` / ``frb_filter_modal``; these functions
are what a user script calls on one cell or when analysing a result, with the reference's argument
order.  Julia's ``f!`` spelling becomes ``f_`` (in place on the first argument)."""
from __future__ import annotations

import numpy as np

__all__ = [
    "L1_error", "L2_error", "Linf_error", "shock_detector", "positive_limiter", "filter_exp", "filter_exp1d",
    "filter_exp2d", "basis_norm", "interp_face_", "poly_derivative_", "rs_jacobi", "rs_ab", "xy_rs", "conserve_prim",
    "prim_conserve", "VSpace1D", "maxwellian", "moments_conserve", "heaviside", "euler_flux", "sound_speed",
]


# ------------------------------------------------------------------ accuracy analysis (tools.jl:43-53)
def L1_error(u, ue, dx):
    return float(np.sum(np.abs(np.asarray(u) - np.asarray(ue)) * dx))


def L2_error(u, ue, dx):
    return float(np.sqrt(np.sum((np.abs(np.asarray(u) - np.asarray(ue)) * dx) ** 2)))


def Linf_error(u, ue, dx):
    return float(np.max(np.abs(np.asarray(u) - np.asarray(ue)) * dx))


# ------------------------------------------------------------------ [KB] variable conversion (lambda form)
def conserve_prim(w, gamma):
    """[KB] conserve_prim: (rho, rho U.., rho E) -> (rho, U.., lambda = rho / 2p); last axis = variables."""
    w = np.asarray(w, dtype=np.float64)
    rho, E = w[..., 0], w[..., -1]
    vel = w[..., 1:-1] / rho[..., None]
    lam = 0.5 * rho / (gamma - 1.0) / (E - 0.5 * rho * np.sum(vel * vel, axis=-1))
    return np.concatenate([rho[..., None], vel, lam[..., None]], axis=-1)


def prim_conserve(prim, gamma):
    """[KB] prim_conserve, the inverse of ``conserve_prim``."""
    p = np.asarray(prim, dtype=np.float64)
    rho, lam = p[..., 0], p[..., -1]
    vel = p[..., 1:-1]
    E = 0.5 * rho / lam / (gamma - 1.0) + 0.5 * rho * np.sum(vel * vel, axis=-1)
    return np.concatenate([rho[..., None], rho[..., None] * vel, E[..., None]], axis=-1)


def sound_speed(prim, gamma):
    """[KB] sound_speed(prim, gamma) = sqrt(gamma / (2 lambda))."""
    return np.sqrt(0.5 * gamma / np.asarray(prim, dtype=np.float64)[..., -1])


def euler_flux(w, gamma):
    """[KB] euler_flux(w, gamma): (F,) for 1-D states (3 variables), (F, G) for 2-D ones (4 variables)."""
    w = np.asarray(w, dtype=np.float64)
    rho, E = w[..., 0], w[..., -1]
    vel = w[..., 1:-1] / rho[..., None]
    p = (gamma - 1.0) * (E - 0.5 * rho * np.sum(vel * vel, axis=-1))
    out = []
    for d in range(vel.shape[-1]):
        vd = vel[..., d]
        mom = [w[..., 1 + c] * vd + (p if c == d else 0.0) for c in range(vel.shape[-1])]
        out.append(np.stack([w[..., 1 + d], *mom, (E + p) * vd], axis=-1))
    return tuple(out)


# ------------------------------------------------------------------ [KB] 1-D velocity space and moments
class VSpace1D:
    """[KB] VSpace1D(u0, u1, nu) (example/bgk_wave.jl:25): midpoint nodes ``u`` and uniform ``weights``."""

    def __init__(self, u0, u1, nu):
        self.u0, self.u1, self.nu = float(u0), float(u1), int(nu)
        du = (self.u1 - self.u0) / self.nu
        self.u = self.u0 + (np.arange(1, self.nu + 1) - 0.5) * du
        self.du = np.full(self.nu, du)
        self.weights = np.full(self.nu, du)


def maxwellian(v, prim):
    """[KB] maxwellian(u, prim), 1-D1V: rho sqrt(lambda / pi) exp(-lambda (u - U)^2); prim[..., 3] broadcasts
    against v[..., nu]."""
    prim = np.asarray(prim, dtype=np.float64)
    rho, U, lam = prim[..., 0:1], prim[..., 1:2], prim[..., 2:3]
    return rho * np.sqrt(lam / np.pi) * np.exp(-lam * (np.asarray(v, dtype=np.float64) - U) ** 2)


def moments_conserve(f, v, w):
    """[KB] moments_conserve(f, u, weights) = [sum w f, sum w u f, 1/2 sum w u^2 f] over the last axis."""
    f, v, w = (np.asarray(a, dtype=np.float64) for a in (f, v, w))
    return np.stack([np.sum(w * f, -1), np.sum(v * w * f, -1), 0.5 * np.sum(v * v * w * f, -1)], axis=-1)


def heaviside(x):
    """[KB] heaviside(x): 1 for x >= 0 (bgk_wave.jl:26 builds the upwind switch with it)."""
    return (np.asarray(x) >= 0).astype(np.float64)


# ------------------------------------------------------------------ dissipation.jl
def shock_detector(Se, deg, S0=None, kappa=4.0):
    """Persson-Peraire switch (dissipation.jl:13-23): True when the element is to be filtered."""
    if S0 is None:
        S0 = -3.0 * np.log10(deg)
    if Se < S0 - kappa:
        sigma = 1.0
    elif Se < S0 + kappa:
        sigma = 0.5 * (1.0 - np.sin(0.5 * np.pi * (Se - S0) / kappa))
    else:
        sigma = 0.0
    return bool(sigma < 0.99)


def _limit_density(rho, mean, floor_candidates, edge_values):
    eps = min([1e-13] + floor_candidates)
    lo = min(float(np.min(edge_values)), float(np.min(rho)))
    t = min((mean - eps) / (mean - lo + 1e-8), 1.0)
    if not (0.0 < t <= 1.0):
        raise AssertionError("incorrect range of limiter parameter t")
    return t, eps


def positive_limiter(u, *args):
    """In place, one cell (dissipation.jl:28-206), dispatching on ``u.ndim`` like the reference's methods:

    * ``positive_limiter(u[nsp], weights, ll, lr, t0=1)``             scalar (:28-59, mean = sum(u w)/sum(w))
    * ``positive_limiter(u[nsp, 3], gamma, weights, ll, lr, t0=1)``   1-D Euler (:61-123, mean = sum(u w))
    * ``positive_limiter(u[nsp, nsp, 4], gamma, weights, ll, lr, t0=1)``  2-D Euler (:125-206)

    The density corrector is the reference's.  Its energy corrector (:100-121, :173-203) ends in
    ``minimum(tj, t0)``, which is a MethodError in Julia, so it has no defined result; like the device
    kernels (``frb_limiter_positivity``) this function applies the density corrector only."""
    u = np.asarray(u)
    if u.dtype != np.float64 or not u.flags.writeable:
        raise TypeError("positive_limiter works in place on a float64 array")
    if u.ndim == 1:
        w, ll, lr = (np.asarray(a, dtype=np.float64) for a in args[:3])
        mean = float(np.sum(u * w) / np.sum(w))
        t, _ = _limit_density(u, mean, [mean], [u @ ll, u @ lr])
        u[:] = t * (u - mean) + mean
        return None
    gamma = float(args[0])
    w, ll, lr = (np.asarray(a, dtype=np.float64) for a in args[1:4])
    if u.ndim == 2:
        mean = np.array([np.sum(u[:, k] * w) for k in range(u.shape[1])])
        edges = np.stack([ll @ u, lr @ u])  # [2, 3]
        points = u.reshape(-1, u.shape[1])
    elif u.ndim == 3:
        # the reference sums over axes(u, 2) = the first nsp variables only when nsp < 4; all 4 are needed
        mean = np.array([np.sum(u[:, :, k] * w) for k in range(u.shape[2])])
        edges = np.concatenate([np.einsum("jlk,l->jk", u, ll), np.einsum("ljk,l->jk", u, lr),
                                np.einsum("jlk,l->jk", u, lr), np.einsum("ljk,l->jk", u, ll)])  # [4 nsp, 4]
        points = u.reshape(-1, u.shape[2])
    else:
        raise ValueError("u must be [nsp], [nsp, 3] or [nsp, nsp, 4]")
    p_mean = 0.5 * mean[0] / conserve_prim(mean, gamma)[-1]
    t1, _ = _limit_density(u[..., 0], float(mean[0]), [float(mean[0]), float(p_mean)], edges[:, 0])
    u[..., 0] = t1 * (u[..., 0] - mean[0]) + mean[0]
    return None


# ------------------------------------------------------------------ poly_filter.jl
def filter_exp1d(N, s, Nc=0):
    """[KB] filter_exp1d: exp(-alpha ((i - Nc)/(N - Nc))^s) for i >= Nc, alpha = -log(eps)."""
    alpha = -np.log(np.finfo(np.float64).eps)
    d = np.ones(N + 1)
    i = np.arange(Nc, N + 1)
    d[Nc:] = np.exp(-alpha * ((i - Nc) / (N - Nc)) ** s)
    return d


def filter_exp2d(N, s, Nc=0):
    """poly_filter.jl:32-49: the same decay in the total degree i + j of the triangle modes."""
    alpha = -np.log(np.finfo(np.float64).eps)
    tot = np.array([i + j for i in range(N + 1) for j in range(N + 1 - i)], dtype=np.float64)
    d = np.ones(tot.size)
    m = tot >= Nc
    d[m] = np.exp(-alpha * ((tot[m] - Nc) / (N - Nc)) ** s)
    return d


def filter_exp(N, s, V, Nc=0, invV=None):
    """poly_filter.jl:11-22: nodal filter matrix V diag(filter) V^-1 for a 1-D or a triangle basis."""
    V = np.asarray(V, dtype=np.float64)
    nv = V.shape[0]
    if nv == N + 1:
        d = filter_exp1d(N, s, Nc)
    elif nv == (N + 1) * (N + 2) // 2:
        d = filter_exp2d(N, s, Nc)
    else:
        raise ValueError("V is neither a 1-D nor a triangle Vandermonde matrix of degree N")
    return (V * d) @ (np.linalg.inv(V) if invV is None else np.asarray(invV))


def basis_norm(deg):
    """poly_filter.jl:54-67: L1 norms of the orthonormal Legendre modes by a 100-point rectangle rule."""
    x = np.linspace(-1.0, 1.0, 100)
    dx = x[1] - x[0]
    out = np.zeros(deg + 1)
    for n in range(deg + 1):
        c = np.zeros(n + 1)
        c[n] = np.sqrt((2 * n + 1) / 2.0)
        out[n] = dx * np.sum(np.abs(np.polynomial.legendre.legval(x, c)))
    return out


# ------------------------------------------------------------------ interpolate.jl, derivative.jl
def interp_face_(fd, f, ll, lr):
    """interp_face!(fδ, f, ll, lr): f[nsp] -> fδ[2], or f[n, nsp] -> fδ[n, 2]."""
    f = np.asarray(f)
    fd[..., 0] = f @ np.asarray(ll)
    fd[..., 1] = f @ np.asarray(lr)
    return None


def poly_derivative_(df, f, pdm):
    """poly_derivative!(df, f, pdm): df[i] = dot(f, pdm[i, :])."""
    pdm = np.asarray(pdm)
    f = np.asarray(f)
    if f.shape[-1] != pdm.shape[0]:
        raise AssertionError("length(f) == size(pdm, 1)")
    df[...] = f @ pdm.T
    return None


# ------------------------------------------------------------------ geometry
def rs_jacobi(r, s=None, vertices=None):
    """Jacobian [xr xs; yr ys] of the bilinear map of a quadrilateral with vertices[4, 2]
    (counter-clockwise from the lower left) at (r, s) (geo_jacobi.jl:77-108).  Scalars give a 2x2
    matrix, vectors r[n], s[n] give [n, 2, 2]; ``rs_jacobi(r, vertices)`` is the reference's shorthand for
    the tensor grid r_i, r_j -> [n, n, 2, 2]; vertices[nx, ny, 4, 2] adds two leading axes."""
    if vertices is None:
        vertices, s = s, None
    v = np.asarray(vertices, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    if s is None:
        rr, ss = np.meshgrid(r, r, indexing="ij")
    else:
        rr, ss = r, np.asarray(s, dtype=np.float64)
    if v.ndim == 4:
        return np.stack([[rs_jacobi(rr, ss, v[i, j]) for j in range(v.shape[1])] for i in range(v.shape[0])])
    a, b = rr[..., None], ss[..., None]
    d_r = ((b - 1.0) * v[0] + (1.0 - b) * v[1] + (b + 1.0) * v[2] - (b + 1.0) * v[3]) / 4.0
    d_s = ((a - 1.0) * v[0] - (a + 1.0) * v[1] + (a + 1.0) * v[2] + (1.0 - a) * v[3]) / 4.0
    return np.stack([d_r, d_s], axis=-1)  # [..., (x, y), (r, s)]


def xy_rs(x, y=None):
    """Equilateral reference triangle -> right triangle (geo_transform.jl:65-95)."""
    if y is None:
        x, y = np.asarray(x)[:, 0], np.asarray(x)[:, 1]
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    q = np.sqrt(3.0) * y
    L1, L2, L3 = (q + 1.0) / 3.0, (-3.0 * x - q + 2.0) / 6.0, (3.0 * x - q + 2.0) / 6.0
    return L3 - L2 - L1, L1 - L2 - L3


def rs_ab(r, s=None):
    """Right triangle -> collapsed square coordinates (geo_transform.jl:102-128)."""
    if s is None:
        r, s = np.asarray(r)[:, 0], np.asarray(r)[:, 1]
    r, s = np.asarray(r, dtype=np.float64), np.asarray(s, dtype=np.float64)
    top = s == 1.0
    a = np.where(top, -1.0, 2.0 * (1.0 + r) / np.where(top, 1.0, 1.0 - s) - 1.0)
    return a, 1.0 * s
