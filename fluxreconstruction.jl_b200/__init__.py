"""frb200 -- Blackwell-native flux-reconstruction residual + explicit step behind the
FluxReconstruction.jl API shapes.  Compute lives in lib/libfrb200.so (hand-written
sm_100a CUDA, C ABI in include/frb200.h); this package is the thin host mirror of the
reference interface.  There is no CPU fallback."""
from ._lib import Context, FRBError, LIB_PATH, SIGNATURES, lib, pinned_empty, pinned_free  # noqa: F401
from .spaces import *  # noqa: F401,F403
from .unstruct import *  # noqa: F401,F403
from .tools import *  # noqa: F401,F403
from .problems import (  # noqa: F401
    BGKProblem, DistributedEuler2D, DistributedEuler2DCurv, DistributedNSCavity, Euler, Euler2DProblem, Euler2DCurvProblem, ExplicitRK, RK4, Tsit5, FRAdvectionProblem, FREulerProblem, Integrator, Midpoint, NSCavityProblem, SSPRK33, TriEulerProblem, init, ref_vhs_vis,
    solve, step_,
)
from . import examples, partition  # noqa: F401

__version__ = "0.1.0"
