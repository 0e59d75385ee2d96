# FRB200.jl -- the thin Julia shim a FluxReconstruction.jl maintainer would add to route the
# semi-discrete residual and the explicit step through libfrb200.so (include/frb200.h).
#
# Julia is not installed in the build image, so this file has not been executed there; it is
# written against the header, and every call below has a ctypes twin in
# fluxreconstruction.jl_b200/_lib.py that IS exercised by tests/.
#
#   using FluxReconstruction, OrdinaryDiffEq
#   include("FRB200.jl"); using .FRB200
#   ps   = FRPSpace2D(0.0, 1.0, 2048, 0.0, 1.0, 2048, 3, 1, 1)
#   prob = FRB200.Euler2DProblem(u0, ps, γ)               # uploads parent(u0)
#   ode  = ODEProblem(FRB200.rhs!(prob), u0, tspan, p)     # drop-in f!(du,u,p,t)
#   # or, device-resident stepping (what example/euler2d_wave.jl:125-135 does, fused):
#   FRB200.set_step_hooks!(prob; ghost = :wave_x)
#   FRB200.step!(prob, :midpoint, dt, nt); u1 = FRB200.download(prob)
module FRB200

const lib = get(ENV, "FRB200_LIB", joinpath(@__DIR__, "..", "fluxreconstruction.jl_b200", "lib", "libfrb200.so"))

struct Operators
    deg::Int32
    ll::Ptr{Float64}; lr::Ptr{Float64}; lpdm::Ptr{Float64}
    dgl::Ptr{Float64}; dgr::Ptr{Float64}; dll::Ptr{Float64}; dlr::Ptr{Float64}
end

last_error() = unsafe_string(ccall((:frb_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL))
check(rc) = rc == 0 ? nothing : error("libfrb200 error $rc: $(last_error())")

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = -1)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:frb_ctx_create, lib), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r))
        finalizer(c -> ccall((:frb_ctx_destroy, lib), Int32, (Ptr{Cvoid},), c.h), new(r[]))
    end
end
const default_ctx = Ref{Union{Nothing,Context}}(nothing)
ctx() = (default_ctx[] === nothing && (default_ctx[] = Context()); default_ctx[])

mutable struct Problem
    h::Ptr{Cvoid}
    dims::Dims
    keep::Vector{Any}      # operator arrays stay rooted while the handle lives
end
destroy!(p::Problem) = (p.h != C_NULL && ccall((:frb_prob_destroy, lib), Int32, (Ptr{Cvoid},), p.h); p.h = C_NULL)

# ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr (struct.jl:49-51,63-64) are passed as they are: FRPSpace
# remains the single source of truth.  ps.dl is column-major [m,k], exactly what the ABI expects.
function operators(ps)
    keep = Any[Vector{Float64}(ps.ll), Vector{Float64}(ps.lr), Matrix{Float64}(ps.dl),
               Vector{Float64}(ps.dhl), Vector{Float64}(ps.dhr), Vector{Float64}(ps.dll), Vector{Float64}(ps.dlr)]
    Operators(Int32(ps.deg), map(pointer, keep)...), keep
end

# FRAdvectionProblem(u, tspan, ps, a, bc)  -- src/Equation/eq_advection.jl:1-26
function AdvectionProblem(u::Matrix{Float64}, ps, a, bc::Symbol; variant = :packaged)
    ops, keep = operators(ps); J = Vector{Float64}(ps.J[1:size(u, 1)]); push!(keep, J)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:frb_advection1d_create, lib), Int32,
        (Ptr{Cvoid}, Int32, Ref{Operators}, Ptr{Float64}, Float64, Int32, Int32, Ref{Ptr{Cvoid}}),
        ctx().h, size(u, 1), ops, J, a, bc == :period ? 1 : 0, variant == :lowlevel ? 1 : 0, r))
    p = finalizer(destroy!, Problem(r[], size(u), keep)); upload!(p, u); p
end

# FREulerProblem(u, tspan, ps, γ, bc)  -- src/Equation/eq_euler.jl:1-27
function EulerProblem(u::Array{Float64,3}, ps, γ, bc::Symbol)
    ops, keep = operators(ps); J = Vector{Float64}(ps.J[1:size(u, 1)]); push!(keep, J)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:frb_euler1d_create, lib), Int32,
        (Ptr{Cvoid}, Int32, Ref{Operators}, Ptr{Float64}, Float64, Int32, Ref{Ptr{Cvoid}}),
        ctx().h, size(u, 1), ops, J, γ, bc == :period ? 1 : 0, r))
    p = finalizer(destroy!, Problem(r[], size(u), keep)); upload!(p, u); p
end

# dudt! of example/euler2d_wave.jl:35-107; u0 is the OffsetArray 0:nx+1 x 0:ny+1 x nsp x nsp x 4
function Euler2DProblem(u0, ps, γ)
    A = parent(u0)::Array{Float64,5}
    ops, keep = operators(ps)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    Jx, Jy = ps.J[1, 1][1, 1][1, 1], ps.J[1, 1][1, 1][2, 2]     # diag of the 2x2 Jacobian (struct.jl:135)
    GC.@preserve keep check(ccall((:frb_euler2d_create, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ref{Operators}, Float64, Float64, Float64, Ref{Ptr{Cvoid}}),
        ctx().h, size(A, 1) - 2, size(A, 2) - 2, ops, Jx, Jy, γ, r))
    p = finalizer(destroy!, Problem(r[], size(A), keep)); upload!(p, A); p
end

# dudt! of dev/parallelogram.jl:80-165 (fp = false: correction factors from ps.iJ) and dev/cylinder2.jl:52-164
# (fp = true: factors from the flux-point ps.Ji; wall = true: mirror wall on x face 1).  u0 and ps.iJ / ps.Ji are
# OffsetArrays over 0:nx+1 x 0:ny+1 (a space without radial ghosts is embedded first, DESIGN.md 4.4);
# n1[i, j], n2[i, j] are the scripts' global tables of unit normals.  literal_fy = true keeps the scripts'
# fy_interaction[i, j, l, m] (parallelogram.jl:147-148).
function Euler2DCurvProblem(u0, ps, γ, n1, n2; fp = false, wall = false, literal_fy = false)
    A = parent(u0)::Array{Float64,5}
    nx, ny, nsp = size(A, 1) - 2, size(A, 2) - 2, size(A, 3)
    ops, keep = operators(ps)
    PJ, PJi = parent(ps.iJ), parent(ps.Ji)
    iJ = [PJ[i, j][k, l][a, b] for i in 1:nx+2, j in 1:ny+2, k in 1:nsp, l in 1:nsp, a in 1:2, b in 1:2]
    N1 = [n1[i, j][c] for i in 1:nx+1, j in 1:ny, c in 1:2]
    N2 = [n2[i, j][c] for i in 1:nx, j in 1:ny+1, c in 1:2]
    fpc = fp ? [q == 1 ? (inv(PJi[i+1, j+1][4, p]) * n1[i, j])[1] : q == 2 ? (inv(PJi[i+1, j+1][2, p]) * n1[i+1, j])[1] :
                q == 3 ? (inv(PJi[i+1, j+1][1, p]) * n2[i, j])[2] : (inv(PJi[i+1, j+1][3, p]) * n2[i, j+1])[2]
                for i in 1:nx, j in 1:ny, p in 1:nsp, q in 1:4] : nothing
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep iJ N1 N2 fpc check(ccall((:frb_euler2d_curv_create, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ref{Operators}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32,
         Float64, Ref{Ptr{Cvoid}}),
        ctx().h, nx, ny, ops, iJ, N1, N2, fp ? pointer(fpc) : C_NULL, (literal_fy ? 1 : 0) | (wall ? 2 : 0), γ, r))
    p = finalizer(destroy!, Problem(r[], size(A), keep)); upload!(p, A); p
end

# dev/sod.jl:124-127: p = (ps.cellType, ps.J, ps.lf, ps.cellNormals, ps.fpn, ps.∂l, ps.ϕ, γ) of a TriFRPSpace
function TriEulerProblem(u0::Array{Float64,3}, ps, γ)
    ncell = size(u0, 1)
    J = [ps.J[i][a, b] for i in 1:ncell, a in 1:2, b in 1:2]                 # [ncell,2,2]
    fpn = Int32[ps.fpn[i, j, k][c] for c in 1:3, i in 1:ncell, j in 1:3, k in 1:ps.deg+1]
    ct, nrm = Int32.(ps.cellType), Array{Float64}(ps.cellNormals)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve ct J nrm fpn check(ccall((:frb_tri_euler_create, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Float64, Ref{Ptr{Cvoid}}),
        ctx().h, ncell, ps.deg, ct, J, nrm, fpn, ps.lf, ps.∂l, ps.ϕ, γ, r))
    p = finalizer(destroy!, Problem(r[], size(u0), Any[ct, J, nrm, fpn])); upload!(p, u0); p
end
# example/advection_kinetic.jl:73-128: the same mol! relaxing towards the Maxwellian of prim = [ρ, a, 1.0]
kinetic_advection!(p::Problem, a = 1.0) = check(ccall((:frb_bgk1d_set_model, lib), Int32,
    (Ptr{Cvoid}, Int32, Float64), p.h, 1, a))
# mol! of example/bgk_wave.jl:69-129
function BGKProblem(f0::Array{Float64,3}, ps, velo, weights, τ = 1e-2)
    ops, keep = operators(ps)
    dx = Vector{Float64}(ps.dx[1:size(f0, 1)]); v = Vector{Float64}(velo); w = Vector{Float64}(weights)
    append!(keep, (dx, v, w)); r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:frb_bgk1d_create, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ref{Operators}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Ptr{Cvoid}}),
        ctx().h, size(f0, 1), size(f0, 2), ops, dx, v, w, τ, r))
    p = finalizer(destroy!, Problem(r[], size(f0), keep)); upload!(p, f0); p
end

# dudt! + boundary! of example/ns_cavity.jl:147-344 (gas-kinetic flux :49-145).  u0 is the OffsetArray
# 4 x nsp x nsp x 0:ny+1 x 0:nx+1 of the script (:33); gas = ks.gas (K, γ, μᵣ, ω); dt enters the time-averaged
# interface flux; lid = pb[2] of :337; λ0 the wall 1/T of boundary!(u, p, 1.0)
function NSCavityProblem(u0, ps, gas, dt; lid = 0.15, λ0 = 1.0)
    A = parent(u0)::Array{Float64,5}
    ops, keep = operators(ps)
    Jx, Jy = ps.J[1, 1][1, 1][1, 1], ps.J[1, 1][1, 1][2, 2]
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:frb_ns2d_create, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ref{Operators}, Float64, Float64, Float64, Float64, Float64, Float64, Float64,
         Float64, Float64, Ref{Ptr{Cvoid}}),
        ctx().h, size(A, 5) - 2, size(A, 4) - 2, ops, Jx, Jy, gas.K, gas.γ, gas.μᵣ, gas.ω, dt, lid, λ0, r))
    p = finalizer(destroy!, Problem(r[], size(A), keep)); upload!(p, A); p
end

# the curvilinear metric evaluated on the fly from ps.base.vertices and ps.xpl instead of the stored ps.iJ
function set_vertices!(p::Problem, ps)
    V = Array{Float64}(parent(ps.base.vertices)); r = Vector{Float64}(ps.xpl)
    GC.@preserve V r check(ccall((:frb_euler2d_curv_set_vertices, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), p.h, V, r))
end

state_len(p::Problem) = ccall((:frb_state_len, lib), Int64, (Ptr{Cvoid},), p.h)
interior_dofs(p::Problem) = ccall((:frb_interior_dofs, lib), Int64, (Ptr{Cvoid},), p.h)
function device_ptr(p::Problem)      # raw device address of the resident state (CUDA.jl interop: unsafe_wrap(CuArray, ...))
    r = Ref{Ptr{Float64}}(C_NULL)
    check(ccall((:frb_state_device_ptr, lib), Int32, (Ptr{Cvoid}, Ref{Ptr{Float64}}), p.h, r)); r[]
end
function device_info(c::Context = ctx())
    sm, ma, mi = Ref{Int32}(0), Ref{Int32}(0), Ref{Int32}(0); name = zeros(UInt8, 128)
    check(ccall((:frb_device_info, lib), Int32, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ptr{UInt8}, Int32),
                c.h, sm, ma, mi, name, 128))
    (sm_count = sm[], cc = (ma[], mi[]), name = unsafe_string(pointer(name)))
end

upload!(p::Problem, u) = GC.@preserve u check(ccall((:frb_state_upload, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), p.h, parent(u)))
function download(p::Problem)
    u = Array{Float64}(undef, p.dims)
    GC.@preserve u check(ccall((:frb_state_download, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), p.h, u)); u
end

# the SciML in-place RHS: ODEProblem(rhs!(prob), u0, tspan, p)
rhs!(prob::Problem) = function (du, u, p, t)
    GC.@preserve du u check(ccall((:frb_rhs, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64),
                                  prob.h, parent(u), parent(du), t))
    nothing
end

# the same with pinned host buffers streamed through the device in row slabs (2-D Euler): upload, residual and
# download overlap
rhs_pipelined!(prob::Problem; nslab = 32) = function (du, u, p, t)
    GC.@preserve du u check(ccall((:frb_rhs_pipelined, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32),
                                  prob.h, parent(u), parent(du), nslab))
    nothing
end
# step!(itg) with itg.u on the host between steps (euler2d_wave.jl:125-135): u_out = one step from u_in, streamed
# through the device in row slabs (upload, stages and download overlap); ghost cells are the caller's
function step_host!(prob::Problem, u_out, u_in, scheme::Symbol, dt; nslab = 32)
    GC.@preserve u_out u_in check(ccall((:frb_step_host, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Int32),
        prob.h, parent(u_in), parent(u_out), Int32(SCHEME[scheme]), dt, nslab))
    u_out
end
# page-locked host arrays for the host-buffer paths
function host_array(dims::Dims)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:frb_host_alloc, lib), Int32, (Int64, Ref{Ptr{Cvoid}}), 8 * prod(dims), r))
    unsafe_wrap(Array, Ptr{Float64}(r[]), dims)
end
host_free(a::Array{Float64}) = check(ccall((:frb_host_free, lib), Int32, (Ptr{Cvoid},), pointer(a)))

const SCHEME = Dict(:euler => 0, :midpoint => 1, :ssprk3 => 2)
const GHOST = Dict(:none => -1, :wave_x => 0, :wave_y => 1, :copy => 2, :periodic => 3, :cylinder => 4)
set_step_hooks!(p::Problem; ghost = :none, limiter_weights = nothing) = check(ccall((:frb_set_step_hooks, lib), Int32,
    (Ptr{Cvoid}, Int32, Ptr{Float64}), p.h, GHOST[ghost], limiter_weights === nothing ? C_NULL : pointer(limiter_weights)))
ghost_fill!(p::Problem, mode::Symbol) = check(ccall((:frb_ghost_fill, lib), Int32, (Ptr{Cvoid}, Int32), p.h, GHOST[mode]))
# shock sensor + modal filter on every element (euler_highlevel.jl:37-52, shock-vortex.jl:308-321):
# F = ps.V * Diagonal(filterdiag) * ps.iV is built here from whatever KitBase filter the caller uses
function modal_filter!(p::Problem, iV::Matrix{Float64}, F::Matrix{Float64}; eps = 1e-6, S0, kappa = 4.0, ghosts = false)
    n = Ref{Int32}(0)
    check(ccall((:frb_filter_modal, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Float64,
        Float64, Int32, Ref{Int32}), p.h, iV, F, size(iV, 1), eps, S0, kappa, ghosts ? 1 : 0, n)); n[]
end
# the same pass as a hook of step!: when = :before (euler_highlevel.jl:37-52), :after (shock-vortex.jl:308-321), :off
function set_filter_hook!(p::Problem, when::Symbol, iV::Matrix{Float64}, F::Matrix{Float64}; eps = 1e-6, S0, kappa = 4.0,
                          ghosts = false)
    check(ccall((:frb_set_filter_hook, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Float64,
        Float64, Float64, Int32), p.h, Dict(:off => 0, :before => 1, :after => 2)[when], iV, F, size(iV, 1), eps, S0,
        kappa, ghosts ? 1 : 0))
end
# common flux of the Euler problems: :hll (the reference's flux_hll!), :lf, :roe
set_flux!(p::Problem, flux::Symbol) = check(ccall((:frb_set_flux, lib), Int32, (Ptr{Cvoid}, Int32), p.h,
    Int32(Dict(:hll => 0, :lf => 1, :roe => 2)[flux])))
step!(p::Problem, scheme::Symbol, dt, nsteps = 1) = check(ccall((:frb_step, lib), Int32,
    (Ptr{Cvoid}, Int32, Float64, Int32), p.h, SCHEME[scheme], dt, nsteps))
# any explicit RK scheme by its tableau, e.g. the fixed-step Tsit5() of advection_highlevel.jl:26: pass
# OrdinaryDiffEq's constructTsitouras5() / constructRK4() tableau (A strictly lower triangular, weights b);
# A is handed over row-major, hence the transpose
step!(p::Problem, A::Matrix{Float64}, b::Vector{Float64}, dt, nsteps = 1) = check(ccall((:frb_step_tableau, lib),
    Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Int32), p.h, length(b), permutedims(A), b, dt, nsteps))
function positive_limiter!(p::Problem, weights)      # src/dissipation.jl:61-206 on every interior cell
    nbad = Ref{Int32}(0)
    check(ccall((:frb_limiter_positivity, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int32}), p.h, weights, nbad)); nbad[]
end


# ---- measurement ----------------------------------------------------------------------------------------------
set_kernel!(p::Problem, kind::Symbol) = check(ccall((:frb_set_kernel, lib), Int32, (Ptr{Cvoid}, Int32), p.h,
    Int32(Dict(:auto => 0, :generic => 1, :march => 2, :rc => 3, :one_pass => 4, :curv_march => 5)[kind])))
function time_stage(p::Problem, stage_kind::Integer = 1, iters::Integer = 10)   # ms per fused stage launch
    ms = Ref{Float32}(0)
    check(ccall((:frb_time_stage, lib), Int32, (Ptr{Cvoid}, Int32, Int32, Ref{Float32}), p.h, stage_kind, iters, ms)); ms[]
end
function last_timing(p::Problem)                                                # (device ms, kernels) of the last call
    ms, n = Ref{Float32}(0), Ref{Int64}(0)
    check(ccall((:frb_last_timing, lib), Int32, (Ptr{Cvoid}, Ref{Float32}, Ref{Int64}), p.h, ms, n)); (ms[], n[])
end
set_profiling!(p::Problem, on::Bool = true) = check(ccall((:frb_set_profiling, lib), Int32, (Ptr{Cvoid}, Int32), p.h, on ? 1 : 0))
function stage_timing(p::Problem)
    ms, n = Ref{Float32}(0), Ref{Int64}(0)
    check(ccall((:frb_stage_timing, lib), Int32, (Ptr{Cvoid}, Ref{Float32}, Ref{Int64}), p.h, ms, n)); (ms[], n[])
end

# ---- multi-GPU: one Julia process per GPU (e.g. under MPI.jl), slabs along the slowest cell index ---------------
# euler2d problems: rows; ns2d problems: columns.  Exchange the 328-byte blobs by any means (MPI.Allgather), then
# connect to the ranks below (rank-1 mod n) and above (rank+1 mod n); per-stage halo traffic never touches the host.
const HALO_BLOB_BYTES = 6 * 64 + 8
function halo_export(p::Problem)
    blob = zeros(UInt8, HALO_BLOB_BYTES)
    check(ccall((:frb_halo_export, lib), Int32, (Ptr{Cvoid}, Ptr{UInt8}), p.h, blob)); blob
end
halo_connect!(p::Problem, rank::Integer, nranks::Integer, blob_lo::Vector{UInt8}, blob_hi::Vector{UInt8}) =
    check(ccall((:frb_halo_connect, lib), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}, Ptr{UInt8}),
                p.h, rank, nranks, blob_lo, blob_hi))
halo_sync!(p::Problem) = check(ccall((:frb_halo_sync, lib), Int32, (Ptr{Cvoid},), p.h))        # after a new upload!
halo_disconnect!(p::Problem) = check(ccall((:frb_halo_disconnect, lib), Int32, (Ptr{Cvoid},), p.h))

end # module
