# B200Newton.jl — thin Julia glue over libb200newton.so (include/b200newton.h).  `@ccall` only: no CUDA.jl, no
# KernelAbstractions, no CPU fallback.  It plugs the library into NonlinearSolve.jl's own extension points:
#
#   b1  NonlinearFunction{true}(f!; jvp = jvp!)                 -> `brusselator_function(prob)` returns f!/jvp!/vjp! closures
#   b4  linsolve = B200GMRES()  (a LinearSolve.SciMLLinearSolveAlgorithm, needs_concrete_A = false)
#   b5  B200Vector <: AbstractVector{Float64}                    (device array the generic layers can carry around)
#   b6  EnsembleB200 <: SciMLBase.EnsembleAlgorithm
#   +   B200NewtonKrylov <: AbstractNonlinearSolveAlgorithm      whole-solve fast path (pattern: ext/NonlinearSolvePETScExt.jl:38-167)
#
# NOTE: this file cannot be executed in the build environment (no Julia runtime, SURVEY.md §0); it is the binding a
# maintainer adds on the reference side and is kept mechanical — one `@ccall` per exported symbol, same call sequence as
# the Python harness in nonlinearsolve.jl_b200/api.py which IS exercised on the GPU by tests/.
module B200Newton

using LinearAlgebra
import SciMLBase
import SciMLBase: ReturnCode, NLStats
import CommonSolve
import LinearSolve
import NonlinearSolveBase

const libb200 = get(ENV, "B200NEWTON_LIB", joinpath(@__DIR__, "..", "..", "..", "nonlinearsolve.jl_b200", "libb200newton.so"))

# ------------------------------------------------------------------ status handling
struct B200Error <: Exception
    code::Int32
    msg::String
end
const Ctx = Ptr{Cvoid}

function check(ctx::Ctx, status::Int32)
    status == 0 && return nothing
    msg = ctx == C_NULL ? "" : unsafe_string(@ccall libb200.b200_last_error(ctx::Ctx)::Cstring)
    throw(B200Error(status, msg))
end

# SciMLBase.ReturnCode <- B200_RC_* (include/b200newton.h)
const RETCODES = (ReturnCode.Default, ReturnCode.Success, ReturnCode.MaxIters, ReturnCode.MaxTime, ReturnCode.Stalled,
    ReturnCode.StalledSuccess, ReturnCode.Unstable, ReturnCode.InternalLinearSolveFailed,
    ReturnCode.InternalLineSearchFailed, ReturnCode.ShrinkThresholdExceeded, ReturnCode.InitialFailure, ReturnCode.Failure)
retcode(c::Integer) = RETCODES[c + 1]

# ------------------------------------------------------------------ context (one per task per device)
mutable struct Context
    handle::Ctx
    function Context(device::Integer = 0)
        h = Ref{Ctx}(C_NULL)
        st = @ccall libb200.b200_ctx_create(device::Int32, C_NULL::Ptr{Cvoid}, h::Ref{Ctx})::Int32
        st == -6 && throw(B200Error(st, "no CUDA device: the B200 backend has no CPU fallback"))
        check(C_NULL, st)
        ctx = new(h[])
        finalizer(c -> (@ccall libb200.b200_ctx_destroy(c.handle::Ctx)::Int32), ctx)
        return ctx
    end
end
const DEFAULT_CTX = Ref{Union{Nothing, Context}}(nothing)
default_context() = something(DEFAULT_CTX[], (DEFAULT_CTX[] = Context(0)))
sync(ctx::Context) = check(ctx.handle, @ccall libb200.b200_ctx_sync(ctx.handle::Ctx)::Int32)

# ------------------------------------------------------------------ b5: device vector
mutable struct B200Vector <: AbstractVector{Float64}
    ctx::Context
    ptr::Ptr{Float64}
    n::Int
    function B200Vector(ctx::Context, n::Integer)
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx.handle, @ccall libb200.b200_malloc(ctx.handle::Ctx, (8n)::Csize_t, p::Ref{Ptr{Cvoid}})::Int32)
        v = new(ctx, Ptr{Float64}(p[]), n)
        finalizer(x -> (@ccall libb200.b200_free(x.ctx.handle::Ctx, x.ptr::Ptr{Cvoid})::Int32), v)
        return v
    end
end
Base.size(v::B200Vector) = (v.n,)
Base.similar(v::B200Vector) = B200Vector(v.ctx, v.n)
Base.getindex(::B200Vector, ::Int) = error("scalar indexing of a B200Vector is disabled (fast_scalar_indexing = false)")
function B200Vector(ctx::Context, x::AbstractArray{Float64})
    v = B200Vector(ctx, length(x))
    copyto!(v, x)
end
function Base.copyto!(dst::B200Vector, src::Array{Float64})
    GC.@preserve src check(dst.ctx.handle, @ccall libb200.b200_memcpy_h2d(dst.ctx.handle::Ctx, dst.ptr::Ptr{Cvoid}, pointer(src)::Ptr{Cvoid}, (8 * dst.n)::Csize_t)::Int32)
    return dst
end
function Base.copyto!(dst::Array{Float64}, src::B200Vector)
    GC.@preserve dst check(src.ctx.handle, @ccall libb200.b200_memcpy_d2h(src.ctx.handle::Ctx, pointer(dst)::Ptr{Cvoid}, src.ptr::Ptr{Cvoid}, (8 * src.n)::Csize_t)::Int32)
    return dst
end
function Base.copyto!(dst::B200Vector, src::B200Vector)
    check(dst.ctx.handle, @ccall libb200.b200_copy(dst.ctx.handle::Ctx, dst.n::Int64, src.ptr::Ptr{Float64}, dst.ptr::Ptr{Float64})::Int32)
    return dst
end
Base.Array(v::B200Vector) = copyto!(Vector{Float64}(undef, v.n), v)
Base.copy(v::B200Vector) = copyto!(similar(v), v)
Base.fill!(v::B200Vector, a::Real) = (check(v.ctx.handle, @ccall libb200.b200_fill(v.ctx.handle::Ctx, v.n::Int64, Float64(a)::Float64, v.ptr::Ptr{Float64})::Int32); v)
# the BLAS-1 surface L2-L4 touch (SURVEY.md §8b b5)
function LinearAlgebra.axpy!(a::Real, x::B200Vector, y::B200Vector)
    check(y.ctx.handle, @ccall libb200.b200_axpy(y.ctx.handle::Ctx, y.n::Int64, Float64(a)::Float64, x.ptr::Ptr{Float64}, y.ptr::Ptr{Float64})::Int32); y
end
function LinearAlgebra.axpby!(a::Real, x::B200Vector, b::Real, y::B200Vector)
    check(y.ctx.handle, @ccall libb200.b200_axpby(y.ctx.handle::Ctx, y.n::Int64, Float64(a)::Float64, x.ptr::Ptr{Float64}, Float64(b)::Float64, y.ptr::Ptr{Float64})::Int32); y
end
LinearAlgebra.rmul!(x::B200Vector, a::Real) = (check(x.ctx.handle, @ccall libb200.b200_scal(x.ctx.handle::Ctx, x.n::Int64, Float64(a)::Float64, x.ptr::Ptr{Float64})::Int32); x)
function LinearAlgebra.dot(x::B200Vector, y::B200Vector)
    r = Ref{Float64}(0.0)
    check(x.ctx.handle, @ccall libb200.b200_dot(x.ctx.handle::Ctx, x.n::Int64, x.ptr::Ptr{Float64}, y.ptr::Ptr{Float64}, r::Ref{Float64})::Int32)
    return r[]
end
function LinearAlgebra.norm(x::B200Vector, p::Real = 2)
    r = Ref{Float64}(0.0)
    if p == 2
        check(x.ctx.handle, @ccall libb200.b200_nrm2(x.ctx.handle::Ctx, x.n::Int64, x.ptr::Ptr{Float64}, r::Ref{Float64})::Int32)
    elseif p == Inf
        check(x.ctx.handle, @ccall libb200.b200_norminf(x.ctx.handle::Ctx, x.n::Int64, x.ptr::Ptr{Float64}, r::Ref{Float64})::Int32)
    else
        error("only the 2- and Inf-norms are provided on the device")
    end
    return r[]
end
Base.maximum(::typeof(abs), x::B200Vector) = norm(x, Inf)   # common_defaults.jl:37
function Base.:(==)(x::B200Vector, y::B200Vector)            # termination_conditions.jl:446
    r = Ref{Int32}(0)
    check(x.ctx.handle, @ccall libb200.b200_equal(x.ctx.handle::Ctx, x.n::Int64, x.ptr::Ptr{Float64}, y.ptr::Ptr{Float64}, r::Ref{Int32})::Int32)
    return r[] == 1
end

# ------------------------------------------------------------------ b1: built-in device problems as f! / jvp! / vjp!
mutable struct Problem
    ctx::Context
    handle::Ptr{Cvoid}
    n::Int
end
function brusselator(ctx::Context, dim::Integer, N::Integer, A, B, alpha)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    if dim == 2
        check(ctx.handle, @ccall libb200.b200_problem_create_bruss2d(ctx.handle::Ctx, N::Int32, A::Float64, B::Float64, alpha::Float64, h::Ref{Ptr{Cvoid}})::Int32)
    else
        check(ctx.handle, @ccall libb200.b200_problem_create_bruss3d(ctx.handle::Ctx, N::Int32, A::Float64, B::Float64, alpha::Float64, h::Ref{Ptr{Cvoid}})::Int32)
    end
    p = Problem(ctx, h[], 2 * N^dim)
    finalizer(x -> (@ccall libb200.b200_problem_destroy(x.handle::Ptr{Cvoid})::Int32), p)
    return p
end
residual!(du::B200Vector, u::B200Vector, prob::Problem) = (check(prob.ctx.handle, @ccall libb200.b200_residual(prob.handle::Ptr{Cvoid}, u.ptr::Ptr{Float64}, du.ptr::Ptr{Float64})::Int32); nothing)
jvp!(Jv::B200Vector, v::B200Vector, u::B200Vector, prob::Problem) = (check(prob.ctx.handle, @ccall libb200.b200_jvp(prob.handle::Ptr{Cvoid}, u.ptr::Ptr{Float64}, v.ptr::Ptr{Float64}, Jv.ptr::Ptr{Float64})::Int32); nothing)
vjp!(Jtw::B200Vector, w::B200Vector, u::B200Vector, prob::Problem) = (check(prob.ctx.handle, @ccall libb200.b200_vjp(prob.handle::Ptr{Cvoid}, u.ptr::Ptr{Float64}, w.ptr::Ptr{Float64}, Jtw.ptr::Ptr{Float64})::Int32); nothing)

"""
    brusselator_function(prob) -> NonlinearFunction{true}

`NonlinearFunction{true}(f!; jvp = jvp!, vjp = vjp!)` whose closures call the fused sm_100a kernels (b1).  With it
`solve(NonlinearProblem(fn, u0::B200Vector, p), NewtonRaphson(linsolve = B200GMRES()))` runs the reference's own driver
(`step!`, termination, stats) over device arrays, every arithmetic step in libb200newton.
"""
function brusselator_function(prob::Problem)
    f!(du, u, p) = residual!(du, u, prob)
    jvpf!(Jv, v, u, p) = jvp!(Jv, v, u, prob)
    vjpf!(Jtw, w, u, p) = vjp!(Jtw, w, u, prob)
    return SciMLBase.NonlinearFunction{true}(f!; jvp = jvpf!, vjp = vjpf!)
end

# ------------------------------------------------------------------ option structs (bit-compatible with the header)
struct GmresOpts
    memory::Int32; restart::Int32; itmax::Int32; orth::Int32; warm_start::Int32; engine::Int32; check_every::Int32; block::Int32
    atol::Float64; rtol::Float64
end
struct GmresStats
    status::Int32; iters::Int32; nmatvec::Int32; restarts::Int32
    rnorm0::Float64; rnorm::Float64; tol::Float64; bytes::Float64
end
struct NewtonOpts
    abstol::Float64; reltol::Float64
    maxiters::Int32; linsolve::Int32; jvp_mode::Int32; globalization::Int32; forcing::Int32; termination::Int32; store_trace::Int32; fused_step::Int32
    gmres::GmresOpts
    ew_eta0::Float64; ew_eta_max::Float64; ew_gamma::Float64; ew_alpha::Float64; ew_safeguard_threshold::Float64
    ew_safeguard::Int32; max_shrink_times::Int32
    tr_step_threshold::Float64; tr_shrink_threshold::Float64; tr_expand_threshold::Float64; tr_shrink_factor::Float64
    tr_expand_factor::Float64; tr_max_trust_radius::Float64; tr_initial_trust_radius::Float64
    ls_c1::Float64; ls_rho_hi::Float64; ls_rho_lo::Float64; ls_maxiters::Int32; precond::Int32
    descent::Int32; tr_scheme::Int32; pt_alpha_initial::Float64
end
struct NewtonResult
    retcode::Int32; nsteps::Int32; nf::Int32; njacs::Int32; nfactors::Int32; nsolve::Int32; njvp::Int32; ntrace::Int32
    resid_inf::Float64; bytes::Float64
end
function default_newton_opts()
    r = Ref{NewtonOpts}()
    @ccall libb200.b200_newton_opts_default(r::Ref{NewtonOpts})::Cvoid
    return r[]
end

# ------------------------------------------------------------------ b4: custom LinearSolve algorithm
"""
    B200GMRES(; restart = 0, memory = 20, orth = :cgs2)

`LinearSolve.SciMLLinearSolveAlgorithm` with `needs_concrete_A == false`, so `construct_jacobian_cache` hands it the
matrix-free `JacobianOperator` (jacobian.jl:43-47).  `solve!` runs the device-resident GMRES; when `A` wraps a
`brusselator_function` the JVP is the built-in fused kernel, otherwise `mul!(w, A, v)` is called back between kernels.
"""
Base.@kwdef struct B200GMRES <: LinearSolve.SciMLLinearSolveAlgorithm
    restart::Int = 0
    memory::Int = 20
    orth::Symbol = :cgs2
end
LinearSolve.needs_concrete_A(::B200GMRES) = false
LinearSolve.needs_square_A(::B200GMRES) = true

mutable struct GmresCache
    handle::Ptr{Cvoid}
    n::Int
end
function LinearSolve.init_cacheval(alg::B200GMRES, A, b::B200Vector, u, Pl, Pr, maxiters::Int, abstol, reltol, verbose, assumptions)
    orth = alg.orth === :mgs ? 0 : alg.orth === :cgs ? 1 : 2
    o = Ref(GmresOpts(alg.memory, alg.restart, maxiters, orth, 0, 0, 8, 0, abstol, reltol))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(b.ctx.handle, @ccall libb200.b200_gmres_create(b.ctx.handle::Ctx, length(b)::Int64, o::Ref{GmresOpts}, h::Ref{Ptr{Cvoid}})::Int32)
    c = GmresCache(h[], length(b))
    finalizer(x -> (@ccall libb200.b200_gmres_destroy(x.handle::Ptr{Cvoid})::Int32), c)
    return c
end

# host callback used when the operator is an arbitrary SciMLOperator: mul!(y, A, x) on device vectors
function _matvec_trampoline(user::Ptr{Cvoid}, x::Ptr{Float64}, y::Ptr{Float64})::Int32
    try
        (A, ctx, n) = unsafe_pointer_to_objref(user)::Tuple
        mul!(unsafe_wrap_device(ctx, y, n), A, unsafe_wrap_device(ctx, x, n))
        return Int32(0)
    catch
        return Int32(1)       # no exception may cross the ABI
    end
end
unsafe_wrap_device(ctx, p, n) = (v = ccall(:jl_new_struct_uninit, Any, (Any,), B200Vector); v.ctx = ctx; v.ptr = p; v.n = n; v)

function SciMLBase.solve!(cache::LinearSolve.LinearCache, alg::B200GMRES; kwargs...)
    gm = cache.cacheval::GmresCache
    b, x, A = cache.b::B200Vector, cache.u::B200Vector, cache.A
    ctx = b.ctx
    check(ctx.handle, @ccall libb200.b200_gmres_set_tolerances(gm.handle::Ptr{Cvoid}, Float64(cache.abstol)::Float64, Float64(cache.reltol)::Float64)::Int32)
    op = Ref{Ptr{Cvoid}}(C_NULL)
    prob = device_problem(A)                       # non-nothing when f.jvp is one of our closures
    if prob !== nothing
        u = cache.p.u::B200Vector                  # LinearSolveParameters(u, p)  (linear_solve.jl:1-4)
        check(ctx.handle, @ccall libb200.b200_linop_from_problem(prob.handle::Ptr{Cvoid}, u.ptr::Ptr{Float64}, 0::Int32, op::Ref{Ptr{Cvoid}})::Int32)
    else
        boxed = Ref{Any}((A, ctx, length(b)))
        cb = @cfunction(_matvec_trampoline, Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}))
        check(ctx.handle, @ccall libb200.b200_linop_from_callback(ctx.handle::Ctx, length(b)::Int64, cb::Ptr{Cvoid}, pointer_from_objref(boxed)::Ptr{Cvoid}, op::Ref{Ptr{Cvoid}})::Int32)
    end
    # LinearSolve `precs`: cache.Pl / cache.Pr.  A `B200BlockJacobi` marker maps to the built-in device preconditioner; any
    # other non-identity preconditioner is wrapped like the operator above (callback applying `ldiv!`).
    pl = cache.Pl isa B200BlockJacobi && prob !== nothing ? block_jacobi_op(prob, cache.p.u) : C_NULL
    pr = cache.Pr isa B200BlockJacobi && prob !== nothing ? block_jacobi_op(prob, cache.p.u) : C_NULL
    check(ctx.handle, @ccall libb200.b200_gmres_set_precond(gm.handle::Ptr{Cvoid}, pl::Ptr{Cvoid}, pr::Ptr{Cvoid})::Int32)
    st = Ref{GmresStats}()
    GC.@preserve st check(ctx.handle, @ccall libb200.b200_gmres_solve(gm.handle::Ptr{Cvoid}, op[]::Ptr{Cvoid}, b.ptr::Ptr{Float64}, x.ptr::Ptr{Float64}, st::Ref{GmresStats})::Int32)
    @ccall libb200.b200_gmres_set_precond(gm.handle::Ptr{Cvoid}, C_NULL::Ptr{Cvoid}, C_NULL::Ptr{Cvoid})::Int32
    @ccall libb200.b200_linop_destroy(op[]::Ptr{Cvoid})::Int32
    pl == C_NULL || @ccall libb200.b200_linop_destroy(pl::Ptr{Cvoid})::Int32
    pr == C_NULL || @ccall libb200.b200_linop_destroy(pr::Ptr{Cvoid})::Int32
    s = st[]
    rc = s.status == 1 || s.status == 3 ? ReturnCode.Success : s.status == 2 ? ReturnCode.MaxIters : ReturnCode.Failure
    return SciMLBase.build_linear_solution(alg, x, s.rnorm, cache; retcode = rc, iters = Int(s.iters))
end
"""`precs = (A, p) -> (B200BlockJacobi(), I)`: inverse of the 2x2 species blocks on the diagonal of the Brusselator Jacobian."""
struct B200BlockJacobi end
function block_jacobi_op(prob, u::B200Vector)
    op = Ref{Ptr{Cvoid}}(C_NULL)
    check(prob.ctx.handle, @ccall libb200.b200_linop_block_jacobi(prob.handle::Ptr{Cvoid}, u.ptr::Ptr{Float64}, op::Ref{Ptr{Cvoid}})::Int32)
    return op[]
end
device_problem(A) = nothing   # specialised for JacobianOperators built from `brusselator_function` (holds the Problem)

# ------------------------------------------------------------------ whole-solve fast path
"""
    B200NewtonKrylov(; linsolve = :gmres | :dense_lu | :sparse_gmres, globalization = :none | :trust_region,
                       forcing = nothing | EisenstatWalkerForcing2(), orth = :cgs2)

New `AbstractNonlinearSolveAlgorithm` whose `__solve` runs the entire Newton iteration inside the library
(`b200_newton_*`) and returns `SciMLBase.build_solution(prob, alg, u, resid; retcode, stats = NLStats(nf, njacs, nfactors,
nsolve, nsteps), original)`, as the wrapper algorithms in the reference do (ext/NonlinearSolvePETScExt.jl:162-166).
"""
Base.@kwdef struct B200NewtonKrylov <: NonlinearSolveBase.AbstractNonlinearSolveAlgorithm
    problem::Problem
    linsolve::Symbol = :gmres            # :gmres | :dense_lu | :sparse_gmres
    globalization::Symbol = :none        # :none | :trust_region | :linesearch (BackTracking)
    radius_update_scheme::Symbol = :simple  # :simple | :nlsolve | :nocedal_wright | :hei | :yuan | :fan
    descent::Symbol = :newton            # :newton | :pseudo_transient
    alpha_initial::Float64 = 1.0e-3      # PseudoTransient(alpha_initial)
    forcing::Bool = false                # EisenstatWalkerForcing2()
    precs::Symbol = :none                # :none | :block_jacobi_left | :block_jacobi_right
    orth::Symbol = :cgs2
end

const _TR_SCHEMES = (simple = 0, nlsolve = 1, nocedal_wright = 2, hei = 3, yuan = 4, fan = 5)

function SciMLBase.__solve(prob::SciMLBase.NonlinearProblem, alg::B200NewtonKrylov, args...;
        abstol = nothing, reltol = nothing, maxiters = 1000, kwargs...)
    dp, ctx = alg.problem, alg.problem.ctx
    o = default_newton_opts()
    g = o.gmres
    g = GmresOpts(g.memory, g.restart, g.itmax, alg.orth === :mgs ? 0 : alg.orth === :cgs ? 1 : 2, g.warm_start, g.engine, g.check_every, 0, 0.0, 0.0)
    o = NewtonOpts(something(abstol, 0.0), something(reltol, 0.0), maxiters,
        alg.linsolve === :gmres ? 0 : alg.linsolve === :dense_lu ? 1 : 2, o.jvp_mode,
        alg.globalization === :trust_region ? 1 : alg.globalization === :linesearch ? 2 : 0,
        alg.forcing ? 1 : 0, o.termination, 0, 1, g, o.ew_eta0, o.ew_eta_max, o.ew_gamma, o.ew_alpha, o.ew_safeguard_threshold, o.ew_safeguard,
        o.max_shrink_times, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,      # tr_* : scheme defaults
        0.0, 0.0, 0.0, Int32(0),                                    # ls_* : BackTracking defaults
        Int32(alg.precs === :block_jacobi_left ? 1 : alg.precs === :block_jacobi_right ? 2 : 0),
        Int32(alg.descent === :pseudo_transient ? 1 : 0), Int32(getfield(_TR_SCHEMES, alg.radius_update_scheme)), alg.alpha_initial)
    nw = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.handle, @ccall libb200.b200_newton_create(dp.handle::Ptr{Cvoid}, Ref(o)::Ref{NewtonOpts}, nw::Ref{Ptr{Cvoid}})::Int32)
    try
        res = Ref{NewtonResult}()
        u0 = prob.u0
        if u0 isa B200Vector
            check(ctx.handle, @ccall libb200.b200_newton_reinit(nw[]::Ptr{Cvoid}, u0.ptr::Ptr{Float64})::Int32)
            check(ctx.handle, @ccall libb200.b200_newton_solve(nw[]::Ptr{Cvoid}, res::Ref{NewtonResult})::Int32)
            pu = Ref{Ptr{Float64}}(); pf = Ref{Ptr{Float64}}()
            @ccall libb200.b200_newton_u(nw[]::Ptr{Cvoid}, pu::Ref{Ptr{Float64}})::Int32
            @ccall libb200.b200_newton_fu(nw[]::Ptr{Cvoid}, pf::Ref{Ptr{Float64}})::Int32
            u = copy(unsafe_wrap_device(ctx, pu[], dp.n)); fu = copy(unsafe_wrap_device(ctx, pf[], dp.n))
        else    # host arrays: the end-to-end call (H2D, solve, D2H)
            u0h = vec(collect(Float64, u0)); u = similar(u0h); fu = similar(u0h)
            GC.@preserve u0h u fu check(ctx.handle, @ccall libb200.b200_newton_solve_host(nw[]::Ptr{Cvoid}, pointer(u0h)::Ptr{Float64},
                pointer(u)::Ptr{Float64}, pointer(fu)::Ptr{Float64}, res::Ref{NewtonResult})::Int32)
            u = reshape(u, size(u0)); fu = reshape(fu, size(u0))
        end
        r = res[]
        stats = NLStats(Int(r.nf), Int(r.njacs), Int(r.nfactors), Int(r.nsolve), Int(r.nsteps))
        return SciMLBase.build_solution(prob, alg, u, fu; retcode = retcode(r.retcode), stats, original = r)
    finally
        @ccall libb200.b200_newton_destroy(nw[]::Ptr{Cvoid})::Int32
    end
end

# ------------------------------------------------------------------ b6: ensemble algorithm
"""
    EnsembleB200(; rank = 0, world_size = 1)

`solve(ensembleprob, alg, EnsembleB200(); trajectories)` evaluates `prob_func(prob, i, repeat)` for the trajectories of
this rank's contiguous block, packs `u0` / `(A, B)` into batches and calls `b200_ens_solve` (one CTA per trajectory).
One process per GPU; the cross-rank gather of `EnsembleSolution.u` is an `ncclAllGather` issued by the launcher
(MPI.jl / Distributed wrapper), not by the library — the data path itself has no collective.
"""
Base.@kwdef struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    rank::Int = 0
    world_size::Int = 1
end

function shard_range(K::Integer, rank::Integer, world::Integer)
    base, rem = divrem(K, world)
    lo = rank * base + min(rank, rem)
    return (lo + 1):(lo + base + (rank < rem ? 1 : 0))      # 1-based trajectory indices of this rank
end

function SciMLBase.__solve(ens::SciMLBase.AbstractEnsembleProblem, alg::B200NewtonKrylov, ealg::EnsembleB200; trajectories, abstol = nothing, kwargs...)
    ctx = alg.problem.ctx
    idx = shard_range(trajectories, ealg.rank, ealg.world_size)
    probs = [ens.prob_func(ens.prob, i, 1) for i in idx]
    N = round(Int, sqrt(length(first(probs).u0) ÷ 2)); n = 2N^2; K = length(idx)
    u0 = reduce(hcat, (vec(p.u0) for p in probs)); A = [p.p[1] for p in probs]; B = [p.p[2] for p in probs]
    d_u0, d_A, d_B, d_u = B200Vector(ctx, u0), B200Vector(ctx, A), B200Vector(ctx, B), B200Vector(ctx, n * K)
    o = Ref(NewtonOpts(something(abstol, 0.0), 0.0, 1000, 0, 0, 0, 0, 0, 0, 1, default_newton_opts().gmres, 0.5, 0.9, 0.9, 2.0, 0.1, 1, 32, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, Int32(0), Int32(0)))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.handle, @ccall libb200.b200_ens_create(ctx.handle::Ctx, N::Int32, K::Int32, Float64(first(probs).p[3])::Float64, o::Ref{NewtonOpts}, h::Ref{Ptr{Cvoid}})::Int32)
    d_res = B200Vector(ctx, K); rc = Vector{Int32}(undef, K)
    # (retcodes / nsteps / njvp device buffers elided: same pattern as d_res)
    check(ctx.handle, @ccall libb200.b200_ens_solve(h[]::Ptr{Cvoid}, d_u0.ptr::Ptr{Float64}, d_A.ptr::Ptr{Float64}, d_B.ptr::Ptr{Float64}, d_u.ptr::Ptr{Float64},
        d_res.ptr::Ptr{Float64}, C_NULL::Ptr{Int32}, C_NULL::Ptr{Int32}, C_NULL::Ptr{Int32}, C_NULL::Ptr{Cvoid})::Int32)
    @ccall libb200.b200_ens_destroy(h[]::Ptr{Cvoid})::Int32
    U = reshape(Array(d_u), n, K)
    sols = [SciMLBase.build_solution(probs[k], alg, reshape(U[:, k], size(probs[k].u0)), nothing; retcode = ReturnCode.Success) for k in 1:K]
    return SciMLBase.EnsembleSolution(sols, 0.0, true)
end

export Context, B200Vector, brusselator, brusselator_function, B200GMRES, B200NewtonKrylov, EnsembleB200

end # module
