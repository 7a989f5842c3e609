# FiniteDiffB200.jl — Julia host side of the B200 drop-in for FiniteDiff.jl's coloured Jacobian path.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` binary (SURVEY.md §8c).  The file is the binding
# a maintainer adds; the same C ABI (include/fdjac_b200.h) is exercised end-to-end through the Python/ctypes mirror
# (finitediff.jl_b200/api.py), which keeps the same names and argument meaning.
#
# What it does: adds METHODS to FiniteDiff.finite_difference_jacobian! (src/jacobians.jl:504-514) that dispatch on a
# device-array `x` (CUDA.CuVector{Float64}) and forward the WHOLE colour loop to libfdjac_b200.so with one ccall — not
# the per-colour hooks (`_colorediteration!`), because a host round trip per colour would forfeit the device residency.
# Everything else of FiniteDiff.jl (CPU arrays, gradients, hessians, jvp, complex step) keeps using the stock package.
module FiniteDiffB200

using FiniteDiff, SparseArrays, CUDA
import FiniteDiff: finite_difference_jacobian!, JacobianCache

const libfdjac = get(ENV, "FDJAC_B200_LIB", "libfdjac_b200.so")

const FDB_FORWARD, FDB_CENTRAL = Cint(0), Cint(1)
const FDB_J_CSC_NZVAL, FDB_J_DENSE, FDB_J_BAND, FDB_J_SLOTS = Cint(0), Cint(1), Cint(2), Cint(3)

# mirror of fdb_plan_opts (include/fdjac_b200.h)
struct PlanOpts
    fdtype::Int32
    device::Int32
    use_current_device::Int32
    no_drift::Int32
    max_batch::Int64
    scratch_bytes::Int64
    rank::Int32
    world::Int32
    partition::Int32
    strategy::Int32
    use_graph::Int32
    reserved::Int32
end
PlanOpts(fd; max_batch = 1, rank = 0, world = 1) =
    PlanOpts(fd, 0, 1, 0, max_batch, 0, rank, world, 0, 0, 0, 0)

struct FdbError <: Exception
    status::Cint
    msg::String
end
function check(st::Cint)
    st == 0 && return nothing
    throw(FdbError(st, unsafe_string(ccall((:fdb_last_error, libfdjac), Cstring, ()))))
end

fdcode(::Val{:forward}) = FDB_FORWARD
fdcode(::Val{:central}) = FDB_CENTRAL
fdcode(::Val{T}) where {T} = FiniteDiff.fdtype_error(Float64)   # src/epsilons.jl:159-167 (complex step: §8f "next")

# ---- plans: keyed on the identity of (sparsity pattern, colorvec, fdtype); the reference redoes this work per call
mutable struct Plan
    handle::Ptr{Cvoid}
    function Plan(h)
        p = new(h)
        finalizer(p -> ccall((:fdb_plan_destroy, libfdjac), Cint, (Ptr{Cvoid},), p.handle), p)
        p
    end
end
const PLANS = IdDict{Any, Plan}()

colorptr(cv::AbstractUnitRange) = (first(cv) == 1 ? C_NULL : pointer(collect(Int64, cv)))   # NULL => 1:n (jacobians.jl:16)
colorptr(cv::Vector{Int64}) = pointer(cv)

# SparseMatrixCSC{Float64,Int64}: colptr / rowval cross the ABI as they are — Int64, 1-based, host memory.
function plan_for(J::SparseMatrixCSC{Float64, Int64}, sparsity::SparseMatrixCSC, colorvec, fd)
    get!(PLANS, (J.colptr, J.rowval, sparsity.colptr, sparsity.rowval, colorvec, fd)) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        opts = Ref(PlanOpts(fd))
        m, n = size(sparsity)
        GC.@preserve J sparsity colorvec begin
            check(ccall((:fdb_plan_create_csc, libfdjac), Cint,
                (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int64}, Ref{PlanOpts}),
                h, m, n, pointer(sparsity.colptr), pointer(sparsity.rowval), FDB_J_CSC_NZVAL,
                J === sparsity ? C_NULL : pointer(J.colptr), J === sparsity ? C_NULL : pointer(J.rowval), 0,
                colorptr(colorvec), opts))
        end
        Plan(h[])
    end
end

# ---- user function: a Julia closure f!(fx::CuVector, x::CuVector) behind the fdb_fn C signature
# int f(void* ctx, double* d_fx, const double* d_x, int64 batch, int64 ldfx, int64 ldx, void* stream)
function f_trampoline(ctx::Ptr{Cvoid}, fx::CuPtr{Float64}, x::CuPtr{Float64}, batch::Int64, ldfx::Int64, ldx::Int64,
        stream::Ptr{Cvoid})::Cint
    st = unsafe_pointer_to_objref(ctx)::FnState
    try
        for b in 0:(batch - 1)          # plans are created with max_batch = 1 for plain closures
            fxv = unsafe_wrap(CuArray, fx + b * ldfx * 8, st.m)
            xv = unsafe_wrap(CuArray, x + b * ldx * 8, st.n)
            st.f(fxv, xv)               # must only enqueue on the task-local CUDA.jl stream (== `stream`)
        end
        return Cint(0)
    catch err
        st.err = err                    # never unwind through C: report, rethrow on the Julia side
        return Cint(1)
    end
end
mutable struct FnState
    f::Any
    m::Int
    n::Int
    err::Any
end

"""
    finite_difference_jacobian!(J::CuSparseJ, f!, x::CuVector{Float64}, cache::JacobianCache, f_in = nothing; ...)

Same signature and keyword meaning as `src/jacobians.jl:504-514`.  `J` is a `SparseMatrixCSC` whose `nzval` lives on the
device (a thin wrapper type `DeviceCSC` below); `x`, `cache.fx`, `f_in` are `CuVector{Float64}`.
"""
struct DeviceCSC
    host::SparseMatrixCSC{Float64, Int64}   # pattern (colptr / rowval) — what the reference dispatches on
    nzval::CuVector{Float64}                # values on the device
end
Base.size(J::DeviceCSC) = size(J.host)

function finite_difference_jacobian!(J::DeviceCSC, f, x::CuVector{Float64},
        cache::JacobianCache{T1, T2, T3, T4, cType, sType, fdtype, returntype}, f_in = nothing;
        relstep = FiniteDiff.default_relstep(fdtype, eltype(x)), absstep = relstep,
        colorvec = cache.colorvec, sparsity = cache.sparsity, dir = true) where {T1, T2, T3, T4, cType, sType, fdtype, returntype}
    sp = sparsity isa DeviceCSC ? sparsity.host : sparsity
    plan = plan_for(J.host, sp, colorvec, fdcode(fdtype))
    st = FnState(f, size(J, 1), length(x), nothing)
    cf = @cfunction(f_trampoline, Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Int64, Int64, Int64, Ptr{Cvoid}))
    GC.@preserve st x J cache f_in begin
        rc = ccall((:fdb_jacobian, libfdjac), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64},
                Float64, Float64, Float64, Ptr{Cvoid}),
            plan.handle, cf, pointer_from_objref(st), pointer(x), pointer(J.nzval), pointer(cache.fx),
            f_in === nothing ? CU_NULL : pointer(f_in), relstep, absstep, Float64(dir), CUDA.stream().handle)
    end
    st.err === nothing || throw(st.err)     # an exception inside f! propagates like in the reference
    check(rc)
    nothing                                  # jacobians.jl:652
end

# ---- column-block shards (few-colour problems on several GPUs): one process per GPU, each owning the columns c0+1:c1.
# `sub` is the block's pattern (colptr slice rebased to 1, rows rebased to the block's first row), `f_rows!` computes
# that row range from the x slice the rows depend on, `eps` holds the step sizes of the FULL x (color_eps below).
function color_eps!(eps::CuVector{Float64}, epsplan::Plan, x::CuVector{Float64}; relstep = 0.0, absstep = 0.0, dir = true)
    check(ccall((:fdb_color_eps, libfdjac), Cint,
        (Ptr{Cvoid}, CuPtr{Float64}, Float64, Float64, Float64, CuPtr{Float64}, Ptr{Cvoid}),
        epsplan.handle, pointer(x), relstep, absstep, Float64(dir), pointer(eps), CUDA.stream().handle))
    eps
end

function eps_plan(n::Integer, colorvec::Vector{Int64}, fd::Cint)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    opts = Ref(PlanOpts(fd))
    check(ccall((:fdb_eps_plan_create, libfdjac), Cint, (Ptr{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{PlanOpts}),
        h, n, colorvec, opts))
    Plan(h[])
end

set_external_eps!(plan::Plan, eps::CuVector{Float64}) =
    check(ccall((:fdb_plan_set_external_eps, libfdjac), Cint, (Ptr{Cvoid}, CuPtr{Float64}), plan.handle, pointer(eps)))

end # module
