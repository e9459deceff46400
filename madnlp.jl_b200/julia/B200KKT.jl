# B200KKT.jl -- the shim a MadNLP.jl maintainer adds to drive libb200kkt.so (include/b200kkt.h) from MadNLP's
# existing plugin surface.  NOT RUN in this repository's CI (no Julia in the build image); it is the reference-side
# binding INTEGRATION.md describes, kept next to the library so that the C ABI and its consumer evolve together.
#
#   using MadNLP, MadNLPGPU, CUDA
#   include("B200KKT.jl"); using .B200KKT
#   madnlp(nlp; kkt_system = MadNLP.SparseCondensedKKTSystem, linear_solver = B200KKT.B200Solver,
#          equality_treatment = MadNLP.RelaxEquality, tol = 1e-4)
#
# Interface implemented (src/LinearSolvers/linearsolvers.jl:13-95; same list every plugin imports,
# lib/MadNLPHSL/src/MadNLPHSL.jl:3-32): constructor(csc; opt, logger), factorize!, solve_linear_system!, is_inertia,
# inertia, improve!, introduce, input_type, default_options, is_supported, is_async.
module B200KKT

import MadNLP
import MadNLP: AbstractLinearSolver, AbstractOptions, MadNLPLogger, SymbolicException, FactorizationException,
    SolveException, factorize!, solve_linear_system!, is_inertia, inertia, improve!, introduce, input_type,
    default_options, is_supported, is_async
using CUDA, CUDA.CUSPARSE

const libb200kkt = get(ENV, "B200KKT_LIB", "libb200kkt.so")

# mirror of `struct b2_options` (include/b200kkt.h)
Base.@kwdef mutable struct B200Options <: AbstractOptions
    b200_ordering::Int32 = 0            # B2_ORDER_METIS_ND
    b200_nemin::Int32 = 16
    b200_relax_zeros::Float64 = 0.25
    b200_pivot_eps::Float64 = 1e-13
    b200_use_cuda_graph::Int32 = 1
    b200_small_front_max::Int32 = 160
    b200_kkt_n_primal::Int32 = 0        # set by the KKT overloads below for SparseKKTSystem
    b200_fuse_max_fronts::Int32 = 16
    b200_dep_schedule::Int32 = 1
end

struct CB2Options
    ordering::Int32; nemin::Int32; relax_zeros::Float64; pivot_eps::Float64
    use_cuda_graph::Int32; small_front_max::Int32; n_parts::Int32; part_rank::Int32
    kkt_n_primal::Int32; fuse_max_fronts::Int32; dep_schedule::Int32; reserved::NTuple{5,Int32}
end
CB2Options(o::B200Options) = CB2Options(o.b200_ordering, o.b200_nemin, o.b200_relax_zeros, o.b200_pivot_eps,
    o.b200_use_cuda_graph, o.b200_small_front_max, 1, 0, o.b200_kkt_n_primal, o.b200_fuse_max_fronts, o.b200_dep_schedule, ntuple(_ -> Int32(0), 5))

last_error() = unsafe_string(ccall((:b2_last_error, libb200kkt), Cstring, ()))
function check(rc::Cint, exc)
    rc == 0 && return
    rc == 3 && throw(SymbolicException())
    rc == 4 && throw(FactorizationException())
    rc == 5 && throw(SolveException())
    error("b200kkt error $rc: $(last_error())")
end

mutable struct B200Solver{T} <: AbstractLinearSolver{T}
    handle::Ptr{Cvoid}
    tril::CuSparseMatrixCSC{T,Int32}    # kept by reference: values are re-read on every factorize! (cudss.jl:154-158)
    opt::B200Options
    logger::MadNLPLogger
end

function B200Solver(csc::CuSparseMatrixCSC{Float64,Int32}; opt = B200Options(), logger = MadNLPLogger())
    n = size(csc, 1)
    colptr = Array(csc.colPtr) .- Int32(1)          # host, 0-based (analysis runs on the host, once)
    rowval = Array(csc.rowVal) .- Int32(1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    copt = Ref(CB2Options(opt))
    rc = ccall((:b2_create, libb200kkt), Cint,
        (Int32, Int64, Ptr{Int32}, Ptr{Int32}, CuPtr{Float64}, Ptr{CB2Options}, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
        n, length(rowval), colptr, rowval, pointer(csc.nzVal), copt, C_NULL, h)
    check(rc, SymbolicException)
    M = B200Solver{Float64}(h[], csc, opt, logger)
    finalizer(m -> ccall((:b2_destroy, libb200kkt), Cint, (Ptr{Cvoid},), m.handle), M)
    return M
end

stream_ptr() = Ptr{Cvoid}(UInt(CUDA.stream().handle))

function factorize!(M::B200Solver)
    check(ccall((:b2_factorize, libb200kkt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), M.handle, stream_ptr()), FactorizationException)
    return M
end

function solve_linear_system!(M::B200Solver{T}, x::CuVector{T}) where T
    check(ccall((:b2_solve, libb200kkt), Cint, (Ptr{Cvoid}, CuPtr{T}, Int32, Ptr{Cvoid}), M.handle, pointer(x), 1, stream_ptr()), SolveException)
    return x
end

is_inertia(::B200Solver) = true
function inertia(M::B200Solver)
    p = Ref{Int64}(0); z = Ref{Int64}(0); n = Ref{Int64}(0)
    check(ccall((:b2_inertia, libb200kkt), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}),
        M.handle, p, z, n, stream_ptr()), FactorizationException)
    return (Int(p[]), Int(z[]), Int(n[]))       # (num_pos, num_zero, num_neg): the order src/IPM/solver.jl:626 destructures
end

function improve!(M::B200Solver)
    ch = Ref{Int32}(0)
    ccall((:b2_improve, libb200kkt), Cint, (Ptr{Cvoid}, Ptr{Int32}), M.handle, ch)
    return ch[] != 0
end

introduce(::B200Solver) = "b200kkt (sm_100a multifrontal LDL')"
input_type(::Type{<:B200Solver}) = :csc
default_options(::Type{<:B200Solver}) = B200Options()
is_supported(::Type{<:B200Solver}, ::Type{Float64}) = true
is_supported(::Type{<:B200Solver}, ::Type{Float32}) = false
is_async(::B200Solver) = true

# ---------------------------------------------------------------------------------------------------------------------
# Assembly overloads on CuVector storage: the same leaf functions MadNLPGPU overloads
# (lib/MadNLPGPU/src/KKT/gpu_sparse.jl:308-382), now one ccall each.  `kkt.ext` holds the native plans created in
# get_sparse_condensed_ext (one-time): ext.cond (b2_condensed_plan), ext.hess_plan / ext.jt_plan (b2_transfer_plan).
# ---------------------------------------------------------------------------------------------------------------------
function MadNLP.build_kkt!(kkt::MadNLP.SparseCondensedKKTSystem{T,VT}) where {T, VT <: CuVector{T}}
    rc = ccall((:b2_condensed_assemble, libb200kkt), Cint,
        (Ptr{Cvoid}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, Ptr{Cvoid}),
        kkt.ext.cond, pointer(MadNLP.nzval(kkt.aug_com)), pointer(kkt.pr_diag), pointer(kkt.du_diag),
        pointer(MadNLP.nzval(kkt.hess_com)), pointer(MadNLP.nzval(kkt.jt_csc)), pointer(kkt.diag_buffer), stream_ptr())
    check(rc, FactorizationException)
end

function MadNLP.compress_hessian!(kkt::MadNLP.SparseCondensedKKTSystem{T,VT}) where {T, VT <: CuVector{T}}
    check(ccall((:b2_transfer, libb200kkt), Cint, (Ptr{Cvoid}, CuPtr{T}, CuPtr{T}, Ptr{Cvoid}),
        kkt.ext.hess_plan, pointer(MadNLP.nzval(kkt.hess_com)), pointer(kkt.hess_raw.V), stream_ptr()), FactorizationException)
end

function MadNLP.compress_jacobian!(kkt::MadNLP.SparseCondensedKKTSystem{T,VT}) where {T, VT <: CuVector{T}}
    check(ccall((:b2_transfer, libb200kkt), Cint, (Ptr{Cvoid}, CuPtr{T}, CuPtr{T}, Ptr{Cvoid}),
        kkt.ext.jt_plan, pointer(MadNLP.nzval(kkt.jt_csc)), pointer(kkt.jt_coo.V), stream_ptr()), FactorizationException)
end

# transfer!(dest::CuSparseMatrixCSC, src::SparseMatrixCOO, plan) for SparseKKTSystem (accumulating, unlike
# lib/MadNLPGPU/ext/MadNLPGPUCUDAExt/cuda_sparse.jl:6-12 which drops duplicates)
function b200_transfer!(dest_nz::CuVector{T}, V::CuVector{T}, plan::Ptr{Cvoid}) where T
    check(ccall((:b2_transfer, libb200kkt), Cint, (Ptr{Cvoid}, CuPtr{T}, CuPtr{T}, Ptr{Cvoid}),
        plan, pointer(dest_nz), pointer(V), stream_ptr()), FactorizationException)
end

# ---------------------------------------------------------------------------------------------------------------------
# Optional: the vector passes of RichardsonIterator (src/LinearSolvers/backsolve.jl:36-52) as single launches.
#   b200_richardson_begin!(b, w, x, norms)   norms[3] = ||b||_inf ; x = 0 ; w = b
#   b200_richardson_update!(b, w, x, norms)  x += w ; w = b ; norms[1] = 0 ; norms[2] = ||x||_inf
# followed by mul!(w, kkt, x, -1, 1) through b2_condensed_kkt_mul_norm (accumulates ||w||_inf into norms[1]); one
# 24-byte D2H copy then carries the three norms of the stopping rule.
# ---------------------------------------------------------------------------------------------------------------------
function b200_richardson_begin!(b::CuVector{T}, w::CuVector{T}, x::CuVector{T}, norms::CuVector{T}) where T
    check(ccall((:b2_richardson_begin, libb200kkt), Cint, (Int64, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, Ptr{Cvoid}),
        length(b), pointer(b), pointer(w), pointer(x), pointer(norms, 3), stream_ptr()), SolveException)
end
function b200_richardson_update!(b::CuVector{T}, w::CuVector{T}, x::CuVector{T}, norms::CuVector{T}) where T
    check(ccall((:b2_richardson_update, libb200kkt), Cint, (Int64, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, Ptr{Cvoid}),
        length(b), pointer(b), pointer(w), pointer(x), pointer(norms), stream_ptr()), SolveException)
end

# inertia read split in two (queue more work behind factorize!, block once): b2_inertia_enqueue / b2_inertia_fetch
inertia_enqueue!(M::B200Solver) = check(ccall((:b2_inertia_enqueue, libb200kkt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), M.handle, stream_ptr()), FactorizationException)
function inertia_fetch(M::B200Solver)
    p = Ref{Int64}(0); z = Ref{Int64}(0); n = Ref{Int64}(0)
    check(ccall((:b2_inertia_fetch, libb200kkt), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), M.handle, p, z, n), FactorizationException)
    return (Int(p[]), Int(z[]), Int(n[]))
end

end # module
