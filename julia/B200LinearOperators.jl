# B200LinearOperators.jl -- the binding a LinearOperators.jl maintainer would add (e.g. as
# ext/LinearOperatorsB200Ext.jl).  UNTESTED HERE: Julia is not installed in the build image or on the GPU
# box; the same C ABI is exercised by the ctypes binding (linearoperators.jl_b200/_lib.py) in tests/.
#
# Each closure body is one `ccall` into libb2o.so (include/b2o.h).  The operator types keep the reference's
# field names (src/abstract.jl:46-59, src/lbfgs.jl:62-75) so Krylov.jl-style callers (mul!, size, eltype,
# adjoint/transpose) work unchanged.
module B200LinearOperators

using LinearOperators, CUDA, LinearAlgebra, SparseArrays
import LinearOperators: AbstractQuasiNewtonOperator, LinearOperatorException, storage_type, has_args5, isallocated5, reset!
import Base: push!
import LinearAlgebra: diag

const libb2o = get(ENV, "LIBB2O", "libb2o.so")
const B2O_F64 = Cint(0)

last_error() = unsafe_string(ccall((:b2o_last_error, libb2o), Cstring, ()))
function check(rc::Cint)
  rc == 0 && return nothing
  msg = last_error()
  rc == 1 && throw(LinearOperatorException(msg))                 # B2O_ESHAPE
  rc == 5 && (startswith(msg, "only the diagonal") ? throw(LinearOperatorException(msg)) : error(msg))  # B2O_ESTATE
  (rc == 2 && startswith(msg, "indices should be")) && throw(LinearOperatorException(msg))
  error("libb2o status $rc: $msg")
end

mutable struct Context
  handle::Ptr{Cvoid}
end
function Context(device::Integer = CUDA.deviceid(CUDA.device()); stream = CUDA.stream())
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_ctx_create, libb2o), Cint, (Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), device, stream.handle, h))
  ctx = Context(h[])
  finalizer(c -> ccall((:b2o_ctx_destroy, libb2o), Cint, (Ptr{Cvoid},), c.handle), ctx)
end
const default_ctx = Ref{Union{Nothing, Context}}(nothing)
ctx() = something(default_ctx[], (default_ctx[] = Context()))

# ---- leaf operators: drop-in closures for LinearOperator{T,S}(…) with S = CuVector{Float64} ------------------
function opDiagonal(d::CuVector{Float64}; c = ctx())
  n = length(d)
  prod! = (res, v, α, β) -> check(ccall((:b2o_diag_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, Int64, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    c.handle, B2O_F64, length(res), length(v), d, n, res, length(res), v, length(v), α, β))
  LinearOperator{Float64, CuVector{Float64}}(n, n, true, true, prod!, prod!, prod!)
end

function opHouseholder(h::CuVector{Float64}; c = ctx())
  n = length(h)
  prod! = (res, v, α, β) -> check(ccall((:b2o_householder_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    c.handle, B2O_F64, n, h, res, length(res), v, length(v), α, β))
  LinearOperator{Float64, CuVector{Float64}}(n, n, true, true, prod!, nothing, prod!)
end
# opEye / opOnes / opZeros / opRestriction / opExtension follow the same pattern with
# b2o_eye_apply / b2o_ones_apply / b2o_zeros_apply / b2o_index_create + b2o_restrict_apply / b2o_extend_apply.

# ---- LinearOperator(M) for a dense CuMatrix (src/constructors.jl:15-29): Float64 and Float32 ------------------
# The closures alias M like the reference's; the handle only owns the partial-sum workspace of split products.
const B2O_DTYPE = Dict(Float64 => Cint(0), Float32 => Cint(1))
function B200DenseOperator(M::CuMatrix{T}; symmetric = false, hermitian = false, c = ctx()) where {T <: Union{Float64, Float32}}
  nrow, ncol = size(M)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_dense_create, libb2o), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, Int64, Int64, Ptr{Ptr{Cvoid}}),
    c.handle, B2O_DTYPE[T], M, nrow, ncol, max(1, stride(M, 2)), h))
  handle = h[]
  run(trans) = (res, v, α, β) -> check(ccall((:b2o_dense_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    handle, trans, res, length(res), v, length(v), α, β))
  op = LinearOperator{T, CuVector{T}}(nrow, ncol, symmetric, hermitian, run(Cint(0)), run(Cint(1)), run(Cint(1)))
  finalizer(_ -> ccall((:b2o_dense_destroy, libb2o), Cint, (Ptr{Cvoid},), handle), op)
  return op
end
# ---- LinearOperator(M) for a sparse matrix (src/constructors.jl:3-5,15-29): colptr / rowval stay on the host (1-based, as
# SparseMatrixCSC stores them), nzval lives on the device and is aliased; call refresh!(op) after changing nzval in place.
function B200SparseOperator(M::SparseMatrixCSC{T, Int64}, nzval::CuVector{T}; symmetric = false, hermitian = false,
                            c = ctx()) where {T <: Union{Float64, Float32}}
  nrow, ncol = size(M)
  length(nzval) == nnz(M) || throw(LinearOperatorException("nzval must hold nnz(M) device values"))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_sparse_create, libb2o), Cint,
    (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, CuPtr{Cvoid}, Ptr{Ptr{Cvoid}}),
    c.handle, B2O_DTYPE[T], Cint(0), nrow, ncol, nnz(M), M.colptr, M.rowval, nzval, h))
  handle = h[]
  run(trans) = (res, v, α, β) -> check(ccall((:b2o_sparse_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    handle, trans, res, length(res), v, length(v), α, β))
  op = LinearOperator{T, CuVector{T}}(nrow, ncol, symmetric, hermitian, run(Cint(0)), run(Cint(1)), run(Cint(1)))
  finalizer(_ -> ccall((:b2o_sparse_destroy, libb2o), Cint, (Ptr{Cvoid},), handle), op)
  return op, () -> check(ccall((:b2o_sparse_refresh, libb2o), Cint, (Ptr{Cvoid},), handle))
end
B200SparseOperator(M::SparseMatrixCSC{T, Int64}; kw...) where {T} = B200SparseOperator(M, CuVector{T}(M.nzval); kw...)

# BlockDiagonalOperator(A, B, C) of CuMatrix blocks (test/gpu/nvidia.jl:8-15): wrap the blocks first, the reference's
# own block loop (src/special-operators.jl:258-267) then calls the closures above on the slab views.
B200BlockDiagonalOperator(Ms::CuMatrix...; kw...) = BlockDiagonalOperator((B200DenseOperator(M; kw...) for M in Ms)...)

# ---- quasi-Newton operators ----------------------------------------------------------------------------------
mutable struct B200LBFGSOperator{T, F} <: AbstractQuasiNewtonOperator{T}
  const nrow::Int
  const ncol::Int
  const symmetric::Bool
  const hermitian::Bool
  const prod!::F
  const tprod!::F
  const ctprod!::F
  const inverse::Bool
  handle::Ptr{Cvoid}
  ctx::Context
  nprod::Int
  ntprod::Int
  nctprod::Int
end

function B200LBFGSOperator(n::Int; mem::Int = 5, scaling::Bool = true, damped::Bool = false, σ₂ = 0.99, σ₃ = 10.0,
                           inverse::Bool = false, c = ctx())
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_lbfgs_create, libb2o), Cint,
    (Ptr{Cvoid}, Cint, Int64, Cint, Cint, Cint, Cdouble, Cdouble, Cint, Ptr{Ptr{Cvoid}}),
    c.handle, B2O_F64, n, mem, scaling, damped, σ₂, σ₃, inverse, h))
  handle = h[]
  prod! = (res, x, α, β) -> check(ccall((:b2o_qn_apply, libb2o), Cint,
    (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    handle, res, length(res), x, length(x), α, β))
  op = B200LBFGSOperator{Float64, typeof(prod!)}(n, n, true, true, prod!, prod!, prod!, inverse, handle, c, 0, 0, 0)
  finalizer(o -> ccall((:b2o_qn_destroy, libb2o), Cint, (Ptr{Cvoid},), o.handle), op)
end
B200InverseLBFGSOperator(n::Int; kw...) = B200LBFGSOperator(n; inverse = true, kw...)

has_args5(::B200LBFGSOperator) = true
isallocated5(::B200LBFGSOperator) = true
storage_type(::B200LBFGSOperator{T}) where {T} = CuVector{T}

function push!(op::B200LBFGSOperator, s::CuVector{Float64}, y::CuVector{Float64})
  acc = Ref{Cint}(0)
  check(ccall((:b2o_qn_push, libb2o), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ptr{Cint}),
              op.handle, s, y, length(s), acc))
  return op
end
function push!(op::B200LBFGSOperator, s::CuVector{Float64}, y::CuVector{Float64}, Bs::CuVector{Float64})
  acc = Ref{Cint}(0)
  check(ccall((:b2o_lbfgs_push_damped_fwd, libb2o), Cint,
              (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ptr{Cint}), op.handle, s, y, Bs, length(s), acc))
  return op
end
function push!(op::B200LBFGSOperator, s::CuVector{Float64}, y::CuVector{Float64}, α::Float64, g::CuVector{Float64},
               Bs::CuVector{Float64} = similar(g))
  acc = Ref{Cint}(0)
  check(ccall((:b2o_lbfgs_push_damped_inv, libb2o), Cint,
              (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cdouble, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ptr{Cint}),
              op.handle, s, y, α, g, Bs, length(s), acc))
  return op
end
# mul!(Res, op, X, α, β) with matrices (src/operations.jl:34-36): the reference hands the matrices to prod!; here the block
# kernel streams every state column once per 8 right-hand sides (b2o_qn_apply_multi)
function LinearAlgebra.mul!(res::CuMatrix{Float64}, op::B200LBFGSOperator, X::CuMatrix{Float64}, α, β)
  (size(X, 1) == op.ncol && size(res, 1) == op.nrow && size(X, 2) == size(res, 2)) || throw(LinearOperatorException("shape mismatch"))
  check(ccall((:b2o_qn_apply_multi, libb2o), Cint,
              (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Int64, Cint, Cdouble, Cdouble),
              op.handle, res, stride(res, 2), X, stride(X, 2), size(X, 1), size(X, 2), α, β))
  return res
end
# solve_shifted_system!(x, B, b, σ) / ldiv!(x, B, b) (src/utilities.jl:207-289) on the device state
function LinearOperators.solve_shifted_system!(x::CuVector{Float64}, op::B200LBFGSOperator, b::CuVector{Float64}, σ::Float64)
  σ >= 0 || throw(ArgumentError("σ must be nonnegative"))
  check(ccall((:b2o_lbfgs_solve_shifted, libb2o), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble),
              op.handle, x, length(x), b, length(b), σ))
  return x
end
LinearAlgebra.ldiv!(x::CuVector{Float64}, op::B200LBFGSOperator, b::CuVector{Float64}) = LinearOperators.solve_shifted_system!(x, op, b, 0.0)
function reset!(op::B200LBFGSOperator)
  check(ccall((:b2o_qn_reset, libb2o), Cint, (Ptr{Cvoid},), op.handle))
  op.nprod = op.ntprod = op.nctprod = 0
  return op
end
function LinearOperators.diag!(op::B200LBFGSOperator, d::CuVector{Float64})
  check(ccall((:b2o_qn_diag, libb2o), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Int64), op.handle, d, length(d)))
  return d
end
diag(op::B200LBFGSOperator) = LinearOperators.diag!(op, CuVector{Float64}(undef, op.nrow))

end # module
