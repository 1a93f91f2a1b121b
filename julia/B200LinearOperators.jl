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
const B2O_F32 = Cint(1)
b2o_dtype(::Type{Float64}) = B2O_F64
b2o_dtype(::Type{Float32}) = B2O_F32   # quasi-Newton operators (test/test_lbfgs.jl:162-178), dense / sparse matrix leaves

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

# mulOpEye! / mulOpOnes! / mulOpZeros! (src/special-operators.jl:36-44, 79-85, 102-108)
function opEye(nrow::Int, ncol::Int = nrow; c = ctx())
  prod! = (res, v, α, β) -> check(ccall((:b2o_eye_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, Int64, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    c.handle, B2O_F64, length(res), length(v), res, length(res), v, length(v), α, β))
  LinearOperator{Float64, CuVector{Float64}}(nrow, ncol, nrow == ncol, nrow == ncol, prod!, prod!, prod!)
end
function opOnes(nrow::Int, ncol::Int; c = ctx())
  prod! = (res, v, α, β) -> check(ccall((:b2o_ones_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, Int64, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    c.handle, B2O_F64, length(res), length(v), res, length(res), v, length(v), α, β))
  LinearOperator{Float64, CuVector{Float64}}(nrow, ncol, nrow == ncol, nrow == ncol, prod!, prod!, prod!)
end
function opZeros(nrow::Int, ncol::Int; c = ctx())
  prod! = (res, v, α, β) -> check(ccall((:b2o_zeros_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, Int64, Int64, CuPtr{Cvoid}, Int64, Int64, Cdouble, Cdouble),
    c.handle, B2O_F64, length(res), length(v), res, length(res), length(v), α, β))
  LinearOperator{Float64, CuVector{Float64}}(nrow, ncol, nrow == ncol, nrow == ncol, prod!, prod!, prod!)
end

# opRestriction / opExtension (src/special-operators.jl:167-221): the index set lives in a handle (duplicates resolved once,
# last occurrence wins); α, β are ignored exactly as the reference ignores them.
mutable struct IndexSet
  handle::Ptr{Cvoid}
end
function IndexSet(idx::AbstractVector{<:Integer}, ncol::Int; c = ctx())
  h = Ref{Ptr{Cvoid}}(C_NULL)
  idx64 = Vector{Int64}(idx)                                    # HOST array, 1-based as in Julia
  check(ccall((:b2o_index_create, libb2o), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int64, Ptr{Ptr{Cvoid}}),
    c.handle, idx64, length(idx64), ncol, h))
  ix = IndexSet(h[])
  finalizer(i -> ccall((:b2o_index_destroy, libb2o), Cint, (Ptr{Cvoid},), i.handle), ix)
end
function opRestriction(idx::AbstractVector{<:Integer}, ncol::Int; c = ctx())
  ix = IndexSet(idx, ncol; c = c)
  prod! = (res, v, α, β) -> check(ccall((:b2o_restrict_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64), ix.handle, B2O_F64, res, length(res), v, length(v)))
  tprod! = (res, u, α, β) -> check(ccall((:b2o_extend_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64), ix.handle, B2O_F64, res, length(res), u, length(u)))
  LinearOperator{Float64, CuVector{Float64}}(length(idx), ncol, false, false, prod!, tprod!, tprod!)
end
opExtension(idx::AbstractVector{<:Integer}, ncol::Int; kw...) = transpose(opRestriction(idx, ncol; kw...))

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

# LBFGSOperator(T, n; ...) src/lbfgs.jl:168: T = Float64 or Float32 (Float32: create / push! / mul! / reset!, see include/b2o.h)
B200LBFGSOperator(n::Int; kw...) = B200LBFGSOperator(Float64, n; kw...)
function B200LBFGSOperator(::Type{T}, n::Int; mem::Int = 5, scaling::Bool = true, damped::Bool = false, σ₂ = 0.99, σ₃ = 10.0,
                           inverse::Bool = false, c = ctx()) where {T <: Union{Float64, Float32}}
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_lbfgs_create, libb2o), Cint,
    (Ptr{Cvoid}, Cint, Int64, Cint, Cint, Cint, Cdouble, Cdouble, Cint, Ptr{Ptr{Cvoid}}),
    c.handle, b2o_dtype(T), n, mem, scaling, damped, σ₂, σ₃, inverse, h))
  handle = h[]
  prod! = (res, x, α, β) -> check(ccall((:b2o_qn_apply, libb2o), Cint,
    (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    handle, res, length(res), x, length(x), α, β))
  op = B200LBFGSOperator{T, typeof(prod!)}(n, n, true, true, prod!, prod!, prod!, inverse, handle, c, 0, 0, 0)
  finalizer(o -> ccall((:b2o_qn_destroy, libb2o), Cint, (Ptr{Cvoid},), o.handle), op)
end
B200InverseLBFGSOperator(n::Int; kw...) = B200LBFGSOperator(Float64, n; inverse = true, kw...)
B200InverseLBFGSOperator(::Type{T}, n::Int; kw...) where {T} = B200LBFGSOperator(T, n; inverse = true, kw...)

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


# ---- L-SR1 (src/lsr1.jl:86-113): same handle type, tprod!/ctprod! left `nothing` like the reference ---------------
mutable struct B200LSR1Operator{T, F} <: AbstractQuasiNewtonOperator{T}
  const nrow::Int
  const ncol::Int
  const symmetric::Bool
  const hermitian::Bool
  const prod!::F
  const tprod!::Nothing
  const ctprod!::Nothing
  handle::Ptr{Cvoid}
  ctx::Context
  nprod::Int
  ntprod::Int
  nctprod::Int
end
B200LSR1Operator(n::Int; kw...) = B200LSR1Operator(Float64, n; kw...)
function B200LSR1Operator(::Type{T}, n::Int; mem::Int = 5, scaling::Bool = true, c = ctx()) where {T <: Union{Float64, Float32}}
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_lsr1_create, libb2o), Cint, (Ptr{Cvoid}, Cint, Int64, Cint, Cint, Ptr{Ptr{Cvoid}}),
    c.handle, b2o_dtype(T), n, mem, scaling, h))
  handle = h[]
  prod! = (res, x, α, β) -> check(ccall((:b2o_qn_apply, libb2o), Cint,
    (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    handle, res, length(res), x, length(x), α, β))
  op = B200LSR1Operator{T, typeof(prod!)}(n, n, true, true, prod!, nothing, nothing, handle, c, 0, 0, 0)
  finalizer(o -> ccall((:b2o_qn_destroy, libb2o), Cint, (Ptr{Cvoid},), o.handle), op)
end
has_args5(::B200LSR1Operator) = true
isallocated5(::B200LSR1Operator) = true
storage_type(::B200LSR1Operator{T}) where {T} = CuVector{T}
const B200QN = Union{B200LBFGSOperator, B200LSR1Operator}
function push!(op::B200LSR1Operator, s::CuVector{Float64}, y::CuVector{Float64})
  acc = Ref{Cint}(0)
  check(ccall((:b2o_qn_push, libb2o), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ptr{Cint}),
              op.handle, s, y, length(s), acc))
  op.nprod += 1                                                   # the reference's push! runs mul!(ymBs, op, s, -1, 1) (src/lsr1.jl:125)
  return op
end
function reset!(op::B200LSR1Operator)
  check(ccall((:b2o_qn_reset, libb2o), Cint, (Ptr{Cvoid},), op.handle))
  op.nprod = op.ntprod = op.nctprod = 0
  return op
end
function LinearOperators.diag!(op::B200LSR1Operator, d::CuVector{Float64})
  check(ccall((:b2o_qn_diag, libb2o), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Int64), op.handle, d, length(d)))
  return d
end
diag(op::B200LSR1Operator) = LinearOperators.diag!(op, CuVector{Float64}(undef, op.nrow))

# mul! with HOST vectors: H2D, apply, D2H inside the library (pipelined in row chunks for the two-phase operators)
function LinearAlgebra.mul!(res::Vector{Float64}, op::B200QN, x::Vector{Float64}, α::Number, β::Number)
  (length(x) == op.ncol && length(res) == op.nrow) || throw(LinearOperatorException("shape mismatch"))
  op.nprod += 1
  check(ccall((:b2o_qn_apply_host, libb2o), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cdouble),
              op.handle, res, x, length(x), α, β))
  return res
end

# op.data access (checkpoint / resume, tests): columns by 0-based ring slot, which = 0 s, 1 y, 2 a, 3 b
function get_col!(dst::CuVector{Float64}, op::B200QN, which::Integer, k0::Integer)
  check(ccall((:b2o_qn_get_col, libb2o), Cint, (Ptr{Cvoid}, Cint, Cint, CuPtr{Cvoid}), op.handle, which, k0, dst))
  return dst
end
function set_col!(op::B200QN, which::Integer, k0::Integer, src::CuVector{Float64})
  check(ccall((:b2o_qn_set_col, libb2o), Cint, (Ptr{Cvoid}, Cint, Cint, CuPtr{Cvoid}), op.handle, which, k0, src))
  return op
end
function get_scalars(op::B200QN, mem::Int)
  ins = Ref{Cint}(0); γ = Ref{Cdouble}(0); ub = Ref{Cdouble}(0)
  ys = Vector{Float64}(undef, mem); aux = Vector{Float64}(undef, mem)
  check(ccall((:b2o_qn_get_scalars, libb2o), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
              op.handle, ins, γ, ub, ys, aux))
  return (insert = Int(ins[]), scaling_factor = γ[], opnorm_upper_bound = ub[], ys = ys, aux = aux)
end
function set_scalars!(op::B200QN, insert::Integer, γ::Float64, ub::Float64, ys::Vector{Float64}, aux::Vector{Float64})
  check(ccall((:b2o_qn_set_scalars, libb2o), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
              op.handle, insert, γ, ub, ys, aux))
  return op
end
# "inverse_mode" / "forward_mode" (compact representations, opt-in), "push_mode"
set_option!(op::B200QN, key::AbstractString, value::Integer) =
  (check(ccall((:b2o_qn_set_option, libb2o), Cint, (Ptr{Cvoid}, Cstring, Int64), op.handle, key, value)); op)
function apply_bytes(op::B200QN, β::Real = 0.0)
  out = Ref{Cdouble}(0)
  check(ccall((:b2o_qn_apply_bytes, libb2o), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}), op.handle, β, out))
  return out[]
end

# ---- diagonal quasi-Newton updates (src/DiagonalHessianApproximation.jl): the operator stays the reference's own
# DiagonalPSB / DiagonalAndrei / DiagonalBFGS / SpectralGradient over a CuVector `d`; only push! is replaced ---------------
const DIAGQN_KIND = Dict(:psb => Cint(0), :andrei => Cint(1), :bfgs => Cint(2), :spectral => Cint(3))
function diagqn_push!(d::CuVector{Float64}, s::CuVector{Float64}, y::CuVector{Float64}, kind::Symbol; c = ctx())
  rc = ccall((:b2o_diagqn_push, libb2o), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64),
             c.handle, DIAGQN_KIND[kind], d, length(d), s, y, length(s))
  rc == 5 && error(last_error())                                  # s == 0: the reference's error(...)
  check(rc)
  return d
end
push!(B::LinearOperators.DiagonalPSB{Float64, I, CuVector{Float64}}, s::CuVector{Float64}, y::CuVector{Float64}) where {I} = (diagqn_push!(B.d, s, y, :psb); B)
push!(B::LinearOperators.DiagonalAndrei{Float64, I, CuVector{Float64}}, s::CuVector{Float64}, y::CuVector{Float64}) where {I} = (diagqn_push!(B.d, s, y, :andrei); B)
push!(B::LinearOperators.DiagonalBFGS{Float64, I, CuVector{Float64}}, s::CuVector{Float64}, y::CuVector{Float64}) where {I} = (diagqn_push!(B.d, s, y, :bfgs); B)
push!(B::LinearOperators.SpectralGradient{Float64, I, CuVector{Float64}}, s::CuVector{Float64}, y::CuVector{Float64}) where {I} = (diagqn_push!(B.d, s, y, :spectral); B)

# ---- fused static operator trees (src/operations.jl:100-234 closure trees collapsed into ONE launch) ---------------------
# Build bottom-up with the node constructors, then `fuse` compiles; the result is an ordinary LinearOperator.
mutable struct Graph
  handle::Ptr{Cvoid}
  n::Int
  keep::Vector{Any}                                               # leaf vectors are aliased by the library: keep them alive
end
function Graph(n::Int; c = ctx())
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_graph_create, libb2o), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), c.handle, n, h))
  g = Graph(h[], n, Any[])
  finalizer(x -> ccall((:b2o_graph_destroy, libb2o), Cint, (Ptr{Cvoid},), x.handle), g)
end
const LEAF_KIND = Dict(:diagonal => Cint(0), :eye => Cint(1), :zeros => Cint(2), :ones => Cint(3), :householder => Cint(4))
function leaf!(g::Graph, kind::Symbol, vec::Union{Nothing, CuVector{Float64}} = nothing)
  node = Ref{Cint}(0)
  vec === nothing || push!(g.keep, vec)
  check(ccall((:b2o_graph_leaf, libb2o), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Ptr{Cint}),
              g.handle, LEAF_KIND[kind], vec === nothing ? CU_NULL : pointer(vec), node))
  return node[]
end
function unary!(g::Graph, kind::Symbol, child::Integer, x::Real = 0.0)      # :scale (op*x), :neg (-op), :transpose
  node = Ref{Cint}(0)
  k = kind === :scale ? Cint(12) : kind === :neg ? Cint(13) : Cint(14)
  check(ccall((:b2o_graph_unary, libb2o), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Ptr{Cint}), g.handle, k, child, x, node))
  return node[]
end
function binary!(g::Graph, kind::Symbol, a::Integer, b::Integer)            # :sum (op1+op2), :prod (op1*op2)
  node = Ref{Cint}(0)
  check(ccall((:b2o_graph_binary, libb2o), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cint}),
              g.handle, kind === :sum ? Cint(10) : Cint(11), a, b, node))
  return node[]
end
function fuse(g::Graph, root::Integer)
  check(ccall((:b2o_graph_compile, libb2o), Cint, (Ptr{Cvoid}, Cint), g.handle, root))
  run(trans) = (res, v, α, β) -> check(ccall((:b2o_graph_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble),
    g.handle, trans, res, length(res), v, length(v), α, β))
  LinearOperator{Float64, CuVector{Float64}}(g.n, g.n, false, false, run(Cint(0)), run(Cint(1)), run(Cint(1)))
end
function graph_info(g::Graph; transposed::Bool = false, β::Real = 0.0)
  np = Ref{Cint}(0); nr = Ref{Cint}(0); bytes = Ref{Cdouble}(0); jit = Ref{Cint}(0)
  check(ccall((:b2o_graph_info, libb2o), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cint}, Ptr{Cint}, Ptr{Cdouble}),
              g.handle, transposed, β, np, nr, bytes))
  check(ccall((:b2o_graph_uses_jit, libb2o), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cint}), g.handle, transposed, β, jit))
  return (passes = Int(np[]), reductions = Int(nr[]), alg_bytes = bytes[], jit = jit[] != 0)
end
# which executor ran the last apply of a variant (:interpreter, :nvrtc, :aot = compiled into libb2o) and the table key
function graph_variant(g::Graph; transposed::Bool = false, β::Real = 0.0)
  ex = Ref{Cint}(0); h = Ref{UInt64}(0)
  check(ccall((:b2o_graph_variant, libb2o), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cint}, Ptr{UInt64}), g.handle, transposed, β, ex, h))
  return (executor = (:interpreter, :nvrtc, :aot)[ex[] + 1], source_hash = h[])
end
function graph_jit_source(g::Graph; transposed::Bool = false, β::Real = 0.0)
  buf = Vector{UInt8}(undef, 1 << 16); len = Ref{Int64}(0)
  check(ccall((:b2o_graph_jit_source, libb2o), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{UInt8}, Int64, Ptr{Int64}),
              g.handle, transposed, β, buf, length(buf), len))
  return String(buf[1:len[]])
end
function graph_jit_check(g::Graph)
  cb = Ref{Int64}(0)
  check(ccall((:b2o_graph_jit_check, libb2o), Cint, (Ptr{Cvoid}, Ptr{Int64}), g.handle, cb))
  return cb[]
end
# BASELINE config 3 in one launch:
#   g = Graph(n); H = leaf!(g, :householder, h); D = leaf!(g, :diagonal, d); E = leaf!(g, :eye)
#   op = fuse(g, binary!(g, :sum, binary!(g, :prod, H, D), unary!(g, :scale, E, 0.1)))

# ---- kron(A, B) on the tensor cores (src/kron.jl:10-49): bf16 column-major CuMatrix operands -----------------------------
const BF16 = Core.BFloat16
function B200Kron(A::CuMatrix{BF16}, B::CuMatrix{BF16}; max_batch::Int = 1, c = ctx())
  m, n = size(A); p, q = size(B)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_kron_create, libb2o), Cint,
    (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, Int64, CuPtr{Cvoid}, Int64, Int64, Cint, Ptr{Ptr{Cvoid}}),
    c.handle, Cint(2), A, m, n, B, p, q, max_batch, h))
  handle = h[]
  dtype_of(res) = eltype(res) === Float32 ? Cint(1) : Cint(2)
  run(trans) = (res, x, α, β) -> check(ccall((:b2o_kron_apply, libb2o), Cint,
    (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Cint, Int64, CuPtr{Cvoid}, Int64, Cint, Cdouble, Cdouble),
    handle, trans, res, dtype_of(res), length(res), x, length(x), Cint(1), α, β))
  op = LinearOperator{BF16, CuVector{BF16}}(m * p, n * q, false, false, run(Cint(0)), run(Cint(1)), run(Cint(1)))
  finalizer(_ -> ccall((:b2o_kron_destroy, libb2o), Cint, (Ptr{Cvoid},), handle), op)
  return op, handle                                               # keep A and B alive as long as op is used (aliased)
end
Base.kron(A::CuMatrix{BF16}, B::CuMatrix{BF16}) = first(B200Kron(A, B))
function kron_flops(handle::Ptr{Cvoid}, nb::Integer = 1)
  out = Ref{Cdouble}(0)
  check(ccall((:b2o_kron_flops, libb2o), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), handle, nb, out))
  return out[]
end
kron_set_option!(handle::Ptr{Cvoid}, key::AbstractString, value::Integer) =
  check(ccall((:b2o_kron_set_option, libb2o), Cint, (Ptr{Cvoid}, Cstring, Int64), handle, key, value))
kron_launch_floor(grid::Integer, cluster::Integer, smem::Integer; c = ctx()) =
  check(ccall((:b2o_kron_launch_floor, libb2o), Cint, (Ptr{Cvoid}, Cint, Cint, Int64), c.handle, grid, cluster, smem))

# ---- roofline helpers ----------------------------------------------------------------------------------------------------
function dense_apply_bytes(handle::Ptr{Cvoid}, trans::Bool = false, β::Real = 0.0)
  out = Ref{Cdouble}(0)
  check(ccall((:b2o_dense_apply_bytes, libb2o), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cdouble}), handle, trans, β, out))
  return out[]
end
function sparse_apply_bytes(handle::Ptr{Cvoid}, trans::Bool = false, β::Real = 0.0)
  out = Ref{Cdouble}(0)
  check(ccall((:b2o_sparse_apply_bytes, libb2o), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cdouble}), handle, trans, β, out))
  return out[]
end

# ---- context plumbing ----------------------------------------------------------------------------------------------------
version() = ccall((:b2o_version, libb2o), Cint, ())
sync(c::Context = ctx()) = check(ccall((:b2o_ctx_sync, libb2o), Cint, (Ptr{Cvoid},), c.handle))
set_option!(c::Context, key::AbstractString, value::Integer) =
  (check(ccall((:b2o_ctx_set_option, libb2o), Cint, (Ptr{Cvoid}, Cstring, Int64), c.handle, key, value)); c)
function debug_read(c::Context, offset::Integer, count::Integer)
  out = Vector{Float64}(undef, count)
  check(ccall((:b2o_ctx_debug_read, libb2o), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}), c.handle, offset, count, out))
  return out
end
function launch_count(c::Context = ctx())
  n = Ref{Int64}(0)
  check(ccall((:b2o_ctx_launch_count, libb2o), Cint, (Ptr{Cvoid}, Ptr{Int64}), c.handle, n))
  return n[]
end
function kernel_time(c::Context = ctx(); reset::Bool = false)
  ms = Ref{Cdouble}(0); n = Ref{Int64}(0)
  check(ccall((:b2o_ctx_kernel_time, libb2o), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Int64}), c.handle, reset, ms, n))
  return ms[], n[]
end
function numa_node(c::Context = ctx())
  node = Ref{Cint}(-1)
  check(ccall((:b2o_ctx_numa_node, libb2o), Cint, (Ptr{Cvoid}, Ptr{Cint}), c.handle, node))
  return Int(node[])
end
# raw memory helpers (a Julia caller has CuArray; these exist for hosts without a CUDA array library)
function device_malloc(bytes::Integer; c = ctx())
  p = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_malloc, libb2o), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}), c.handle, bytes, p))
  return p[]
end
device_free(p::Ptr{Cvoid}; c = ctx()) = check(ccall((:b2o_free, libb2o), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c.handle, p))
function host_alloc(bytes::Integer; c = ctx())                   # pinned, on the GPU's own NUMA node when known
  p = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:b2o_host_alloc, libb2o), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}), c.handle, bytes, p))
  return p[]
end
host_free(p::Ptr{Cvoid}; c = ctx()) = check(ccall((:b2o_host_free, libb2o), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c.handle, p))
memcpy_h2d(dst::Ptr{Cvoid}, src::Ptr{Cvoid}, bytes::Integer; c = ctx()) =
  check(ccall((:b2o_memcpy_h2d, libb2o), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), c.handle, dst, src, bytes))
memcpy_d2h(dst::Ptr{Cvoid}, src::Ptr{Cvoid}, bytes::Integer; c = ctx()) =
  check(ccall((:b2o_memcpy_d2h, libb2o), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), c.handle, dst, src, bytes))
memset_zero(p::Ptr{Cvoid}, bytes::Integer; c = ctx()) =
  check(ccall((:b2o_memset_zero, libb2o), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), c.handle, p, bytes))
# x[i] = lo + (hi - lo) * u(seed, i): the counter-based generator shared bit for bit with the CPU oracle
function fill_uniform!(x::CuVector{Float64}, seed::Integer, lo::Real = 0.0, hi::Real = 1.0; c = ctx())
  check(ccall((:b2o_fill_uniform, libb2o), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, Int64, UInt64, Cdouble, Cdouble),
              c.handle, B2O_F64, x, length(x), seed, lo, hi))
  return x
end
function device_dot(a::CuVector{Float64}, b::CuVector{Float64}; c = ctx())
  out = Ref{Cdouble}(0)
  check(ccall((:b2o_dot, libb2o), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ptr{Cdouble}),
              c.handle, B2O_F64, a, b, length(a), out))
  return out[]
end

# ---- ComplexF64 leaves and the conj-sandwich primitives (src/adjtrans.jl:128-136, src/special-operators.jl:140, src/linalg.jl:79)
const CVec = CuVector{ComplexF64}
function opDiagonal(d::CVec; c = ctx())
  n = length(d)
  run(conj_d) = (res, v, α, β) -> check(ccall((:b2o_cdiag_apply, libb2o), Cint,
    (Ptr{Cvoid}, Int64, Int64, CuPtr{Cvoid}, Int64, Cint, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, Cdouble),
    c.handle, length(res), length(v), d, n, conj_d, res, length(res), v, length(v), real(α), imag(α), real(β), imag(β)))
  LinearOperator{ComplexF64, CVec}(n, n, true, isreal(d), run(Cint(0)), run(Cint(0)), run(Cint(1)))
end
function opHouseholder(h::CVec; c = ctx())
  n = length(h)
  prod! = (res, v, α, β) -> check(ccall((:b2o_chouseholder_apply, libb2o), Cint,
    (Ptr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, Cdouble),
    c.handle, n, h, res, length(res), v, length(v), real(α), imag(α), real(β), imag(β)))
  LinearOperator{ComplexF64, CVec}(n, n, isreal(h), true, prod!, nothing, prod!)
end
function opEye(::Type{ComplexF64}, nrow::Int, ncol::Int = nrow; c = ctx())
  prod! = (res, v, α, β) -> check(ccall((:b2o_ceye_apply, libb2o), Cint,
    (Ptr{Cvoid}, Int64, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, Cdouble),
    c.handle, length(res), length(v), res, length(res), v, length(v), real(α), imag(α), real(β), imag(β)))
  LinearOperator{ComplexF64, CVec}(nrow, ncol, nrow == ncol, nrow == ncol, prod!, prod!, prod!)
end
function opZeros(::Type{ComplexF64}, nrow::Int, ncol::Int; c = ctx())
  prod! = (res, v, α, β) -> check(ccall((:b2o_czeros_apply, libb2o), Cint,
    (Ptr{Cvoid}, Int64, Int64, CuPtr{Cvoid}, Int64, Int64, Cdouble, Cdouble),
    c.handle, length(res), length(v), res, length(res), length(v), real(β), imag(β)))
  LinearOperator{ComplexF64, CVec}(nrow, ncol, nrow == ncol, nrow == ncol, prod!, prod!, prod!)
end
# conj!(res) and conj.(v) of the wrappers' generic mul! then run on the device through CUDA.jl broadcasts; the library's own
# kernel is available as conj_device!(dst, src) (dst === src is conj!)
conj_device!(dst::CVec, src::CVec; c = ctx()) =
  (check(ccall((:b2o_conj, libb2o), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64), c.handle, dst, src, length(src))); dst)

# ---- row-partitioned multi-GPU: one Julia process per GPU (MPI.jl or Distributed for the rendezvous) --------------------
# After comm_init! every vector argument is the calling rank's contiguous row slab and every inner product is all-reduced.
function comm_unique_id()
  id = Vector{UInt8}(undef, 128)
  check(ccall((:b2o_comm_unique_id, libb2o), Cint, (Ptr{UInt8},), id))
  return id                                                       # broadcast from rank 0 (e.g. MPI.Bcast!)
end
comm_init!(c::Context, id::Vector{UInt8}, nranks::Integer, rank::Integer) =
  check(ccall((:b2o_comm_init, libb2o), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), c.handle, id, nranks, rank))
comm_destroy!(c::Context) = check(ccall((:b2o_comm_destroy, libb2o), Cint, (Ptr{Cvoid},), c.handle))
# NVLink peer mailbox: all-gather the 64-byte IPC handles (rank order), then connect; the applies then all-reduce their dots
# INSIDE the persistent kernel.  set_option!(c, "use_mailbox", 0 | 1) switches between the mailbox and NCCL afterwards.
function mbox_local_handle(c::Context)
  hdl = Vector{UInt8}(undef, 64)
  check(ccall((:b2o_mbox_local_handle, libb2o), Cint, (Ptr{Cvoid}, Ptr{UInt8}), c.handle, hdl))
  return hdl
end
mbox_connect!(c::Context, handles::Vector{UInt8}, nranks::Integer, rank::Integer) =
  check(ccall((:b2o_mbox_connect, libb2o), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), c.handle, handles, nranks, rank))
mbox_disconnect!(c::Context) = check(ccall((:b2o_mbox_disconnect, libb2o), Cint, (Ptr{Cvoid},), c.handle))

end # module
