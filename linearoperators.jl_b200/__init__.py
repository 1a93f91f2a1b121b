"""linearoperators.jl_b200 -- B200-native operator-apply engine with the operator surface of
JuliaSmoothOptimizers/LinearOperators.jl (mul!, prod!/tprod!/ctprod!, +, *, adjoint/transpose, hcat/vcat,
BlockDiagonalOperator, push!/reset!) over hand-written sm_100a kernels behind the C ABI in include/b2o.h.

The directory name contains a dot, so import it through the top-level shim:  `import linearoperators_jl_b200 as lo`.
"""
from ._lib import B2OError, ErrorException, LinearOperatorException  # noqa: F401
from .abstract import (AbstractLinearOperator, AdjointLinearOperator, ConjugateLinearOperator, Hermitian,  # noqa: F401
                       LinearOperator, Matrix, Storage, Symmetric, TransposeLinearOperator, adjoint, apply, conj, eltype,
                       has_args5, isallocated5, ishermitian, issymmetric, mul_, nctprod, nprod, ntprod, reset_, size,
                       storage_type, transpose)
from .cat import hcat, hvcat, vcat  # noqa: F401
from .constructors import DenseMatrixOperator, SparseMatrixOperator  # noqa: F401
from .context import Context, default_context  # noqa: F401
from .diagqn import (AbstractDiagonalQuasiNewtonOperator, DiagonalAndrei, DiagonalBFGS, DiagonalPSB, ShiftedOperator,  # noqa: F401
                     SpectralGradient)
from .graph import FusedOperator, fuse  # noqa: F401
from .kron import KronOperator, kron  # noqa: F401
from .qn import (InverseLBFGSOperator, LBFGSOperator, LSR1Operator, diag, diag_, ldiv_, push_, solve_shifted_system_)  # noqa: F401
from .special_operators import (BlockDiagonalOperator, LocalBlockOfDiagonal, getindex, opDiagonal, opExtension, opEye, opHouseholder,  # noqa: F401
                                opOnes, opRestriction, opZeros)

__version__ = "0.1.0"
from .utilities import check_ctranspose, check_hermitian, check_positive_definite, normest  # noqa: F401,E402
