// Stand-in for dune/fem/solver/diagonalpreconditioner.hh when oracle/_ref is built (the real header needs dune-istl and the discrete
// function stack): solver/cginverseoperator.hh names DiagonalPreconditioner only for ASSEMBLED operators (:581-585); the matrix-free
// operators bound here never take that branch, so a declaration suffices.
#ifndef B200FEM_REF_SHIM_DIAGONALPRECONDITIONER_HH
#define B200FEM_REF_SHIM_DIAGONALPRECONDITIONER_HH
namespace Dune { namespace Fem {
template <class DiscreteFunction, class LinearOperator> class DiagonalPreconditioner;
}}
#endif
