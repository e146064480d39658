// Shadows dune/fem/solver/parameter.hh when oracle/_ref is built: the real header drags in the whole parameter/IO stack
// (dune-common), while the Krylov loops of dune/fem/solver/linear/{cg,bicgstab,gmres}.hh -- the code under test -- only
// use the ToleranceCriteria constants (reference: dune/fem/solver/parameter.hh:14-18) and forEachPrimaryDof
// (reference: dune/fem/space/common/auxiliarydofs.hh:302-315; restated here because that header needs dune-grid).
#ifndef B200FEM_REF_SHIM_SOLVERPARAMETER_HH
#define B200FEM_REF_SHIM_SOLVERPARAMETER_HH
#include <cstddef>
namespace Dune { namespace Fem {
namespace LinearSolver {
struct ToleranceCriteria { static const int absolute = 0; static const int relative = 1; static const int residualReduction = 2; };
}
template <class AuxiliaryDofs, class F>
static void forEachPrimaryDof(const AuxiliaryDofs& aux, F&& f) {
  const std::size_t na = aux.size();
  for (std::size_t a = 0, dof = 0; a < na; ++a, ++dof) { const std::size_t next = aux[a]; for (; dof < next; ++dof) f(dof); }
}
}}
#endif
