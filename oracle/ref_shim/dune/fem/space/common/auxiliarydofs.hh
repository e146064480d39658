// Stand-in for dune/fem/space/common/auxiliarydofs.hh when oracle/_ref is built (the real header needs dune-grid): the Krylov loops
// of dune/fem/solver/linear/{cg,bicgstab,gmres}.hh -- the code under test -- only use forEachPrimaryDof
// (reference: dune/fem/space/common/auxiliarydofs.hh:302-315), restated here over a sorted list of auxiliary dofs terminated by the
// vector size.  ref_bind.cpp includes this file before the loops.
#ifndef B200FEM_REF_SHIM_AUXILIARYDOFS_HH
#define B200FEM_REF_SHIM_AUXILIARYDOFS_HH
#include <cstddef>
namespace Dune { namespace Fem {
template <class AuxiliaryDofs, class F>
static void forEachPrimaryDof(const AuxiliaryDofs& aux, F&& f) {
  const std::size_t na = aux.size();
  for (std::size_t a = 0, dof = 0; a < na; ++a, ++dof) { const std::size_t next = aux[a]; for (; dof < next; ++dof) f(dof); }
}
}}
#endif
