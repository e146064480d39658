// Stand-in for dune/fem/solver/inverseoperatorinterface.hh when oracle/_ref is built (the real header needs
// function/common/discretefunction.hh): dune/fem/solver/newtoninverseoperator.hh uses only the plain record Impl::SolverInfo from it
// (reference: inverseoperatorinterface.hh:13-31, returned by info()); its linear inverse operator is a template argument.
#ifndef B200FEM_REF_SHIM_INVERSEOPERATORINTERFACE_HH
#define B200FEM_REF_SHIM_INVERSEOPERATORINTERFACE_HH
#include <vector>
namespace Dune { namespace Fem { namespace Impl {
struct SolverInfo {
  SolverInfo(bool c, int l, int nl, const std::vector<double>& t) : converged(c), linearIterations(l), nonlinearIterations(nl), timing(t) {}
  bool converged; int linearIterations; int nonlinearIterations; std::vector<double> timing;
};
}}}
#endif
