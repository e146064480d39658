// generic DG quadrature kernel, order 3
#include "launch_dgq.hpp"
namespace b200fem {
int launch_dg_quadrature_n4(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data, int mi, int ms) {
  return launch_dg_quadrature_n<4>(op, u, w, bvec, with_data, mi, ms);
}
}  // namespace b200fem
