// generic DG quadrature kernel, orders 1 and 2
#include "launch_dgq.hpp"
namespace b200fem {
int launch_dg_quadrature_n23(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data, int mi, int ms) {
  return op->sp->n1 == 2 ? launch_dg_quadrature_n<2>(op, u, w, bvec, with_data, mi, ms) : launch_dg_quadrature_n<3>(op, u, w, bvec, with_data, mi, ms);
}
}  // namespace b200fem
