// generic DG quadrature kernel, orders 4 and 5
#include "launch_dgq.hpp"
namespace b200fem {
int launch_dg_quadrature_n56(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data, int mi, int ms) {
  return op->sp->n1 == 5 ? launch_dg_quadrature_n<5>(op, u, w, bvec, with_data, mi, ms) : launch_dg_quadrature_n<6>(op, u, w, bvec, with_data, mi, ms);
}
}  // namespace b200fem
