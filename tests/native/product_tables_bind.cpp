// tests/native/product_tables_bind.cpp -- TEST INFRASTRUCTURE.  C bindings around the PRODUCT's own host tables
// (dune_fem_b200/csrc/tables.hpp: the Gauss rules, the rule selection, the Legendre / Lagrange 1-D bases and the local numbering of the
// DG spaces -- the header capi.cu / launch_*.cu build the device tables from), so that tests/test_product_tables.py can check them on
// the CPU against the vectors the compiled reference produced (tests/golden/reference_pieces.json) without going through the oracle.
// Built by the test with g++ (host code only, no CUDA).
#include <cstdint>
#include "../../dune_fem_b200/csrc/tables.hpp"

extern "C" {
int pt_gauss_points_for_order(int order) { try { return b200fem::gauss_points_for_order(order); } catch (...) { return -1; } }
void pt_gauss_rule(int m, double* x, double* w) { const b200fem::Rule1D r = b200fem::gauss_rule(m); for (int i = 0; i < m; ++i) { x[i] = r.x[i]; w[i] = r.w[i]; } }
// what the kernels read: values B[q*n+i], derivatives G[q*n+i] of the n = order+1 functions at the m rule points
void pt_tabulate_1d(int legendre, int order, int m, double* B, double* G) {
  const b200fem::Tab1D t = b200fem::tabulate_1d(legendre ? b200fem::Basis::Legendre : b200fem::Basis::Lagrange, order, m);
  for (std::size_t i = 0; i < t.B.size(); ++i) { B[i] = t.B[i]; G[i] = t.G[i]; }
}
double pt_basis_1d(int legendre, int order, int i, double x, int derivative) {
  static const b200fem::Legendre1D leg;
  if (legendre) return derivative ? leg.derivative(i, x) : leg.value(i, x);
  return derivative ? b200fem::lagrange_derivative(order, i, x) : b200fem::lagrange_value(order, i, x);
}
// map[(m0*n + m1)*n + m2] = stored local index or -1; returns the number of local dofs
int pt_dg_tensor_map(int dim, int order, int kind, int32_t* map) {
  int nb = 0; const std::vector<int> m = b200fem::dg_tensor_map(dim, order, kind, &nb);
  for (std::size_t i = 0; i < m.size(); ++i) map[i] = m[i];
  return nb;
}
}

// ---- the 1-D operator matrices of the product's Kronecker form (dune_fem_b200/csrc/kron_tables.hpp: what the marching / slab /
// tensor-core kernels multiply with), for a DG Legendre space of the given order on cells of size h.
// params = eps, b0, b1, b2, c, beta ; out = 5 blocks (S, Dlo, Dhi, L, R) of 3 axes of n*n doubles (row = test, column = trial function)
#include "../../dune_fem_b200/csrc/kron_tables.hpp"
extern "C" int pt_kron_tables(int dim, int order, const double* h, const double* params, int dirichlet_mask, int has_skeleton, int has_boundary,
                              double scale, double* out) {
  b200fem_model m{}; m.eps = params[0]; m.b[0] = params[1]; m.b[1] = params[2]; m.b[2] = params[3]; m.c = params[4]; m.gamma = 0; m.beta = params[5];
  m.dirichlet_mask = dirichlet_mask; m.data = 0; m.has_skeleton = has_skeleton; m.has_boundary = has_boundary; m.strong_dirichlet = 0;
  const b200fem::Tab1D t = b200fem::tabulate_1d(b200fem::Basis::Legendre, order, b200fem::gauss_points_for_order(2 * order));
  const b200fem::KronHost k = b200fem::build_kron_tables(t, m, dim, h, scale);
  const int nn = k.n * k.n;
  const std::vector<double>* blocks[5] = {k.S, k.Dlo, k.Dhi, k.L, k.R};
  for (int b = 0; b < 5; ++b) for (int d = 0; d < 3; ++d) for (int i = 0; i < nn; ++i) out[(b * 3 + d) * nn + i] = blocks[b][d][i];
  return k.n;
}

// ---- the banded 1-D row tables of the product's Lagrange Kronecker form (kron_tables.hpp: build_lagrange_rows; lagrange_kronecker.cuh /
// the lattice kernel's stencil is the same operator by node type).  out = M[0], M[1], M[2], T[0], T[1], T[2], each L_d * (2k+1) doubles
// with L_d = k n_d + 1 (1 for axes >= dim); returns the total number of doubles written (call with out = NULL for the size).
extern "C" long long pt_lagrange_rows(int dim, int order, const int* n, const double* h, const double* params, int dirichlet_mask, int has_boundary, double* out) {
  b200fem_model m{}; m.eps = params[0]; m.b[0] = params[1]; m.b[1] = params[2]; m.b[2] = params[3]; m.c = params[4]; m.gamma = 0; m.beta = params[5];
  m.dirichlet_mask = dirichlet_mask; m.has_skeleton = 0; m.has_boundary = has_boundary;
  const b200fem::Tab1D t = b200fem::tabulate_1d(b200fem::Basis::Lagrange, order, b200fem::gauss_points_for_order(2 * order));
  const int origin[3] = {0, 0, 0};
  const b200fem::LagRowsHost r = b200fem::build_lagrange_rows(t, m, dim, order, n, origin, n, h);
  long long k = 0;
  for (int d = 0; d < 3; ++d) for (double v : r.M[d]) { if (out) out[k] = v; ++k; }
  for (int d = 0; d < 3; ++d) for (double v : r.T[d]) { if (out) out[k] = v; ++k; }
  return k;
}

// the same operators in the by-node-type form the lattice kernel reads (kron_tables.hpp: build_lagrange_stencil; orders 1, 2).
// out = M[3][2][5], T[3][2][5], Mlo[3], Mhi[3], Tlo[3], Thi[3] (72 doubles)
extern "C" void pt_lagrange_stencil(int dim, int order, const int* n, const double* h, const double* params, int dirichlet_mask, int has_boundary, double* out) {
  b200fem_model m{}; m.eps = params[0]; m.b[0] = params[1]; m.b[1] = params[2]; m.b[2] = params[3]; m.c = params[4]; m.gamma = 0; m.beta = params[5];
  m.dirichlet_mask = dirichlet_mask; m.has_skeleton = 0; m.has_boundary = has_boundary;
  const b200fem::Tab1D t = b200fem::tabulate_1d(b200fem::Basis::Lagrange, order, b200fem::gauss_points_for_order(2 * order));
  const int origin[3] = {0, 0, 0};
  const b200fem::LagStencilHost s = b200fem::build_lagrange_stencil(t, m, dim, order, n, origin, n, h);
  int k = 0;
  for (int d = 0; d < 3; ++d) for (int ty = 0; ty < 2; ++ty) for (int j = 0; j < 5; ++j) out[k++] = s.M[d][ty][j];
  for (int d = 0; d < 3; ++d) for (int ty = 0; ty < 2; ++ty) for (int j = 0; j < 5; ++j) out[k++] = s.T[d][ty][j];
  for (int d = 0; d < 3; ++d) out[k++] = s.Mlo[d];
  for (int d = 0; d < 3; ++d) out[k++] = s.Mhi[d];
  for (int d = 0; d < 3; ++d) out[k++] = s.Tlo[d];
  for (int d = 0; d < 3; ++d) out[k++] = s.Thi[d];
}
