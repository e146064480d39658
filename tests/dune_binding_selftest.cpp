// Compiles dune/fem/schemes/b200galerkin.hh (the reference-side binding) against tests/dune_stub and runs it: an operator
// built from "generated" integrands (source text + constants), applied through Dune::Fem::Operator::operator(), and solved with
// B200KrylovInverseOperator.  Prints the results for tests/test_host_mirror.py to compare with the CPU oracle.
#include <cmath>
#include <cstdio>
#include <dune/grid/yaspgrid_stub.hh>
#include <dune/fem/schemes/b200galerkin.hh>

// what the UFL code generator would emit for  eps grad u . grad v + c u v - f v  with SIPG faces and weak Dirichlet data
struct GeneratedIntegrands {
  static constexpr bool hasSkeleton = true, hasBoundary = true;
  double eps = 0.5, c = 0.3, penalty = 80.0;
  bool nonlinear() const { return false; }
  void b200Constants(std::vector<double>& k) const { k = {eps, c, penalty}; }
  static const char* b200Source() {
    return "__device__ void interior(const double* x, const PointValue& u, PointRange& r, const double* c, int dim) {\n"
           "  r.s = c[1] * u.u - (1.0 + x[0] * x[1]); for (int d = 0; d < dim; ++d) r.F[d] = c[0] * u.du[d]; }\n"
           "__device__ void skeleton(const double* x, int axis, double sign, double ihe, const PointValue& in, const PointValue& out,\n"
           "                         PointRange& rin, PointRange& rout, const double* c, int dim) {\n"
           "  const double jump = in.u - out.u; const double cj = c[0] * c[2] * ihe * jump - c[0] * 0.5 * (in.du[axis] + out.du[axis]) * sign;\n"
           "  rin.s = cj; rout.s = -cj; rin.F[axis] = rout.F[axis] = -c[0] * jump * 0.5 * sign; }\n"
           "__device__ void boundary(const double* x, int axis, int side, double ihbnd, const PointValue& u, PointRange& r, const double* c, int dim) {\n"
           "  const double sign = side ? 1.0 : -1.0, g = sin(x[0] * x[1]);\n"
           "  r.s = c[0] * c[2] * ihbnd * (u.u - g) - c[0] * u.du[axis] * sign; r.F[axis] = -c[0] * (u.u - g) * sign; }\n";
  }
};

// a second "generated" form, interior terms only (continuous space): eps grad u . grad v + c u^3 v - (1 + x0) v
struct GeneratedInteriorIntegrands {
  static constexpr bool hasSkeleton = false, hasBoundary = false;
  bool nonlinear() const { return true; }
  void b200Constants(std::vector<double>& k) const { k = {0.7, 0.4}; }
  static const char* b200Source() {
    return "__device__ void interior(const double* x, const PointValue& u, PointRange& r, const double* c, int dim) {\n"
           "  r.s = c[1] * u.u * u.u * u.u - (1.0 + x[0]); for (int d = 0; d < dim; ++d) r.F[d] = c[0] * u.du[d]; }\n";
  }
};

// unstructured path of the binding: a 3 x 3 patch of distorted quadrilaterals handed over by a (stub) ALUGrid-like grid part
static int unstructuredPart() {
  using namespace Dune; using namespace Dune::Fem;
  typedef StubCubeGrid<2> Grid; typedef StubLeafGridPart<Grid> GridPart; typedef StubLagrangeSpace<GridPart> Space; typedef StubDiscreteFunction<Space> DF;
  const int n = 3; std::vector<std::array<double, 2>> vx; std::vector<std::array<int, 4>> cubes;
  for (int j = 0; j <= n; ++j) for (int i = 0; i <= n; ++i) { const double x = double(i) / n, y = double(j) / n; vx.push_back({x + 0.05 * std::sin(5.0 * y) * x * (1 - x), y + 0.04 * std::sin(4.0 * x) * y * (1 - y)}); }
  for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) { const int v = i + (n + 1) * j; cubes.push_back({v, v + 1, v + n + 1, v + n + 2}); }
  Grid grid(vx, cubes); GridPart gridPart(grid);
  Space space(gridPart, 2, std::size_t((2 * n + 1) * (2 * n + 1)));            // P2 on 3 x 3 quadrilaterals: 49 nodes
  B200GalerkinOperator<GeneratedInteriorIntegrands, DF> op(space, space);
  DF u(space), w(space);
  for (std::size_t i = 0; i < u.dofVector().size(); ++i) u.dofVector()[i] = std::sin(0.37 * double(i));
  op(u, w);
  double s = 0; for (double v : w.dofVector()) s += v * v;
  std::printf("unstructured_apply_norm2 %.17g\n", s);
  for (std::size_t i = 0; i < 4; ++i) std::printf("uw%zu %.17g\n", i, w.dofVector()[i]);
  return 0;
}

int main() {
  using namespace Dune; using namespace Dune::Fem;
  typedef StubYaspGrid<2> Grid; typedef StubGridPart<Grid> GridPart; typedef StubDGSpace<GridPart, B200FEM_DG_ONB> Space; typedef StubDiscreteFunction<Space> DF;
  Grid grid({-1.0, -1.0}, {1.0, 1.0}, {8, 8}); GridPart gridPart(grid); Space space(gridPart, 2, 6);
  try {
    B200GalerkinOperator<GeneratedIntegrands, DF> op(space, space);
    const Operator<DF, DF>& base = op;                      // used through the abstract interface, as the solvers do
    DF u(space), w(space);
    for (std::size_t i = 0; i < u.dofVector().size(); ++i) u.dofVector()[i] = std::sin(0.37 * double(i));
    base(u, w);
    double s = 0; for (double v : w.dofVector()) s += v * v;
    std::printf("apply_norm2 %.17g\n", s);
    // solve L[x] = 0  <=>  A x = -L[0] with GMRES on the device
    DF zero(space), rhs(space), x(space);
    base(zero, rhs); for (double& v : rhs.dofVector()) v = -v;
    B200KrylovInverseOperator<DF> inv(B200KrylovInverseOperator<DF>::gmres, 1e-11, 3000, B200FEM_TOL_ABSOLUTE, 40);
    inv.bind(op); inv(rhs, x);
    base(x, w);
    double r = 0; for (double v : w.dofVector()) r = std::fmax(r, std::fabs(v));
    std::printf("iterations %d residual %.3e\n", inv.iterations(), r);
    for (std::size_t i = 0; i < 6; ++i) std::printf("x%zu %.17g\n", i, x.dofVector()[i]);
    if (!(inv.iterations() > 0 && r < 1e-8)) return 2;
    return unstructuredPart();
  } catch (const Dune::Exception& e) { std::printf("exception: %s\n", e.what()); return 1; }
}
