// selftest.cpp -- exercises the C++ host mirror (b200fem.hh) end to end.  Built by __graft_entry__.build().
// On a machine without a GPU it verifies the error convention (exception, no CPU fallback) and exits 0.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "b200fem.hh"

using namespace B200Fem;

int main() {
  try {
    Context ctx(0);
    typedef CartesianGridPart<3> GridPartType;
    typedef DiscreteFunctionSpace<GridPartType> SpaceType;
    typedef DiscreteFunction<SpaceType> DiscreteFunctionType;
    GridPartType gridPart(ctx, {8, 8, 8}, {0, 0, 0}, {1, 1, 1});

    // (1) DG Q2 advection-diffusion apply: L[u] - (A u - b) must vanish
    SpaceType dg(gridPart, B200FEM_DG_LEGENDRE_HIER, 2);
    Integrands adv; adv.eps = 1e-2; adv.b[0] = 1; adv.beta = 80; adv.dirichlet_mask = 3; adv.data = 1; adv.has_skeleton = 1; adv.has_boundary = 1;
    GalerkinOperator<DiscreteFunctionType> op(dg, dg, adv);
    DiscreteFunctionType u("u", dg), w("w", dg), wl("wl", dg), b("b", dg);
    for (std::size_t i = 0; i < dg.size(); ++i) u.dofVector()[i] = std::sin(0.37 * i);
    op(u, w); op.applyLinear(u, wl); op.loadVector(b);
    double err = 0, scale = 0;
    for (std::size_t i = 0; i < dg.size(); ++i) { err = std::fmax(err, std::fabs(w.dofVector()[i] - (wl.dofVector()[i] - b.dofVector()[i]))); scale = std::fmax(scale, std::fabs(w.dofVector()[i])); }
    std::printf("affine consistency: max |L[u] - (Au - b)| / max|L[u]| = %.3e\n", err / scale);
    if (!(err <= 1e-12 * scale)) return 1;

    // (2) Poisson P2 Lagrange with strong Dirichlet data, CG through the bound operator
    SpaceType p2(gridPart, B200FEM_LAGRANGE, 2);
    Integrands poisson; poisson.dirichlet_mask = 63; poisson.data = 2; poisson.strong_dirichlet = 1;
    GalerkinOperator<DiscreteFunctionType> lap(p2, p2, poisson);
    DiscreteFunctionType rhs("rhs", p2), x("x", p2), r("r", p2);
    lap.loadVector(rhs);
    SolverParameter par; par.tolerance = 1e-10; par.maxIterations = 500;
    CgInverseOperator<DiscreteFunctionType> cg(par);
    cg.bind(lap);
    cg(rhs, x);
    lap.applyLinear(x, r);
    double res = 0; for (std::size_t i = 0; i < p2.size(); ++i) res += (r.dofVector()[i] - rhs.dofVector()[i]) * (r.dofVector()[i] - rhs.dofVector()[i]);
    std::printf("CG: %d iterations, |Ax-b| = %.3e\n", cg.iterations(), std::sqrt(res));
    if (!(cg.converged() && std::sqrt(res) < 2e-10)) return 1;
    std::printf("host selftest OK\n");
    return 0;
  } catch (const InvalidStateException& e) {
    const std::string msg = e.what();
    if (msg.find("no CPU fallback") != std::string::npos) { std::printf("no CUDA device: %s (expected on a CPU-only box)\n", e.what()); return 0; }
    std::fprintf(stderr, "error: %s\n", e.what()); return 2;
  }
}
