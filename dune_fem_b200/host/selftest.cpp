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

    // (3) method of lines: MOLGalerkinOperator = M^-1 L; for the orthonormal Legendre basis on this uniform grid M^-1 = 1 / detJ
    MOLGalerkinOperator<DiscreteFunctionType> mol(dg, dg, adv);
    DiscreteFunctionType wm("wm", dg);
    mol(u, wm);
    const double detJ = 1.0 / (8.0 * 8.0 * 8.0);
    double merr = 0;
    for (std::size_t i = 0; i < dg.size(); ++i) merr = std::fmax(merr, std::fabs(wm.dofVector()[i] * detJ - w.dofVector()[i]));
    std::printf("MOL: max |detJ M^-1 L[u] - L[u]| / max|L[u]| = %.3e\n", merr / scale);
    if (!(merr <= 1e-12 * scale)) return 1;

    // (4) non-symmetric system through KrylovInverseOperator<bicgstab> and <gmres>
    Integrands ad = adv; ad.eps = 0.1; ad.b[1] = 0.5; ad.c = 1.0; ad.dirichlet_mask = 63; ad.data = 2;
    GalerkinOperator<DiscreteFunctionType> adop(dg, dg, ad);
    DiscreteFunctionType rb("rb", dg), xb("xb", dg), xg("xg", dg), rr("rr", dg);
    adop.loadVector(rb);
    SolverParameter kp; kp.tolerance = 1e-8; kp.maxIterations = 5000; kp.gmresRestart = 50;
    BicgstabInverseOperator<DiscreteFunctionType> bicg(kp); bicg.bind(adop); bicg(rb, xb);
    GmresInverseOperator<DiscreteFunctionType> gm(kp); gm.bind(adop); gm(rb, xg);
    double dmax = 0, xmax = 0;
    for (std::size_t i = 0; i < dg.size(); ++i) { dmax = std::fmax(dmax, std::fabs(xb.dofVector()[i] - xg.dofVector()[i])); xmax = std::fmax(xmax, std::fabs(xg.dofVector()[i])); }
    std::printf("BiCGStab: %d iterations, GMRES(50): %d iterations, max |x_bicgstab - x_gmres| / max|x| = %.3e\n", bicg.iterations(), gm.iterations(), dmax / xmax);
    if (!(bicg.converged() && gm.converged() && dmax <= 1e-6 * xmax)) return 1;

    // (5) round 2: an unstructured cube mesh (the unit square as 4 x 4 distorted quadrilaterals), P2, Jacobi-CG against CG
    {
      const int n = 4; std::vector<double> vx; std::vector<std::int64_t> cubes;
      for (int j = 0; j <= n; ++j) for (int i = 0; i <= n; ++i) { const double x = double(i) / n, y = double(j) / n; vx.push_back(x + 0.03 * std::sin(7.0 * y) * x * (1 - x)); vx.push_back(y + 0.03 * std::sin(5.0 * x) * y * (1 - y)); }
      for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) { const std::int64_t v = i + (n + 1) * j; cubes.insert(cubes.end(), {v, v + 1, v + n + 1, v + n + 2}); }
      typedef UnstructuredGridPart<2> UGridPart; typedef DiscreteFunctionSpace<UGridPart> USpace; typedef DiscreteFunction<USpace> UFunction;
      UGridPart ugp(ctx, vx, cubes);
      USpace usp(ugp, B200FEM_LAGRANGE, 2);
      Integrands pu; pu.c = 0.5; pu.dirichlet_mask = 1; pu.data = 2; pu.strong_dirichlet = 1;
      GalerkinOperator<UFunction> uop(usp, usp, pu);
      UFunction ub("b", usp), x1("x1", usp), x2("x2", usp);
      uop.loadVector(ub);
      SolverParameter up; up.tolerance = 1e-12; up.maxIterations = 500;
      CgInverseOperator<UFunction> ucg(up); ucg.bind(uop); ucg(ub, x1);
      JacobiCgInverseOperator<UFunction> upcg(up); upcg.bind(uop); upcg(ub, x2);
      double d = 0, m = 0; for (std::size_t i = 0; i < usp.size(); ++i) { d = std::fmax(d, std::fabs(x1.dofVector()[i] - x2.dofVector()[i])); m = std::fmax(m, std::fabs(x1.dofVector()[i])); }
      std::printf("unstructured P2 (%zu dofs): CG %d, Jacobi-CG %d iterations, max |x_cg - x_pcg| / max|x| = %.3e\n", usp.size(), ucg.iterations(), upcg.iterations(), d / m);
      if (!(ucg.converged() && upcg.converged() && d <= 1e-8 * m)) return 1;
    }

    // (6) round 2: a vector-valued space (dimRange 2) with run-time compiled integrands, solved by NewtonInverseOperator
    {
      SpaceType v2(gridPart, B200FEM_DG_LEGENDRE_HIER, 1, B200FEM_NUMBERING_YASP, 2);
      if (v2.size() != 2 * 8 * 8 * 8 * 8 || v2.localBlockSize() != 2) return 1;
      CompiledIntegrands ci; ci.hasSkeleton = true; ci.hasBoundary = true; ci.constants = {0.5, 40.0, 0.3};
      ci.source =
        "__device__ void interior(const double* x, const VectorValue& u, VectorRange& r, const double* c, int dim) {\n"
        "  for (int i = 0; i < dimRange; ++i) { const int j = (i + 1) % dimRange; r.s[i] = u.u[i] + c[2] * u.u[i] * u.u[j] - (1.0 + i + x[0]);\n"
        "    for (int d = 0; d < dim; ++d) r.F[i][d] = c[0] * u.du[i][d]; } }\n"
        "__device__ void skeleton(const double* x, int axis, double sign, double ihe, const VectorValue& in, const VectorValue& out, VectorRange& rin, VectorRange& rout, const double* c, int dim) {\n"
        "  for (int i = 0; i < dimRange; ++i) { const double jump = in.u[i] - out.u[i];\n"
        "    const double cj = c[0] * c[1] * ihe * jump - 0.5 * c[0] * (in.du[i][axis] + out.du[i][axis]) * sign;\n"
        "    rin.s[i] = cj; rout.s[i] = -cj; rin.F[i][axis] = rout.F[i][axis] = -0.5 * c[0] * jump * sign; } }\n"
        "__device__ void boundary(const double* x, int axis, int side, double ihbnd, const VectorValue& u, VectorRange& r, const double* c, int dim) {\n"
        "  const double sign = side ? 1.0 : -1.0;\n"
        "  for (int i = 0; i < dimRange; ++i) { r.s[i] = c[0] * c[1] * ihbnd * u.u[i] - c[0] * u.du[i][axis] * sign; r.F[i][axis] = -c[0] * u.u[i] * sign; } }\n";
      GalerkinOperator<DiscreteFunctionType> sys(v2, v2, ci);
      DiscreteFunctionType wv("w", v2), res("res", v2);
      NewtonParameter np; np.tolerance = 1e-7; np.linear.tolerance = 1e-7; np.linear.errorMeasure = B200FEM_TOL_RESIDUAL_REDUCTION; np.linear.maxIterations = 4000; np.linear.gmresRestart = 30;
      NewtonInverseOperator<DiscreteFunctionType> newton(np); newton.bind(sys);
      newton(wv);
      sys(wv, res);
      double rn = 0; for (double v : res.dofVector()) rn += v * v;
      std::printf("vector-valued DG (dimRange 2), compiled integrands: Newton %d iterations (%d linear), |L[w]| = %.3e\n", newton.iterations(), newton.linearIterations(), std::sqrt(rn));
      if (!(newton.converged() && std::sqrt(rn) < 2e-7)) return 1;
    }

    // (7) FemScheme::solve: constraints on the target, then the inverse operator (linear model: one exact Newton step by CG)
    {
      NewtonParameter sp; sp.linearMethod = B200Fem::cg; sp.linear.tolerance = 1e-12; sp.linear.maxIterations = 1000;
      FemScheme<DiscreteFunctionType> scheme(p2, poisson, sp);
      DiscreteFunctionType uh("uh", p2), res("res", p2);
      for (double& v : uh.dofVector()) v = 0.3;
      const SolverInfo info = scheme.solve(uh);
      scheme(uh, res);
      double rn = 0; for (double v : res.dofVector()) rn = std::fmax(rn, std::fabs(v));
      std::printf("FemScheme::solve: converged %d, %d linear iterations, max |L[uh]| = %.3e\n", (int)info.converged, info.linearIterations, rn);
      if (!(info.converged && rn < 1e-9)) return 1;
    }
    std::printf("host selftest OK\n");
    return 0;
  } catch (const InvalidStateException& e) {
    const std::string msg = e.what();
    if (msg.find("no CPU fallback") != std::string::npos) { std::printf("no CUDA device: %s (expected on a CPU-only box)\n", e.what()); return 0; }
    std::fprintf(stderr, "error: %s\n", e.what()); return 2;
  }
}
