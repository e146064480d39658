// oracle/ref_bind.cpp -- TEST INFRASTRUCTURE.  C bindings around the pieces of the reference that compile from their own
// source files, where they lie under /root/reference (nothing is copied into this repository):
//
//   dune/fem/solver/linear/cg.hh, bicgstab.hh, gmres.hh        the Krylov loops (templates on operator / discrete function)
//   dune/fem/solver/cginverseoperator.hh                       ConjugateGradientSolver, the CG behind the legacy CGInverseOperator
//   dune/fem/operator/common/automaticdifferenceoperator.hh    the Jacobian-free linearisation (difference quotient, choice of eps)
//   dune/fem/solver/newtoninverseoperator.hh, solver/parameter.hh   the Newton loop (line search, Eisenstat-Walker forcing, failure
//                                                              codes, shared linear-iteration budget) and the parameter keys it reads
//   dune/fem/quadrature/gausspoints{,_implementation}.hh       the 1-D Gauss tables
//   dune/fem/space/shapefunctionset/legendrepolynomials.{hh,cc} the Legendre coefficient table and its Horner evaluation
//   dune/fem/space/shapefunctionset/orthonormal/orthonormalbase_{1,2,3}d.hh   the orthonormal P_k bases behind `dgonb`
//   dune/fem/quadrature/femquadratures{,_inline}.hh            CubeQuadrature: tensor construction of the Gauss rules, order selection
//   dune/fem/space/shapefunctionset/legendre.hh                the Legendre shape function SET (multi-index order, hierarchical sort)
//   dune/fem/space/lagrange/generic{geometry,lagrangepoints,basefunctions}.hh   the Lagrange points of the cube (local numbering,
//                                                              sub-entity and dof-in-entity of every node) and the Lagrange basis
//
// Built by oracle/Makefile (target `ref`) into oracle/_ref/libdunefem_ref.so with `-I oracle/ref_shim -I /root/reference`;
// the shim directory only supplies stand-ins for headers of dune-common that are absent from this image (see the files
// there).  The element loop of schemes/galerkin.hh needs dune-grid/-geometry and cannot be compiled here.
//
// tests/test_reference_pieces.py uses this library to pin the oracle's restatements (fo_cg, fo_bicgstab, fo_gmres,
// fo_quadrature, fo_legendre, the dgonb shape functions) against the reference code itself.  Only tests/ may load it.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>   // std::real(double): dune-common pulls it in for the reference (cg.hh:64)
#include <iostream>
#include <map>
#include <memory>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include <dune/fem/space/common/auxiliarydofs.hh>   // stand-in: forEachPrimaryDof
#include <dune/fem/solver/linear/cg.hh>
#include <dune/fem/solver/linear/bicgstab.hh>
#include <dune/fem/solver/linear/gmres.hh>
#include <dune/fem/quadrature/gausspoints.hh>
#include <dune/fem/space/shapefunctionset/legendrepolynomials.hh>
#include <dune/fem/space/shapefunctionset/orthonormal/orthonormalbase_1d.hh>
#include <dune/fem/space/shapefunctionset/orthonormal/orthonormalbase_2d.hh>
#include <dune/fem/space/shapefunctionset/orthonormal/orthonormalbase_3d.hh>
#include <dune/fem/space/lagrange/genericbasefunctions.hh>
#include <dune/fem/space/shapefunctionset/legendre.hh>
#include <dune/fem/quadrature/femquadratures.hh>
#include <dune/fem/operator/common/automaticdifferenceoperator.hh>
#include <dune/fem/solver/newtoninverseoperator.hh>
#include <dune/fem/solver/cginverseoperator.hh>

namespace {

// The slice of the DiscreteFunction interface the reference loops use (function/common/discretefunction.hh,
// blockvectors/defaultblockvectors.hh:39-150, scalarproducts.hh:115-127), over a plain array on one rank.
struct Comm { template <class T> void sum(T*, int) const {} int size() const { return 1; } };
struct GridPart { Comm c; const Comm& comm() const { return c; } };
struct Communicator { double exchangeTime() const { return 0.0; } };
struct Space {
  Communicator communicator() const { return Communicator(); }
  GridPart gp; std::vector<std::size_t> aux;   // sorted auxiliary dofs, terminated by the vector size (auxiliarydofs.hh)
  const GridPart& gridPart() const { return gp; }
  const std::vector<std::size_t>& auxiliaryDofs() const { return aux; }
};

struct Vec {
  typedef double RangeFieldType;
  const Space* sp; std::vector<double> d;
  Vec(const Space& s, std::size_t n) : sp(&s), d(n, 0.0) {}
  const Space& space() const { return *sp; }
  std::vector<double>& dofVector() { return d; }
  const std::vector<double>& dofVector() const { return d; }
  void clear() { std::fill(d.begin(), d.end(), 0.0); }
  void assign(const Vec& o) { d = o.d; }
  Vec& operator+=(const Vec& o) { for (std::size_t i = 0; i < d.size(); ++i) d[i] += o.d[i]; return *this; }
  Vec& operator-=(const Vec& o) { for (std::size_t i = 0; i < d.size(); ++i) d[i] -= o.d[i]; return *this; }
  Vec& operator*=(double s) { for (std::size_t i = 0; i < d.size(); ++i) d[i] *= s; return *this; }
  void axpy(double s, const Vec& o) { for (std::size_t i = 0; i < d.size(); ++i) d[i] += s * o.d[i]; }
  double scalarProductDofs(const Vec& o) const {
    double scp = 0;
    Dune::Fem::forEachPrimaryDof(sp->aux, [&](std::size_t i) { scp += d[i] * o.d[i]; });
    return scp;
  }
  double normSquaredDofs() const { return scalarProductDofs(*this); }
};

typedef void (*ApplyFn)(const double* u, double* w, void* ctx);
struct Op {
  ApplyFn fn; void* ctx;
  void operator()(const Vec& u, Vec& w) const { fn(u.d.data(), w.d.data(), ctx); }
};

// ---- AutomaticDifferenceOperator / AutomaticDifferenceLinearOperator (operator/common/automaticdifferenceoperator.hh:25-166):
// the discrete function is the Vec above behind the (name, space) constructor the class uses for its temporaries
struct SizedSpace : Space { std::size_t n = 0; };
struct NamedVec : Vec {
  typedef SizedSpace DiscreteFunctionSpaceType;
  NamedVec(const std::string&, const SizedSpace& s) : Vec(s, s.n) {}
  const SizedSpace& space() const { return static_cast<const SizedSpace&>(*sp); }
};
struct CallbackDifferenceOperator : Dune::Fem::AutomaticDifferenceOperator<NamedVec> {
  ApplyFn fn; void* ctx;
  CallbackDifferenceOperator(ApplyFn f, void* c, double eps) : Dune::Fem::AutomaticDifferenceOperator<NamedVec>(eps), fn(f), ctx(c) {}
  CallbackDifferenceOperator(ApplyFn f, void* c) : fn(f), ctx(c) {}      // eps from the (empty) parameter file: 0 = chosen per argument
  void operator()(const NamedVec& u, NamedVec& w) const override { fn(u.d.data(), w.d.data(), ctx); }
};

// ---- NewtonInverseOperator< JacobianOperator, LInvOp > (solver/newtoninverseoperator.hh:418-803) around the two classes above.
// Jacobian operator: the reference's difference quotient behind the four-argument constructor the loop uses (:703)
struct DifferenceJacobian : Dune::Fem::AutomaticDifferenceLinearOperator<NamedVec> {
  typedef Dune::Fem::AutomaticDifferenceLinearOperator<NamedVec> Base;
  DifferenceJacobian(const std::string& name, const SizedSpace& d, const SizedSpace& r, const Dune::Fem::SolverParameter&) : Base(name, d, r) {}
  using Base::set;      // (a friend of the base class only; AutomaticDifferenceOperator< ..., DifferenceJacobian > calls it)
};
struct CallbackDifferentiableOperator : Dune::Fem::AutomaticDifferenceOperator<NamedVec, NamedVec, DifferenceJacobian> {
  ApplyFn fn; void* ctx; bool nonlin;
  CallbackDifferentiableOperator(ApplyFn f, void* c, bool nl) : fn(f), ctx(c), nonlin(nl) {}
  void operator()(const NamedVec& u, NamedVec& w) const override { fn(u.d.data(), w.d.data(), ctx); }
  bool nonlinear() const override { return nonlin; }
};
// Linear inverse operator: the slice of KrylovInverseOperator the Newton loop touches (solver/krylovinverseoperators.hh:120-205 -- the
// header itself needs function/common/discretefunction.hh): method, tolerance, budget and criterion come from the reference's
// SolverParameter, the loops are the reference's LinearSolver::{cg,bicgstab,gmres} with the temporaries laid out as there
struct RefKrylov {
  typedef Dune::Fem::SolverParameter SolverParameterType;
  std::shared_ptr<SolverParameterType> par; const DifferenceJacobian* op = nullptr; int method; mutable int its = 0;
  explicit RefKrylov(const SolverParameterType& p) : par(p.clone()) {
    method = par->solverMethod({SolverParameterType::gmres, SolverParameterType::bicgstab, SolverParameterType::cg});   // the default list of :97
  }
  SolverParameterType& parameter() const { return *par; }
  void bind(const DifferenceJacobian& j) { op = &j; }
  void unbind() { op = nullptr; }
  void setMaxIterations(int n) { par->setMaxIterations(n); }
  int iterations() const { return its; }
  void operator()(const NamedVec& u, NamedVec& w) const {
    const SizedSpace& sp = u.space();
    const DifferenceJacobian* none = nullptr;
    if (method == SolverParameterType::gmres) {
      std::vector<NamedVec> v(par->gmresRestart() + 1, NamedVec("GMRes::v", sp));
      its = Dune::Fem::LinearSolver::gmres(*op, none, v, w, u, par->gmresRestart(), par->tolerance(), par->maxIterations(), par->errorMeasure(), (std::ostream*)nullptr);
    } else if (method == SolverParameterType::bicgstab) {
      std::vector<NamedVec> v(5, NamedVec("BiCGStab::r", sp));
      its = Dune::Fem::LinearSolver::bicgstab(*op, none, v, w, u, par->tolerance(), par->maxIterations(), par->errorMeasure(), (std::ostream*)nullptr);
    } else {
      std::vector<NamedVec> v(3, NamedVec("CG::h", sp));
      its = Dune::Fem::LinearSolver::cg(*op, none, v, w, u, par->tolerance(), par->maxIterations(), par->errorMeasure(), (std::ostream*)nullptr);
    }
  }
};

Space makeSpace(std::int64_t n, const std::int64_t* aux, std::int64_t naux) {
  Space s; for (std::int64_t i = 0; i < naux; ++i) s.aux.push_back(std::size_t(aux[i])); s.aux.push_back(std::size_t(n)); return s;
}

// the loops report residuals only through their verbose stream: parse it back (17 significant digits round-trip a double)
int parseHistory(const std::string& log, const char* key, double* hist, int maxHist) {
  std::istringstream in(log); std::string line; int k = 0;
  while (std::getline(in, line)) {
    if (line.find(key) == std::string::npos) continue;
    std::string tail = line.substr(line.rfind(':') + 1);            // " residual 1.2e-3" (cg.hh:110) or " 1.2e-3"
    const std::size_t r = tail.find("residual"); if (r != std::string::npos) tail = tail.substr(r + 8);
    if (hist && k < maxHist) hist[k] = std::stod(tail);
    ++k;
  }
  return k;
}

// ---- generic Lagrange points / base functions of the dim-cube (space/lagrange/shapefunctionset.hh:88 instantiates them like this)
template <int dim> struct ScalarFunctionSpace {
  static const int dimDomain = dim, dimRange = 1; typedef double DomainFieldType; typedef double RangeFieldType;
  typedef Dune::FieldVector<double, dim> DomainType; typedef Dune::FieldVector<double, 1> RangeType;
};
template <int dim, unsigned order> struct LagrangeCube {
  typedef typename Dune::Fem::GeometryWrapper<(1u << dim) - 1u, dim>::ImplType Geometry;
  typedef Dune::Fem::GenericLagrangeBaseFunction<ScalarFunctionSpace<dim>, Geometry, order> BaseFunction;
  typedef Dune::Fem::GenericLagrangePoint<Geometry, order> Point;
  static int points(double* x, int* codim, int* sub, int* dofnum) {
    const int n = (int)BaseFunction::numBaseFunctions;
    for (int b = 0; b < n; ++b) {
      Point pt(b); Dune::FieldVector<double, dim> xl; pt.local(xl);
      unsigned int c, s, k; pt.dofSubEntity(c, s, k);
      if (x) for (int d = 0; d < dim; ++d) x[b * dim + d] = xl[d];
      if (codim) { codim[b] = (int)c; sub[b] = (int)s; dofnum[b] = (int)k; }
    }
    return n;
  }
  static void evaluate(int base, const double* x, double* phi, double* dphi) {
    BaseFunction bf(base); Dune::FieldVector<double, dim> xl; for (int d = 0; d < dim; ++d) xl[d] = x[d];
    Dune::FieldVector<double, 1> v; Dune::FieldVector<int, 0> d0; bf.evaluate(d0, xl, v); *phi = v[0];
    for (int d = 0; d < dim; ++d) { Dune::FieldVector<int, 1> d1; d1[0] = d; bf.evaluate(d1, xl, v); dphi[d] = v[0]; }
  }
};

// ---- LegendreShapeFunctionSet< FunctionSpace, hierarchicalOrdering > (space/shapefunctionset/legendre.hh:216-360)
template <int dim> struct LegendreFunctionSpace : ScalarFunctionSpace<dim> {
  typedef Dune::FieldVector<Dune::FieldVector<double, dim>, 1> JacobianRangeType;
  typedef Dune::FieldVector<Dune::FieldVector<Dune::FieldVector<double, dim>, dim>, 1> HessianRangeType;
};
template <int dim, bool hier> int legendreSet(int order, const double* x, double* phi, double* dphi) {
  Dune::Fem::LegendreShapeFunctionSet<LegendreFunctionSpace<dim>, hier> sfs(order);
  Dune::FieldVector<double, dim> xl; for (int d = 0; d < dim; ++d) xl[d] = x[d];
  if (phi) sfs.evaluateEach(xl, [&](std::size_t i, const Dune::FieldVector<double, 1>& v) { phi[i] = v[0]; });
  if (dphi) sfs.jacobianEach(xl, [&](std::size_t i, const typename LegendreFunctionSpace<dim>::JacobianRangeType& j) { for (int d = 0; d < dim; ++d) dphi[i * dim + d] = j[0][d]; });
  return (int)sfs.size();
}

}  // namespace

extern "C" {

// LegendreShapeFunctionSet (plain and hierarchical ordering): values phi[n] and reference gradients dphi[n][dim] of all shape
// functions at x, in the set's own order; returns n
int ref_legendre_set(int dim, int order, int hierarchical, const double* x, double* phi, double* dphi) {
  if (dim == 2) return hierarchical ? legendreSet<2, true>(order, x, phi, dphi) : legendreSet<2, false>(order, x, phi, dphi);
  if (dim == 3) return hierarchical ? legendreSet<3, true>(order, x, phi, dphi) : legendreSet<3, false>(order, x, phi, dphi);
  return 0;
}

// CubeQuadrature< double, dim >( cube, order ) (quadrature/femquadratures_inline.hh:33-95): points x[n][dim], weights w[n] in the
// rule's own order, *exact receives the order of the rule that was selected; returns n (x, w may be NULL)
int ref_cube_quadrature(int dim, int order, double* x, double* w, int* exact) {
#define B200_CASE(D) if (dim == D) { Dune::Fem::CubeQuadrature<double, D> q(Dune::GeometryTypes::cube(D), order, 0); if (exact) *exact = q.order(); \
    if (x) for (int i = 0; i < (int)q.nop(); ++i) { for (int d = 0; d < D; ++d) x[i * D + d] = q.point(i)[d]; w[i] = q.weight(i); } return (int)q.nop(); }
  B200_CASE(1) B200_CASE(2) B200_CASE(3)
#undef B200_CASE
  return 0;
}

// GenericLagrangePoint of the dim-cube (space/lagrange/genericlagrangepoints.hh): local coordinates, codimension / number of the
// sub-entity and number of the dof inside it, for every local node; returns the number of nodes (0: unsupported dim / order)
int ref_lagrange_cube_points(int dim, int order, double* x, int* codim, int* sub, int* dofnum) {
#define B200_CASE(D, O) if (dim == D && order == O) return LagrangeCube<D, O>::points(x, codim, sub, dofnum);
  B200_CASE(2, 1) B200_CASE(2, 2) B200_CASE(2, 3) B200_CASE(3, 1) B200_CASE(3, 2) B200_CASE(3, 3)
#undef B200_CASE
  return 0;
}
// GenericLagrangeBaseFunction::evaluate (space/lagrange/genericbasefunctions.hh): value and reference gradient of local basis function `base`
int ref_lagrange_cube_evaluate(int dim, int order, int base, const double* x, double* phi, double* dphi) {
#define B200_CASE(D, O) if (dim == D && order == O) { LagrangeCube<D, O>::evaluate(base, x, phi, dphi); return 0; }
  B200_CASE(2, 1) B200_CASE(2, 2) B200_CASE(2, 3) B200_CASE(3, 1) B200_CASE(3, 2) B200_CASE(3, 3)
#undef B200_CASE
  return -1;
}

// LinearSolver::cg (dune/fem/solver/linear/cg.hh:18-117); precon may be NULL
int ref_cg(ApplyFn apply, void* ctx, ApplyFn precon, void* pctx, std::int64_t n, const std::int64_t* aux, std::int64_t naux,
           double* x, const double* b, double eps, int maxit, int crit, double* hist, int maxHist, int* nHist) {
  Space sp = makeSpace(n, aux, naux);
  Vec X(sp, n), B(sp, n); std::memcpy(X.d.data(), x, n * 8); std::memcpy(B.d.data(), b, n * 8);
  std::vector<Vec> tmp(precon ? 5 : 3, Vec(sp, n));
  Op op{apply, ctx}, pre{precon, pctx};
  std::ostringstream os; os << std::setprecision(17);
  const int it = Dune::Fem::LinearSolver::cg(op, precon ? &pre : (Op*)nullptr, tmp, X, B, eps, maxit, crit, &os);
  const int k = parseHistory(os.str(), "Fem::CG it:", hist, maxHist); if (nHist) *nHist = k;
  std::memcpy(x, X.d.data(), n * 8);
  return it;
}

// AutomaticDifferenceOperator::jacobian at u, then nArgs applications of the linear operator (automaticdifferenceoperator.hh:110-166);
// eps <= 0: the reference's dynamic choice; useParameter != 0: eps read from the parameter file (absent -> 0)
int ref_difference_quotient(ApplyFn apply, void* ctx, std::int64_t n, const std::int64_t* aux, std::int64_t naux, const double* u,
                            double eps, int useParameter, const double* args, int nArgs, double* dest) {
  SizedSpace sp; static_cast<Space&>(sp) = makeSpace(n, aux, naux); sp.n = std::size_t(n);
  NamedVec U("u", sp), A("arg", sp), D("dest", sp); std::memcpy(U.d.data(), u, n * 8);
  std::unique_ptr<CallbackDifferenceOperator> op(useParameter ? new CallbackDifferenceOperator(apply, ctx) : new CallbackDifferenceOperator(apply, ctx, eps));
  Dune::Fem::AutomaticDifferenceLinearOperator<NamedVec> jac("jac", sp, sp);
  op->jacobian(U, jac);
  for (int k = 0; k < nArgs; ++k) {
    std::memcpy(A.d.data(), args + std::size_t(k) * n, n * 8);
    jac(A, D);
    std::memcpy(dest + std::size_t(k) * n, D.d.data(), n * 8);
  }
  return 0;
}

// NewtonInverseOperator::operator()(u, w) (solver/newtoninverseoperator.hh:690-803) on a callback operator, configured like the reference
// is: through parameter KEYS ("fem.solver.nonlinear.tolerance", "...linear.method", ... ; keys / values = '\n'-separated "key: value" lines).
// out = {iterations, linearIterations, failure code (NewtonFailure, :389-400)}, *residual = |L[w] - u| of the last iterate.
int ref_newton(ApplyFn apply, void* ctx, int nonlinear, std::int64_t n, const std::int64_t* aux, std::int64_t naux, const double* u, double* w,
               const char* parameters, int* out, double* residual) {
  std::map<std::string, std::string>& table = Dune::Fem::RefShim::table();
  table.clear();
  { std::istringstream in(parameters ? parameters : ""); std::string line;
    while (std::getline(in, line)) { const std::size_t c = line.find(": "); if (c != std::string::npos) table[line.substr(0, c)] = line.substr(c + 2); } }
  int rc = 0;
  try {
    SizedSpace sp; static_cast<Space&>(sp) = makeSpace(n, aux, naux); sp.n = std::size_t(n);
    NamedVec U("u", sp), W("w", sp); if (u) std::memcpy(U.d.data(), u, n * 8); std::memcpy(W.d.data(), w, n * 8);
    CallbackDifferentiableOperator op(apply, ctx, nonlinear != 0);
    Dune::Fem::NewtonInverseOperator<DifferenceJacobian, RefKrylov> newton;      // everything from the parameter table, as FemScheme does
    newton.bind(op);
    newton(U, W);
    out[0] = newton.iterations(); out[1] = newton.linearIterations(); out[2] = (int)newton.failed(); *residual = newton.residual();
    newton.unbind();
    std::memcpy(w, W.d.data(), n * 8);
  } catch (const std::exception& e) { std::cerr << "ref_newton: " << e.what() << std::endl; rc = 1; }
  table.clear();
  return rc;
}

// ConjugateGradientSolver::solve (dune/fem/solver/cginverseoperator.hh:595-650; preconditioned :653-720), the class behind the legacy
// CGInverseOperator: errorMeasure 0 absolute / 1 relative to |b|; returns iterations()
struct LegacyOp : Dune::Fem::Operator<Vec, Vec> {
  ApplyFn fn; void* ctx;
  LegacyOp(ApplyFn f, void* c) : fn(f), ctx(c) {}
  void operator()(const Vec& u, Vec& w) const override { fn(u.d.data(), w.d.data(), ctx); }
};
int ref_legacy_cg(ApplyFn apply, void* ctx, ApplyFn precon, void* pctx, std::int64_t n, const std::int64_t* aux, std::int64_t naux,
                  double* x, const double* b, double eps, int maxit, int errorMeasure) {
  Space sp = makeSpace(n, aux, naux);
  Vec X(sp, n), B(sp, n); std::memcpy(X.d.data(), x, n * 8); std::memcpy(B.d.data(), b, n * 8);
  LegacyOp op(apply, ctx), pre(precon, pctx);
  Dune::Fem::ConjugateGradientSolver<Dune::Fem::Operator<Vec, Vec>> solver(eps, (unsigned)maxit, errorMeasure, false);
  if (precon) solver.solve(op, pre, B, X); else solver.solve(op, B, X);
  std::memcpy(x, X.d.data(), n * 8);
  return (int)solver.iterations();
}

// LinearSolver::bicgstab (dune/fem/solver/linear/bicgstab.hh:63-214)
int ref_bicgstab(ApplyFn apply, void* ctx, std::int64_t n, const std::int64_t* aux, std::int64_t naux,
                 double* x, const double* b, double tol, int maxit, int crit, double* hist, int maxHist, int* nHist) {
  Space sp = makeSpace(n, aux, naux);
  Vec X(sp, n), B(sp, n); std::memcpy(X.d.data(), x, n * 8); std::memcpy(B.d.data(), b, n * 8);
  std::vector<Vec> tmp(5, Vec(sp, n));
  Op op{apply, ctx};
  std::ostringstream os; os << std::setprecision(17);
  const int it = Dune::Fem::LinearSolver::bicgstab(op, (Op*)nullptr, tmp, X, B, tol, maxit, crit, &os);
  const int k = parseHistory(os.str(), "Fem::BiCGstab it:", hist, maxHist); if (nHist) *nHist = k;
  std::memcpy(x, X.d.data(), n * 8);
  return it;
}

// LinearSolver::gmres (dune/fem/solver/linear/gmres.hh:116-301)
int ref_gmres(ApplyFn apply, void* ctx, std::int64_t n, const std::int64_t* aux, std::int64_t naux,
              double* x, const double* b, int restart, double tol, int maxit, int crit, double* hist, int maxHist, int* nHist) {
  Space sp = makeSpace(n, aux, naux);
  Vec X(sp, n), B(sp, n); std::memcpy(X.d.data(), x, n * 8); std::memcpy(B.d.data(), b, n * 8);
  std::vector<Vec> v(restart + 1, Vec(sp, n));
  Op op{apply, ctx};
  std::ostringstream os; os << std::setprecision(17);
  const int it = Dune::Fem::LinearSolver::gmres(op, (Op*)nullptr, v, X, B, restart, tol, maxit, crit, &os);
  const int k = parseHistory(os.str(), "Fem::GMRES it:", hist, maxHist); if (nHist) *nHist = k;
  std::memcpy(x, X.d.data(), n * 8);
  return it;
}

// GaussPts (dune/fem/quadrature/gausspoints_implementation.hh): m-point rule on [0,1]
int ref_gauss_maxp() { return Dune::Fem::GaussPts::MAXP; }
int ref_gauss_rule(int m, double* x, double* w) {
  static const Dune::Fem::GaussPts g;
  for (int i = 0; i < m; ++i) { x[i] = g.point(m, i); w[i] = g.weight(m, i); }
  return g.order(m);
}

// LegendrePolynomials (shapefunctionset/legendrepolynomials.hh:15-62, table in legendrepolynomials.cc)
int ref_legendre_max_order() { return Dune::Fem::LegendrePolynomials::maxOrder; }
double ref_legendre(int num, double x, int deriv) {
  typedef Dune::Fem::LegendrePolynomials L;
  return deriv == 0 ? L::evaluate(num, x) : deriv == 1 ? L::jacobian(num, x) : L::hessian(num, x);
}

// OrthonormalBase_{1,2,3}D on the cube (line / quadrilateral / hexahedron): value and gradient of shape function i
double ref_onb_cube(int dim, int i, const double* x, double* grad) {
  if (dim == 1) { typedef Dune::Fem::OrthonormalBase_1D<double, double> B; if (grad) B::grad_line(i, x, grad); return B::eval_line(i, x); }
  if (dim == 2) { typedef Dune::Fem::OrthonormalBase_2D<double, double> B; if (grad) B::grad_quadrilateral_2d(i, x, grad); return B::eval_quadrilateral_2d(i, x); }
  typedef Dune::Fem::OrthonormalBase_3D<double, double> B; if (grad) B::grad_hexahedron_3d(i, x, grad); return B::eval_hexahedron_3d(i, x);
}

}  // extern "C"
