// b200fem.hh -- C++ host-side mirror of the reference's operator / solver interface for the GPU hot path.
//
// DUNE-FEM is C++ (header templates); its drop-in point for this path is the abstract operator
//   Dune::Fem::Operator< DomainFunction, RangeFunction >::operator()( u, w )   (dune/fem/operator/common/operator.hh:32-65)
// implemented by Dune::Fem::GalerkinOperator (dune/fem/schemes/galerkin.hh:1383-1504) and consumed by
// Dune::Fem::CgInverseOperator / KrylovInverseOperator (dune/fem/solver/krylovinverseoperators.hh:46-281) through
// bind( op ) and operator()( rhs, x ).  DUNE itself is not available in this image, so this header mirrors those
// interfaces (same names, argument meaning, error behaviour: exceptions on the C++ side, negative iteration counts
// for non-converged solves) on top of the C ABI in include/b200fem.h, without any DUNE dependency.  INTEGRATION.md shows
// the adapter a DUNE-FEM maintainer adds (dune/fem/schemes/b200galerkin.hh) -- it is the same code with DUNE's own
// DiscreteFunction in place of the one below.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200fem.h"

namespace B200Fem {

struct InvalidStateException : std::runtime_error { using std::runtime_error::runtime_error; };   // DUNE_THROW(InvalidStateException, ...)
struct NotImplemented : std::runtime_error { using std::runtime_error::runtime_error; };          // DUNE_THROW(NotImplemented, ...)

inline void check(int rc) {
  if (rc == B200FEM_OK) return;
  const std::string msg = b200fem_last_error();
  if (rc == B200FEM_ERR_NOT_IMPLEMENTED) throw NotImplemented(msg);
  throw InvalidStateException(msg);
}

// MPIManager analogue: device context (misc/mpimanager.hh:352-461)
class Context {
 public:
  explicit Context(int device = 0, void* stream = nullptr) { check(b200fem_ctx_create(device, stream, &h_)); }
  ~Context() { b200fem_ctx_destroy(h_); }
  Context(const Context&) = delete;
  b200fem_ctx* handle() const { return h_; }
 private:
  b200fem_ctx* h_ = nullptr;
};

// GridPart over a Cartesian YaspGrid (gridpart/common/gridpart.hh)
template <int dim>
class CartesianGridPart {
 public:
  static constexpr int dimension = dim;
  CartesianGridPart(Context& ctx, const std::array<int, dim>& cells, const std::array<double, dim>& lo, const std::array<double, dim>& hi) {
    int32_t n[3] = {1, 1, 1}; double l[3] = {0, 0, 0}, h[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) { n[d] = cells[d]; l[d] = lo[d]; h[d] = hi[d]; }
    check(b200fem_mesh_cartesian(ctx.handle(), dim, n, l, h, &h_));
  }
  ~CartesianGridPart() { b200fem_mesh_destroy(h_); }
  b200fem_mesh* handle() const { return h_; }
 private:
  b200fem_mesh* h_ = nullptr;
};

// the same on this rank's box of a global grid split into proc[0] x proc[1] x proc[2] boxes (YaspGrid's torus; the context carries
// the NCCL communicator, b200fem_nccl_init), optionally periodic (YaspGrid's periodic bitset: bit d = periodic along axis d)
template <int dim>
class DistributedCartesianGridPart {
 public:
  static constexpr int dimension = dim;
  DistributedCartesianGridPart(Context& ctx, const std::array<int, dim>& cells, const std::array<double, dim>& lo, const std::array<double, dim>& hi,
                               const std::array<int, dim>& proc, int rank, int periodic = 0) {
    int32_t n[3] = {1, 1, 1}, p[3] = {1, 1, 1}; double l[3] = {0, 0, 0}, h[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) { n[d] = cells[d]; l[d] = lo[d]; h[d] = hi[d]; p[d] = proc[d]; }
    check(b200fem_mesh_cartesian_distributed(ctx.handle(), dim, n, l, h, p, rank, &h_));
    if (periodic) { const int rc = b200fem_mesh_set_periodic(h_, periodic); if (rc) { b200fem_mesh_destroy(h_); check(rc); } }
  }
  ~DistributedCartesianGridPart() { b200fem_mesh_destroy(h_); }
  b200fem_mesh* handle() const { return h_; }
 private:
  b200fem_mesh* h_ = nullptr;
};

// AdaptiveLeafGridPart over an unstructured conforming cube grid (ALUGrid< dim, dim, cube, conforming >): vertex coordinates
// [nv][dim] and element -> vertex numbers [ne][2^dim] in the cube reference element's order (GridFactory::insertVertex / insertElement)
template <int dim>
class UnstructuredGridPart {
 public:
  static constexpr int dimension = dim;
  UnstructuredGridPart(Context& ctx, const std::vector<double>& vertices, const std::vector<std::int64_t>& cubes) {
    if (vertices.size() % dim != 0 || cubes.size() % (1u << dim) != 0) throw InvalidStateException("UnstructuredGridPart: vertices [nv][dim], cubes [ne][2^dim]");
    check(b200fem_mesh_unstructured(ctx.handle(), dim, (std::int64_t)(vertices.size() / dim), vertices.data(), (std::int64_t)(cubes.size() >> dim), cubes.data(), &h_));
  }
  ~UnstructuredGridPart() { b200fem_mesh_destroy(h_); }
  b200fem_mesh* handle() const { return h_; }
 private:
  b200fem_mesh* h_ = nullptr;
};

// DiscreteFunctionSpace: LagrangeDiscreteFunctionSpace / LegendreDiscontinuousGalerkinSpace / Hierarchic...
template <class GridPart>
class DiscreteFunctionSpace {
 public:
  typedef GridPart GridPartType;
  // dimRange > 1: FunctionSpace< ..., dimRange >, dof (block, c) = block * dimRange + c (localBlockSize = dimRange)
  DiscreteFunctionSpace(const GridPart& gp, b200fem_space_kind kind, int order, b200fem_numbering numbering = B200FEM_NUMBERING_YASP, int dimRange = 1)
      : gridPart_(gp), order_(order), dimRange_(dimRange) {
    check(b200fem_space_create_vector(gp.handle(), kind, order, numbering, dimRange, &h_));
    int64_t s = 0; check(b200fem_space_size(h_, &s)); size_ = (std::size_t)s;
  }
  ~DiscreteFunctionSpace() { b200fem_space_destroy(h_); }
  std::size_t size() const { return size_; }
  int order() const { return order_; }
  int localBlockSize() const { return dimRange_; }
  const GridPart& gridPart() const { return gridPart_; }
  b200fem_space* handle() const { return h_; }
 private:
  const GridPart& gridPart_; int order_, dimRange_; b200fem_space* h_ = nullptr; std::size_t size_ = 0;
};

// AdaptiveDiscreteFunction: host-owned dof vector in the reference layout (function/adaptivefunction/adaptivefunction.hh:45-203)
template <class Space>
class DiscreteFunction {
 public:
  typedef Space DiscreteFunctionSpaceType;
  typedef double RangeFieldType;
  DiscreteFunction(const std::string& name, const Space& space) : name_(name), space_(space), dofs_(space.size(), 0.0) {}
  const Space& space() const { return space_; }
  double* leakPointer() { return dofs_.data(); }                     // adaptivefunction.hh:130-131
  const double* leakPointer() const { return dofs_.data(); }
  std::vector<double>& dofVector() { return dofs_; }
  const std::vector<double>& dofVector() const { return dofs_; }
  void clear() { dofs_.assign(dofs_.size(), 0.0); }
  void assign(const DiscreteFunction& o) { dofs_ = o.dofs_; }
  void axpy(double a, const DiscreteFunction& o) { for (std::size_t i = 0; i < dofs_.size(); ++i) dofs_[i] += a * o.dofs_[i]; }
  const std::string& name() const { return name_; }
 private:
  std::string name_; const Space& space_; std::vector<double> dofs_;
};

// Dune::Fem::Operator (operator/common/operator.hh:32-65)
template <class DomainFunction, class RangeFunction = DomainFunction>
struct Operator {
  typedef DomainFunction DomainFunctionType;
  typedef RangeFunction RangeFunctionType;
  virtual ~Operator() = default;
  virtual void operator()(const DomainFunctionType& u, RangeFunctionType& w) const = 0;
  virtual void finalize() {}
  virtual bool nonlinear() const { return false; }
};

// Integrands of the advection-diffusion-reaction family (see include/b200fem.h, b200fem_model)
struct Integrands : b200fem_model {
  Integrands() : b200fem_model{} { eps = 1.0; }
};

// Integrands generated from a UFL form (python/dune/models/integrands/model.py:72-106): the bodies of interior / skeleton / boundary as
// CUDA C++ source (interface: include/b200fem.h, b200fem_operator_create_jit) + the dune.ufl.Constant coefficients
struct CompiledIntegrands {
  std::string source; std::vector<double> constants; bool hasSkeleton = true, hasBoundary = true;
};

// Dune::Fem::GalerkinOperator< Integrands, DomainFunction, RangeFunction > (schemes/galerkin.hh:1383-1504), optionally
// wrapped like DirichletWrapperOperator (schemes/dirichletwrapper.hh:29-165) when integrands.strong_dirichlet is set
template <class DiscreteFunctionT>
class GalerkinOperator : public Operator<DiscreteFunctionT, DiscreteFunctionT> {
 public:
  typedef typename DiscreteFunctionT::DiscreteFunctionSpaceType DiscreteFunctionSpaceType;
  GalerkinOperator(const DiscreteFunctionSpaceType& dSpace, const DiscreteFunctionSpaceType& rSpace, const Integrands& integrands)
      : space_(dSpace), integrands_(integrands) {
    if (&dSpace != &rSpace) throw NotImplemented("domain and range space must coincide");
    check(b200fem_operator_create(dSpace.handle(), &integrands_, &h_));
  }
  // ... over run-time compiled integrands (one specialised kernel per form, like the reference's JIT-compiled operator)
  GalerkinOperator(const DiscreteFunctionSpaceType& dSpace, const DiscreteFunctionSpaceType& rSpace, const CompiledIntegrands& integrands)
      : space_(dSpace), integrands_(), compiled_(true) {
    if (&dSpace != &rSpace) throw NotImplemented("domain and range space must coincide");
    check(b200fem_operator_create_jit(dSpace.handle(), integrands.source.c_str(), integrands.constants.empty() ? nullptr : integrands.constants.data(),
                                      (int)integrands.constants.size(), integrands.hasSkeleton, integrands.hasBoundary, &h_));
  }
  void setConstants(const std::vector<double>& c) { check(b200fem_operator_set_constants(h_, c.empty() ? nullptr : c.data(), (int)c.size())); }
  ~GalerkinOperator() override { b200fem_operator_destroy(h_); }
  void operator()(const DiscreteFunctionT& u, DiscreteFunctionT& w) const override { check(b200fem_operator_apply(h_, u.leakPointer(), w.leakPointer())); }
  // homogeneous linear part (what the Krylov solvers apply)
  void applyLinear(const DiscreteFunctionT& u, DiscreteFunctionT& w) const { check(b200fem_operator_apply_linear(h_, u.leakPointer(), w.leakPointer())); }
  void loadVector(DiscreteFunctionT& b) const { check(b200fem_operator_load_vector(h_, b.leakPointer())); }
  bool nonlinear() const override { return compiled_ || integrands_.gamma != 0.0; }     // (compiled forms are not known to be linear)
  // diag(A) of the homogeneous linear part, matrix-free (what DiagonalPreconditioner needs, solver/diagonalpreconditioner.hh)
  void diagonal(DiscreteFunctionT& d) const { check(b200fem_operator_diagonal(h_, d.leakPointer())); }
  // AutomaticDifferenceOperator::jacobian: linearise at u; applyLinear and the Krylov solvers then act on J(u) (automaticdifferenceoperator.hh:108-166)
  void linearize(const DiscreteFunctionT& u, double eps = 0.0) { check(b200fem_operator_linearize(h_, u.leakPointer(), eps)); }
  void dropLinearization() { check(b200fem_operator_linearize(h_, nullptr, 0.0)); }
  void setCommunicate(bool communicate) { check(b200fem_operator_set_communicate(h_, communicate)); }                 // galerkin.hh:1409
  void setQuadratureOrders(unsigned interior, unsigned surface) { check(b200fem_operator_set_quadrature_orders(h_, interior, surface)); }   // :1418-1423
  const DiscreteFunctionSpaceType& domainSpace() const { return space_; }
  const DiscreteFunctionSpaceType& rangeSpace() const { return space_; }
  const Integrands& model() const { return integrands_; }
  b200fem_operator* handle() const { return h_; }
 private:
  const DiscreteFunctionSpaceType& space_; Integrands integrands_; bool compiled_ = false; b200fem_operator* h_ = nullptr;
};

// Dune::Fem::MOLGalerkinOperator (schemes/molgalerkin.hh:37-209): same constructor and interface, applies the inverse local
// mass matrix after the evaluate (w = M^-1 L[u])
template <class DiscreteFunctionT>
class MOLGalerkinOperator : public GalerkinOperator<DiscreteFunctionT> {
 public:
  typedef typename DiscreteFunctionT::DiscreteFunctionSpaceType DiscreteFunctionSpaceType;
  MOLGalerkinOperator(const DiscreteFunctionSpaceType& dSpace, const DiscreteFunctionSpaceType& rSpace, const Integrands& integrands)
      : GalerkinOperator<DiscreteFunctionT>(dSpace, rSpace, integrands) { check(b200fem_operator_set_inverse_mass(this->handle(), 1)); }
};

// Dune::Fem::CgInverseOperator = KrylovInverseOperator< DF, SolverParameter::cg > (solver/krylovinverseoperators.hh:46-281)
struct SolverParameter {                       // solver/parameter.hh:21-295, keys fem.solver.*
  double tolerance = 1e-8; int errorMeasure = B200FEM_TOL_ABSOLUTE; int maxIterations = 1000; bool verbose = false;
  int gmresRestart = 20;                        // fem.solver.gmres.restart (parameter.hh:197-201)
};
// KrylovInverseOperator< DF, method > (solver/krylovinverseoperators.hh:46-288): method cg -> linear/cg.hh, bicgstab -> linear/bicgstab.hh
enum SolverMethod { cg = 0, bicgstab = 1, gmres = 2 };
template <class DiscreteFunctionT, int method = cg>
class KrylovInverseOperator {
 public:
  typedef GalerkinOperator<DiscreteFunctionT> OperatorType;
  explicit KrylovInverseOperator(const SolverParameter& p = SolverParameter()) : parameter_(p) {}
  void bind(const OperatorType& op) { op_ = &op; }                                    // inverseoperatorinterface.hh:109-113
  void unbind() { op_ = nullptr; }
  void operator()(const DiscreteFunctionT& rhs, DiscreteFunctionT& x) const {        // inverseoperatorinterface.hh:81-84
    if (!op_) throw InvalidStateException("KrylovInverseOperator: no operator bound");
    residuals_.assign((std::size_t)std::max(parameter_.maxIterations, 1), 0.0);
    if (method == gmres)
      check(b200fem_gmres_solve(op_->handle(), rhs.leakPointer(), x.leakPointer(), parameter_.gmresRestart, parameter_.tolerance,
                                parameter_.maxIterations, parameter_.errorMeasure, &iterations_, residuals_.data()));
    else {
      auto solve = method == bicgstab ? b200fem_bicgstab_solve : b200fem_cg_solve;
      check(solve(op_->handle(), rhs.leakPointer(), x.leakPointer(), parameter_.tolerance, parameter_.maxIterations,
                  parameter_.errorMeasure, &iterations_, residuals_.data()));
    }
  }
  int iterations() const { return iterations_; }                                       // negative: not converged (linear/cg.hh:116, bicgstab.hh:208-211)
  bool converged() const { return iterations_ >= 0; }
  const std::vector<double>& residuals() const { return residuals_; }
  SolverParameter& parameter() { return parameter_; }
 private:
  SolverParameter parameter_; const OperatorType* op_ = nullptr; mutable int iterations_ = 0; mutable std::vector<double> residuals_;
};
// CG with "fem.solver.preconditioning.method: jacobi" (solver/linear/cg.hh:52-56, 72-107): B = diag(A)^-1 built matrix-free
template <class DiscreteFunctionT>
class JacobiCgInverseOperator {
 public:
  typedef GalerkinOperator<DiscreteFunctionT> OperatorType;
  explicit JacobiCgInverseOperator(const SolverParameter& p = SolverParameter()) : parameter_(p) {}
  void bind(const OperatorType& op) { op_ = &op; }
  void unbind() { op_ = nullptr; }
  void operator()(const DiscreteFunctionT& rhs, DiscreteFunctionT& x) const {
    if (!op_) throw InvalidStateException("JacobiCgInverseOperator: no operator bound");
    residuals_.assign((std::size_t)std::max(parameter_.maxIterations, 1), 0.0);
    check(b200fem_pcg_solve(op_->handle(), rhs.leakPointer(), x.leakPointer(), parameter_.tolerance, parameter_.maxIterations, parameter_.errorMeasure,
                            &iterations_, residuals_.data()));
  }
  int iterations() const { return iterations_; }
  bool converged() const { return iterations_ >= 0; }
 private:
  SolverParameter parameter_; const OperatorType* op_ = nullptr; mutable int iterations_ = 0; mutable std::vector<double> residuals_;
};

// Dune::Fem::NewtonInverseOperator (solver/newtoninverseoperator.hh:423-803): bind( op ); operator()( u, w ) solves L[w] = u from the
// initial guess in w.  Parameters fem.solver.nonlinear.* (NewtonParameter, :37-268)
struct NewtonParameter {
  double tolerance = 1e-6; int maxIterations = 0x7fffffff; bool simpleLineSearch = false;
  int linearMethod = gmres; SolverParameter linear;
};
enum class NewtonFailure { Success = 0, InvalidResidual = 1, IterationsExceeded = 2, LinearIterationsExceeded = 3, LineSearchFailed = 4, TooManyIterations = 5, TooManyLinearIterations = 6, LinearSolverFailed = 7 };
template <class DiscreteFunctionT>
class NewtonInverseOperator {
 public:
  typedef GalerkinOperator<DiscreteFunctionT> OperatorType;
  explicit NewtonInverseOperator(const NewtonParameter& p = NewtonParameter()) : parameter_(p) {}
  void bind(const OperatorType& op) { op_ = &op; }
  void unbind() { op_ = nullptr; }
  void operator()(const DiscreteFunctionT& u, DiscreteFunctionT& w) const { solve(u.leakPointer(), w); }
  void operator()(DiscreteFunctionT& w) const { solve(nullptr, w); }                     // L[w] = 0
  int iterations() const { return iterations_; }
  int linearIterations() const { return linearIterations_; }
  double residual() const { return residual_; }
  NewtonFailure failed() const { return static_cast<NewtonFailure>(failure_); }
  bool converged() const { return failure_ == 0; }
 private:
  void solve(const double* u, DiscreteFunctionT& w) const {
    if (!op_) throw InvalidStateException("NewtonInverseOperator: no operator bound");
    check(b200fem_newton_solve(op_->handle(), u, w.leakPointer(), parameter_.tolerance, parameter_.maxIterations, parameter_.linearMethod, parameter_.linear.tolerance,
                               parameter_.linear.maxIterations, parameter_.linear.errorMeasure, parameter_.linear.gmresRestart, parameter_.simpleLineSearch ? 1 : 0,
                               &iterations_, &linearIterations_, &residual_, &failure_));
  }
  NewtonParameter parameter_; const OperatorType* op_ = nullptr; mutable int iterations_ = 0, linearIterations_ = 0, failure_ = 0; mutable double residual_ = 0;
};

// Dune::Fem::FemScheme< Operator, InverseOperator > (schemes/femscheme.hh:60-260; Python dune.fem.scheme.galerkin): the Galerkin
// operator with its inverse operator.  solve( rhs, solution ): solution = g (+ rhs) on the Dirichlet boundary, then invOp( rhs, solution )
// (femscheme.hh:194-221, 248-254); returns the solver info (converged, nonlinear / linear iterations).
struct SolverInfo { bool converged; int nonlinearIterations, linearIterations; };
template <class DiscreteFunctionT>
class FemScheme {
 public:
  typedef typename DiscreteFunctionT::DiscreteFunctionSpaceType DiscreteFunctionSpaceType;
  typedef GalerkinOperator<DiscreteFunctionT> DifferentiableOperatorType;
  template <class IntegrandsT>
  FemScheme(const DiscreteFunctionSpaceType& space, const IntegrandsT& integrands, const NewtonParameter& parameter = NewtonParameter())
      : space_(space), op_(space, space, integrands), parameter_(parameter), mask_(space.size(), 0), values_(space.size(), 0.0) {
    if (b200fem_operator_dirichlet(op_.handle(), mask_.data(), values_.data()) != B200FEM_OK) mask_.assign(space.size(), 0);   // (no constraints)
  }
  void operator()(const DiscreteFunctionT& arg, DiscreteFunctionT& dest) const { op_(arg, dest); }
  void setConstraints(DiscreteFunctionT& u) const { for (std::size_t i = 0; i < mask_.size(); ++i) if (mask_[i]) u.dofVector()[i] = values_[i]; }
  SolverInfo solve(DiscreteFunctionT& solution) const { return solve(nullptr, solution); }
  SolverInfo solve(const DiscreteFunctionT& rhs, DiscreteFunctionT& solution) const { return solve(&rhs, solution); }
  const DiscreteFunctionSpaceType& space() const { return space_; }
  const DifferentiableOperatorType& fullOperator() const { return op_; }
 private:
  SolverInfo solve(const DiscreteFunctionT* rhs, DiscreteFunctionT& solution) const {
    setConstraints(solution);
    if (rhs) for (std::size_t i = 0; i < mask_.size(); ++i) if (mask_[i]) solution.dofVector()[i] += rhs->dofVector()[i];
    if (!op_.nonlinear()) {      // a linear operator: the one Newton step with the exact Jacobian, A x = b + rhs (newtoninverseoperator.hh:761, 791-792)
      DiscreteFunctionT b("b", space_); op_.loadVector(b);
      if (rhs) for (std::size_t i = 0; i < mask_.size(); ++i) b.dofVector()[i] = mask_[i] ? solution.dofVector()[i] : b.dofVector()[i] + rhs->dofVector()[i];
      int it = 0; std::vector<double> hist((std::size_t)std::max(parameter_.linear.maxIterations, 1));
      const SolverParameter& lp = parameter_.linear;
      if (parameter_.linearMethod == gmres) check(b200fem_gmres_solve(op_.handle(), b.leakPointer(), solution.leakPointer(), lp.gmresRestart, lp.tolerance, lp.maxIterations, lp.errorMeasure, &it, hist.data()));
      else check((parameter_.linearMethod == bicgstab ? b200fem_bicgstab_solve : b200fem_cg_solve)(op_.handle(), b.leakPointer(), solution.leakPointer(), lp.tolerance, lp.maxIterations, lp.errorMeasure, &it, hist.data()));
      return SolverInfo{it >= 0, 1, it >= 0 ? it : -it};
    }
    NewtonInverseOperator<DiscreteFunctionT> invOp(parameter_);
    invOp.bind(op_);
    if (rhs) invOp(*rhs, solution); else invOp(solution);
    return SolverInfo{invOp.converged(), invOp.iterations(), invOp.linearIterations()};
  }
  const DiscreteFunctionSpaceType& space_; DifferentiableOperatorType op_; NewtonParameter parameter_;
  std::vector<std::uint8_t> mask_; std::vector<double> values_;
};

template <class DiscreteFunctionT> using CgInverseOperator = KrylovInverseOperator<DiscreteFunctionT, cg>;                // krylovinverseoperators.hh:284
template <class DiscreteFunctionT> using BicgstabInverseOperator = KrylovInverseOperator<DiscreteFunctionT, bicgstab>;    // :288
template <class DiscreteFunctionT> using GmresInverseOperator = KrylovInverseOperator<DiscreteFunctionT, gmres>;          // :295

}  // namespace B200Fem
