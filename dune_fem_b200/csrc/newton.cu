// newton.cu -- NewtonInverseOperator (dune/fem/solver/newtoninverseoperator.hh:690-803) over the device-resident pieces:
// the caller directly above the Krylov loop (FemScheme::solve -> NewtonInverseOperator -> KrylovInverseOperator -> operator apply).
// The whole iteration stays on the device: residual = L[w] - u, the Jacobian is the difference quotient of
// AutomaticDifferenceLinearOperator (b200fem_operator_linearize_dev), the linear solve is one of the Krylov drivers of solvers.cu,
// w -= dw, optional "simple" line search (:588-629).  Only scalars (norms, iteration counts) cross to the host.
#include <cmath>
#include <limits>

#include "internal.hpp"

using namespace b200fem;

namespace {
int norm_dev(b200fem_operator* op, const double* x, double* out) { double s = 0; int rc = b200fem_dot_dev(op, x, x, &s); *out = std::sqrt(s); return rc; }
// NewtonFailure (newtoninverseoperator.hh:389-400, failed(): 568-584)
int newton_failed(double delta, int it, int maxit, int lit, int maxlit, bool step_completed) {
  if (!(delta < std::numeric_limits<double>::max()) || std::isnan(delta)) return 1;      // InvalidResidual
  if (it >= maxit) return 5;                                                             // TooManyIterations
  if (lit >= maxlit) return 6;                                                           // TooManyLinearIterations
  if (lit < 0) return 7;                                                                 // LinearSolverFailed
  if (!step_completed) return 4;                                                         // LineSearchFailed
  return 0;
}
}  // namespace

extern "C" int b200fem_newton_solve_dev(b200fem_operator* op, const double* u, double* w, double tolerance, int max_iterations, int linear_method,
                                        double linear_tolerance, int linear_max_iterations, int linear_tolerance_criteria, int gmres_restart,
                                        int line_search, int* iterations, int* linear_iterations, double* residual_norm, int* failure) {
  REQUIRE(op && w && iterations && linear_iterations && residual_norm && failure, B200FEM_ERR_INVALID, "newton: null argument");
  REQUIRE(linear_method >= 0 && linear_method <= 2, B200FEM_ERR_INVALID, "newton: linear method must be 0 (cg), 1 (bicgstab) or 2 (gmres)");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(c->device));
  if (!op->d_nw_res) { CUDA_OK(cudaMalloc(&op->d_nw_res, bytes)); CUDA_OK(cudaMalloc(&op->d_nw_dw, bytes)); }
  double* res = op->d_nw_res; double* dw = op->d_nw_dw;
  const bool nonlinear = op->jit != nullptr || op->model.gamma != 0.0;        // op_->nonlinear() (:696)
  auto eval_residual = [&]() -> int {                                          // residual = S[w] - u
    op->jac_mode = false;
    int rc = apply_dev_impl(op, w, res, false); if (rc) return rc;
    if (u) { rc = b200fem_axpy_dev(op, -1.0, u, res); if (rc) return rc; }
    return B200FEM_OK;
  };
  int it = 0, lit = 0; bool step_completed = true; double delta = 0;
  int rc = eval_residual(); if (rc) return rc;
  rc = norm_dev(op, res, &delta); if (rc) return rc;
  while (true) {
    rc = b200fem_operator_linearize_dev(op, w, 0.0); if (rc) return rc;        // (*op_).jacobian( w, jOp ) (:735)
    if (linear_max_iterations - lit <= 0) break;
    CUDA_OK(cudaMemsetAsync(dw, 0, bytes, st));                                // dw.clear()
    int li = 0; const int budget = linear_max_iterations - lit;
    if (linear_method == 0) rc = b200fem_cg_solve_dev(op, res, dw, linear_tolerance, budget, linear_tolerance_criteria, &li, nullptr);
    else if (linear_method == 1) rc = b200fem_bicgstab_solve_dev(op, res, dw, linear_tolerance, budget, linear_tolerance_criteria, &li, nullptr);
    else rc = b200fem_gmres_solve_dev(op, res, dw, gmres_restart, linear_tolerance, budget, linear_tolerance_criteria, &li, nullptr);
    if (rc) { op->jac_mode = false; return rc; }
    if (li < 0) { lit = li; break; }                                           // (:752-756)
    lit += li;
    rc = b200fem_axpy_dev(op, -1.0, dw, w); if (rc) return rc;                 // w -= dw
    if (!nonlinear) break;
    rc = eval_residual(); if (rc) return rc;
    // lineSearch (:588-629)
    const double delta_old = delta; int ls = 0;
    rc = norm_dev(op, res, &delta); if (rc) return rc;
    if (line_search) {
      if (newton_failed(delta, it, max_iterations, lit, linear_max_iterations, step_completed) == 1) {
        double test = 0; rc = b200fem_dot_dev(op, dw, dw, &test); if (rc) return rc;
        if (!(test < std::numeric_limits<double>::max() && !std::isnan(test))) delta = 2.0 * delta_old;
      }
      double factor = 1.0; ls = delta < delta_old ? 1 : 0; int lsit = 0;
      while (delta >= delta_old) {
        const double delta_prev = delta;
        factor *= 0.5;
        if (std::fabs(delta - delta_old) < 1e-5 * delta) { ls = -1; break; }
        rc = b200fem_axpy_dev(op, factor, dw, w); if (rc) return rc;
        rc = eval_residual(); if (rc) return rc;
        rc = norm_dev(op, res, &delta); if (rc) return rc;
        if (std::fabs(delta - delta_prev) < 1e-15) { ls = -1; break; }
        if (newton_failed(delta, it, max_iterations, lit, linear_max_iterations, step_completed) == 1) delta = 2.0 * delta_old;
        if (++lsit >= 1000) { ls = -1; break; }
      }
    }
    step_completed = ls >= 0;
    ++it;
    if (delta < tolerance || newton_failed(delta, it, max_iterations, lit, linear_max_iterations, step_completed) != 0) break;
  }
  op->jac_mode = false; op->state_version += 1;                                 // jInv_.unbind(): the operator is its plain self again
  *iterations = it; *linear_iterations = lit; *residual_norm = delta;
  *failure = newton_failed(delta, it, max_iterations, lit, linear_max_iterations, step_completed);
  return check_comm_error(c);
}

extern "C" int b200fem_newton_solve(b200fem_operator* op, const double* u_host, double* w_host, double tolerance, int max_iterations, int linear_method,
                                    double linear_tolerance, int linear_max_iterations, int linear_tolerance_criteria, int gmres_restart,
                                    int line_search, int* iterations, int* linear_iterations, double* residual_norm, int* failure) {
  REQUIRE(op && w_host, B200FEM_ERR_INVALID, "newton: null argument");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(c->device));
  if (!op->d_nw_w) { CUDA_OK(cudaMalloc(&op->d_nw_w, bytes)); CUDA_OK(cudaMalloc(&op->d_nw_u, bytes)); }
  CUDA_OK(cudaMemcpyAsync(op->d_nw_w, w_host, bytes, cudaMemcpyHostToDevice, st));
  if (u_host) CUDA_OK(cudaMemcpyAsync(op->d_nw_u, u_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = b200fem_newton_solve_dev(op, u_host ? op->d_nw_u : nullptr, op->d_nw_w, tolerance, max_iterations, linear_method, linear_tolerance, linear_max_iterations,
                                    linear_tolerance_criteria, gmres_restart, line_search, iterations, linear_iterations, residual_norm, failure);
  if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(w_host, op->d_nw_w, bytes, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  return B200FEM_OK;
}
