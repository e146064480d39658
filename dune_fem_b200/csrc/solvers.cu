// solvers.cu -- device-resident Krylov solvers on the matrix-free operator (solver/linear/{cg,bicgstab,gmres}.hh,
// solver/krylovinverseoperators.hh), BLAS-1 entry points, the difference-quotient Jacobian.
#include <algorithm>
#include <cmath>

#include "internal.hpp"
#include "vec_kernels.cuh"

using namespace b200fem;

namespace b200fem {

static PeerScalarsDev scalars_dev(b200fem_ctx* c) {
  if (c->world > 1 && c->scalars.ok) return c->scalars.dev;
  PeerScalarsDev A; std::memset(&A, 0, sizeof(A)); A.world = 1; return A;
}
// true when global sums run inside the reduction kernels (one rank, or peer-memory all-reduce): no library call in an iteration
static bool fused_sums(b200fem_ctx* c) { return c->world == 1 || c->scalars.ok; }
// true when an apply contains no library call either (peer-memory halo exchange)
static bool graphable(b200fem_operator* op) {
  b200fem_ctx* c = op->sp->mesh->ctx;
  if (c->world == 1) return true;
  if (!c->scalars.ok) return false;
  return op->sp->kind == B200FEM_LAGRANGE ? op->halo_add.built : op->halo_p2p.built;
}

int negate_dev(double* x, long long n, cudaStream_t st) { negate_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, n); CUDA_OK(cudaGetLastError()); return B200FEM_OK; }
int dirichlet_sub_dev(const double* u, double* w, const uint8_t* mask, const double* vals, long long n, cudaStream_t st) {
  dirichlet_sub_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(u, w, mask, vals, n); CUDA_OK(cudaGetLastError()); return B200FEM_OK;
}

// second reduction stage of `count` partial arrays (kRedBlocks entries each, contiguous) + global sum -> out[0..count)
static int reduce_to(b200fem_operator* op, const double* partial, int count, double* out) {
  b200fem_ctx* c = op->sp->mesh->ctx; cudaStream_t st = c->stream;
  if (fused_sums(c)) {
    for (int k0 = 0; k0 < count; k0 += kArMax)
      reduce_final_allreduce_kernel<<<1, kRedThreads, 0, st>>>(partial + (size_t)k0 * kRedBlocks, kRedBlocks, std::min(kArMax, count - k0), out + k0, scalars_dev(c));
    CUDA_OK(cudaGetLastError()); return B200FEM_OK;
  }
  for (int i = 0; i < count; ++i) reduce_final_kernel<<<1, kRedThreads, 0, st>>>(partial + (size_t)i * kRedBlocks, kRedBlocks, out + i);
  CUDA_OK(cudaGetLastError());
  if (c->nccl.AllReduce(out, out, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, st) != 0) return fail(B200FEM_ERR_COMM, "ncclAllReduce failed");
  return B200FEM_OK;
}
// dot over primary dofs + global sum (function/common/scalarproducts.hh:115-127)
int reduce_sums(b200fem_operator* op, int count) { return reduce_to(op, op->d_partial, count, op->d_sums); }

int ensure_cg_buffers(b200fem_operator* op, int maxit) {
  const size_t bytes = sizeof(double) * (size_t)op->sp->size;
  if (!op->d_h) { CUDA_OK(cudaMalloc(&op->d_h, bytes)); CUDA_OK(cudaMalloc(&op->d_r, bytes)); CUDA_OK(cudaMalloc(&op->d_p, bytes)); }
  if (!op->d_partial) { CUDA_OK(cudaMalloc(&op->d_partial, sizeof(double) * 2 * kRedBlocks)); CUDA_OK(cudaMalloc(&op->d_sums, sizeof(double) * 4)); CUDA_OK(cudaMalloc(&op->d_cg, sizeof(CgState))); CUDA_OK(cudaMalloc(&op->d_counter, 2 * sizeof(unsigned int))); CUDA_OK(cudaMemset(op->d_counter, 0, 2 * sizeof(unsigned int))); }
  if (maxit > op->hist_cap) { if (op->d_hist) cudaFree(op->d_hist); CUDA_OK(cudaMalloc(&op->d_hist, sizeof(double) * (size_t)std::max(maxit, 1))); op->hist_cap = std::max(maxit, 1); }
  return B200FEM_OK;
}

// AutomaticDifferenceLinearOperator::operator() (automaticdifferenceoperator.hh:124-149), everything on the device and on the
// operator's stream (no host round trip: the difference quotient can sit inside a captured CG graph)
int apply_fd_jacobian(b200fem_operator* op, const double* arg, double* dest) {
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const long long n = s->size;
  int rc = ensure_cg_buffers(op, 1); if (rc) return rc;
  dot_partial_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(arg, arg, op->d_aux, n, op->d_partial + kRedBlocks);          // arg.normSquaredDofs()
  rc = reduce_to(op, op->d_partial + kRedBlocks, 1, op->d_sums + 3); if (rc) return rc;
  fd_eps_kernel<<<1, 32, 0, st>>>(op->d_sums + 3, op->d_fd);
  fd_perturb_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_jac_b, op->d_jac_u, arg, n, op->d_fd);
  const bool want_dot = op->want_dot; op->want_dot = false;   // (a fused <u, w> of the perturbed apply is not <arg, J arg>)
  op->jac_mode = false; rc = apply_dev_impl(op, op->d_jac_b, dest, false); op->jac_mode = true; op->want_dot = want_dot; op->dot_parts = 0; if (rc) return rc;   // (*op_)(b_, dest)
  fd_quotient_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(dest, op->d_jac_opu, n, op->d_fd);
  CUDA_OK(cudaGetLastError()); return B200FEM_OK;
}

}  // namespace b200fem

extern "C" int b200fem_operator_linearize_dev(b200fem_operator* op, const double* u, double eps) {
  REQUIRE(op, B200FEM_ERR_INVALID, "linearize: null operator");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const long long n = s->size; const size_t bytes = sizeof(double) * (size_t)n;
  CUDA_OK(cudaSetDevice(c->device));
  op->state_version += 1;                                                                  // a captured iteration applied another operator
  if (!u) { op->jac_mode = false; return B200FEM_OK; }
  int rc = ensure_cg_buffers(op, 1); if (rc) return rc;
  if (!op->d_jac_u) { CUDA_OK(cudaMalloc(&op->d_jac_u, bytes)); CUDA_OK(cudaMalloc(&op->d_jac_opu, bytes)); CUDA_OK(cudaMalloc(&op->d_jac_b, bytes)); CUDA_OK(cudaMalloc(&op->d_fd, sizeof(FdState))); }
  // jOp.set(u, op, eps) (automaticdifferenceoperator.hh:152-166): u_, op_u_ = op(u), norm_u_ = sqrt(u.u) when eps is dynamic
  if (u != op->d_jac_u) CUDA_OK(cudaMemcpyAsync(op->d_jac_u, u, bytes, cudaMemcpyDeviceToDevice, st));
  op->jac_mode = false;
  rc = apply_dev_impl(op, op->d_jac_u, op->d_jac_opu, false); if (rc) return rc;
  FdState h{}; h.eps_given = eps; h.norm_u = 0; h.eps = eps;
  if (eps <= 0) { double uu = 0; rc = b200fem_dot_dev(op, op->d_jac_u, op->d_jac_u, &uu); if (rc) return rc; h.norm_u = std::sqrt(uu); }
  CUDA_OK(cudaMemcpyAsync(op->d_fd, &h, sizeof(FdState), cudaMemcpyHostToDevice, st)); CUDA_OK(cudaStreamSynchronize(st));
  op->jac_mode = true;
  return B200FEM_OK;
}
extern "C" int b200fem_operator_linearize(b200fem_operator* op, const double* u_host, double eps) {
  REQUIRE(op, B200FEM_ERR_INVALID, "linearize: null operator");
  if (!u_host) return b200fem_operator_linearize_dev(op, nullptr, eps);
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  if (!op->d_jac_u) { CUDA_OK(cudaMalloc(&op->d_jac_u, bytes)); CUDA_OK(cudaMalloc(&op->d_jac_opu, bytes)); CUDA_OK(cudaMalloc(&op->d_jac_b, bytes)); CUDA_OK(cudaMalloc(&op->d_fd, sizeof(FdState))); }
  CUDA_OK(cudaMemcpyAsync(op->d_jac_u, u_host, bytes, cudaMemcpyHostToDevice, st));
  return b200fem_operator_linearize_dev(op, op->d_jac_u, eps);
}

extern "C" int b200fem_dot_dev(b200fem_operator* op, const double* x, const double* y, double* result) {
  REQUIRE(op && x && y && result, B200FEM_ERR_INVALID, "dot: null argument");
  b200fem_ctx* c = op->sp->mesh->ctx; CUDA_OK(cudaSetDevice(c->device));
  int rc = ensure_cg_buffers(op, 1); if (rc) return rc;
  dot_partial_kernel<<<kRedBlocks, kRedThreads, 0, c->stream>>>(x, y, op->d_aux, op->sp->size, op->d_partial);
  rc = reduce_sums(op, 1); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(result, op->d_sums, sizeof(double), cudaMemcpyDeviceToHost, c->stream)); CUDA_OK(cudaStreamSynchronize(c->stream));
  return check_comm_error(c);
}
extern "C" int b200fem_axpy_dev(b200fem_operator* op, double alpha, const double* x, double* y) {
  REQUIRE(op && x && y, B200FEM_ERR_INVALID, "axpy: null argument");
  axpy_kernel<<<kRedBlocks, kRedThreads, 0, op->sp->mesh->ctx->stream>>>(alpha, x, y, op->sp->size); CUDA_OK(cudaGetLastError()); return B200FEM_OK;
}

// LinearSolver::cg (solver/linear/cg.hh:18-117), unpreconditioned, on the homogeneous linear part of the operator
extern "C" int b200fem_cg_solve_dev(b200fem_operator* op, const double* b, double* x, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b && x && iterations, B200FEM_ERR_INVALID, "cg: null argument");
  REQUIRE(tolcrit >= 0 && tolcrit <= 2, B200FEM_ERR_INVALID, "cg: unknown tolerance criterion");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const long long n = s->size;
  CUDA_OK(cudaSetDevice(c->device));
  int rc = ensure_cg_buffers(op, maxit); if (rc) return rc;
  CgState init{}; init.epsilon = epsilon; init.max_iterations = maxit; init.tol_criteria = tolcrit;
  CUDA_OK(cudaMemcpyAsync(op->d_cg, &init, sizeof(CgState), cudaMemcpyHostToDevice, st));
  op->want_dot = c->world == 1;                        // (also allocates the partial buffer of the fused <q,h> before any graph capture)
  rc = apply_dev_impl(op, x, op->d_h, true); op->want_dot = false; if (rc) return rc;                          // h = A x
  cg_init_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_h, b, op->d_r, op->d_p, op->d_aux, n, op->d_partial, op->d_partial + kRedBlocks);
  rc = reduce_sums(op, 2); if (rc) return rc;
  cg_init_final_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg);
  CgState host{}; const int chunk = 16;
  // One CG iteration, enqueued on the stream: 4 launches (the block that finishes a reduction last also does the second
  // stage, the global sum over peer memory on several ranks, and the scalar update) + the halo exchange on several ranks.
  // Without peer memory the partial sums go through ncclAllReduce between two kernels.
  const bool single = c->world == 1, fused = fused_sums(c);
  const PeerScalarsDev A = scalars_dev(c);
  auto enqueue_iteration = [&]() -> int {
    cg_update_p_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_p, op->d_r, n, op->d_cg);                      // no-op in iteration 0
    op->want_dot = single; op->dot_parts = 0;
    int e = apply_dev_impl(op, op->d_p, op->d_h, true); op->want_dot = false; if (e) return e;                  // h = A q (+ <q,h> partials when the kernel can)
    if (fused) {
      if (op->dot_parts > 0) cg_alpha_partials_kernel<<<1, kRedThreads, 0, st>>>(op->d_dot_partial, op->dot_parts, op->d_cg);
      else cg_dot_alpha_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_p, op->d_h, op->d_aux, n, op->d_partial, op->d_cg, op->d_counter, A);
      cg_update_xr_residual_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, op->d_r, op->d_p, op->d_h, op->d_aux, n, op->d_partial, op->d_cg, op->d_hist, op->d_counter + 1, A);
    } else {
      cg_dot_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_p, op->d_h, op->d_aux, n, op->d_partial, op->d_cg);
      e = reduce_sums(op, 1); if (e) return e;
      cg_alpha_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg);
      cg_update_xr_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, op->d_r, op->d_p, op->d_h, op->d_aux, n, op->d_partial, op->d_cg);
      e = reduce_sums(op, 1); if (e) return e;
      cg_residual_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg, op->d_hist);
    }
    return B200FEM_OK;
  };
  // A chunk of 16 iterations is captured once into a CUDA graph and replayed (launch-bound sizes spend their time in launch
  // gaps otherwise).  Iterations past convergence / max_iterations are no-ops on the device (every kernel checks the
  // device-resident `done` flag), so whole chunks can always be replayed.  On several ranks all sequence numbers of the
  // exchange protocols live in device memory, so every rank replays its own graph; the residuals -- and therefore the
  // decision to stop -- are bit-identical on all ranks.
  bool use_graph = graphable(op) && maxit >= chunk;
  // Launch-bound sizes on a 2-D Lagrange lattice (BASELINE config 1): a chunk of iterations is ONE cooperative launch with
  // grid-wide barriers instead of kernel boundaries (cg_coop2d.cuh).
  int coop_grid = 0;
  const bool use_coop = single && !op->jac_mode && s->kind == B200FEM_LAGRANGE && !s->unst && s->order <= 2 && s->box.dim == 2 && op->model.gamma == 0.0 && !op->model.has_skeleton &&
                        default_quadrature(op) && n <= (1 << 20) && (!op->model.strong_dirichlet || op->d_dmask) && op->kernel_pref != B200FEM_KERNEL_QUADRATURE;
  if (use_coop) { rc = coop_cg_chunk(op, x, 0, &coop_grid); if (rc) return rc; if (coop_grid > 0) use_graph = false; }
  if (use_graph && !(op->cg_graph && op->cg_graph_version == op->state_version && op->cg_graph_key[0] == (const void*)x && op->cg_graph_key[1] == (const void*)b && op->cg_graph_key[2] == (const void*)op->d_hist)) {
    if (op->cg_graph) { cudaGraphExecDestroy(op->cg_graph); op->cg_graph = nullptr; }
    cudaGraph_t graph = nullptr;
    CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    op->capturing = true; int e = B200FEM_OK;
    for (int k = 0; k < chunk && !e; ++k) e = enqueue_iteration();
    op->capturing = false;
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (e) { if (graph) cudaGraphDestroy(graph); return e; }
    if (ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail(B200FEM_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce)); }
    ce = cudaGraphInstantiate(&op->cg_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { op->cg_graph = nullptr; return fail(B200FEM_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
    op->cg_graph_key[0] = x; op->cg_graph_key[1] = b; op->cg_graph_key[2] = op->d_hist; op->cg_graph_version = op->state_version;
  }
  for (int it = 0; it < maxit;) {
    const int upto = std::min(maxit, it + chunk);
    if (coop_grid > 0) { rc = coop_cg_chunk(op, x, chunk, &coop_grid); if (rc) return rc; it += chunk; }
    else if (use_graph) { CUDA_OK(cudaGraphLaunch(op->cg_graph, st)); it += chunk; }
    else for (; it < upto; ++it) { rc = enqueue_iteration(); if (rc) return rc; }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(&host, op->d_cg, sizeof(CgState), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
    rc = check_comm_error(c); if (rc) return rc;
    if (host.done) break;
  }
  if (maxit <= 0) { CUDA_OK(cudaMemcpyAsync(&host, op->d_cg, sizeof(CgState), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st)); }
  REQUIRE(std::isfinite(host.residual), B200FEM_ERR_INVALID, "cg: residual is not finite (alpha/beta NaN, cf. cg.hh:74,91)");
  if (history && host.iterations > 0) { CUDA_OK(cudaMemcpyAsync(history, op->d_hist, sizeof(double) * host.iterations, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st)); }
  *iterations = (host.iterations < maxit) ? host.iterations : -host.iterations;                                // cg.hh:116
  return B200FEM_OK;
}
extern "C" int b200fem_cg_solve(b200fem_operator* op, const double* b_host, double* x_host, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b_host && x_host && iterations, B200FEM_ERR_INVALID, "cg: null argument");
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  if (!op->d_x) { CUDA_OK(cudaMalloc(&op->d_x, bytes)); CUDA_OK(cudaMalloc(&op->d_b, bytes)); }
  CUDA_OK(cudaMemcpyAsync(op->d_x, x_host, bytes, cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(op->d_b, b_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = b200fem_cg_solve_dev(op, op->d_b, op->d_x, epsilon, maxit, tolcrit, iterations, history); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(x_host, op->d_x, bytes, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  return B200FEM_OK;
}

// LinearSolver::cg, preconditioned branch (solver/linear/cg.hh:52-56, 72-107) with the Jacobi preconditioner
int host_diagonal(b200fem_operator* op, std::vector<double>& diag, bool dirichlet_rows);     // capi.cu
extern "C" int b200fem_pcg_solve_dev(b200fem_operator* op, const double* b, double* x, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b && x && iterations, B200FEM_ERR_INVALID, "pcg: null argument");
  REQUIRE(tolcrit >= 0 && tolcrit <= 2, B200FEM_ERR_INVALID, "pcg: unknown tolerance criterion");
  REQUIRE(!op->jac_mode, B200FEM_ERR_NOT_IMPLEMENTED, "pcg: no diagonal for a difference-quotient linearisation");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const long long n = s->size; const size_t bytes = sizeof(double) * (size_t)n;
  CUDA_OK(cudaSetDevice(c->device));
  int rc = ensure_cg_buffers(op, maxit); if (rc) return rc;
  if (!op->d_dinv || op->dinv_mass != op->inverse_mass || op->dinv_version != op->state_version) {
    // several ranks, continuous space: interface nodes hold partial sums (like the apply), completed by the Add exchange before
    // the Dirichlet rows are set to one
    const bool shared_nodes = c->world > 1 && s->kind == B200FEM_LAGRANGE;
    std::vector<double> d; rc = host_diagonal(op, d, !shared_nodes); if (rc) return rc;
    if (!op->d_dinv) { CUDA_OK(cudaMalloc(&op->d_dinv, bytes)); CUDA_OK(cudaMalloc(&op->d_pq, bytes)); CUDA_OK(cudaMalloc(&op->d_ps, bytes)); }
    CUDA_OK(cudaMemcpyAsync(op->d_dinv, d.data(), bytes, cudaMemcpyHostToDevice, st));
    if (shared_nodes) {
      rc = exchange(op, op->d_dinv, st); if (rc) return rc;
      if (op->d_dmask) set_masked_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_dinv, op->d_dmask, 1.0, n);
    }
    invert_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_dinv, n);
    CUDA_OK(cudaStreamSynchronize(st)); op->dinv_mass = op->inverse_mass; op->dinv_version = op->state_version;
  }
  double* p = op->d_p; double* q = op->d_pq; double* sv = op->d_ps; double* h = op->d_h;
  CgState init{}; init.epsilon = epsilon; init.max_iterations = maxit; init.tol_criteria = tolcrit;
  CUDA_OK(cudaMemcpyAsync(op->d_cg, &init, sizeof(CgState), cudaMemcpyHostToDevice, st));
  rc = apply_dev_impl(op, x, h, true); if (rc) return rc;                                                       // h = A x
  pcg_init_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(h, b, op->d_dinv, p, q, sv, op->d_aux, n, op->d_partial, op->d_partial + kRedBlocks);
  rc = reduce_sums(op, 2); if (rc) return rc;
  cg_init_final_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg);
  const bool fused = fused_sums(c); const PeerScalarsDev A = scalars_dev(c);
  CgState host{}; const int chunk = 16;
  for (int it = 0; it < maxit;) {
    const int upto = std::min(maxit, it + chunk);
    for (; it < upto; ++it) {
      pcg_update_q_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(q, sv, n, op->d_cg);
      rc = apply_dev_impl(op, q, h, true); if (rc) return rc;                                                   // h = A q
      if (fused) {
        cg_dot_alpha_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(q, h, op->d_aux, n, op->d_partial, op->d_cg, op->d_counter, A);
        pcg_update_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, p, sv, q, h, op->d_dinv, op->d_aux, n, op->d_partial, op->d_cg, op->d_hist, op->d_counter + 1, A);
      } else {
        cg_dot_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(q, h, op->d_aux, n, op->d_partial, op->d_cg);
        rc = reduce_sums(op, 1); if (rc) return rc;
        cg_alpha_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg);
        pcg_update_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, p, sv, q, h, op->d_dinv, op->d_aux, n, op->d_partial, op->d_cg, nullptr, nullptr, A);
        rc = reduce_sums(op, 1); if (rc) return rc;
        cg_residual_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg, op->d_hist);
      }
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(&host, op->d_cg, sizeof(CgState), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
    rc = check_comm_error(c); if (rc) return rc;
    if (host.done) break;
  }
  if (maxit <= 0) { CUDA_OK(cudaMemcpyAsync(&host, op->d_cg, sizeof(CgState), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st)); }
  REQUIRE(std::isfinite(host.residual), B200FEM_ERR_INVALID, "pcg: residual is not finite");
  if (history && host.iterations > 0) { CUDA_OK(cudaMemcpyAsync(history, op->d_hist, sizeof(double) * host.iterations, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st)); }
  *iterations = (host.iterations < maxit) ? host.iterations : -host.iterations;
  return B200FEM_OK;
}
extern "C" int b200fem_pcg_solve(b200fem_operator* op, const double* b_host, double* x_host, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b_host && x_host && iterations, B200FEM_ERR_INVALID, "pcg: null argument");
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  if (!op->d_x) { CUDA_OK(cudaMalloc(&op->d_x, bytes)); CUDA_OK(cudaMalloc(&op->d_b, bytes)); }
  CUDA_OK(cudaMemcpyAsync(op->d_x, x_host, bytes, cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(op->d_b, b_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = b200fem_pcg_solve_dev(op, op->d_b, op->d_x, epsilon, maxit, tolcrit, iterations, history); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(x_host, op->d_x, bytes, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  return B200FEM_OK;
}

// LinearSolver::gmres (solver/linear/gmres.hh:117-301), unpreconditioned, on the homogeneous linear part of the operator
extern "C" int b200fem_gmres_solve_dev(b200fem_operator* op, const double* b, double* u, int m, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b && u && iterations, B200FEM_ERR_INVALID, "gmres: null argument");
  REQUIRE(tolcrit >= 0 && tolcrit <= 2, B200FEM_ERR_INVALID, "gmres: unknown tolerance criterion");
  REQUIRE(m >= 1 && m <= 200, B200FEM_ERR_INVALID, "gmres: restart must be in [1, 200]");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const long long n = s->size;
  CUDA_OK(cudaSetDevice(c->device));
  int rc0 = ensure_cg_buffers(op, 1); if (rc0) return rc0;
  const size_t bytes = sizeof(double) * (size_t)n;
  while ((int)op->gmres_v.size() < m + 1) { double* q = nullptr; CUDA_OK(cudaMalloc(&q, bytes)); op->gmres_v.push_back(q); }
  if (op->gm_cap < m + 2) {
    if (op->d_gm_partial) cudaFree(op->d_gm_partial); if (op->d_gm_sums) cudaFree(op->d_gm_sums);
    CUDA_OK(cudaMalloc(&op->d_gm_partial, sizeof(double) * (size_t)(m + 2) * kRedBlocks)); CUDA_OK(cudaMalloc(&op->d_gm_sums, sizeof(double) * (size_t)(m + 2)));
    op->gm_cap = m + 2;
  }
  std::vector<double*>& v = op->gmres_v;
  // device scalar products of `count` (vector, v_l) pairs -> d_gm_sums[offset ..], globally reduced
  auto reduce = [&](int offset, int count) -> int { return reduce_to(op, op->d_gm_partial + (size_t)offset * kRedBlocks, count, op->d_gm_sums + offset); };
  auto norm2 = [&](const double* x, double* out) -> int {            // <x,x> over primary dofs, on the host
    dot_partial_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, x, op->d_aux, n, op->d_gm_partial);
    int e = reduce(0, 1); if (e) return e;
    CUDA_OK(cudaMemcpyAsync(out, op->d_gm_sums, sizeof(double), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
    return check_comm_error(c);
  };
  std::vector<double> H((size_t)(m + 1) * m, 0.0), g(m + 1, 0.0), sn(m, 0.0), cs(m, 0.0), y(m + 1, 0.0), gd(m + 2, 0.0);
  auto Hm = [&](int i, int j) -> double& { return H[(size_t)i * m + j]; };
  auto rotate = [](double& x, double& yy, double cc, double ss) { const double _x = x, _y = yy; x = cc * _x + ss * _y; yy = cc * _y - ss * _x; };
  double tol = epsilon, t = 0;
  int rc;
  if (tolcrit == B200FEM_TOL_RELATIVE) { rc = norm2(b, &t); if (rc) return rc; tol *= std::sqrt(t); }
  int it = 0;
  while (true) {
    rc = apply_dev_impl(op, u, v[0], true); if (rc) return rc;                                                  // v0 = A u - b
    axpy_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(-1.0, b, v[0], n);
    rc = norm2(v[0], &t); if (rc) return rc;
    const double res = std::sqrt(t);
    REQUIRE(std::isfinite(res), B200FEM_ERR_INVALID, "gmres: residual is not finite");
    if (tolcrit == B200FEM_TOL_RESIDUAL_REDUCTION && it == 0) tol *= res;
    if (res <= tol * (1 + 1e-15)) break;
    g[0] = -res; for (int i = 1; i <= m; ++i) g[i] = 0.0;
    scale_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(v[0], 1.0 / res, n);
    for (int j = 0; j < m; ++j) {
      double* vjp = v[j + 1];
      rc = apply_dev_impl(op, v[j], vjp, true); if (rc) return rc;
      // classical Gram-Schmidt: all j+1 scalar products of vjp in one (chunked) sweep, then the axpys, then the norm -- the
      // coefficients never leave the device; ONE device->host copy per iteration brings H(0..j, j) and H(j+1, j)^2
      for (int l0 = 0; l0 <= j; l0 += kGemvChunk) {
        GmresVecs V; const int cnt = std::min(kGemvChunk, j + 1 - l0); for (int q = 0; q < kGemvChunk; ++q) V.v[q] = v[std::min(l0 + q, j)];
        gmres_gemv_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(vjp, V, cnt, op->d_aux, n, op->d_gm_partial + (size_t)(1 + l0) * kRedBlocks);
      }
      rc = reduce(1, j + 1); if (rc) return rc;
      for (int l0 = 0; l0 <= j; l0 += kGemvChunk) {
        GmresVecs V; const int cnt = std::min(kGemvChunk, j + 1 - l0); for (int q = 0; q < kGemvChunk; ++q) V.v[q] = v[std::min(l0 + q, j)];
        gmres_axpys_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(vjp, V, cnt, op->d_gm_sums + 1 + l0, -1.0, n);
      }
      dot_partial_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(vjp, vjp, op->d_aux, n, op->d_gm_partial);
      rc = reduce(0, 1); if (rc) return rc;
      CUDA_OK(cudaMemcpyAsync(gd.data(), op->d_gm_sums, sizeof(double) * (size_t)(j + 2), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
      rc = check_comm_error(c); if (rc) return rc;
      for (int i = 0; i <= j; ++i) Hm(i, j) = gd[1 + i];
      Hm(j + 1, j) = std::sqrt(gd[0]);
      scale_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(vjp, 1.0 / Hm(j + 1, j), n);
      for (int i = 0; i < j; ++i) rotate(Hm(i + 1, j), Hm(i, j), cs[i], sn[i]);                                 // Givens rotations, gmres.hh:227-239
      const double hjj = Hm(j, j), hjpj = Hm(j + 1, j), nrm = std::sqrt(hjj * hjj + hjpj * hjpj);
      cs[j] = hjj / nrm; sn[j] = -hjpj / nrm;
      rotate(Hm(j + 1, j), Hm(j, j), cs[j], sn[j]);
      rotate(g[j + 1], g[j], cs[j], sn[j]);
      REQUIRE(std::isfinite(g[j + 1]), B200FEM_ERR_INVALID, "gmres: breakdown (non-finite Hessenberg entry)");
      if (history && it < std::max(maxit, 1)) history[it] = std::fabs(g[j + 1]);
      ++it;
      if (std::fabs(g[j + 1]) < tol || it >= maxit) break;
    }
    int last = it % m; if (last == 0) last = m;
    for (int i = last - 1; i >= 0; --i) {                                                                       // back substitution, :255-260
      double d = 0; for (int k = 0; k < last - (i + 1); ++k) d += Hm(i, i + 1 + k) * y[i + 1 + k];
      y[i] = (g[i] - d) / Hm(i, i);
    }
    CUDA_OK(cudaMemcpyAsync(op->d_gm_sums, y.data(), sizeof(double) * (size_t)last, cudaMemcpyHostToDevice, st));
    for (int l0 = 0; l0 < last; l0 += kGemvChunk) {                                                             // u += (v_0 .. v_last-1) y
      GmresVecs V; const int cnt = std::min(kGemvChunk, last - l0); for (int q = 0; q < kGemvChunk; ++q) V.v[q] = v[std::min(l0 + q, last - 1)];
      gmres_axpys_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(u, V, cnt, op->d_gm_sums + l0, 1.0, n);
    }
    CUDA_OK(cudaStreamSynchronize(st));          // y is a host vector that is rewritten in the next cycle
    if (std::fabs(g[last]) < tol || it >= maxit) break;
  }
  CUDA_OK(cudaGetLastError());
  *iterations = (it < maxit) ? it : -it;
  return B200FEM_OK;
}
extern "C" int b200fem_gmres_solve(b200fem_operator* op, const double* b_host, double* x_host, int restart, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b_host && x_host && iterations, B200FEM_ERR_INVALID, "gmres: null argument");
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  if (!op->d_x) { CUDA_OK(cudaMalloc(&op->d_x, bytes)); CUDA_OK(cudaMalloc(&op->d_b, bytes)); }
  CUDA_OK(cudaMemcpyAsync(op->d_x, x_host, bytes, cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(op->d_b, b_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = b200fem_gmres_solve_dev(op, op->d_b, op->d_x, restart, epsilon, maxit, tolcrit, iterations, history); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(x_host, op->d_x, bytes, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  return B200FEM_OK;
}

// LinearSolver::bicgstab (solver/linear/bicgstab.hh:64-214), unpreconditioned, on the homogeneous linear part of the operator
extern "C" int b200fem_bicgstab_solve_dev(b200fem_operator* op, const double* b, double* x, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b && x && iterations, B200FEM_ERR_INVALID, "bicgstab: null argument");
  REQUIRE(tolcrit >= 0 && tolcrit <= 2, B200FEM_ERR_INVALID, "bicgstab: unknown tolerance criterion");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const long long n = s->size;
  CUDA_OK(cudaSetDevice(c->device));
  int rc = ensure_cg_buffers(op, std::max(maxit, 1)); if (rc) return rc;
  const size_t bytes = sizeof(double) * (size_t)n;
  if (!op->d_rstar) {
    CUDA_OK(cudaMalloc(&op->d_rstar, bytes)); CUDA_OK(cudaMalloc(&op->d_s, bytes)); CUDA_OK(cudaMalloc(&op->d_tmp, bytes));
    CUDA_OK(cudaMalloc(&op->d_partial5, sizeof(double) * 5 * kRedBlocks)); CUDA_OK(cudaMalloc(&op->d_sums5, sizeof(double) * 8)); CUDA_OK(cudaMalloc(&op->d_bicg, sizeof(BicgState)));
  }
  double* r = op->d_r; double* p = op->d_p; double* rstar = op->d_rstar; double* sv = op->d_s; double* tmp = op->d_tmp;
  BicgState init{}; init.epsilon = epsilon; init.max_iterations = maxit; init.tol_criteria = tolcrit;
  CUDA_OK(cudaMemcpyAsync(op->d_bicg, &init, sizeof(BicgState), cudaMemcpyHostToDevice, st));
  rc = apply_dev_impl(op, x, r, true); if (rc) return rc;                                                      // r = A x
  bicg_init_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(r, b, p, rstar, op->d_aux, n, op->d_partial, op->d_partial + kRedBlocks);
  rc = reduce_sums(op, 2); if (rc) return rc;
  bicg_init_final_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_bicg);
  const bool fused = fused_sums(c); const PeerScalarsDev A = scalars_dev(c);
  BicgState host{}; const int chunk = 8; int issued = 0;
  do {
    for (int k = 0; k < chunk; ++k, ++issued) {
      rc = apply_dev_impl(op, p, tmp, true); if (rc) return rc;                                                // tmp = A p
      if (fused) bicg_dot_alpha_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(tmp, rstar, op->d_aux, n, op->d_partial, op->d_bicg, op->d_counter, A);
      else {
        bicg_dot_alpha_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(tmp, rstar, op->d_aux, n, op->d_partial, op->d_bicg, nullptr, A);
        rc = reduce_sums(op, 1); if (rc) return rc;
        bicg_alpha_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_bicg);
      }
      bicg_s_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(sv, r, tmp, n, op->d_bicg);
      rc = apply_dev_impl(op, sv, r, true); if (rc) return rc;                                                 // r = A s
      if (fused) bicg_dots5_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(r, sv, rstar, op->d_aux, n, op->d_partial5, op->d_bicg, op->d_hist, op->d_counter + 1, A);
      else {
        bicg_dots5_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(r, sv, rstar, op->d_aux, n, op->d_partial5, op->d_bicg, op->d_hist, nullptr, A);
        rc = reduce_to(op, op->d_partial5, 5, op->d_sums5); if (rc) return rc;
        bicg_scalars_kernel<<<1, 32, 0, st>>>(op->d_sums5, op->d_bicg, op->d_hist);
      }
      bicg_update_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, r, p, sv, tmp, n, op->d_bicg, op->d_counter);
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(&host, op->d_bicg, sizeof(BicgState), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
    rc = check_comm_error(c); if (rc) return rc;
  } while (!host.done);
  REQUIRE(std::isfinite(host.res), B200FEM_ERR_INVALID, "bicgstab: residual is not finite (breakdown: <tmp,r*> or <r,r> vanished)");
  if (history && host.iterations > 0) { CUDA_OK(cudaMemcpyAsync(history, op->d_hist, sizeof(double) * host.iterations, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st)); }
  *iterations = (host.iterations >= maxit) ? -host.iterations : host.iterations;                               // bicgstab.hh:208-211
  return B200FEM_OK;
}
extern "C" int b200fem_bicgstab_solve(b200fem_operator* op, const double* b_host, double* x_host, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b_host && x_host && iterations, B200FEM_ERR_INVALID, "bicgstab: null argument");
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  if (!op->d_x) { CUDA_OK(cudaMalloc(&op->d_x, bytes)); CUDA_OK(cudaMalloc(&op->d_b, bytes)); }
  CUDA_OK(cudaMemcpyAsync(op->d_x, x_host, bytes, cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(op->d_b, b_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = b200fem_bicgstab_solve_dev(op, op->d_b, op->d_x, epsilon, maxit, tolcrit, iterations, history); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(x_host, op->d_x, bytes, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  return B200FEM_OK;
}
