// apply.cu -- GalerkinOperator::evaluate + w.communicate() (dune/fem/schemes/galerkin.hh:1459-1496) on device vectors:
// kernel choice, load vector, halo exchange, DirichletWrapperOperator, and the host-pointer entry points with their
// copy/compute pipeline.
#include <algorithm>

#include "internal.hpp"

namespace b200fem {

int check_comm_error(b200fem_ctx* c) {
  if (c && c->h_comm_error) {
    const int e = *reinterpret_cast<volatile int*>(c->h_comm_error);
    if (e != 0) return fail(B200FEM_ERR_COMM, e == kCommTimeoutScalars ? "time-out in the peer-memory all-reduce (a rank did not arrive)" : "time-out in the peer-memory halo exchange (a neighbour did not arrive)");
  }
  return B200FEM_OK;
}

// one operator application on device vectors, without the stand-alone halo exchange (a launcher may do the exchange itself:
// op->exchange_fused)
int apply_local(b200fem_operator* op, const double* u, double* w, bool linear) {
  b200fem_space* s = op->sp; const int N = s->n1;
  op->exchange_fused = false;
  if (op->jit) return apply_jit(op, u, w, linear);       // run-time compiled integrands: always the generic quadrature kernel
  if (s->kind == B200FEM_LAGRANGE) {
    REQUIRE(!op->model.has_skeleton, B200FEM_ERR_NOT_IMPLEMENTED, "skeleton integrands on continuous spaces");
    REQUIRE(default_quadrature(op), B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange spaces: only quadrature orders that select the (order+1)-point Gauss rule");
    if (s->unst) {            // unstructured cube mesh: index arrays + per-element geometry, colour-ordered scatter
      REQUIRE(op->kernel_pref == B200FEM_KERNEL_AUTO || op->kernel_pref == B200FEM_KERNEL_QUADRATURE, B200FEM_ERR_NOT_IMPLEMENTED, "unstructured meshes: the quadrature kernel only (the Kronecker form needs a Cartesian mesh)");
      int rc = launch_lagrange_unstructured(op, u, w, !linear); if (rc) return rc;
      op->timing.kernel = B200FEM_KERNEL_QUADRATURE;
      return B200FEM_OK;
    }
    // linear models: Kronecker form (one launch, every node written once); otherwise the generic quadrature kernel with
    // colour-ordered scatter
    const bool lag_kron_ok = op->model.gamma == 0.0 && s->order <= 2;      // (the lattice kernels carry the two row types of orders 1, 2)
    int lk = op->kernel_pref;
    if (lk == B200FEM_KERNEL_AUTO) lk = lag_kron_ok ? B200FEM_KERNEL_KRONECKER : B200FEM_KERNEL_QUADRATURE;
    if (lk == B200FEM_KERNEL_KRONECKER || lk == B200FEM_KERNEL_KRONECKER_TILE) {
      REQUIRE(lag_kron_ok, B200FEM_ERR_INVALID, "Kronecker kernel needs a linear model and a Lagrange space of order 1 or 2");
      const double* bvec = nullptr;
      if (!linear && op->model.data) { int rc = ensure_bvec(op); if (rc) return rc; bvec = op->d_bvec; }
      int rc = launch_lagrange_kronecker(op, u, w, bvec); if (rc) return rc;
      op->timing.kernel = B200FEM_KERNEL_KRONECKER;
      return B200FEM_OK;
    }
    int rc = launch_lagrange_quadrature(op, u, w, !linear); if (rc) return rc;
    op->timing.kernel = B200FEM_KERNEL_QUADRATURE;
    return B200FEM_OK;
  }
  // The Kronecker form rests on the default rules (the (k+1)-point Gauss rules integrate the 1-D mass matrices of the
  // orthonormal basis exactly); other quadrature orders (setQuadratureOrders) go through the generic quadrature kernel
  const bool kron_ok = op->model.gamma == 0.0 && N >= 2 && N <= 6 && s->tensor_full && s->box.periodic == 0 && default_quadrature(op);
  int kernel = op->kernel_pref;
  if (kernel == B200FEM_KERNEL_AUTO) kernel = kron_ok ? B200FEM_KERNEL_KRONECKER : B200FEM_KERNEL_QUADRATURE;
  int rc;
  // the data terms (sources, boundary values) do not depend on u: L[u] = N(u) - l with l = -L[0] evaluated once
  // (python/dune/fem/operator/__init__.py:347-350), so every apply runs the homogeneous integrands and subtracts the stored
  // load vector instead of evaluating the analytic data at every quadrature point again
  const double* bvec = nullptr;
  if (!linear && op->model.data && !op->in_bvec) { rc = ensure_bvec(op); if (rc) return rc; bvec = op->d_bvec; }
  if (kernel == B200FEM_KERNEL_KRONECKER || kernel == B200FEM_KERNEL_KRONECKER_TILE) {
    REQUIRE(op->model.gamma == 0.0, B200FEM_ERR_INVALID, "Kronecker kernel needs a linear model");
    REQUIRE(kron_ok, B200FEM_ERR_NOT_IMPLEMENTED, "Kronecker kernel: 3-D Q_k Legendre spaces of order 1..5 with the default quadrature orders (2-D and dgonb spaces run through the quadrature kernel)");
    // marching kernel for Q2, slab kernel for Q3..Q5, tile kernel for Q1 and for Q2 boxes / vectors the TMA views cannot take
    if (N >= 4) rc = launch_dg_slab(op, u, w, bvec);
    else if (kernel == B200FEM_KERNEL_KRONECKER && dg_march_ok(op, u, w, bvec)) rc = launch_dg_march(op, u, w, bvec, op->want_exchange);
    else rc = launch_dg_kronecker_v1(op, u, w, bvec);
    kernel = B200FEM_KERNEL_KRONECKER;
  } else {
    rc = launch_dg_quadrature_any(op, u, w, bvec, op->in_bvec);
  }
  if (rc) return rc;
  op->timing.kernel = kernel;
  return B200FEM_OK;
}

// b = -L[0], evaluated once by the quadrature kernel with the data terms switched on
int ensure_bvec(b200fem_operator* op) {
  if (op->d_bvec) return B200FEM_OK;
  REQUIRE(!op->capturing, B200FEM_ERR_INVALID, "the load vector must exist before graph capture");
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  double* zero_u = nullptr; double* bv = nullptr;
  CUDA_OK(cudaMalloc(&zero_u, bytes)); CUDA_OK(cudaMalloc(&bv, bytes));
  CUDA_OK(cudaMemsetAsync(zero_u, 0, bytes, st)); CUDA_OK(cudaMemsetAsync(bv, 0, bytes, st));
  const int saved = op->kernel_pref; const bool want = op->want_exchange; const BoxDev* ab = op->active_box;
  op->kernel_pref = B200FEM_KERNEL_QUADRATURE; op->want_exchange = false; op->active_box = nullptr;
  const bool fd = op->fuse_dirichlet; op->fuse_dirichlet = false; op->in_bvec = true;
  int rc = apply_local(op, zero_u, bv, /*linear=*/false);
  op->kernel_pref = saved; op->want_exchange = want; op->active_box = ab; op->fuse_dirichlet = fd; op->in_bvec = false;
  if (rc) { cudaFree(zero_u); cudaFree(bv); return rc; }
  rc = negate_dev(bv, s->size, st);
  if (rc) { cudaFree(zero_u); cudaFree(bv); return rc; }
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(zero_u));
  op->d_bvec = bv; return B200FEM_OK;
}

int exchange(b200fem_operator* op, double* v, cudaStream_t st) {
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx;
  int rc;
  if (s->kind == B200FEM_LAGRANGE) rc = op->halo_add.built ? halo_exchange_add_p2p(op->halo_add, v, c->d_comm_error, st) : halo_exchange(op->halo, c->nccl, c->comm, v, true, st);
  else rc = op->halo_p2p.built ? halo_exchange_p2p(op->halo_p2p, v, c->d_comm_error, st) : halo_exchange_dg(op->halo_dg, c->nccl, c->comm, v, st);
  return rc ? fail(B200FEM_ERR_COMM, "halo exchange failed") : B200FEM_OK;
}

int apply_fd_jacobian(b200fem_operator* op, const double* arg, double* dest);     // solvers.cu

int apply_dev_impl(b200fem_operator* op, const double* u, double* w, bool linear) {
  if (linear && op->jac_mode) return apply_fd_jacobian(op, u, w);
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream;
  const bool timing_events = op->timing_enabled && !op->capturing;
  if (timing_events) CUDA_OK(cudaEventRecord(op->ev0, st));
  const bool distributed = op->communicate && c->world > 1;
  // Single rank: the Dirichlet wrapper can ride along in the store of the Lagrange Kronecker kernel (with several ranks it has
  // to follow the Add exchange).  Several ranks: the marching DG kernel performs the Copy exchange itself (rows on rank
  // interfaces leave for the neighbours' mailboxes as soon as they are complete, the receive part runs in the kernel's
  // tail); every other kernel is followed by the stand-alone send / receive kernels.
  op->dirichlet_fused = false;
  op->fuse_dirichlet = !distributed && op->model.strong_dirichlet && op->d_dmask != nullptr; op->fuse_linear = linear;
  op->want_exchange = distributed;
  int rc = apply_local(op, u, w, linear);
  op->fuse_dirichlet = false; op->want_exchange = false;
  if (rc) return rc;
  if (distributed) {
    if (timing_events) CUDA_OK(cudaEventRecord(op->evx0, st));
    if (!op->exchange_fused) { rc = exchange(op, w, st); if (rc) return rc; op->timing.launches_per_apply += 2; }
    if (timing_events) CUDA_OK(cudaEventRecord(op->evx1, st));
  }
  // DirichletWrapperOperator: op_(u,w) (communication included) first, then subConstraints (dirichletwrapper.hh:101-105)
  if (op->model.strong_dirichlet && op->d_dmask && !op->dirichlet_fused) {
    rc = dirichlet_sub_dev(u, w, op->d_dmask, linear ? nullptr : op->d_dvals, s->size, st); if (rc) return rc;
    op->timing.launches_per_apply += 1;
  }
  if (timing_events) CUDA_OK(cudaEventRecord(op->ev1, st));
  op->timing.applies += 1;
  return B200FEM_OK;
}

static int ensure_staging(b200fem_operator* op) {
  const size_t bytes = sizeof(double) * (size_t)op->sp->size;
  if (!op->d_u) CUDA_OK(cudaMalloc(&op->d_u, bytes));
  if (!op->d_w) CUDA_OK(cudaMalloc(&op->d_w, bytes));
  return B200FEM_OK;
}
// Host-pointer apply of a DG space on one rank, pipelined over z-slabs: the element-major dof vector is contiguous per
// z-plane, so slab c+1 travels host->device while slab c is computed and slab c-1 travels device->host.  PCIe is full
// duplex: the end-to-end time drops from H2D + kernel + D2H to about max(H2D, D2H).  A slab needs one plane of u beyond
// each end (face neighbours), so the H2D pieces are shifted by one plane against the compute slabs.
static int apply_host_pipelined(b200fem_operator* op, const double* u, double* w, bool linear, int nchunks) {
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const BoxDev& b = s->box;
  const int nz = b.n[2]; const size_t plane = (size_t)b.n[0] * b.n[1] * s->nb;
  if (!op->h2d_stream) {
    CUDA_OK(cudaStreamCreateWithFlags(&op->h2d_stream, cudaStreamNonBlocking)); CUDA_OK(cudaStreamCreateWithFlags(&op->d2h_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : op->pipe_ev) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (!linear && op->model.data && op->kernel_pref != B200FEM_KERNEL_QUADRATURE) { int rc = ensure_bvec(op); if (rc) return rc; }
  cudaEvent_t* ev_h = op->pipe_ev; cudaEvent_t* ev_c = op->pipe_ev + 16; cudaEvent_t ev_start = op->pipe_ev[32], ev_done = op->pipe_ev[33];
  CUDA_OK(cudaEventRecord(ev_start, st)); CUDA_OK(cudaStreamWaitEvent(op->h2d_stream, ev_start, 0)); CUDA_OK(cudaStreamWaitEvent(op->d2h_stream, ev_start, 0));
  int launches = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int z0 = (int)((long long)nz * c / nchunks), z1 = (int)((long long)nz * (c + 1) / nchunks);
    const int h0 = c == 0 ? 0 : z0 + 1, h1 = c == nchunks - 1 ? nz : z1 + 1;
    CUDA_OK(cudaMemcpyAsync(op->d_u + h0 * plane, u + h0 * plane, sizeof(double) * (h1 - h0) * plane, cudaMemcpyHostToDevice, op->h2d_stream));
    CUDA_OK(cudaEventRecord(ev_h[c], op->h2d_stream)); CUDA_OK(cudaStreamWaitEvent(st, ev_h[c], 0));
    BoxDev sub = b; sub.own_lo[2] = z0; sub.own_hi[2] = z1;
    op->active_box = &sub; const int rc = apply_local(op, op->d_u, op->d_w, linear); op->active_box = nullptr; if (rc) return rc;
    launches += op->timing.launches_per_apply;
    CUDA_OK(cudaEventRecord(ev_c[c], st)); CUDA_OK(cudaStreamWaitEvent(op->d2h_stream, ev_c[c], 0));
    CUDA_OK(cudaMemcpyAsync(w + z0 * plane, op->d_w + z0 * plane, sizeof(double) * (z1 - z0) * plane, cudaMemcpyDeviceToHost, op->d2h_stream));
  }
  CUDA_OK(cudaEventRecord(ev_done, op->d2h_stream)); CUDA_OK(cudaStreamWaitEvent(st, ev_done, 0));
  CUDA_OK(cudaStreamSynchronize(st));
  op->timing.launches_per_apply = launches; op->timing.applies += 1;
  return B200FEM_OK;
}
int apply_host(b200fem_operator* op, const double* u, double* w, bool linear) {
  REQUIRE(op && u && w, B200FEM_ERR_INVALID, "apply: null argument");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(c->device));
  int rc = ensure_staging(op); if (rc) return rc;
  if (op->host_pipeline_chunks >= 2 && !(linear && op->jac_mode) && s->kind != B200FEM_LAGRANGE && s->dim_range == 1 && c->world == 1 && s->box.dim == 3 && bytes >= (8u << 20) && s->box.n[2] >= 16 &&
      default_quadrature(op))
    return apply_host_pipelined(op, u, w, linear, std::min(std::min(op->host_pipeline_chunks, 16), s->box.n[2] / 4));
  CUDA_OK(cudaMemcpyAsync(op->d_u, u, bytes, cudaMemcpyHostToDevice, st));
  rc = apply_dev_impl(op, op->d_u, op->d_w, linear); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(w, op->d_w, bytes, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return check_comm_error(c);
}

}  // namespace b200fem
