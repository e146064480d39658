// launch_dg.cu -- host side of the generic DG quadrature kernel (dg_quadrature.cuh) and of the tile Kronecker kernel
// (dg_kronecker.cuh: Q1, and Q2 boxes the marching kernel cannot take)
#include "dg_kronecker.cuh"
#include "launch_dgq.hpp"
#include "kron_tables.hpp"

using namespace b200fem;

namespace b200fem {

// generic quadrature kernel: the Gauss rules follow the operator's quadrature orders (galerkin.hh:131-132, 1418-1423; rule =
// smallest Gauss rule of at least the requested order, femquadratures_inline.hh:59-70)
int launch_dg_quadrature_any(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data) {
  const int k = op->sp->order, N = op->sp->n1;
  int mi, ms;
  try { mi = gauss_points_for_order(op->q_interior ? (int)op->q_interior : 2 * k); ms = gauss_points_for_order(op->q_surface ? (int)op->q_surface : 2 * k + 1); }
  catch (const std::exception& ex) { return fail(B200FEM_ERR_NOT_IMPLEMENTED, ex.what()); }
  // a rule with fewer points than basis functions per axis is under-integration the reference would perform as asked
  if (N <= 3) return launch_dg_quadrature_n23(op, u, w, bvec, with_data, mi, ms);
  if (N == 4) return launch_dg_quadrature_n4(op, u, w, bvec, with_data, mi, ms);
  if (N <= 6) return launch_dg_quadrature_n56(op, u, w, bvec, with_data, mi, ms);
  return fail(B200FEM_ERR_NOT_IMPLEMENTED, "DG order > 5");
}

template <int N, int TX, int TY, int TZ> static int launch_v1(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  using Cfg = KronCfg<N, TX, TY, TZ>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  b200fem_ctx* ctx = op->sp->mesh->ctx;
  KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
  KronTabDev<N> K;
  for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K.S[d][i] = kh.S[d][i]; K.Dlo[d][i] = kh.Dlo[d][i]; K.Dhi[d][i] = kh.Dhi[d][i]; K.L[d][i] = kh.L[d][i]; K.R[d][i] = kh.R[d][i]; }
  const int tx = (b.own_hi[0] - b.own_lo[0] + TX - 1) / TX, ty = (b.own_hi[1] - b.own_lo[1] + TY - 1) / TY, tz = (b.own_hi[2] - b.own_lo[2] + TZ - 1) / TZ;
  auto kern = dg_kronecker_kernel<N, TX, TY, TZ>;
  int rc = ensure_smem_attr(ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
  kern<<<(unsigned)(tx * ty * tz), Cfg::kThreads, Cfg::smem_bytes(), ctx->stream>>>(K, b, op->d_perm, u, w, bvec, tx, ty);
  CUDA_OK(cudaGetLastError());
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}
int launch_dg_kronecker_v1(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  if (op->sp->n1 == 2) return launch_v1<2, 8, 8, 4>(op, u, w, bvec);
  if (op->sp->n1 == 3) return launch_v1<3, 8, 4, 4>(op, u, w, bvec);
  return fail(B200FEM_ERR_NOT_IMPLEMENTED, "tile Kronecker kernel: orders 1 and 2");
}

}  // namespace b200fem
