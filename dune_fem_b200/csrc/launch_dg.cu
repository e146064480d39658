// launch_dg.cu -- host side of the generic DG quadrature kernel (dg_quadrature.cuh) and of the tile Kronecker kernel
// (dg_kronecker.cuh: Q1, and Q2 boxes the marching kernel cannot take)
#include "dg_kronecker.cuh"
#include "dg_quadrature.cuh"
#include "integrands.cuh"
#include "internal.hpp"
#include "kron_tables.hpp"

using namespace b200fem;

namespace b200fem {

template <int N> static DgTabDev<N> make_tab(const Tab1D& t) {
  DgTabDev<N> T;
  for (int i = 0; i < N * N; ++i) { T.B[i] = t.B[i]; T.G[i] = t.G[i]; }
  for (int i = 0; i < N; ++i) { T.x[i] = t.x[i]; T.w[i] = t.w[i]; T.phi[0][i] = t.phi0[i]; T.phi[1][i] = t.phi1[i]; T.dphi[0][i] = t.dphi0[i]; T.dphi[1][i] = t.dphi1[i]; }
  return T;
}

template <int N> static int launch_dg_quadrature(b200fem_operator* op, const double* u, double* w, bool with_data) {
  using Cfg = DgQuadCfg<N>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  b200fem_ctx* ctx = op->sp->mesh->ctx;
  const long long n_owned = (long long)(b.own_hi[0] - b.own_lo[0]) * (b.own_hi[1] - b.own_lo[1]) * (b.own_hi[2] - b.own_lo[2]);
  auto kern = dg_quadrature_kernel<N, AdrIntegrands>;
  int rc = ensure_smem_attr(ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
  AdrIntegrands I; I.m = op->model; I.dim = b.dim; I.with_data = with_data;
  const unsigned grid = (unsigned)((n_owned + Cfg::EB - 1) / Cfg::EB);
  kern<<<grid, Cfg::kThreads, Cfg::smem_bytes(), ctx->stream>>>(make_tab<N>(op->sp->tab), b, I, op->d_perm, u, w, nullptr, n_owned, mass_scale(op));
  CUDA_OK(cudaGetLastError());
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}
int launch_dg_quadrature_any(b200fem_operator* op, const double* u, double* w, bool with_data) {
  switch (op->sp->n1) {
    case 2: return launch_dg_quadrature<2>(op, u, w, with_data);
    case 3: return launch_dg_quadrature<3>(op, u, w, with_data);
    case 4: return launch_dg_quadrature<4>(op, u, w, with_data);
    case 5: return launch_dg_quadrature<5>(op, u, w, with_data);
    case 6: return launch_dg_quadrature<6>(op, u, w, with_data);
  }
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
