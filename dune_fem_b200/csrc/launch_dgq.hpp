// launch_dgq.hpp -- host side of the generic DG quadrature kernel (dg_quadrature.cuh), shared by the translation units that
// instantiate it for the different orders (launch_dgq_*.cu: the instantiations are heavy, so they compile in parallel).
#pragma once
#include "dg_quadrature.cuh"
#include "integrands.cuh"
#include "internal.hpp"

namespace b200fem {

template <int N, int MI, int MS> inline QuadTabDev<N, MI, MS> make_quad_tab(bool legendre, int order) {
  const Tab1D ti = tabulate_1d(legendre ? Basis::Legendre : Basis::Lagrange, order, MI), ts = tabulate_1d(legendre ? Basis::Legendre : Basis::Lagrange, order, MS);
  QuadTabDev<N, MI, MS> T;
  for (int i = 0; i < MI * N; ++i) { T.Bi[i] = ti.B[i]; T.Gi[i] = ti.G[i]; }
  for (int i = 0; i < MS * N; ++i) { T.Bs[i] = ts.B[i]; T.Gs[i] = ts.G[i]; }
  for (int i = 0; i < MI; ++i) { T.xi[i] = ti.x[i]; T.wi[i] = ti.w[i]; }
  for (int i = 0; i < MS; ++i) { T.xs[i] = ts.x[i]; T.ws[i] = ts.w[i]; }
  for (int i = 0; i < N; ++i) { T.phi[0][i] = ti.phi0[i]; T.phi[1][i] = ti.phi1[i]; T.dphi[0][i] = ti.dphi0[i]; T.dphi[1][i] = ti.dphi1[i]; }
  return T;
}

template <int N, int MI, int MS> inline int launch_dg_quadrature(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data) {
  using Cfg = DgQuadCfg<N, MI, MS>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  b200fem_ctx* ctx = op->sp->mesh->ctx;
  const long long n_owned = (long long)(b.own_hi[0] - b.own_lo[0]) * (b.own_hi[1] - b.own_lo[1]) * (b.own_hi[2] - b.own_lo[2]);
  const unsigned grid = (unsigned)((n_owned + Cfg::EB - 1) / Cfg::EB);
  const auto tab = make_quad_tab<N, MI, MS>(true, N - 1);
  const bool gen = !op->sp->tensor_full || b.periodic != 0;      // sub-basis space or periodic grid: the general instantiation
  auto launch = [&](auto kern, auto I) -> int {
    int rc = ensure_smem_attr(ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
    I.m = op->model; I.dim = b.dim; I.with_data = with_data;
    kern<<<grid, Cfg::kThreads, Cfg::smem_bytes(), ctx->stream>>>(tab, b, I, op->d_perm, op->sp->nb, u, w, bvec, n_owned, mass_scale(op));
    return B200FEM_OK;
  };
  int rc;
  // the load-vector pass (once per operator) is the instantiation that carries the analytic data; it always takes the general kernel
  if (with_data) rc = launch(dg_quadrature_kernel<N, MI, MS, AdrIntegrands, true>, AdrIntegrands{});
  else if (gen) rc = launch(dg_quadrature_kernel<N, MI, MS, AdrIntegrandsHom, true>, AdrIntegrandsHom{});
  else rc = launch(dg_quadrature_kernel<N, MI, MS, AdrIntegrandsHom, false>, AdrIntegrandsHom{});
  if (rc) return rc;
  CUDA_OK(cudaGetLastError());
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}
// the rules a kernel exists for: MI, MS in {N, N+1} independently and MI = MS = N+2, for N <= 4; MI = MS in {N, N+1} for N = 5, 6
template <int N> inline int launch_dg_quadrature_n(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data, int mi, int ms) {
  if (mi == N && ms == N) return launch_dg_quadrature<N, N, N>(op, u, w, bvec, with_data);
  if (mi == N + 1 && ms == N + 1) return launch_dg_quadrature<N, N + 1, N + 1>(op, u, w, bvec, with_data);
  if constexpr (N <= 4) {
    if (mi == N + 1 && ms == N) return launch_dg_quadrature<N, N + 1, N>(op, u, w, bvec, with_data);
    if (mi == N && ms == N + 1) return launch_dg_quadrature<N, N, N + 1>(op, u, w, bvec, with_data);
    if (mi == N + 2 && ms == N + 2) return launch_dg_quadrature<N, N + 2, N + 2>(op, u, w, bvec, with_data);
  }
  return fail(B200FEM_ERR_NOT_IMPLEMENTED, "quadrature orders: no device kernel for this pair of Gauss rules (interior / surface points per axis within order+1 .. order+3)");
}

int launch_dg_quadrature_n23(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data, int mi, int ms);
int launch_dg_quadrature_n4(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data, int mi, int ms);
int launch_dg_quadrature_n56(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data, int mi, int ms);

}  // namespace b200fem
