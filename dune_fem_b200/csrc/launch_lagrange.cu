// launch_lagrange.cu -- host side of the continuous Lagrange kernels: the sum-factorised lattice kernel
// (lagrange_kronecker.cuh), the generic quadrature kernels with colour-ordered scatter (lagrange_quadrature.cuh) and the
// cooperative CG kernel for launch-bound 2-D problems (cg_coop2d.cuh)
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "cg_coop2d.cuh"
#include "integrands.cuh"
#include "internal.hpp"
#include "kron_tables.hpp"
#include "lagrange_kronecker.cuh"
#include "lagrange_lattice.cuh"
#include "lagrange_quadrature.cuh"

using namespace b200fem;

namespace b200fem {

template <int N> static DgTabDev<N> make_tab(const Tab1D& t) {
  DgTabDev<N> T;
  for (int i = 0; i < N * N; ++i) { T.B[i] = t.B[i]; T.G[i] = t.G[i]; }
  for (int i = 0; i < N; ++i) { T.x[i] = t.x[i]; T.w[i] = t.w[i]; T.phi[0][i] = t.phi0[i]; T.phi[1][i] = t.phi1[i]; T.dphi[0][i] = t.dphi0[i]; T.dphi[1][i] = t.dphi1[i]; }
  return T;
}
template <int N> static QuadTabDev<N, N, N> make_quad_tab3(const Tab1D& t) {
  QuadTabDev<N, N, N> T;
  for (int i = 0; i < N * N; ++i) { T.Bi[i] = T.Bs[i] = t.B[i]; T.Gi[i] = T.Gs[i] = t.G[i]; }
  for (int i = 0; i < N; ++i) { T.xi[i] = T.xs[i] = t.x[i]; T.wi[i] = T.ws[i] = t.w[i]; T.phi[0][i] = t.phi0[i]; T.phi[1][i] = t.phi1[i]; T.dphi[0][i] = t.dphi0[i]; T.dphi[1][i] = t.dphi1[i]; }
  return T;
}

// generic element integrals, 2^dim colours = 2^dim launches with plain read-modify-write (deterministic)
template <int N> static int launch_lagrange(b200fem_operator* op, const double* u, double* w, bool with_data) {
  const BoxDev& b = op->active_box ? *op->active_box : op->sp->box; b200fem_ctx* ctx = op->sp->mesh->ctx; cudaStream_t st = ctx->stream;
  CUDA_OK(cudaMemsetAsync(w, 0, sizeof(double) * (size_t)op->sp->size, st));                 // w.clear() (galerkin.hh:1463)
  AdrIntegrands I; I.m = op->model; I.dim = b.dim; I.with_data = with_data;
  int launches = 1;
  if (b.dim == 3) {
    using Cfg = DgQuadCfg<N, N, N>; auto kern = lagrange3d_quadrature_kernel<N, AdrIntegrands>;
    int rc = ensure_smem_attr(ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
    for (int c = 0; c < 8; ++c) {
      const int c0 = c & 1, c1 = (c >> 1) & 1, c2 = c >> 2;
      const int m0 = (b.n[0] - c0 + 1) / 2, m1 = (b.n[1] - c1 + 1) / 2, m2 = (b.n[2] - c2 + 1) / 2;
      const long long nc = (long long)m0 * m1 * m2; if (nc <= 0) continue;
      kern<<<(unsigned)((nc + Cfg::EB - 1) / Cfg::EB), Cfg::kThreads, Cfg::smem_bytes(), st>>>(make_quad_tab3<N>(op->sp->tab), b, I, op->sp->lay, u, w, c0, c1, c2, m0, m1, nc);
      ++launches;
    }
  } else {
    for (int c = 0; c < 4; ++c) {
      const int c0 = c & 1, c1 = c >> 1; const int m0 = (b.n[0] - c0 + 1) / 2, m1 = (b.n[1] - c1 + 1) / 2;
      const long long nc = (long long)m0 * m1; if (nc <= 0) continue;
      lagrange2d_quadrature_kernel<N, AdrIntegrands><<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(make_tab<N>(op->sp->tab), b, I, op->sp->lay, u, w, c0, c1, m0, nc);
      ++launches;
    }
  }
  CUDA_OK(cudaGetLastError());
  op->timing.launches_per_apply = launches;
  return B200FEM_OK;
}
int launch_lagrange_quadrature(b200fem_operator* op, const double* u, double* w, bool with_data) {
  return op->sp->n1 == 2 ? launch_lagrange<2>(op, u, w, with_data) : op->sp->n1 == 3 ? launch_lagrange<3>(op, u, w, with_data) : launch_lagrange<4>(op, u, w, with_data);
}

static int ensure_lag_rows(b200fem_operator* op) {
  if (op->d_lag_rows) return B200FEM_OK;
  b200fem_space* s = op->sp; const BoxDev& b = s->box;
  LagRowsHost rh = build_lagrange_rows(s->tab, op->model, b.dim, s->order, b.n, b.origin, b.gn, b.h);
  size_t total = 0; for (int d = 0; d < 3; ++d) total += 2 * rh.M[d].size();
  std::vector<double> flat; flat.reserve(total); size_t offM[3], offT[3];
  for (int d = 0; d < 3; ++d) { offM[d] = flat.size(); flat.insert(flat.end(), rh.M[d].begin(), rh.M[d].end()); offT[d] = flat.size(); flat.insert(flat.end(), rh.T[d].begin(), rh.T[d].end()); }
  CUDA_OK(cudaMalloc(&op->d_lag_rows, sizeof(double) * flat.size()));
  CUDA_OK(cudaMemcpy(op->d_lag_rows, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice));
  for (int d = 0; d < 3; ++d) { op->lag_rows.M[d] = op->d_lag_rows + offM[d]; op->lag_rows.T[d] = op->d_lag_rows + offT[d]; }
  return B200FEM_OK;
}

// first-generation lattice kernel (one node per thread, per-row coefficient tables): kept selectable through
// B200FEM_KERNEL_KRONECKER_TILE for A/B checks of the second generation
static int launch_lagrange_kronecker_v1(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  b200fem_space* s = op->sp; const int k = s->order; b200fem_ctx* ctx = s->mesh->ctx;
  int rc = ensure_lag_rows(op); if (rc) return rc;
  const LagrangeLayoutDev& L = s->lay; const bool mapped = L.lattice_map != nullptr;
  constexpr int HY = 16, ctas_per_sm = 2;
  const int TX = 32 - 2 * k, TY = HY - 2 * k;
  const int tx = (int)((L.lattice[0] + TX - 1) / TX), ty = (int)((L.lattice[1] + TY - 1) / TY);
  const int L2 = (int)L.lattice[2], slots = ctas_per_sm * ctx->sms, tiles = tx * ty;
  int best_nseg = 1; double best_cost = 1e300;
  for (int ns = 1; ns <= 64; ++ns) {
    const int zs = (L2 + ns - 1) / ns; if (zs > 128) continue;
    const int nse = (L2 + zs - 1) / zs;
    const double waves = std::ceil((double)tiles * nse / slots), cost = waves * (zs + 2 * k + 6);
    if (cost < best_cost) { best_cost = cost; best_nseg = nse; }
    if (zs <= 4) break;
  }
  const int zseg = (L2 + best_nseg - 1) / best_nseg, nseg = (L2 + zseg - 1) / zseg;
  const unsigned grid = (unsigned)(tiles * nseg); cudaStream_t st = ctx->stream;
  const unsigned char* dmask = op->fuse_dirichlet ? op->d_dmask : nullptr; const double* dvals = op->fuse_dirichlet && !op->fuse_linear ? op->d_dvals : nullptr;
  double* dotp = nullptr; op->dot_parts = 0;
  if (op->want_dot && op->fuse_dirichlet == (op->model.strong_dirichlet && op->d_dmask != nullptr) && ctx->world == 1) {
    if ((int)grid > op->dot_cap) { if (op->capturing) return fail(B200FEM_ERR_INVALID, "dot partial buffer must exist before graph capture"); if (op->d_dot_partial) cudaFree(op->d_dot_partial); CUDA_OK(cudaMalloc(&op->d_dot_partial, sizeof(double) * grid)); op->dot_cap = (int)grid; }
    dotp = op->d_dot_partial; op->dot_parts = (int)grid;
  }
#define B200FEM_LAGK(KK, MM) lagrange_kronecker_kernel<KK, MM, HY><<<grid, 32 * HY, 0, st>>>(L, op->lag_rows, u, w, bvec, dmask, dvals, tx, ty, zseg, dotp)
  if (k == 1) { if (mapped) B200FEM_LAGK(1, true); else B200FEM_LAGK(1, false); }
  else        { if (mapped) B200FEM_LAGK(2, true); else B200FEM_LAGK(2, false); }
#undef B200FEM_LAGK
  op->dirichlet_fused = op->fuse_dirichlet;
  CUDA_OK(cudaGetLastError());
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}

// second-generation lattice kernel (lagrange_lattice.cuh): four nodes per thread, two row types in the constant bank
template <int K, int LX, bool MAPPED, int WARPS = 16>
static int launch_lattice(b200fem_operator* op, const LagStencilDev<K>& S, const double* u, double* w, const double* bvec, const double* dvals) {
  using Cfg = LagLatCfg<K, LX, WARPS>;
  b200fem_space* s = op->sp; b200fem_ctx* ctx = s->mesh->ctx; const LagrangeLayoutDev& L = s->lay;
  const int tx = (int)((L.lattice[0] + Cfg::TXO - 1) / Cfg::TXO), ty = (int)((L.lattice[1] + Cfg::TYO - 1) / Cfg::TYO), tiles = tx * ty;
  // z-segments: every segment re-reads 2k planes; the number of segments is chosen so that the grid fills whole waves of the
  // resident CTA slots (one CTA per SM)
  const int L2 = (int)L.lattice[2], slots = ctx->sms * Cfg::kCtasPerSm;
  int best_nseg = 1; double best_cost = 1e300;
  for (int ns = 1; ns <= 64 && ns <= L2; ++ns) {
    const int zs = (L2 + ns - 1) / ns, nse = (L2 + zs - 1) / zs;
    const double waves = std::ceil((double)tiles * nse / slots), cost = waves * (zs + 2 * K + 4);
    if (cost < best_cost) { best_cost = cost; best_nseg = nse; }
  }
  const int zseg = (L2 + best_nseg - 1) / best_nseg, nseg = (L2 + zseg - 1) / zseg;
  const unsigned grid = (unsigned)(tiles * nseg);
  double* dotp = nullptr; op->dot_parts = 0;
  if (op->want_dot && op->fuse_dirichlet == (op->model.strong_dirichlet && op->d_dmask != nullptr) && ctx->world == 1) {
    if ((int)grid > op->dot_cap) { if (op->capturing) return fail(B200FEM_ERR_INVALID, "dot partial buffer must exist before graph capture"); if (op->d_dot_partial) cudaFree(op->d_dot_partial); CUDA_OK(cudaMalloc(&op->d_dot_partial, sizeof(double) * grid)); op->dot_cap = (int)grid; }
    dotp = op->d_dot_partial; op->dot_parts = (int)grid;
  }
  auto kern = (bvec != nullptr || dvals != nullptr) ? lagrange_lattice_kernel<K, LX, MAPPED, WARPS, true> : lagrange_lattice_kernel<K, LX, MAPPED, WARPS, false>;
  int rc = ensure_smem_attr(ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
  kern<<<grid, Cfg::kThreads, Cfg::smem_bytes(), ctx->stream>>>(L, S, u, w, bvec, dvals, tx, ty, zseg, dotp);
  CUDA_OK(cudaGetLastError());
  return B200FEM_OK;
}
template <int K> static int launch_lattice_k(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  b200fem_space* s = op->sp; const BoxDev& b = s->box; const LagrangeLayoutDev& L = s->lay;
  const LagStencilHost h = build_lagrange_stencil(s->tab, op->model, b.dim, K, b.n, b.origin, b.gn, b.h);
  LagStencilDev<K> S; std::memset(&S, 0, sizeof(S));
  for (int d = 0; d < 3; ++d) {
    for (int ty = 0; ty < 2; ++ty) for (int t = 0; t < 2 * K + 1; ++t) { S.M[d][ty][t] = h.M[d][ty][t]; S.T[d][ty][t] = h.T[d][ty][t]; }
    S.Mlo[d] = h.Mlo[d]; S.Mhi[d] = h.Mhi[d]; S.Tlo[d] = h.Tlo[d]; S.Thi[d] = h.Thi[d];
    S.glo[d] = d < b.dim ? K * b.origin[d] : 0; S.gend[d] = d < b.dim ? K * b.gn[d] : 0;
  }
  S.dirichlet_bits = op->fuse_dirichlet ? (op->model.dirichlet_mask & ((1 << (2 * b.dim)) - 1)) : 0;
  S.affine = op->fuse_dirichlet && !op->fuse_linear ? 1 : 0;
  const double* dvals = S.affine ? op->d_dvals : nullptr;
  const bool mapped = L.lattice_map != nullptr;
  // tile: 64 nodes wide (16 lanes x 4 nodes), 16 rows (8 warps x 2 rows), two CTAs per SM.  Measured on P2 128^3 (profiles/
  // r02_lagrange_lattice.md): 259 us against 295 us for 16-warp CTAs with 32 rows -- the smaller CTAs run out of phase (one
  // CTA's shared-memory phase under the other's FMA phase), which outweighs their larger halo share
  // (round 2, measured: 32 x 32 tiles (LX = 8) 273 us, 32 x 64 tiles (LX = 8, 16 warps) 300 us -- better tile efficiency on paper,
  // slower in fact)
  int rc = mapped ? launch_lattice<K, 16, true, 8>(op, S, u, w, bvec, dvals) : launch_lattice<K, 16, false, 8>(op, S, u, w, bvec, dvals);
  if (rc) return rc;
  op->dirichlet_fused = op->fuse_dirichlet;
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}
int launch_lagrange_kronecker(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  b200fem_space* s = op->sp;
  REQUIRE(s->lay.lattice[0] * s->lay.lattice[1] * s->lay.lattice[2] < (1ll << 31) && s->size < (1ll << 31), B200FEM_ERR_NOT_IMPLEMENTED, "lattice kernel: 32-bit dof cursors");
  if (op->kernel_pref == B200FEM_KERNEL_KRONECKER_TILE) return launch_lagrange_kronecker_v1(op, u, w, bvec);
  return s->order == 1 ? launch_lattice_k<1>(op, u, w, bvec) : launch_lattice_k<2>(op, u, w, bvec);
}

// Launch-bound sizes on a 2-D Lagrange lattice (BASELINE config 1): a chunk of CG iterations is ONE cooperative launch with
// grid-wide barriers instead of kernel boundaries.  *coop_grid_inout == 0: decide whether the path applies (and with which
// grid); > 0: launch `iters` iterations.
int coop_cg_chunk(b200fem_operator* op, double* x, int iters, int* coop_grid_inout) {
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream;
  if (*coop_grid_inout == 0) {
    int rc = ensure_lag_rows(op); if (rc) return rc;
    int per_sm = 0, coop_ok = 0;
    CUDA_OK(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, c->device));
    if (s->order == 1) CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_coop2d_kernel<1>, kCoopThreads, 0));
    else CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_coop2d_kernel<2>, kCoopThreads, 0));
    const long long nodes = s->lay.lattice[0] * s->lay.lattice[1];
    *coop_grid_inout = coop_ok ? (int)std::min<long long>(std::min<long long>((long long)per_sm * c->sms, kRedBlocks), (nodes + kCoopThreads - 1) / kCoopThreads) : 0;
    return B200FEM_OK;
  }
  const unsigned char* dm = op->model.strong_dirichlet ? op->d_dmask : nullptr; double* xx = x; int it = iters;
  void* args[] = {(void*)&s->lay, (void*)&op->lag_rows, (void*)&xx, (void*)&op->d_r, (void*)&op->d_p, (void*)&op->d_h, (void*)&dm, (void*)&op->d_partial,
                  (void*)&op->d_cg, (void*)&op->d_hist, (void*)&it};
  if (s->order == 1) CUDA_OK(cudaLaunchCooperativeKernel((const void*)cg_coop2d_kernel<1>, dim3((unsigned)*coop_grid_inout), dim3(kCoopThreads), args, 0, st));
  else CUDA_OK(cudaLaunchCooperativeKernel((const void*)cg_coop2d_kernel<2>, dim3((unsigned)*coop_grid_inout), dim3(kCoopThreads), args, 0, st));
  return B200FEM_OK;
}

}  // namespace b200fem
