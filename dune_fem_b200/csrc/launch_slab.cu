// launch_slab.cu -- host side of the slab Kronecker DG kernel for Q3..Q5 (dg_kronecker_slab.cuh)
#include <algorithm>

#include "dg_kronecker_slab.cuh"
#include "internal.hpp"
#include "kron_tables.hpp"

using namespace b200fem;

namespace b200fem {

// one CTA per TX x TY x TZ tile (persistent over tiles), n threads per element
template <int N, int TX, int TY, int TZ, int MINB> static int launch_slab(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  using Cfg = KronSlabCfg<N, TX, TY, TZ, 1>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  b200fem_ctx* ctx = op->sp->mesh->ctx;
  if (!op->kron_ready) {
    KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
    op->kron_tab.resize(sizeof(KronTabDev<N>));
    KronTabDev<N>& K0 = *reinterpret_cast<KronTabDev<N>*>(op->kron_tab.data());
    for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K0.S[d][i] = kh.S[d][i]; K0.Dlo[d][i] = kh.Dlo[d][i]; K0.Dhi[d][i] = kh.Dhi[d][i]; K0.L[d][i] = kh.L[d][i]; K0.R[d][i] = kh.R[d][i]; }
    op->kron_ready = true;
  }
  const KronTabDev<N>& K = *reinterpret_cast<const KronTabDev<N>*>(op->kron_tab.data());
  const int tx = (b.own_hi[0] - b.own_lo[0] + TX - 1) / TX, ty = (b.own_hi[1] - b.own_lo[1] + TY - 1) / TY, tz = (b.own_hi[2] - b.own_lo[2] + TZ - 1) / TZ;
  auto kern = dg_kronecker_slab_kernel<N, TX, TY, TZ, MINB, 1>;
  int rc = ensure_smem_attr(ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
  const long long ntiles = (long long)tx * ty * tz;
  REQUIRE(ntiles < (1ll << 31), B200FEM_ERR_NOT_IMPLEMENTED, "slab kernel: too many tiles");
  const long long grid = std::min<long long>(ntiles, (long long)MINB * ctx->sms);
  kern<<<(unsigned)grid, Cfg::kThreads, Cfg::smem_bytes(), ctx->stream>>>(K, b, op->d_perm, u, w, bvec, tx, ty, (int)ntiles, nullptr);
  CUDA_OK(cudaGetLastError());
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}

// tile shapes: 4x4x4 (Q3), 4x2x2 (Q4, Q5), two CTAs per SM.  Measured alternatives (4x4x2 with 4 CTAs, 4x4x3 with 3, 2x2x2 with 3
// for Q5) were within 2 % or slower (profiles/r01_dg_kronecker_slab_q3.md)
int launch_dg_slab(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  switch (op->sp->n1) {
    case 4: return launch_slab<4, 4, 4, 4, 2>(op, u, w, bvec);
    case 5: return launch_slab<5, 4, 2, 2, 2>(op, u, w, bvec);
    case 6: return launch_slab<6, 4, 2, 2, 2>(op, u, w, bvec);
  }
  return fail(B200FEM_ERR_NOT_IMPLEMENTED, "slab kernel: orders 3..5");
}

}  // namespace b200fem
