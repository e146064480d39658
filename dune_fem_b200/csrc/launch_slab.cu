// launch_slab.cu -- host side of the slab Kronecker DG kernel for Q3..Q5 (dg_kronecker_slab.cuh)
#include <algorithm>

#include <cstdio>
#include <functional>
#include <vector>
#include <cstdlib>

#include "dg_kronecker_mma.cuh"
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

typedef CUresult (*EncodeTiledQ3Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledQ3Fn g_encode_tiled_q3 = nullptr;     // driver entry point, resolved at run time (no link-time dependency on libcuda)
static bool ensure_encode_tiled_q3() {
  if (g_encode_tiled_q3) return true;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
  g_encode_tiled_q3 = (EncodeTiledQ3Fn)fn; return true;
}

// KronMmaOrder (dg_kronecker_mma.cuh): split the 64 dofs of an element into four groups of 16 such that inside a group all source banks
// (dof % 16) and all destination banks (swizzled offset % 16) are different -- four perfect matchings of a 4-regular bipartite
// multigraph on 16 + 16 banks, found by augmenting paths
static KronMmaOrder make_mma_order(const std::vector<int>& perm) {
  int off_of[64];                                      // stored dof -> offset inside a swizzled element
  for (int i = 0; i < 64; ++i) { const int a = i >> 4, b = (i >> 2) & 3, c = i & 3; off_of[perm[(size_t)i]] = KronMmaCfg::eoff(a, b, c >> 1) + (c & 1); }
  std::vector<int> left(64); for (int j = 0; j < 64; ++j) left[(size_t)j] = j;        // remaining edges (dofs)
  KronMmaOrder O;
  for (int m = 0; m < 4; ++m) {
    int match_r[16]; std::fill(match_r, match_r + 16, -1);                               // destination bank -> edge
    for (int l = 0; l < 16; ++l) {
      bool seen[16] = {};
      std::function<bool(int)> aug = [&](int lb) {
        for (int e : left) {
          if (e % 16 != lb) continue;
          const int rb = off_of[e] % 16; if (seen[rb]) continue; seen[rb] = true;
          if (match_r[rb] < 0 || aug(match_r[rb] % 16)) { match_r[rb] = e; return true; }
        }
        return false;
      };
      aug(l);
    }
    for (int r = 0; r < 16; ++r) {
      const int e = match_r[r]; O.dof[m * 16 + r] = e; O.off[m * 16 + r] = off_of[e];
      left.erase(std::find(left.begin(), left.end(), e));
    }
  }
  return O;
}

// Q3 on the FP64 tensor cores (dg_kronecker_mma.cuh): persistent CTAs, one per SM, each marching through its share of the
// (8 x 8 column, z) plane steps
static int launch_mma_q3(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  using Cfg = KronMmaCfg; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  b200fem_ctx* ctx = op->sp->mesh->ctx;
  if (!op->kron_ready) {
    KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
    op->kron_tab.resize(sizeof(KronTabDev<4>));
    KronTabDev<4>& K0 = *reinterpret_cast<KronTabDev<4>*>(op->kron_tab.data());
    for (int d = 0; d < 3; ++d) for (int i = 0; i < 16; ++i) { K0.S[d][i] = kh.S[d][i]; K0.Dlo[d][i] = kh.Dlo[d][i]; K0.Dhi[d][i] = kh.Dhi[d][i]; K0.L[d][i] = kh.L[d][i]; K0.R[d][i] = kh.R[d][i]; }
    op->kron_ready = true;
  }
  const KronTabDev<4>& K = *reinterpret_cast<const KronTabDev<4>*>(op->kron_tab.data());
  const int on[3] = {b.own_hi[0] - b.own_lo[0], b.own_hi[1] - b.own_lo[1], b.own_hi[2] - b.own_lo[2]};
  const int tx = (on[0] + Cfg::TX - 1) / Cfg::TX, ty = (on[1] + Cfg::TY - 1) / Cfg::TY;
  const long long steps = (long long)tx * ty * on[2];
  // tensor maps (host-side encoding, ~1 us each; an apply of this kernel takes >= 100 us)
  KronMmaMaps M;
  {
    const cuuint64_t du[4] = {64, (cuuint64_t)b.n[0], (cuuint64_t)b.n[1], (cuuint64_t)b.n[2]}, dw[4] = {64, (cuuint64_t)on[0], (cuuint64_t)on[1], (cuuint64_t)on[2]};
    const cuuint64_t st[3] = {512, 512ull * b.n[0], 512ull * b.n[0] * b.n[1]};
    const cuuint32_t bu[4] = {64, (cuuint32_t)Cfg::PX, (cuuint32_t)Cfg::PY, 1}, bw[4] = {64, (cuuint32_t)Cfg::TX, (cuuint32_t)Cfg::TY, 1}, es[4] = {1, 1, 1, 1};
    const long long own_off = ((long long)b.own_lo[0] + (long long)b.n[0] * (b.own_lo[1] + (long long)b.n[1] * b.own_lo[2])) * 64;
    auto enc = [&](CUtensorMap* m, const double* base, const cuuint64_t* d, const cuuint32_t* bx) {
      return g_encode_tiled_q3(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    REQUIRE(enc(&M.u_plane, u, du, bu) && enc(&M.w_tile, w + own_off, dw, bw) && enc(&M.b_tile, (bvec ? bvec : w) + own_off, dw, bw), B200FEM_ERR_CUDA, "cuTensorMapEncodeTiled failed (Q3 tensor-core kernel)");
  }
  auto kern = bvec ? dg_kronecker_mma_kernel<true, Cfg::TY> : dg_kronecker_mma_kernel<false, Cfg::TY>;
  int rc = ensure_smem_attr(ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
  const unsigned grid = (unsigned)std::max(1ll, std::min<long long>(steps, (long long)ctx->sms * Cfg::kCtasPerSm));
  static long long* d_dbg = nullptr; static const bool want_dbg = std::getenv("B200FEM_MMA_TIMELINE") != nullptr;   // diagnostics only
  if (want_dbg && !d_dbg) { CUDA_OK(cudaMalloc(&d_dbg, 16 * sizeof(long long))); CUDA_OK(cudaMemset(d_dbg, 0, 16 * sizeof(long long))); }
  kern<<<grid, Cfg::kThreads, Cfg::smem_bytes(), ctx->stream>>>(K, b, M, make_mma_order(op->sp->perm), tx, ty, want_dbg ? d_dbg : nullptr);
  CUDA_OK(cudaGetLastError());
  if (want_dbg) {
    long long h[16]; CUDA_OK(cudaStreamSynchronize(ctx->stream)); CUDA_OK(cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
    for (int w2 = 0; w2 < 2; ++w2) { const long long* t = h + 8 * w2; const double n = (double)std::max(1ll, t[6]);
      std::fprintf(stderr, "[mma timeline, CTA 1 thread %d, clocks per plane step over %lld steps] wait u %.0f | stage %.0f | sync %.0f | compute+out %.0f | sync %.0f | store/issue %.0f\n", w2 ? 255 : 0, t[6], t[0] / n, t[1] / n, t[2] / n, t[3] / n, t[4] / n, t[5] / n); }
  }
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}

// tile shapes: 4x4x4 (Q3), 4x2x2 (Q4, Q5), two CTAs per SM.  Measured alternatives (4x4x2 with 4 CTAs, 4x4x3 with 3, 2x2x2 with 3
// for Q5) were within 2 % or slower (profiles/r01_dg_kronecker_slab_q3.md)
int launch_dg_slab(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  switch (op->sp->n1) {
    case 4: {
      auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
      static const bool force_slab = std::getenv("B200FEM_Q3_SLAB") != nullptr;      // A/B switch (read once)
      if (!force_slab && al16(u) && al16(w) && (!bvec || al16(bvec)) && ensure_encode_tiled_q3()) return launch_mma_q3(op, u, w, bvec);
      return launch_slab<4, 4, 4, 4, 2>(op, u, w, bvec);
    }
    case 5: return launch_slab<5, 4, 2, 2, 2>(op, u, w, bvec);
    case 6: return launch_slab<6, 4, 2, 2, 2>(op, u, w, bvec);
  }
  return fail(B200FEM_ERR_NOT_IMPLEMENTED, "slab kernel: orders 3..5");
}

}  // namespace b200fem
