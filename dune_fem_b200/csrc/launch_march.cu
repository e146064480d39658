// launch_march.cu -- host side of the z-marching Kronecker DG kernel (dg_kronecker_march.cuh): tensor maps, persistent
// grid, programmatic dependent launch, and the fused Copy exchange of w on several ranks.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dg_kronecker_march.cuh"
#include "internal.hpp"
#include "kron_tables.hpp"
#include "march_schedule.hpp"

using namespace b200fem;

// ---- TMA tensor maps (driver entry point resolved at run time: no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;     // a driver entry point: the same for every device of the process
static bool ensure_encode_tiled() {
  if (g_encode_tiled) return true;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
  g_encode_tiled = (EncodeTiledFn)fn; return true;
}
// 4-D tensor of doubles [d3][d2][d1][d0] (d0 contiguous) with byte strides s1..s3 and box b0 x b1 x b2 x b3
static bool make_map4(CUtensorMap* m, const double* base, const uint64_t (&d)[4], const uint64_t (&s)[3], const uint32_t (&bx)[4]) {
  const cuuint64_t dims[4] = {d[0], d[1], d[2], d[3]}; const cuuint64_t strides[3] = {s[0], s[1], s[2]};
  const cuuint32_t boxd[4] = {bx[0], bx[1], bx[2], bx[3]}; const cuuint32_t estr[4] = {1, 1, 1, 1};
  return g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Host-side cost matters at 30 us per apply: the 1-D operator tables are built once per operator state and encoded tensor
// maps are cached per (u, w, b, owned range).
struct KronMapKey { const void *u, *w, *b; int lo[3], hi[3]; bool operator==(const KronMapKey& o) const { return std::memcmp(this, &o, sizeof(KronMapKey)) == 0; } };
struct MarchSchedule { MarchScheduleKey key; MarchRun* d_runs = nullptr; int* d_begin = nullptr; };
struct MarchMapCache { static constexpr int kSlots = 32; KronMapKey key[kSlots]; KronMarchMaps maps[kSlots]; bool valid[kSlots] = {}; int next = 0;
                       std::vector<MarchSchedule> schedules;
                       unsigned long long* d_ts = nullptr; };      // diagnostics (B200FEM_MARCH_TS): globaltimer stamps of CTA 0

namespace b200fem {

void free_march_cache(b200fem_operator* op) {
  if (op->march_cache && op->march_cache->d_ts) {
    std::vector<unsigned long long> h(16 + 4 * 160); cudaDeviceSynchronize(); cudaMemcpy(h.data(), op->march_cache->d_ts, h.size() * 8, cudaMemcpyDeviceToHost); cudaFree(op->march_cache->d_ts);
    std::fprintf(stderr, "[b200fem march, one rank, ns] CTA 0: start -> loop end %lld | stores drained +%lld | previous kernel's end -> this start %lld | previous loop end -> this loop end %lld\n",
                 (long long)(h[0] - h[8]), (long long)(h[1] - h[0]), (long long)(h[8] - h[9]), (long long)(h[0] - h[10]));
    unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0, p1 = 0; int n = 0;
    for (int b = 0; b < 160; ++b) { const unsigned long long st = h[16 + 4 * b], en = h[16 + 4 * b + 1], pe = h[16 + 4 * b + 2]; if (!st) continue; ++n; s0 = std::min(s0, st); s1 = std::max(s1, st); e0 = std::min(e0, en); e1 = std::max(e1, en); p1 = std::max(p1, pe); }
    if (std::getenv("B200FEM_MARCH_TS_DUMP")) for (int b = 0; b < 160; ++b) { const unsigned long long st = h[16 + 4 * b], en = h[16 + 4 * b + 1]; if (st) std::fprintf(stderr, "cta %d start %lld dur %lld\n", b, (long long)(st - s0), (long long)(en - st)); }
    std::fprintf(stderr, "[b200fem march, one rank, ns] %d CTAs: starts spread %lld | ends spread %lld | first start -> last end %lld | previous launch's last end -> first start %lld\n",
                 n, (long long)(s1 - s0), (long long)(e1 - e0), (long long)(e1 - s0), (long long)(s0 - p1));
  }
  if (op->march_cache) for (MarchSchedule& sc : op->march_cache->schedules) { cudaFree(sc.d_runs); cudaFree(sc.d_begin); }
  delete op->march_cache; op->march_cache = nullptr;
}

// Q2 on a 3-D box whose x extents are even (element pairs make TMA rows multiples of 16 bytes), 16-byte aligned vectors
bool dg_march_ok(const b200fem_operator* op, const double* u, const double* w, const double* bvec) {
  const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return op->sp->n1 == 3 && b.dim == 3 && b.n[0] % 2 == 0 && b.own_lo[0] % 2 == 0 && (b.own_hi[0] - b.own_lo[0]) % 2 == 0 &&
         al16(u) && al16(w) && (!bvec || al16(bvec)) && ensure_encode_tiled();
}

template <bool HIER> static int launch_march(b200fem_operator* op, const double* u, double* w, const double* bvec, bool fuse_exchange) {
  constexpr int N = 3, TX = 16, TY = 16, N3 = N * N * N;
  using Cfg = KronMarchCfg<N, TX, TY>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  b200fem_ctx* ctx = op->sp->mesh->ctx;
  if (!op->kron_ready) {
    KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
    op->kron_tab.resize(sizeof(KronTabDev<N>));
    KronTabDev<N>& K0 = *reinterpret_cast<KronTabDev<N>*>(op->kron_tab.data());
    for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K0.S[d][i] = kh.S[d][i]; K0.Dlo[d][i] = kh.Dlo[d][i]; K0.Dhi[d][i] = kh.Dhi[d][i]; K0.L[d][i] = kh.L[d][i]; K0.R[d][i] = kh.R[d][i]; }
    op->kron_ready = true; op->kron_chk = -1;
  }
  const KronTabDev<N>& K = *reinterpret_cast<const KronTabDev<N>*>(op->kron_tab.data());
  const int on[3] = {b.own_hi[0] - b.own_lo[0], b.own_hi[1] - b.own_lo[1], b.own_hi[2] - b.own_lo[2]};
  const int tx = (on[0] + TX - 1) / TX, ty = (on[1] + TY - 1) / TY, ncols = tx * ty;
  if (!op->march_cache) op->march_cache = new MarchMapCache;
  MarchMapCache& mc = *op->march_cache;
  KronMapKey key; std::memset(&key, 0, sizeof(key)); key.u = u; key.w = w; key.b = bvec;
  for (int d = 0; d < 3; ++d) { key.lo[d] = b.own_lo[d]; key.hi[d] = b.own_hi[d]; }
  int slot = -1;
  for (int i = 0; i < MarchMapCache::kSlots; ++i) if (mc.valid[i] && mc.key[i] == key) { slot = i; break; }
  if (slot < 0) {
    slot = mc.next; mc.next = (mc.next + 1) % MarchMapCache::kSlots;
    const uint64_t sp = 2ull * N3 * 8, s1 = (uint64_t)b.n[0] * N3 * 8, s2 = s1 * b.n[1];
    const long long own_off = ((long long)b.own_lo[0] + (long long)b.n[0] * (b.own_lo[1] + (long long)b.n[1] * b.own_lo[2])) * N3;
    KronMarchMaps& M = mc.maps[slot];
    const uint64_t du[4] = {2ull * N3, (uint64_t)b.n[0] / 2, (uint64_t)b.n[1], (uint64_t)b.n[2]};
    const uint64_t dw[4] = {2ull * N3, (uint64_t)on[0] / 2, (uint64_t)on[1], (uint64_t)on[2]};
    const uint64_t st[3] = {sp, s1, s2};
    const uint32_t bplane[4] = {2u * N3, (TX + 4) / 2, TY + 2, 1}, btile[4] = {2u * N3, TX / 2, TY, 1}, bwarp[4] = {2u * N3, TX / 2, 32 / TX, 1};
    bool ok = make_map4(&M.u_plane, u, du, st, bplane) && make_map4(&M.u_edge, u, du, st, btile) &&
              make_map4(&M.w_tile, w + own_off, dw, st, bwarp) && make_map4(&M.b_tile, (bvec ? bvec : w) + own_off, dw, st, bwarp);
    REQUIRE(ok, B200FEM_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed");
    mc.key[slot] = key; mc.valid[slot] = true;
  }
  // checkerboard self matrices on the y and z axes (no advection there: even and odd Legendre modes decouple); entries
  // that are zero up to quadrature rounding (<= 1e-14 of the matrix norm) are not multiplied at all
  if (op->kron_chk < 0) {
    bool chk = true;
    for (int d = 1; d < 3; ++d) {
      double mx = 0; for (int i = 0; i < N * N; ++i) mx = std::max(mx, std::fabs(K.S[d][i]));
      for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) if (((i + j) & 1) && std::fabs(K.S[d][i * N + j]) > 1e-14 * mx) chk = false;
    }
    op->kron_chk = chk ? 1 : 0;
  }
  const int variant = (bvec ? 1 : 0) + (op->kron_chk ? 2 : 0);
  using KernT = void (*)(const KronTabDev<N>, const BoxDev, const KronMarchMaps, const MarchCommDev, const MarchRun*, const int*, const int);
  const KernT kerns[4] = {dg_kronecker_march_kernel<N, HIER, TX, TY, false, false>, dg_kronecker_march_kernel<N, HIER, TX, TY, true, false>,
                          dg_kronecker_march_kernel<N, HIER, TX, TY, false, true>, dg_kronecker_march_kernel<N, HIER, TX, TY, true, true>};
  KernT kern = kerns[variant];
  int rc = ensure_smem_attr(ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
  const long long total = (long long)ncols * on[2];
  const int grid = (int)std::max(1ll, std::min(total, (long long)ctx->sms));
  const bool fuse = fuse_exchange && !op->active_box && op->halo_p2p.built && op->halo_p2p.march_ok;
  // run list of this launch shape
  MarchScheduleKey sk; std::memset(&sk, 0, sizeof(sk));
  for (int d = 0; d < 3; ++d) sk.on[d] = on[d];
  sk.grid = grid; sk.ifz_lo = fuse && op->halo_p2p.march.enabled[1 + 3 * 0]; sk.ifz_hi = fuse && op->halo_p2p.march.enabled[1 + 3 * 2];
  sk.gx_lo = b.origin[0] + b.own_lo[0] == 0; sk.gx_hi = b.origin[0] + b.own_hi[0] == b.gn[0];
  sk.gy_lo = b.origin[1] + b.own_lo[1] == 0; sk.gy_hi = b.origin[1] + b.own_hi[1] == b.gn[1];
  sk.ify_lo = fuse && op->halo_p2p.march.enabled[0 + 3 * 1]; sk.ify_hi = fuse && op->halo_p2p.march.enabled[2 + 3 * 1];
  const MarchSchedule* sched = nullptr;
  for (const MarchSchedule& sc : mc.schedules) if (sc.key == sk) { sched = &sc; break; }
  if (!sched) {
    REQUIRE(!op->capturing, B200FEM_ERR_INVALID, "marching kernel: the run list must exist before graph capture (apply once before capturing)");
    std::vector<MarchRun> runs; std::vector<int> begin; march_schedule(sk, tx, ty, runs, begin);
    MarchSchedule sc; sc.key = sk;
    CUDA_OK(cudaMalloc(&sc.d_runs, sizeof(MarchRun) * std::max<size_t>(runs.size(), 1))); CUDA_OK(cudaMalloc(&sc.d_begin, sizeof(int) * begin.size()));
    CUDA_OK(cudaMemcpyAsync(sc.d_runs, runs.data(), sizeof(MarchRun) * runs.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(sc.d_begin, begin.data(), sizeof(int) * begin.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));            // (the host vectors go out of scope)
    mc.schedules.push_back(sc); sched = &mc.schedules.back();
  }
  // the Copy exchange of w rides along when this launch covers the whole owned box of a rank whose neighbours all lie in
  // the y-z plane of the process grid (comm.cuh: MarchCommDev)
  MarchCommDev C; std::memset(&C, 0, sizeof(C));
  op->exchange_fused = false;
  if (fuse) { C = op->halo_p2p.march; C.w = w; op->exchange_fused = true; }
  else if (std::getenv("B200FEM_MARCH_TS") && !op->capturing) {
    if (!mc.d_ts) { CUDA_OK(cudaMalloc(&mc.d_ts, (16 + 4 * 160) * sizeof(unsigned long long))); CUDA_OK(cudaMemset(mc.d_ts, 0, (16 + 4 * 160) * sizeof(unsigned long long))); }
    C.ts = mc.d_ts;
  }
  // programmatic dependent launch: the CTAs of this launch may be scheduled while the previous kernel of the stream drains;
  // the kernel itself waits (griddepcontrol.wait) before it touches global memory
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(Cfg::kThreads); cfg.dynamicSmemBytes = Cfg::smem_bytes(); cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
  // (not for the launches that carry the exchange: their CTAs all stay until the receive is done, nothing of the next launch can move
  // in early, and the programmatic release after such a grid was measured 1.5-2.3 us SLOWER per step than plain stream order at N = 2)
  cfg.attrs = attr; cfg.numAttrs = fuse ? 0 : 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, K, b, mc.maps[slot], C, (const MarchRun*)sched->d_runs, (const int*)sched->d_begin, tx));
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}

int launch_dg_march(b200fem_operator* op, const double* u, double* w, const double* bvec, bool fuse_exchange) {
  return op->sp->kind == B200FEM_DG_LEGENDRE_HIER ? launch_march<true>(op, u, w, bvec, fuse_exchange) : launch_march<false>(op, u, w, bvec, fuse_exchange);
}

}  // namespace b200fem

// host-only: the run list the marching kernel would use (b200fem.h)
extern "C" int b200fem_march_schedule(const int32_t* on, int grid, int flags, int32_t* runs_out, int32_t cap, int32_t* begin_out, int32_t* nruns_out) {
  REQUIRE(on && begin_out && nruns_out && grid >= 1 && on[0] >= 1 && on[1] >= 1 && on[2] >= 1, B200FEM_ERR_INVALID, "march_schedule: bad argument");
  MarchScheduleKey k; std::memset(&k, 0, sizeof(k));
  for (int d = 0; d < 3; ++d) k.on[d] = on[d];
  k.grid = grid; k.ifz_lo = flags & 1; k.ifz_hi = (flags >> 1) & 1; k.gx_lo = (flags >> 2) & 1; k.gx_hi = (flags >> 3) & 1; k.gy_lo = (flags >> 4) & 1; k.gy_hi = (flags >> 5) & 1; k.ify_lo = (flags >> 6) & 1; k.ify_hi = (flags >> 7) & 1;
  std::vector<MarchRun> runs; std::vector<int> begin;
  march_schedule(k, (on[0] + 15) / 16, (on[1] + 15) / 16, runs, begin);
  *nruns_out = (int32_t)runs.size();
  for (int i = 0; i <= grid; ++i) begin_out[i] = begin[(size_t)i];
  if (runs_out) for (size_t i = 0; i < runs.size() && (int32_t)i < cap; ++i) { runs_out[4 * i] = runs[i].col; runs_out[4 * i + 1] = runs[i].za; runs_out[4 * i + 2] = runs[i].zb; runs_out[4 * i + 3] = runs[i].flush; }
  return B200FEM_OK;
}
