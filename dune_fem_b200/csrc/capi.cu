// capi.cu -- C ABI (include/b200fem.h) over the CUDA kernels.  Host logic only: handles, tables, launch
// configuration, the CG driver, halo exchange.  No CPU compute fallback exists: every compute entry point needs a
// CUDA device and reports B200FEM_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/b200fem.h"
#include "dg_kronecker.cuh"
#include "dg_kronecker_pipe.cuh"
#include "dg_kronecker_march.cuh"
#include "dg_kronecker_slab.cuh"
#include "dg_kronecker_tensor.cuh"
#include "dg_kronecker_tma.cuh"
#include "dg_quadrature.cuh"
#include "halo.cuh"
#include "integrands.cuh"
#include "kron_tables.hpp"
#include "lagrange_kronecker.cuh"
#include "lagrange_quadrature.cuh"
#include "tables.hpp"
#include "vec_kernels.cuh"
#include "cg_coop2d.cuh"

using namespace b200fem;

static thread_local std::string g_error;
static int fail(int code, const std::string& msg) { g_error = msg; return code; }
#define CUDA_OK(expr)                                                                                   \
  do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(B200FEM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)
#define REQUIRE(cond, code, msg) do { if (!(cond)) return fail(code, msg); } while (0)

// ---------------------------------------------------------------------------------------------------------------
struct b200fem_ctx {
  int device = 0; cudaStream_t stream = nullptr; bool own_stream = false;
  NcclApi nccl; void* comm = nullptr; bool own_comm = false; int rank = 0, world = 1;
};
struct b200fem_mesh {
  b200fem_ctx* ctx; int dim; int gn[3]; double lo[3], hi[3], h[3];
  int proc[3], pc[3];            // process grid and this rank's coordinates
  BoxDev box;                    // local box incl. ghost layers (ghost layers only used by DG spaces)
  int olo[3], ohi[3];            // owned range in global element coordinates
};
struct b200fem_space {
  b200fem_mesh* mesh; int kind, order, numbering, n1, nb; long long size, elements;
  BoxDev box;                    // DG: mesh box with ghosts; Lagrange: owned elements only
  Tab1D tab; std::vector<int> perm;
  LagrangeLayoutDev lay; long long* d_lattice_map = nullptr; std::vector<long long> lattice_map;
};
struct b200fem_operator {
  b200fem_space* sp; b200fem_model model; int kernel_pref = B200FEM_KERNEL_AUTO; bool communicate = true;
  unsigned q_interior = 0, q_surface = 0; bool inverse_mass = false;
  int* d_perm = nullptr; double* d_bvec = nullptr; uint8_t* d_dmask = nullptr; double* d_dvals = nullptr; uint8_t* d_aux = nullptr;
  std::vector<uint8_t> h_dmask; std::vector<double> h_dvals;
  double *d_u = nullptr, *d_w = nullptr;                       // staging for the host-pointer API
  double *d_h = nullptr, *d_r = nullptr, *d_p = nullptr, *d_x = nullptr, *d_b = nullptr, *d_partial = nullptr, *d_sums = nullptr, *d_hist = nullptr;
  CgState* d_cg = nullptr; int hist_cap = 0; unsigned int* d_counter = nullptr;
  bool jac_mode = false; double *d_jac_u = nullptr, *d_jac_opu = nullptr, *d_jac_b = nullptr; FdState* d_fd = nullptr;   // AutomaticDifferenceLinearOperator
  double* d_dinv = nullptr; double *d_pq = nullptr, *d_ps = nullptr; bool dinv_mass = false;   // Jacobi preconditioner: 1 / diag(A) (+ the inverse-mass state it was built for), PCG work vectors
  std::vector<double*> gmres_v; double* d_gm_partial = nullptr; double* d_gm_sums = nullptr; int gm_cap = 0;   // GMRES basis and reduction scratch
  double *d_rstar = nullptr, *d_s = nullptr, *d_tmp = nullptr, *d_partial5 = nullptr, *d_sums5 = nullptr; BicgState* d_bicg = nullptr;   // BiCGStab work vectors
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr; cudaEvent_t pipe_ev[2 * 16 + 2] = {};   // host-pointer apply: copy/compute pipeline
  bool want_dot = false; int dot_parts = 0; double* d_dot_partial = nullptr; int dot_cap = 0;     // <u, A u> fused into the lattice kernel (CG)
  cudaGraphExec_t cg_graph = nullptr; const void* cg_graph_key[3] = {nullptr, nullptr, nullptr}; bool capturing = false;
  bool kron_ready = false; int kron_chk = -1; bool fuse_dirichlet = false, fuse_linear = false, dirichlet_fused = false; double* d_lag_rows = nullptr; LagKronRows lag_rows{}; std::vector<unsigned char> kron_tab; struct KronMapCache* map_cache = nullptr; struct MarchMapCache* march_cache = nullptr;
  HaloPlan halo; HaloPlanDG halo_dg; HaloPlanP2P halo_p2p; const BoxDev* active_box = nullptr; int reserve_sms = 0; unsigned long long fused_seq = 0; bool last_launch_tensor = false;      // active_box: sub-box override for split launches
  cudaStream_t comm_stream = nullptr; cudaEvent_t ev_bnd = nullptr, ev_comm = nullptr, dbg_ev[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evx0 = nullptr, evx1 = nullptr; b200fem_timing timing{};
  bool timing_enabled = false;     // event records cost ~1.5 us each on the host: only after b200fem_operator_timing was asked for
};

extern "C" const char* b200fem_last_error(void) { return g_error.c_str(); }
extern "C" int b200fem_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------------------------
extern "C" int b200fem_ctx_create(int device, void* stream, b200fem_ctx** out) {
  REQUIRE(out, B200FEM_ERR_INVALID, "ctx_create: out is null");
  int count = 0; cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(B200FEM_ERR_CUDA, "no CUDA device available: this library has no CPU fallback");
  REQUIRE(device >= 0 && device < count, B200FEM_ERR_INVALID, "ctx_create: bad device index");
  CUDA_OK(cudaSetDevice(device));
  auto* c = new b200fem_ctx; c->device = device;
  if (stream) c->stream = (cudaStream_t)stream; else { CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  *out = c; return B200FEM_OK;
}
extern "C" int b200fem_ctx_destroy(b200fem_ctx* c) {
  if (!c) return B200FEM_OK;
  if (c->own_comm && c->comm && c->nccl.ok()) c->nccl.CommDestroy(c->comm);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c; return B200FEM_OK;
}
extern "C" int b200fem_ctx_synchronize(b200fem_ctx* c) { REQUIRE(c, B200FEM_ERR_INVALID, "null ctx"); CUDA_OK(cudaSetDevice(c->device)); CUDA_OK(cudaStreamSynchronize(c->stream)); return B200FEM_OK; }
extern "C" int b200fem_malloc(b200fem_ctx* c, int64_t bytes, void** dev) { REQUIRE(c && dev, B200FEM_ERR_INVALID, "malloc: null"); CUDA_OK(cudaSetDevice(c->device)); CUDA_OK(cudaMalloc(dev, (size_t)bytes)); return B200FEM_OK; }
extern "C" int b200fem_free(b200fem_ctx* c, void* dev) { REQUIRE(c, B200FEM_ERR_INVALID, "free: null"); CUDA_OK(cudaSetDevice(c->device)); CUDA_OK(cudaFree(dev)); return B200FEM_OK; }
extern "C" int b200fem_memcpy_h2d(b200fem_ctx* c, void* dev, const void* host, int64_t bytes) {
  REQUIRE(c, B200FEM_ERR_INVALID, "null ctx"); CUDA_OK(cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, c->stream)); CUDA_OK(cudaStreamSynchronize(c->stream)); return B200FEM_OK; }
extern "C" int b200fem_memcpy_d2h(b200fem_ctx* c, void* host, const void* dev, int64_t bytes) {
  REQUIRE(c, B200FEM_ERR_INVALID, "null ctx"); CUDA_OK(cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream)); CUDA_OK(cudaStreamSynchronize(c->stream)); return B200FEM_OK; }

// ---------------------------------------------------------------------------------------------------------------
static int make_mesh(b200fem_ctx* ctx, int dim, const int32_t* n, const double* lo, const double* hi, const int32_t* proc, int rank, b200fem_mesh** out) {
  REQUIRE(ctx && n && lo && hi && out, B200FEM_ERR_INVALID, "mesh: null argument");
  REQUIRE(dim == 2 || dim == 3, B200FEM_ERR_NOT_IMPLEMENTED, "mesh: dim must be 2 or 3");
  auto* m = new b200fem_mesh; m->ctx = ctx; m->dim = dim;
  int world = 1;
  for (int d = 0; d < 3; ++d) {
    m->gn[d] = d < dim ? n[d] : 1; m->lo[d] = d < dim ? lo[d] : 0.0; m->hi[d] = d < dim ? hi[d] : 1.0;
    m->proc[d] = (proc && d < dim) ? proc[d] : 1; world *= m->proc[d];
    if (m->gn[d] < 1 || !(m->hi[d] > m->lo[d]) || m->proc[d] < 1 || m->proc[d] > m->gn[d]) { delete m; return fail(B200FEM_ERR_INVALID, "mesh: bad extents"); }
    m->h[d] = (m->hi[d] - m->lo[d]) / m->gn[d];
  }
  if (rank < 0 || rank >= world) { delete m; return fail(B200FEM_ERR_INVALID, "mesh: rank outside process grid"); }
  m->pc[0] = rank % m->proc[0]; m->pc[1] = (rank / m->proc[0]) % m->proc[1]; m->pc[2] = rank / (m->proc[0] * m->proc[1]);
  BoxDev& b = m->box; b.dim = dim;
  for (int d = 0; d < 3; ++d) {
    // block distribution: the first (gn % proc) ranks along an axis get one extra cell
    const int q = m->gn[d] / m->proc[d], r = m->gn[d] % m->proc[d], c = m->pc[d];
    m->olo[d] = c * q + std::min(c, r); m->ohi[d] = m->olo[d] + q + (c < r ? 1 : 0);
    const int glo = m->olo[d] > 0 ? 1 : 0, ghi = m->ohi[d] < m->gn[d] ? 1 : 0;      // overlap 1 where a neighbour rank exists
    b.origin[d] = m->olo[d] - glo; b.n[d] = (m->ohi[d] - m->olo[d]) + glo + ghi;
    b.own_lo[d] = glo; b.own_hi[d] = glo + (m->ohi[d] - m->olo[d]);
    b.gn[d] = m->gn[d]; b.lo[d] = m->lo[d]; b.h[d] = m->h[d];
  }
  *out = m; return B200FEM_OK;
}
extern "C" int b200fem_mesh_cartesian(b200fem_ctx* ctx, int dim, const int32_t* n, const double* lo, const double* hi, b200fem_mesh** out) {
  return make_mesh(ctx, dim, n, lo, hi, nullptr, 0, out);
}
extern "C" int b200fem_mesh_cartesian_distributed(b200fem_ctx* ctx, int dim, const int32_t* n, const double* lo, const double* hi, const int32_t* proc, int rank, b200fem_mesh** out) {
  REQUIRE(proc, B200FEM_ERR_INVALID, "mesh: proc is null");
  return make_mesh(ctx, dim, n, lo, hi, proc, rank, out);
}
extern "C" int b200fem_mesh_destroy(b200fem_mesh* m) { delete m; return B200FEM_OK; }
extern "C" int b200fem_partition_box(int dim, const int32_t* n, const int32_t* proc, int rank, int overlap, int32_t* out) {
  REQUIRE(n && proc && out && (dim == 2 || dim == 3), B200FEM_ERR_INVALID, "partition_box: bad argument");
  int p[3] = {1, 1, 1}, g[3] = {1, 1, 1}, world = 1;
  for (int d = 0; d < dim; ++d) { p[d] = proc[d]; g[d] = n[d]; REQUIRE(p[d] >= 1 && p[d] <= g[d], B200FEM_ERR_INVALID, "partition_box: bad process grid"); world *= p[d]; }
  REQUIRE(rank >= 0 && rank < world, B200FEM_ERR_INVALID, "partition_box: rank outside process grid");
  const int pc[3] = {rank % p[0], (rank / p[0]) % p[1], rank / (p[0] * p[1])};
  for (int d = 0; d < 3; ++d) {
    const int q = g[d] / p[d], r = g[d] % p[d], c = pc[d];
    const int olo = c * q + std::min(c, r), ohi = olo + q + (c < r ? 1 : 0);
    const int glo = (overlap && olo > 0) ? 1 : 0, ghi = (overlap && ohi < g[d]) ? 1 : 0;
    out[d] = olo - glo; out[3 + d] = (ohi - olo) + glo + ghi; out[6 + d] = glo; out[9 + d] = glo + (ohi - olo);
  }
  return B200FEM_OK;
}
extern "C" int b200fem_mesh_local_box(b200fem_mesh* m, int overlap, int32_t* out) {
  REQUIRE(m && out, B200FEM_ERR_INVALID, "mesh_local_box: null argument");
  const int rank = m->pc[0] + m->proc[0] * (m->pc[1] + m->proc[1] * m->pc[2]);
  return b200fem_partition_box(m->dim, m->gn, m->proc, rank, overlap, out);
}

// ---------------------------------------------------------------------------------------------------------------
// AdaptiveLeafIndexSet first-touch numbering of the Lagrange lattice (gridpart/adaptiveleafindexset.hh:884-906)
static void build_adaptive_leaf_map(b200fem_space* s) {
  const BoxDev& b = s->box; const int dim = b.dim, k = s->order;
  const long long L0 = s->lay.lattice[0], L1 = s->lay.lattice[1], L2 = s->lay.lattice[2];
  s->lattice_map.assign((size_t)(L0 * L1 * L2), -1);
  long long cnt[4] = {0, 0, 0, 0}, type_off[4] = {0, 0, 0, 0}, counter[4] = {0, 0, 0, 0};
  for (int sft = 0; sft < (1 << dim); ++sft) { if (s->lay.group_offset[sft] < 0) continue; long long c = 1; for (int d = 0; d < 3; ++d) c *= s->lay.group_dims[sft][d]; cnt[__builtin_popcount(sft)] += c; }
  for (int p = 1; p <= dim; ++p) type_off[p] = type_off[p - 1] + cnt[p - 1];
  // sub-entities of the cube in reference-element order, as lattice offsets in {0,1,2}
  std::vector<std::array<int, 3>> subs[4];
  for (int v = 0; v < (1 << dim); ++v) { std::array<int, 3> a = {0, 0, 0}; for (int d = 0; d < dim; ++d) a[d] = 2 * ((v >> d) & 1); subs[0].push_back(a); }
  if (dim == 2) { subs[1] = {{0, 1, 0}, {2, 1, 0}, {1, 0, 0}, {1, 2, 0}}; subs[2] = {{1, 1, 0}}; }
  if (dim == 3) {
    subs[1] = {{0, 0, 1}, {2, 0, 1}, {0, 2, 1}, {2, 2, 1}, {0, 1, 0}, {2, 1, 0}, {1, 0, 0}, {1, 2, 0}, {0, 1, 2}, {2, 1, 2}, {1, 0, 2}, {1, 2, 2}};
    subs[2] = {{0, 1, 1}, {2, 1, 1}, {1, 0, 1}, {1, 2, 1}, {1, 1, 0}, {1, 1, 2}};
    subs[3] = {{1, 1, 1}};
  }
  for (int e2 = 0; e2 < b.n[2]; ++e2) for (int e1 = 0; e1 < b.n[1]; ++e1) for (int e0 = 0; e0 < b.n[0]; ++e0) {
    const int ec[3] = {e0, e1, e2};
    for (int cd = 0; cd <= dim; ++cd) { const int pd = dim - cd; if (k == 1 && pd != 0) continue;
      for (auto& a : subs[pd]) {
        long long g[3] = {0, 0, 0}; for (int d = 0; d < dim; ++d) g[d] = (long long)k * ec[d] + (k == 1 ? a[d] / 2 : a[d]);
        long long& slot = s->lattice_map[(size_t)(g[0] + L0 * (g[1] + L1 * g[2]))];
        if (slot < 0) slot = type_off[pd] + counter[pd]++;
      } }
  }
}

extern "C" int b200fem_space_create(b200fem_mesh* mesh, int kind, int order, int numbering, b200fem_space** out) {
  REQUIRE(mesh && out, B200FEM_ERR_INVALID, "space_create: null argument");
  REQUIRE(kind >= 0 && kind <= 2, B200FEM_ERR_INVALID, "space_create: unknown space kind");
  try {
    auto s = std::unique_ptr<b200fem_space>(new b200fem_space);
    s->mesh = mesh; s->kind = kind; s->order = order; s->numbering = numbering; s->n1 = order + 1;
    const int dim = mesh->dim; s->nb = 1; for (int d = 0; d < dim; ++d) s->nb *= s->n1;
    s->box = mesh->box;
    if (kind == B200FEM_LAGRANGE) {
      REQUIRE(order == 1 || order == 2, B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange spaces: order 1 and 2 only");
      BoxDev& b = s->box;      // continuous spaces need no ghost elements: the local box is the owned box
      for (int d = 0; d < 3; ++d) { b.origin[d] = mesh->olo[d]; b.n[d] = mesh->ohi[d] - mesh->olo[d]; b.own_lo[d] = 0; b.own_hi[d] = b.n[d]; }
      LagrangeLayoutDev& L = s->lay; L.order = order; L.lattice_map = nullptr;
      for (int d = 0; d < 3; ++d) L.lattice[d] = d < dim ? (long long)order * b.n[d] + 1 : 1;
      long long off = 0;
      for (int sft = 0; sft < 8; ++sft) { L.group_offset[sft] = -1; for (int d = 0; d < 3; ++d) L.group_dims[sft][d] = 1; }
      for (int pc = 0; pc <= dim; ++pc) for (int sft = 0; sft < (1 << dim); ++sft) {
        if (__builtin_popcount(sft) != pc || (order == 1 && sft != 0)) continue;
        L.group_offset[sft] = off; long long c = 1;
        for (int d = 0; d < 3; ++d) { L.group_dims[sft][d] = d < dim ? b.n[d] + (((sft >> d) & 1) ? 0 : 1) : 1; c *= L.group_dims[sft][d]; }
        off += c;
      }
      s->size = off; s->elements = (long long)b.n[0] * b.n[1] * b.n[2];
      s->tab = tabulate_1d(Basis::Lagrange, order, gauss_points_for_order(2 * order));
      if (numbering == B200FEM_NUMBERING_ADAPTIVE_LEAF) {
        build_adaptive_leaf_map(s.get());
        CUDA_OK(cudaSetDevice(mesh->ctx->device));
        CUDA_OK(cudaMalloc(&s->d_lattice_map, s->lattice_map.size() * sizeof(long long)));
        CUDA_OK(cudaMemcpy(s->d_lattice_map, s->lattice_map.data(), s->lattice_map.size() * sizeof(long long), cudaMemcpyHostToDevice));
        L.lattice_map = s->d_lattice_map;
      }
    } else {
      REQUIRE(order >= 1 && order <= 5, B200FEM_ERR_NOT_IMPLEMENTED, "DG Legendre spaces: orders 1..5");
      REQUIRE(dim == 3, B200FEM_ERR_NOT_IMPLEMENTED, "DG Legendre spaces: 3-D boxes only");
      const BoxDev& b = s->box;
      s->elements = (long long)b.n[0] * b.n[1] * b.n[2]; s->size = s->elements * s->nb;
      s->tab = tabulate_1d(Basis::Legendre, order, gauss_points_for_order(2 * order));
      s->perm = legendre_local_permutation(dim, order, kind == B200FEM_DG_LEGENDRE_HIER);
    }
    *out = s.release(); return B200FEM_OK;
  } catch (const std::exception& ex) { return fail(B200FEM_ERR_INVALID, ex.what()); }
}
extern "C" int b200fem_space_destroy(b200fem_space* s) { if (s && s->d_lattice_map) cudaFree(s->d_lattice_map); delete s; return B200FEM_OK; }
extern "C" int b200fem_space_size(b200fem_space* s, int64_t* size) { REQUIRE(s && size, B200FEM_ERR_INVALID, "null"); *size = s->size; return B200FEM_OK; }
extern "C" int b200fem_space_local_size(b200fem_space* s, int32_t* nb) { REQUIRE(s && nb, B200FEM_ERR_INVALID, "null"); *nb = s->nb; return B200FEM_OK; }
extern "C" int b200fem_space_elements(b200fem_space* s, int64_t* n) { REQUIRE(s && n, B200FEM_ERR_INVALID, "null"); *n = s->elements; return B200FEM_OK; }
extern "C" int b200fem_space_dofmap(b200fem_space* s, int64_t e, int64_t* out) {
  REQUIRE(s && out, B200FEM_ERR_INVALID, "null"); REQUIRE(e >= 0 && e < s->elements, B200FEM_ERR_INVALID, "dofmap: element out of range");
  if (s->kind != B200FEM_LAGRANGE) { for (int j = 0; j < s->nb; ++j) out[j] = e * s->nb + j; return B200FEM_OK; }
  const BoxDev& b = s->box; const int n1 = s->n1, k = s->order;
  const int ec[3] = {(int)(e % b.n[0]), (int)((e / b.n[0]) % b.n[1]), (int)(e / ((long long)b.n[0] * b.n[1]))};
  LagrangeLayoutDev L = s->lay; L.lattice_map = s->lattice_map.empty() ? nullptr : s->lattice_map.data();
  for (int l = 0; l < s->nb; ++l) {                      // local numbering: coordinate 0 fastest (genericlagrangepoints.hh:862-876)
    const int a0 = l % n1, a1 = (l / n1) % n1, a2 = b.dim == 3 ? l / (n1 * n1) : 0;
    out[l] = lagrange_dof(L, (long long)k * ec[0] + a0, (long long)k * ec[1] + a1, b.dim == 3 ? (long long)k * ec[2] + a2 : 0);
  }
  return B200FEM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
template <int N> static DgTabDev<N> make_tab(const Tab1D& t) {
  DgTabDev<N> T;
  for (int i = 0; i < N * N; ++i) { T.B[i] = t.B[i]; T.G[i] = t.G[i]; }
  for (int i = 0; i < N; ++i) { T.x[i] = t.x[i]; T.w[i] = t.w[i]; T.phi[0][i] = t.phi0[i]; T.phi[1][i] = t.phi1[i]; T.dphi[0][i] = t.dphi0[i]; T.dphi[1][i] = t.dphi1[i]; }
  return T;
}
static AdrIntegrands make_integrands(const b200fem_operator* op, bool with_data) { AdrIntegrands I; I.m = op->model; I.dim = op->sp->box.dim; I.with_data = with_data; return I; }

// factor applied to every element result: 1, or referenceVolume / volume when the operator acts as MOLGalerkinOperator
static double mass_scale(const b200fem_operator* op) {
  if (!op->inverse_mass) return 1.0;
  const BoxDev& b = op->sp->box; double vol = 1; for (int d = 0; d < b.dim; ++d) vol *= b.h[d];
  return 1.0 / vol;
}
static bool default_quadrature(const b200fem_operator* op) {
  const int k = op->sp->order;
  const int mi = gauss_points_for_order(op->q_interior ? (int)op->q_interior : 2 * k), ms = gauss_points_for_order(op->q_surface ? (int)op->q_surface : 2 * k + 1);
  return mi == k + 1 && ms == k + 1;
}

template <int N> static int launch_dg_quadrature(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data) {
  using Cfg = DgQuadCfg<N>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  const long long n_owned = (long long)(b.own_hi[0] - b.own_lo[0]) * (b.own_hi[1] - b.own_lo[1]) * (b.own_hi[2] - b.own_lo[2]);
  auto kern = dg_quadrature_kernel<N, AdrIntegrands>;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes()));
  const unsigned grid = (unsigned)((n_owned + Cfg::EB - 1) / Cfg::EB);
  kern<<<grid, Cfg::kThreads, Cfg::smem_bytes(), op->sp->mesh->ctx->stream>>>(make_tab<N>(op->sp->tab), b, make_integrands(op, with_data), op->d_perm, u, w, bvec, n_owned, mass_scale(op));
  CUDA_OK(cudaGetLastError()); return B200FEM_OK;
}
template <int N, int TX, int TY, int TZ> static int launch_dg_kronecker(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  using Cfg = KronCfg<N, TX, TY, TZ>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
  KronTabDev<N> K;
  for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K.S[d][i] = kh.S[d][i]; K.Dlo[d][i] = kh.Dlo[d][i]; K.Dhi[d][i] = kh.Dhi[d][i]; K.L[d][i] = kh.L[d][i]; K.R[d][i] = kh.R[d][i]; }
  const int tx = (b.own_hi[0] - b.own_lo[0] + TX - 1) / TX, ty = (b.own_hi[1] - b.own_lo[1] + TY - 1) / TY, tz = (b.own_hi[2] - b.own_lo[2] + TZ - 1) / TZ;
  auto kern = dg_kronecker_kernel<N, TX, TY, TZ>;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes()));
  kern<<<(unsigned)(tx * ty * tz), Cfg::kThreads, Cfg::smem_bytes(), op->sp->mesh->ctx->stream>>>(K, b, op->d_perm, u, w, bvec, tx, ty);
  CUDA_OK(cudaGetLastError()); return B200FEM_OK;
}
template <int N, bool HIER> static int launch_dg_kronecker_tma(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  constexpr int TX = 8, TY = 4, TZ = 4;
  using Cfg = KronTmaCfg<N, TX, TY, TZ>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
  KronTabDev<N> K;
  for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K.S[d][i] = kh.S[d][i]; K.Dlo[d][i] = kh.Dlo[d][i]; K.Dhi[d][i] = kh.Dhi[d][i]; K.L[d][i] = kh.L[d][i]; K.R[d][i] = kh.R[d][i]; }
  const int tx = (b.own_hi[0] - b.own_lo[0] + TX - 1) / TX, ty = (b.own_hi[1] - b.own_lo[1] + TY - 1) / TY, tz = (b.own_hi[2] - b.own_lo[2] + TZ - 1) / TZ;
  auto kern = dg_kronecker_tma_kernel<N, HIER, TX, TY, TZ>;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes()));
  kern<<<(unsigned)(tx * ty * tz), Cfg::kThreads, Cfg::smem_bytes(), op->sp->mesh->ctx->stream>>>(K, b, u, w, bvec, tx, ty);
  CUDA_OK(cudaGetLastError()); return B200FEM_OK;
}
// ---- TMA tensor maps (driver entry point resolved at run time: no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static bool ensure_encode_tiled() {
  if (g_encode_tiled) return true;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
  g_encode_tiled = (EncodeTiledFn)fn; return true;
}
// 3-D tensor of doubles [d2][d1][d0] with byte strides s1, s2 and box b0 x b1 x b2
static bool make_map3(CUtensorMap* m, const double* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1, uint32_t b2) {
  const cuuint64_t dims[3] = {d0, d1, d2}; const cuuint64_t strides[2] = {s1, s2};
  const cuuint32_t boxd[3] = {b0, b1, b2}; const cuuint32_t estr[3] = {1, 1, 1};
  return g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <int N> static bool tensor_path_ok(const b200fem_operator* op, const double* u, const double* w, const double* bvec) {
  const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return N == 3 && b.n[0] % 2 == 0 && b.own_lo[0] % 2 == 0 && al16(u) && al16(w) && (!bvec || al16(bvec)) && ensure_encode_tiled();
}
// Host-side cost matters at 40 us per apply: the 1-D operator tables are built once per operator, the kernel attribute
// is set once per instantiation, and encoded tensor maps are cached per (u, w, b, owned range).
struct KronMapKey { const void *u, *w, *b; int lo[3], hi[3]; bool operator==(const KronMapKey& o) const { return std::memcmp(this, &o, sizeof(KronMapKey)) == 0; } };
struct KronMapCache { static constexpr int kSlots = 32; KronMapKey key[kSlots]; KronTensorMaps maps[kSlots]; bool valid[kSlots] = {}; int next = 0; };
template <int N, bool HIER, bool SPLIT> static int launch_dg_kronecker_tensor(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  constexpr int TX = 8, TY = 4, TZ = 4, N3 = N * N * N;
  using Cfg = KronTensorCfg<N, TX, TY, TZ, SPLIT>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  if (!op->kron_ready) {
    KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
    op->kron_tab.resize(sizeof(KronTabDev<N>));
    KronTabDev<N>& K0 = *reinterpret_cast<KronTabDev<N>*>(op->kron_tab.data());
    for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K0.S[d][i] = kh.S[d][i]; K0.Dlo[d][i] = kh.Dlo[d][i]; K0.Dhi[d][i] = kh.Dhi[d][i]; K0.L[d][i] = kh.L[d][i]; K0.R[d][i] = kh.R[d][i]; }
    op->kron_ready = true;
  }
  const KronTabDev<N>& K = *reinterpret_cast<const KronTabDev<N>*>(op->kron_tab.data());
  const int on[3] = {b.own_hi[0] - b.own_lo[0], b.own_hi[1] - b.own_lo[1], b.own_hi[2] - b.own_lo[2]};
  const int tx = (on[0] + TX - 1) / TX, ty = (on[1] + TY - 1) / TY, tz = (on[2] + TZ - 1) / TZ, ntiles = tx * ty * tz;
  if (!op->map_cache) op->map_cache = new KronMapCache;
  KronMapCache& mc = *op->map_cache;
  KronMapKey key; std::memset(&key, 0, sizeof(key)); key.u = u; key.w = w; key.b = bvec;
  for (int d = 0; d < 3; ++d) { key.lo[d] = b.own_lo[d]; key.hi[d] = b.own_hi[d]; }
  int slot = -1;
  for (int i = 0; i < KronMapCache::kSlots; ++i) if (mc.valid[i] && mc.key[i] == key) { slot = i; break; }
  if (slot < 0) {
    slot = mc.next; mc.next = (mc.next + 1) % KronMapCache::kSlots;
    const uint64_t s1 = (uint64_t)b.n[0] * N3 * 8, s2 = s1 * b.n[1];
    const long long own_off = ((long long)b.own_lo[0] + (long long)b.n[0] * (b.own_lo[1] + (long long)b.n[1] * b.own_lo[2])) * N3;
    KronTensorMaps& M = mc.maps[slot];
    bool ok = make_map3(&M.u_tile, u, (uint64_t)b.n[0] * N3, b.n[1], b.n[2], s1, s2, TX * N3, TY, TZ) &&
              make_map3(&M.u_xhalo, u, (uint64_t)b.n[0] * N3, b.n[1], b.n[2], s1, s2, 2 * N3, TY, TZ) &&
              make_map3(&M.u_yhalo, u, (uint64_t)b.n[0] * N3, b.n[1], b.n[2], s1, s2, TX * N3, 1, TZ) &&
              make_map3(&M.u_zhalo, u, (uint64_t)b.n[0] * N3, b.n[1], b.n[2], s1, s2, TX * N3, TY, 1) &&
              make_map3(&M.w_tile, w + own_off, (uint64_t)on[0] * N3, on[1], on[2], s1, s2, TX * N3, TY, TZ) &&
              make_map3(&M.b_tile, (bvec ? bvec : w) + own_off, (uint64_t)on[0] * N3, on[1], on[2], s1, s2, TX * N3, TY, TZ);
    REQUIRE(ok, B200FEM_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    mc.key[slot] = key; mc.valid[slot] = true;
  }
  static int sms = 0;
  if (!sms) CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, op->sp->mesh->ctx->device));
  auto kern = dg_kronecker_tensor_kernel<N, HIER, TX, TY, TZ, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) { CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes())); attr_set = true; }
  // a persistent grid fills every SM for its whole run time; while a halo exchange is in flight a few SMs are left free so
  // that the exchange kernels are guaranteed to run concurrently
  const int grid = std::max(1, std::min(ntiles, sms - op->reserve_sms));
  KronSendDev snd; std::memset(&snd, 0, sizeof(snd));
  if (op->fused_seq) {          // this launch also sends the y/z face halos of w (see apply_dev_impl)
    HaloPlanP2P& hp = op->halo_p2p;
    for (int i = 0; i < hp.nnb; ++i) {
      const int c = hp.dir_code[i], dx = c % 3 - 1, dy = (c / 3) % 3 - 1, dz = c / 9 - 1;
      if (dx != 0) continue;
      const int d = (dy + 1) + 3 * (dz + 1);
      snd.enabled[d] = 1; snd.any = 1;
      snd.remote[d][0] = hp.host_nb[i].remote_data[0]; snd.remote[d][1] = hp.host_nb[i].remote_data[1];
      snd.remote_ready[d] = hp.host_nb[i].remote_ready; snd.local_ack[d] = hp.host_nb[i].local_ack;
      snd.expected[d] = (unsigned)(tx * (dy == 0 ? ty : 1) * (dz == 0 ? tz : 1));
    }
    snd.dir_counter = hp.d_cta_counter; snd.seq = op->fused_seq; snd.error = hp.d_error;
  }
  kern<<<(unsigned)grid, Cfg::kThreads, Cfg::smem_bytes(), op->sp->mesh->ctx->stream>>>(K, b, mc.maps[slot], snd, bvec ? 1 : 0, tx, ty, ntiles);
  CUDA_OK(cudaGetLastError()); return B200FEM_OK;
}

// 4-D tensor of doubles [d3][d2][d1][d0] (d0 contiguous) with byte strides s1..s3 and box b0 x b1 x b2 x b3
static bool make_map4(CUtensorMap* m, const double* base, const uint64_t (&d)[4], const uint64_t (&s)[3], const uint32_t (&bx)[4]) {
  const cuuint64_t dims[4] = {d[0], d[1], d[2], d[3]}; const cuuint64_t strides[3] = {s[0], s[1], s[2]};
  const cuuint32_t boxd[4] = {bx[0], bx[1], bx[2], bx[3]}; const cuuint32_t estr[4] = {1, 1, 1, 1};
  return g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <int N> static bool march_path_ok(const b200fem_operator* op, const double* u, const double* w, const double* bvec) {
  const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  return N == 3 && b.dim == 3 && (b.own_hi[0] - b.own_lo[0]) % 2 == 0 && tensor_path_ok<N>(op, u, w, bvec);
}
struct MarchMapCache { static constexpr int kSlots = 32; KronMapKey key[kSlots]; KronMarchMaps maps[kSlots]; bool valid[kSlots] = {}; int next = 0; };
// z-marching Kronecker kernel (dg_kronecker_march.cuh): persistent grid, every CTA gets the same number of plane-tiles
template <int N, bool HIER> static int launch_dg_kronecker_march(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  constexpr int TX = 16, TY = 16, N3 = N * N * N;
  using Cfg = KronMarchCfg<N, TX, TY>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  if (!op->kron_ready) {
    KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
    op->kron_tab.resize(sizeof(KronTabDev<N>));
    KronTabDev<N>& K0 = *reinterpret_cast<KronTabDev<N>*>(op->kron_tab.data());
    for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K0.S[d][i] = kh.S[d][i]; K0.Dlo[d][i] = kh.Dlo[d][i]; K0.Dhi[d][i] = kh.Dhi[d][i]; K0.L[d][i] = kh.L[d][i]; K0.R[d][i] = kh.R[d][i]; }
    op->kron_ready = true;
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
  static int sms = 0;
  if (!sms) CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, op->sp->mesh->ctx->device));
  // checkerboard self matrices on the y and z axes (no advection there: even and odd Legendre modes decouple); entries
  // that are zero up to quadrature rounding (<= 1e-14 of the matrix norm) are not multiplied at all
  if (op->kron_chk < 0) {
    bool chk = true;
    for (int d = 1; d < 3; ++d) {
      double mx = 0; for (int i = 0; i < N * N; ++i) mx = std::max(mx, std::fabs(K.S[d][i]));
      for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) if (((i + j) & 1) && std::fabs(K.S[d][i * N + j]) > 1e-14 * mx) chk = false;
    }
    op->kron_chk = chk && !std::getenv("B200FEM_NO_CHK") ? 1 : 0;
  }
  const int variant = (bvec ? 1 : 0) + (op->kron_chk ? 2 : 0);
  using KernT = void (*)(const KronTabDev<N>, const BoxDev, const KronMarchMaps, const int, const int);
  const KernT kerns[4] = {dg_kronecker_march_kernel<N, HIER, TX, TY, false, false>, dg_kronecker_march_kernel<N, HIER, TX, TY, true, false>,
                          dg_kronecker_march_kernel<N, HIER, TX, TY, false, true>, dg_kronecker_march_kernel<N, HIER, TX, TY, true, true>};
  KernT kern = kerns[variant];
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[variant]) { CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes())); attr_set[variant] = true; }
  const long long total = (long long)ncols * on[2];
  const int grid = (int)std::max(1ll, std::min(total, (long long)(sms - op->reserve_sms)));
  // programmatic dependent launch: the CTAs of this launch may be scheduled while the previous kernel of the stream drains;
  // the kernel itself waits (griddepcontrol.wait) before it touches global memory
  static const bool no_pdl = std::getenv("B200FEM_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(Cfg::kThreads); cfg.dynamicSmemBytes = Cfg::smem_bytes(); cfg.stream = op->sp->mesh->ctx->stream;
  cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, K, b, mc.maps[slot], tx, ncols));
  return B200FEM_OK;
}

// Kronecker kernel of the higher orders (dg_kronecker_slab.cuh): one CTA per TX x TY x TZ tile, n threads per element
template <int N, int TX, int TY, int TZ, int MINB, int SPLIT> static int launch_dg_kronecker_slab(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  using Cfg = KronSlabCfg<N, TX, TY, TZ, SPLIT>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  if (!op->kron_ready) {
    KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
    op->kron_tab.resize(sizeof(KronTabDev<N>));
    KronTabDev<N>& K0 = *reinterpret_cast<KronTabDev<N>*>(op->kron_tab.data());
    for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K0.S[d][i] = kh.S[d][i]; K0.Dlo[d][i] = kh.Dlo[d][i]; K0.Dhi[d][i] = kh.Dhi[d][i]; K0.L[d][i] = kh.L[d][i]; K0.R[d][i] = kh.R[d][i]; }
    op->kron_ready = true;
  }
  const KronTabDev<N>& K = *reinterpret_cast<const KronTabDev<N>*>(op->kron_tab.data());
  const int tx = (b.own_hi[0] - b.own_lo[0] + TX - 1) / TX, ty = (b.own_hi[1] - b.own_lo[1] + TY - 1) / TY, tz = (b.own_hi[2] - b.own_lo[2] + TZ - 1) / TZ;
  auto kern = dg_kronecker_slab_kernel<N, TX, TY, TZ, MINB, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) { CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes())); attr_set = true; }
  static int sms = 0;
  if (!sms) CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, op->sp->mesh->ctx->device));
  const long long ntiles = (long long)tx * ty * tz;
  REQUIRE(ntiles < (1ll << 31), B200FEM_ERR_NOT_IMPLEMENTED, "slab kernel: too many tiles");
  static const char* pw = std::getenv("B200FEM_SLAB_WAVES");     // CTAs per resident slot (0: one CTA per tile, the non-persistent schedule)
  const int waves = pw ? std::atoi(pw) : 1;
  const long long grid = waves > 0 ? std::min<long long>(ntiles, (long long)waves * MINB * sms) : ntiles;
  static long long* d_tl = nullptr; static int tl_calls = 0;
  if (!d_tl && std::getenv("B200FEM_SLAB_TIMELINE")) { CUDA_OK(cudaMalloc(&d_tl, 64)); CUDA_OK(cudaMemset(d_tl, 0, 64)); }
  kern<<<(unsigned)grid, Cfg::kThreads, Cfg::smem_bytes(), op->sp->mesh->ctx->stream>>>(K, b, op->d_perm, u, w, bvec, tx, ty, (int)ntiles, d_tl);
  CUDA_OK(cudaGetLastError());
  if (d_tl && ++tl_calls == 12) {
    long long h[8]; cudaDeviceSynchronize(); cudaMemcpy(h, d_tl, 64, cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[b200fem slab timeline, clocks] staging issued %lld | data landed %lld | phase A %lld | phase B %lld | b rows + barrier + combine %lld | store %lld\n",
                 h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5]);
  }
  return B200FEM_OK;
}

static long long* g_dbg = nullptr; static int g_dbg_calls = 0;
template <int N, bool HIER, bool SPLIT> static int launch_dg_kronecker_pipe(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  constexpr int TX = 8, TY = 4, TZ = 4;
  using Cfg = KronPipeCfg<N, TX, TY, TZ, SPLIT>; const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  KronHost kh = build_kron_tables(op->sp->tab, op->model, b.dim, b.h, mass_scale(op));
  KronTabDev<N> K;
  for (int d = 0; d < 3; ++d) for (int i = 0; i < N * N; ++i) { K.S[d][i] = kh.S[d][i]; K.Dlo[d][i] = kh.Dlo[d][i]; K.Dhi[d][i] = kh.Dhi[d][i]; K.L[d][i] = kh.L[d][i]; K.R[d][i] = kh.R[d][i]; }
  const int tx = (b.own_hi[0] - b.own_lo[0] + TX - 1) / TX, ty = (b.own_hi[1] - b.own_lo[1] + TY - 1) / TY, tz = (b.own_hi[2] - b.own_lo[2] + TZ - 1) / TZ;
  const int ntiles = tx * ty * tz;
  if (!g_dbg && std::getenv("B200FEM_DEBUG_TIMELINE")) { cudaMalloc(&g_dbg, 8 * 8 * 32); cudaMemset(g_dbg, 0, 8 * 8 * 32); }
  static int sms = 0;
  if (!sms) CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, op->sp->mesh->ctx->device));
  auto kern = dg_kronecker_pipe_kernel<N, HIER, TX, TY, TZ, SPLIT>;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes()));
  kern<<<(unsigned)std::min(ntiles, sms), Cfg::kThreads, Cfg::smem_bytes(), op->sp->mesh->ctx->stream>>>(K, b, u, w, bvec, tx, ty, ntiles, std::getenv("B200FEM_DEBUG_SKIP") ? 1 : 0, g_dbg);
  CUDA_OK(cudaGetLastError());
  if (g_dbg && ++g_dbg_calls == 40) {      // dump the timeline of CTA 0 for one warm call
    std::vector<long long> h(8 * 32); cudaDeviceSynchronize(); cudaMemcpy(h.data(), g_dbg, h.size() * 8, cudaMemcpyDeviceToHost);
    long long t0 = h[8 * 1 + 0];
    for (int it = 0; it < 16; ++it) { std::fprintf(stderr, "it %2d:", it); for (int k = 0; k < 8; ++k) std::fprintf(stderr, " %8lld", h[8 * it + k] ? h[8 * it + k] - t0 : -1); std::fprintf(stderr, "\n"); }
  }
  return B200FEM_OK;
}
template <int N> static int launch_lagrange(b200fem_operator* op, const double* u, double* w, bool with_data) {
  const BoxDev& b = op->active_box ? *op->active_box : op->sp->box; cudaStream_t st = op->sp->mesh->ctx->stream;
  CUDA_OK(cudaMemsetAsync(w, 0, sizeof(double) * (size_t)op->sp->size, st));                 // w.clear() (galerkin.hh:1463)
  AdrIntegrands I = make_integrands(op, with_data);
  int launches = 1;
  if (b.dim == 3) {
    using Cfg = DgQuadCfg<N>; auto kern = lagrange3d_quadrature_kernel<N, AdrIntegrands>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes()));
    for (int c = 0; c < 8; ++c) {
      const int c0 = c & 1, c1 = (c >> 1) & 1, c2 = c >> 2;
      const int m0 = (b.n[0] - c0 + 1) / 2, m1 = (b.n[1] - c1 + 1) / 2, m2 = (b.n[2] - c2 + 1) / 2;
      const long long nc = (long long)m0 * m1 * m2; if (nc <= 0) continue;
      kern<<<(unsigned)((nc + Cfg::EB - 1) / Cfg::EB), Cfg::kThreads, Cfg::smem_bytes(), st>>>(make_tab<N>(op->sp->tab), b, I, op->sp->lay, u, w, c0, c1, c2, m0, m1, nc);
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

// Lagrange Kronecker (sum-factorised lattice stencil) kernel, lagrange_kronecker.cuh
static int launch_lagrange_kronecker(b200fem_operator* op, const double* u, double* w, const double* bvec) {
  b200fem_space* s = op->sp; const BoxDev& b = s->box; const int k = s->order, W = 2 * k + 1;
  REQUIRE(s->lay.lattice[0] * s->lay.lattice[1] * s->lay.lattice[2] < (1ll << 31) && s->size < (1ll << 31), B200FEM_ERR_NOT_IMPLEMENTED, "lattice kernel: 32-bit dof cursors");
  if (!op->d_lag_rows) {
    LagRowsHost rh = build_lagrange_rows(s->tab, op->model, b.dim, k, b.n, b.origin, b.gn, b.h);
    size_t total = 0; for (int d = 0; d < 3; ++d) total += 2 * rh.M[d].size();
    std::vector<double> flat; flat.reserve(total); size_t offM[3], offT[3];
    for (int d = 0; d < 3; ++d) { offM[d] = flat.size(); flat.insert(flat.end(), rh.M[d].begin(), rh.M[d].end()); offT[d] = flat.size(); flat.insert(flat.end(), rh.T[d].begin(), rh.T[d].end()); }
    CUDA_OK(cudaMalloc(&op->d_lag_rows, sizeof(double) * flat.size()));
    CUDA_OK(cudaMemcpy(op->d_lag_rows, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice));
    for (int d = 0; d < 3; ++d) { op->lag_rows.M[d] = op->d_lag_rows + offM[d]; op->lag_rows.T[d] = op->d_lag_rows + offT[d]; }
  }
  (void)W;
  const LagrangeLayoutDev& L = s->lay; const bool mapped = L.lattice_map != nullptr;
  static const char* hy_env = std::getenv("B200FEM_LAG_HY");
  const int HY = hy_env ? std::atoi(hy_env) : 16, ctas_per_sm = HY <= 16 ? 2 : 1;
  const int TX = 32 - 2 * k, TY = HY - 2 * k;
  const int tx = (int)((L.lattice[0] + TX - 1) / TX), ty = (int)((L.lattice[1] + TY - 1) / TY);
  // z-segments: every segment re-reads 2k planes and stages its z-rows (<= kMaxSeg planes); the number of segments is chosen
  // so that the grid fills whole waves of the resident CTA slots
  static int sms = 0;
  if (!sms) CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->mesh->ctx->device));
  const int L2 = (int)L.lattice[2], slots = ctas_per_sm * sms, tiles = tx * ty;
  int best_nseg = 1; double best_cost = 1e300;
  for (int ns = 1; ns <= 64; ++ns) {
    const int zs = (L2 + ns - 1) / ns; if (zs > 128) continue;
    const int nse = (L2 + zs - 1) / zs;
    const double waves = std::ceil((double)tiles * nse / slots), cost = waves * (zs + 2 * k + 6);
    if (cost < best_cost) { best_cost = cost; best_nseg = nse; }
    if (zs <= 4) break;
  }
  const int zseg = (L2 + best_nseg - 1) / best_nseg, nseg = (L2 + zseg - 1) / zseg;
  const unsigned grid = (unsigned)(tiles * nseg); cudaStream_t st = s->mesh->ctx->stream;
  const unsigned char* dmask = op->fuse_dirichlet ? op->d_dmask : nullptr; const double* dvals = op->fuse_dirichlet && !op->fuse_linear ? op->d_dvals : nullptr;
  // fused <u, w> partials (requested by the CG driver on one rank, where every dof is primary and the Dirichlet rows are fused too)
  double* dotp = nullptr; op->dot_parts = 0;
  if (op->want_dot && op->fuse_dirichlet == (op->model.strong_dirichlet && op->d_dmask != nullptr) && s->mesh->ctx->world == 1) {
    if ((int)grid > op->dot_cap) { if (op->capturing) return fail(B200FEM_ERR_INVALID, "dot partial buffer must exist before graph capture"); if (op->d_dot_partial) cudaFree(op->d_dot_partial); CUDA_OK(cudaMalloc(&op->d_dot_partial, sizeof(double) * grid)); op->dot_cap = (int)grid; }
    dotp = op->d_dot_partial; op->dot_parts = (int)grid;
  }
#define B200FEM_LAGK(KK, MM, HH) lagrange_kronecker_kernel<KK, MM, HH><<<grid, 32 * HH, 0, st>>>(L, op->lag_rows, u, w, bvec, dmask, dvals, tx, ty, zseg, dotp)
#define B200FEM_LAGK_HY(KK, MM) do { if (HY == 24) B200FEM_LAGK(KK, MM, 24); else B200FEM_LAGK(KK, MM, 16); } while (0)
  if (k == 1) { if (mapped) B200FEM_LAGK_HY(1, true); else B200FEM_LAGK_HY(1, false); }
  else        { if (mapped) B200FEM_LAGK_HY(2, true); else B200FEM_LAGK_HY(2, false); }
#undef B200FEM_LAGK_HY
#undef B200FEM_LAGK
  op->dirichlet_fused = op->fuse_dirichlet;
  CUDA_OK(cudaGetLastError());
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}

static int ensure_bvec(b200fem_operator* op);

// one operator application on device vectors, without halo exchange
static int apply_local(b200fem_operator* op, const double* u, double* w, bool linear) {
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream;
  const long long n = s->size; const int N = s->n1;
  REQUIRE(default_quadrature(op), B200FEM_ERR_NOT_IMPLEMENTED, "only quadrature orders that select the (order+1)-point Gauss rule are implemented on the device");
  if (s->kind == B200FEM_LAGRANGE) {
    REQUIRE(!op->model.has_skeleton, B200FEM_ERR_NOT_IMPLEMENTED, "skeleton integrands on continuous spaces");
    // linear models: Kronecker form (one launch, every node written once); otherwise the generic quadrature kernel with
    // colour-ordered scatter
    const bool lag_kron_ok = op->model.gamma == 0.0;
    int lk = op->kernel_pref;
    if (lk == B200FEM_KERNEL_AUTO) lk = lag_kron_ok ? B200FEM_KERNEL_KRONECKER : B200FEM_KERNEL_QUADRATURE;
    if (lk == B200FEM_KERNEL_KRONECKER) {
      REQUIRE(lag_kron_ok, B200FEM_ERR_INVALID, "Kronecker kernel needs a linear model");
      const double* bvec = nullptr;
      if (!linear && op->model.data) { int rc = ensure_bvec(op); if (rc) return rc; bvec = op->d_bvec; }
      int rc = launch_lagrange_kronecker(op, u, w, bvec); if (rc) return rc;
      op->timing.kernel = B200FEM_KERNEL_KRONECKER;
      return B200FEM_OK;
    }
    int rc = N == 2 ? launch_lagrange<2>(op, u, w, !linear) : launch_lagrange<3>(op, u, w, !linear);
    if (rc) return rc;
    op->timing.kernel = B200FEM_KERNEL_QUADRATURE;
    return B200FEM_OK;
  }
  const bool kron_ok = op->model.gamma == 0.0 && N >= 2 && N <= 6;
  int kernel = op->kernel_pref;
  if (kernel == B200FEM_KERNEL_AUTO) kernel = kron_ok ? B200FEM_KERNEL_KRONECKER : B200FEM_KERNEL_QUADRATURE;
  if (kernel == B200FEM_KERNEL_KRONECKER) {
    REQUIRE(kron_ok, B200FEM_ERR_INVALID, "Kronecker kernel needs a linear model");
    const double* bvec = nullptr;
    if (!linear && op->model.data) { int rc = ensure_bvec(op); if (rc) return rc; bvec = op->d_bvec; }
    // v2 (bulk-copy staged) needs 8-byte aligned vectors whose w / b share the 16-byte phase; otherwise v1
    const char* variant_env = std::getenv("B200FEM_KRON_VARIANT");      // v1 | tma | pipe | split | tensor | tensor3 | march (default), for A/B measurements
    const std::string variant = variant_env ? variant_env : "march";
    const bool phase_ok = !bvec || ((reinterpret_cast<uintptr_t>(bvec) ^ reinterpret_cast<uintptr_t>(w)) & 8) == 0;
    const bool hier = s->kind == B200FEM_DG_LEGENDRE_HIER;
    int rc;
    if (N >= 4) {
      // tile shapes: 4x4x4 (Q3), 4x2x2 (Q4, Q5), two CTAs per SM.  Measured alternatives (4x4x2 with 4 CTAs, 4x4x3 with 3, 2x2x2 with 3
      // for Q5) were within 2 % or slower: the kernel moves 2.5-3.5 x 8 B/dof of halo'd input through the L2->SM fabric and sits
      // at ~75 % of that path's throughput (profiles/r01_dg_kronecker_slab_q3.md)
      // one thread per slab.  SPLIT = 2 (two threads per slab, 12 instead of 6 warps per SM for Q5; the kernel template still
      // carries it) measured 97 vs 101 GDoF/s for Q5 and 99 vs 144 for Q3: the kernel is not short of warps, it waits for its
      // staging phase -- two CTAs per SM is all the 111 KB tiles allow
      if (N == 4) rc = launch_dg_kronecker_slab<4, 4, 4, 4, 2, 1>(op, u, w, bvec);
      else if (N == 5) rc = launch_dg_kronecker_slab<5, 4, 2, 2, 2, 1>(op, u, w, bvec);
      else rc = launch_dg_kronecker_slab<6, 4, 2, 2, 2, 1>(op, u, w, bvec);
      if (rc) return rc;
      op->timing.kernel = kernel; op->timing.launches_per_apply = 1;
      return B200FEM_OK;
    }
    const bool use_march = variant == "march" && N == 3 && !op->fused_seq && march_path_ok<3>(op, u, w, bvec);
    const bool use_tensor = !use_march && (variant == "tensor" || variant == "march") && N == 3 && tensor_path_ok<3>(op, u, w, bvec);
    op->last_launch_tensor = use_tensor;
    if (use_march) rc = hier ? launch_dg_kronecker_march<3, true>(op, u, w, bvec) : launch_dg_kronecker_march<3, false>(op, u, w, bvec);
    else if (use_tensor) rc = hier ? launch_dg_kronecker_tensor<3, true, false>(op, u, w, bvec) : launch_dg_kronecker_tensor<3, false, false>(op, u, w, bvec);
    else if (variant == "tensor3" && tensor_path_ok<3>(op, u, w, bvec) && N == 3) rc = hier ? launch_dg_kronecker_tensor<3, true, true>(op, u, w, bvec) : launch_dg_kronecker_tensor<3, false, true>(op, u, w, bvec);
    else if (N == 3 && phase_ok && (variant == "pipe" || variant == "tensor" || variant == "tensor3" || variant == "march")) rc = hier ? launch_dg_kronecker_pipe<3, true, false>(op, u, w, bvec) : launch_dg_kronecker_pipe<3, false, false>(op, u, w, bvec);
    else if (N == 3 && phase_ok && variant == "split") rc = hier ? launch_dg_kronecker_pipe<3, true, true>(op, u, w, bvec) : launch_dg_kronecker_pipe<3, false, true>(op, u, w, bvec);
    else if (N == 3 && phase_ok && variant == "tma") rc = hier ? launch_dg_kronecker_tma<3, true>(op, u, w, bvec) : launch_dg_kronecker_tma<3, false>(op, u, w, bvec);
    else rc = N == 2 ? launch_dg_kronecker<2, 8, 8, 4>(op, u, w, bvec) : launch_dg_kronecker<3, 8, 4, 4>(op, u, w, bvec);
    if (rc) return rc;
  } else {
    int rc = B200FEM_ERR_NOT_IMPLEMENTED;
    const bool with_data = !linear;
    switch (N) {
      case 2: rc = launch_dg_quadrature<2>(op, u, w, nullptr, with_data); break;
      case 3: rc = launch_dg_quadrature<3>(op, u, w, nullptr, with_data); break;
      case 4: rc = launch_dg_quadrature<4>(op, u, w, nullptr, with_data); break;
      case 5: rc = launch_dg_quadrature<5>(op, u, w, nullptr, with_data); break;
      case 6: rc = launch_dg_quadrature<6>(op, u, w, nullptr, with_data); break;
      default: return fail(B200FEM_ERR_NOT_IMPLEMENTED, "DG order > 5");
    }
    if (rc) return rc;
  }
  (void)n; (void)st;
  op->timing.kernel = kernel; op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}

// b = -L[0], evaluated once by the quadrature kernel with the data terms switched on
static int ensure_bvec(b200fem_operator* op) {
  if (op->d_bvec) return B200FEM_OK;
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  double* zero_u = nullptr; double* bv = nullptr;
  CUDA_OK(cudaMalloc(&zero_u, bytes)); CUDA_OK(cudaMalloc(&bv, bytes));
  CUDA_OK(cudaMemsetAsync(zero_u, 0, bytes, st)); CUDA_OK(cudaMemsetAsync(bv, 0, bytes, st));
  const int saved = op->kernel_pref; op->kernel_pref = B200FEM_KERNEL_QUADRATURE;
  int rc = apply_local(op, zero_u, bv, /*linear=*/false);
  op->kernel_pref = saved;
  if (rc) { cudaFree(zero_u); cudaFree(bv); return rc; }
  negate_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(bv, s->size);
  CUDA_OK(cudaGetLastError()); CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(zero_u));
  op->d_bvec = bv; return B200FEM_OK;
}

static int exchange(b200fem_operator* op, double* v, cudaStream_t st) {
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx;
  int rc = s->kind == B200FEM_LAGRANGE ? halo_exchange(op->halo, c->nccl, c->comm, v, true, st)
           : op->halo_p2p.built        ? halo_exchange_p2p(op->halo_p2p, v, st)
                                       : halo_exchange_dg(op->halo_dg, c->nccl, c->comm, v, st);
  return rc ? fail(B200FEM_ERR_COMM, "halo exchange failed") : B200FEM_OK;
}

// GalerkinOperator::evaluate + w.communicate() (galerkin.hh:1459-1496).  On several ranks the DG apply is split: the owned
// layers next to rank interfaces are computed first, their Copy exchange then runs on a second stream while the interior
// is computed -- the exchange the reference performs serially after the loop is hidden behind the interior elements.
static int reduce_sums(b200fem_operator* op, int count);
static int ensure_cg_buffers(b200fem_operator* op, int maxit);
static int apply_dev_impl(b200fem_operator* op, const double* u, double* w, bool linear);
// AutomaticDifferenceLinearOperator::operator() (automaticdifferenceoperator.hh:124-149), everything on the device and on the
// operator's stream (no host round trip: the difference quotient can sit inside a captured CG graph)
static int apply_fd_jacobian(b200fem_operator* op, const double* arg, double* dest) {
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const long long n = s->size;
  int rc = ensure_cg_buffers(op, 1); if (rc) return rc;
  dot_partial_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(arg, arg, op->d_aux, n, op->d_partial + kRedBlocks);          // arg.normSquaredDofs()
  reduce_final_kernel<<<1, kRedThreads, 0, st>>>(op->d_partial + kRedBlocks, kRedBlocks, op->d_sums + 3);
  b200fem_ctx* c = s->mesh->ctx;
  if (c->world > 1 && c->nccl.AllReduce(op->d_sums + 3, op->d_sums + 3, 1, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, st) != 0) return fail(B200FEM_ERR_COMM, "ncclAllReduce failed");
  fd_eps_kernel<<<1, 32, 0, st>>>(op->d_sums + 3, op->d_fd);
  fd_perturb_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_jac_b, op->d_jac_u, arg, n, op->d_fd);
  const bool want_dot = op->want_dot; op->want_dot = false;   // (a fused <u, w> of the perturbed apply is not <arg, J arg>)
  op->jac_mode = false; rc = apply_dev_impl(op, op->d_jac_b, dest, false); op->jac_mode = true; op->want_dot = want_dot; op->dot_parts = 0; if (rc) return rc;   // (*op_)(b_, dest)
  fd_quotient_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(dest, op->d_jac_opu, n, op->d_fd);
  CUDA_OK(cudaGetLastError()); return B200FEM_OK;
}
static int apply_dev_impl(b200fem_operator* op, const double* u, double* w, bool linear) {
  if (linear && op->jac_mode) return apply_fd_jacobian(op, u, w);
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream;
  const bool timing_events = op->timing_enabled && !op->capturing;
  if (timing_events) CUDA_OK(cudaEventRecord(op->ev0, st));
  const bool distributed = op->communicate && s->mesh->ctx->world > 1;
  int rc;
  // Measured on 2 x B200 (profiles/r01_multigpu.md): exchange kernels queued on a second stream do not start before the
  // persistent compute kernel retires, even with SMs left free, so splitting only adds 12 us of small launches; the
  // peer-memory exchange itself takes 3-4 us.  Default: one compute launch, then the exchange on the same stream.  The
  // split/overlap schedule stays selectable for experiments.
  static const bool split_overlap = std::getenv("B200FEM_SPLIT_OVERLAP") != nullptr;
  if (split_overlap && distributed && s->kind != B200FEM_LAGRANGE && op->comm_stream) {
    // boundary sub-boxes: per split axis, one slab at each interface (thickness = one tile layer), carved successively
    // out of the owned box so that the pieces do not overlap
    const BoxDev& full = s->box; BoxDev rest = full; std::vector<BoxDev> bnd;
    const int thick[3] = {8, 4, 4};
    for (int a = 2; a >= 0; --a) {
      const bool lo_if = full.own_lo[a] > 0, hi_if = full.own_hi[a] < full.n[a];
      if (lo_if && rest.own_hi[a] - rest.own_lo[a] > 0) { BoxDev b = rest; b.own_hi[a] = std::min(rest.own_hi[a], rest.own_lo[a] + thick[a]); bnd.push_back(b); rest.own_lo[a] = b.own_hi[a]; }
      if (hi_if && rest.own_hi[a] - rest.own_lo[a] > 0) { BoxDev b = rest; b.own_lo[a] = std::max(rest.own_lo[a], rest.own_hi[a] - thick[a]); bnd.push_back(b); rest.own_hi[a] = b.own_lo[a]; }
    }
    int launches = 0;
    for (const BoxDev& b : bnd) { op->active_box = &b; rc = apply_local(op, u, w, linear); op->active_box = nullptr; if (rc) return rc; launches += op->timing.launches_per_apply; }
    CUDA_OK(cudaEventRecord(op->ev_bnd, st));
    if (op->dbg_ev[0]) CUDA_OK(cudaEventRecord(op->dbg_ev[0], st));
    CUDA_OK(cudaStreamWaitEvent(op->comm_stream, op->ev_bnd, 0));
    CUDA_OK(cudaEventRecord(op->evx0, op->comm_stream));
    rc = exchange(op, w, op->comm_stream); if (rc) return rc;
    CUDA_OK(cudaEventRecord(op->evx1, op->comm_stream));
    CUDA_OK(cudaEventRecord(op->ev_comm, op->comm_stream));
    bool has_rest = true; for (int a = 0; a < 3; ++a) has_rest = has_rest && rest.own_hi[a] > rest.own_lo[a];
    if (has_rest) { op->active_box = &rest; op->reserve_sms = 8; rc = apply_local(op, u, w, linear); op->active_box = nullptr; op->reserve_sms = 0; if (rc) return rc; launches += op->timing.launches_per_apply; }
    if (op->dbg_ev[1]) CUDA_OK(cudaEventRecord(op->dbg_ev[1], st));
    CUDA_OK(cudaStreamWaitEvent(st, op->ev_comm, 0));
    op->timing.launches_per_apply = launches + 3 * (int)op->halo_dg.nb.size();
  } else {
    // fused send: when the peer-memory mailboxes exist the TMA kernel itself stores boundary rows of w into the
    // neighbours' mailboxes while it is still computing; afterwards only the receive part runs
    // measured (profiles/r01_multigpu.md): marching kernel + one exchange kernel beats the tensor kernel with fused sends
    // (53.5 vs 56.0 us at 2 GPUs, 61.2 vs 73.6 us at 4) -- the fused schedule is opt-in
    static const bool no_fused = std::getenv("B200FEM_FUSED_SEND") == nullptr || std::getenv("B200FEM_NO_FUSED_SEND") != nullptr;
    const bool try_fused = distributed && !no_fused && s->kind != B200FEM_LAGRANGE && op->halo_p2p.built && op->halo_p2p.nnb > 0 &&
                           s->box.own_lo[0] == 0 && s->box.own_hi[0] == s->box.n[0];
    op->last_launch_tensor = false;
    if (try_fused) op->fused_seq = op->halo_p2p.seq + 1;
    // single rank: the Dirichlet wrapper can ride along in the store of the Lagrange Kronecker kernel (with several ranks it
    // has to follow the Add exchange)
    op->dirichlet_fused = false;
    op->fuse_dirichlet = !distributed && op->model.strong_dirichlet && op->d_dmask != nullptr; op->fuse_linear = linear;
    rc = apply_local(op, u, w, linear);
    op->fuse_dirichlet = false;
    const bool fused = try_fused && op->last_launch_tensor;
    op->fused_seq = 0;
    if (rc) return rc;
    if (distributed) {
      if (timing_events) CUDA_OK(cudaEventRecord(op->evx0, st));
      if (fused) { const unsigned long long seq = ++op->halo_p2p.seq; if (halo_exchange_p2p_fused_tail(op->halo_p2p, w, seq, st) != 0) return fail(B200FEM_ERR_COMM, "halo exchange failed"); op->timing.launches_per_apply += 1; }
      else { rc = exchange(op, w, st); if (rc) return rc; }
      if (timing_events) CUDA_OK(cudaEventRecord(op->evx1, st));
    }
  }
  // DirichletWrapperOperator: op_(u,w) (communication included) first, then subConstraints (dirichletwrapper.hh:101-105)
  if (op->model.strong_dirichlet && op->d_dmask && !op->dirichlet_fused) {
    dirichlet_sub_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(u, w, op->d_dmask, linear ? nullptr : op->d_dvals, s->size);
    CUDA_OK(cudaGetLastError()); op->timing.launches_per_apply += 1;
  }
  if (timing_events) CUDA_OK(cudaEventRecord(op->ev1, st));
  op->timing.applies += 1;
  if (op->dbg_ev[0] && op->timing.applies == 300) {
    CUDA_OK(cudaEventSynchronize(op->ev1)); float a, b, c, d, e;
    cudaEventElapsedTime(&a, op->ev0, op->dbg_ev[0]); cudaEventElapsedTime(&b, op->ev0, op->evx0); cudaEventElapsedTime(&c, op->ev0, op->evx1);
    cudaEventElapsedTime(&d, op->ev0, op->dbg_ev[1]); cudaEventElapsedTime(&e, op->ev0, op->ev1);
    std::fprintf(stderr, "[b200fem timeline us] boundary done %.1f | exchange start %.1f end %.1f | interior done %.1f | apply done %.1f\n", a * 1e3, b * 1e3, c * 1e3, d * 1e3, e * 1e3);
  }
  return B200FEM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
static void mark_dirichlet(b200fem_operator* op) {
  // DirichletConstraints::updateDirichletDofs (schemes/dirichletconstraints.hh:435-554): all Lagrange nodes on boundary
  // faces whose side is flagged; values g(x_node).  Host-side, closed form over the boundary lattice.
  b200fem_space* s = op->sp; const BoxDev& b = s->box; const int dim = b.dim, k = s->order;
  op->h_dmask.assign((size_t)s->size, 0); op->h_dvals.assign((size_t)s->size, 0.0);
  LagrangeLayoutDev L = s->lay; L.lattice_map = s->lattice_map.empty() ? nullptr : s->lattice_map.data();
  const long long L0 = L.lattice[0], L1 = L.lattice[1], L2 = L.lattice[2];
  for (long long g2 = 0; g2 < L2; ++g2) for (long long g1 = 0; g1 < L1; ++g1) for (long long g0 = 0; g0 < L0; ++g0) {
    const long long g[3] = {g0, g1, g2}; bool on = false;
    for (int d = 0; d < dim; ++d) {
      const long long gg = (long long)k * b.origin[d] + g[d];            // global lattice coordinate
      if (gg == 0 && ((op->model.dirichlet_mask >> (2 * d)) & 1)) on = true;
      if (gg == (long long)k * b.gn[d] && ((op->model.dirichlet_mask >> (2 * d + 1)) & 1)) on = true;
    }
    if (!on) continue;
    double x[3] = {0, 0, 0}; for (int d = 0; d < dim; ++d) x[d] = b.lo[d] + b.h[d] * (b.origin[d] + (double)g[d] / k);
    double val = 0;
    if (op->model.data == 1) val = std::sin(x[0] * x[1]);
    else if (op->model.data == 2) { val = 1; for (int d = 0; d < dim; ++d) val *= std::sin(M_PI * x[d]); }
    const long long dof = lagrange_dof(L, g0, g1, g2);
    op->h_dmask[(size_t)dof] = 1; op->h_dvals[(size_t)dof] = val;
  }
}

extern "C" int b200fem_operator_create(b200fem_space* s, const b200fem_model* model, b200fem_operator** out) {
  REQUIRE(s && model && out, B200FEM_ERR_INVALID, "operator_create: null argument");
  REQUIRE(!(model->strong_dirichlet && s->kind != B200FEM_LAGRANGE), B200FEM_ERR_INVALID, "strong Dirichlet constraints need a Lagrange space");
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  auto* op = new b200fem_operator; op->sp = s; op->model = *model;
  if (s->kind != B200FEM_LAGRANGE) {
    CUDA_OK(cudaMalloc(&op->d_perm, sizeof(int) * s->perm.size()));
    CUDA_OK(cudaMemcpy(op->d_perm, s->perm.data(), sizeof(int) * s->perm.size(), cudaMemcpyHostToDevice));
  }
  if (model->strong_dirichlet) {
    mark_dirichlet(op);
    CUDA_OK(cudaMalloc(&op->d_dmask, (size_t)s->size)); CUDA_OK(cudaMalloc(&op->d_dvals, sizeof(double) * (size_t)s->size));
    CUDA_OK(cudaMemcpy(op->d_dmask, op->h_dmask.data(), (size_t)s->size, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(op->d_dvals, op->h_dvals.data(), sizeof(double) * (size_t)s->size, cudaMemcpyHostToDevice));
  }
  CUDA_OK(cudaEventCreate(&op->ev0)); CUDA_OK(cudaEventCreate(&op->ev1)); CUDA_OK(cudaEventCreate(&op->evx0)); CUDA_OK(cudaEventCreate(&op->evx1));
  if (s->mesh->ctx->world > 1) {
    int rc = halo_plan_build(op->halo, s->mesh->proc, s->mesh->pc, s->box, s->kind == B200FEM_LAGRANGE, s->kind == B200FEM_LAGRANGE ? s->order : 0, s->nb, s->lay, s->size, &op->d_aux);
    if (!rc && s->kind != B200FEM_LAGRANGE) rc = halo_plan_dg_build(op->halo_dg, s->mesh->proc, s->mesh->pc, s->box, s->nb);
    if (rc) { delete op; return fail(B200FEM_ERR_COMM, "halo plan failed"); }
    // peer-memory mailboxes (all ranks must take the same decision: it depends only on the environment and on CUDA IPC
    // working on this box); falls back to NCCL send/recv
    if (s->kind != B200FEM_LAGRANGE && !std::getenv("B200FEM_NO_P2P")) {
      b200fem_ctx* c = s->mesh->ctx;
      if (halo_plan_p2p_build(op->halo_p2p, op->halo_dg, c->nccl, c->comm, c->rank, c->world, s->mesh->proc, s->mesh->pc, s->mesh->gn, s->nb, c->stream) != 0) {
        halo_plan_p2p_free(op->halo_p2p); cudaGetLastError();
      }
      // agree on the outcome: one failing rank switches everybody to NCCL
      int ok = op->halo_p2p.built ? 1 : 0; int* d_ok = nullptr; CUDA_OK(cudaMalloc(&d_ok, sizeof(int)));
      CUDA_OK(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
      if (c->nccl.AllReduce(d_ok, d_ok, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, c->comm, c->stream) != 0) { delete op; return fail(B200FEM_ERR_COMM, "ncclAllReduce failed"); }
      CUDA_OK(cudaStreamSynchronize(c->stream)); CUDA_OK(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost)); cudaFree(d_ok);
      if (!ok) halo_plan_p2p_free(op->halo_p2p);
    }
    { int lo_p = 0, hi_p = 0; CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p)); CUDA_OK(cudaStreamCreateWithPriority(&op->comm_stream, cudaStreamNonBlocking, hi_p)); }
    CUDA_OK(cudaEventCreateWithFlags(&op->ev_bnd, cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&op->ev_comm, cudaEventDisableTiming));
    if (std::getenv("B200FEM_DEBUG_EVENTS")) { CUDA_OK(cudaEventCreate(&op->dbg_ev[0])); CUDA_OK(cudaEventCreate(&op->dbg_ev[1])); op->timing_enabled = true; }
  }
  *out = op; return B200FEM_OK;
}
static void free_map_cache(b200fem_operator* op) { delete op->map_cache; op->map_cache = nullptr; delete op->march_cache; op->march_cache = nullptr; }
extern "C" int b200fem_operator_destroy(b200fem_operator* op) {
  if (!op) return B200FEM_OK;
  for (void* p : {(void*)op->d_perm, (void*)op->d_bvec, (void*)op->d_dmask, (void*)op->d_dvals, (void*)op->d_aux, (void*)op->d_u, (void*)op->d_w, (void*)op->d_h, (void*)op->d_r,
                  (void*)op->d_p, (void*)op->d_x, (void*)op->d_b, (void*)op->d_partial, (void*)op->d_sums, (void*)op->d_hist, (void*)op->d_cg, (void*)op->d_lag_rows, (void*)op->d_counter, (void*)op->d_rstar, (void*)op->d_s, (void*)op->d_tmp, (void*)op->d_partial5, (void*)op->d_sums5, (void*)op->d_bicg, (void*)op->d_dot_partial}) if (p) cudaFree(p);
  if (op->cg_graph) cudaGraphExecDestroy(op->cg_graph);
  for (double* q : op->gmres_v) if (q) cudaFree(q);
  for (void* q : {(void*)op->d_dinv, (void*)op->d_pq, (void*)op->d_ps}) if (q) cudaFree(q);
  for (void* q : {(void*)op->d_jac_u, (void*)op->d_jac_opu, (void*)op->d_jac_b, (void*)op->d_fd}) if (q) cudaFree(q);
  if (op->d_gm_partial) cudaFree(op->d_gm_partial);
  if (op->d_gm_sums) cudaFree(op->d_gm_sums);
  halo_plan_p2p_free(op->halo_p2p); halo_plan_free(op->halo); halo_plan_dg_free(op->halo_dg); free_map_cache(op);
  if (op->comm_stream) cudaStreamDestroy(op->comm_stream);
  if (op->h2d_stream) cudaStreamDestroy(op->h2d_stream);
  if (op->d2h_stream) cudaStreamDestroy(op->d2h_stream);
  for (cudaEvent_t e : op->pipe_ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {op->ev_bnd, op->ev_comm}) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {op->ev0, op->ev1, op->evx0, op->evx1}) if (e) cudaEventDestroy(e);
  delete op; return B200FEM_OK;
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
static int apply_host(b200fem_operator* op, const double* u, double* w, bool linear) {
  REQUIRE(op && u && w, B200FEM_ERR_INVALID, "apply: null argument");
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  int rc = ensure_staging(op); if (rc) return rc;
  const bool no_pipeline = std::getenv("B200FEM_NO_PIPELINE") != nullptr;     // (read per call: tests toggle it)
  if (!no_pipeline && !(linear && op->jac_mode) && s->kind != B200FEM_LAGRANGE && s->mesh->ctx->world == 1 && s->box.dim == 3 && bytes >= (8u << 20) && s->box.n[2] >= 16 && default_quadrature(op))
  {
    static const char* ch = std::getenv("B200FEM_PIPE_CHUNKS");
    const int want = ch ? std::max(2, std::min(16, std::atoi(ch))) : 8;
    return apply_host_pipelined(op, u, w, linear, std::min(want, s->box.n[2] / 4));
  }
  CUDA_OK(cudaMemcpyAsync(op->d_u, u, bytes, cudaMemcpyHostToDevice, st));
  rc = apply_dev_impl(op, op->d_u, op->d_w, linear); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(w, op->d_w, bytes, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return B200FEM_OK;
}
extern "C" int b200fem_operator_apply(b200fem_operator* op, const double* u, double* w) { return apply_host(op, u, w, false); }
extern "C" int b200fem_operator_apply_linear(b200fem_operator* op, const double* u, double* w) { return apply_host(op, u, w, true); }
extern "C" int b200fem_operator_apply_dev(b200fem_operator* op, const double* u, double* w, int linear) {
  REQUIRE(op && u && w, B200FEM_ERR_INVALID, "apply_dev: null argument");
  int cur = -1; cudaGetDevice(&cur);
  if (cur != op->sp->mesh->ctx->device) CUDA_OK(cudaSetDevice(op->sp->mesh->ctx->device));
  return apply_dev_impl(op, u, w, linear != 0);
}
extern "C" int b200fem_operator_load_vector(b200fem_operator* op, double* b_host) {
  REQUIRE(op && b_host, B200FEM_ERR_INVALID, "load_vector: null argument");
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  int rc = ensure_staging(op); if (rc) return rc;
  CUDA_OK(cudaMemsetAsync(op->d_u, 0, bytes, st));
  const int saved = op->kernel_pref; op->kernel_pref = B200FEM_KERNEL_QUADRATURE;
  rc = apply_dev_impl(op, op->d_u, op->d_w, false); op->kernel_pref = saved; if (rc) return rc;
  negate_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_w, s->size);
  CUDA_OK(cudaMemcpyAsync(b_host, op->d_w, bytes, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  return B200FEM_OK;
}
extern "C" int b200fem_operator_set_communicate(b200fem_operator* op, int c) { REQUIRE(op, B200FEM_ERR_INVALID, "null"); op->communicate = c != 0; return B200FEM_OK; }
extern "C" int b200fem_operator_set_quadrature_orders(b200fem_operator* op, unsigned qi, unsigned qs) { REQUIRE(op, B200FEM_ERR_INVALID, "null"); op->q_interior = qi; op->q_surface = qs; return B200FEM_OK; }
extern "C" int b200fem_operator_set_kernel(b200fem_operator* op, int k) { REQUIRE(op && k >= 0 && k <= 2, B200FEM_ERR_INVALID, "set_kernel: bad kernel id"); op->kernel_pref = k; return B200FEM_OK; }
extern "C" int b200fem_operator_set_inverse_mass(b200fem_operator* op, int on) {
  REQUIRE(op, B200FEM_ERR_INVALID, "null");
  REQUIRE(op->sp->kind != B200FEM_LAGRANGE, B200FEM_ERR_NOT_IMPLEMENTED, "inverse mass (MOLGalerkinOperator): DG spaces only");
  if (op->inverse_mass == (on != 0)) return B200FEM_OK;
  op->inverse_mass = on != 0;
  // the scaled 1-D operators and the scaled load vector are rebuilt on the next apply
  op->kron_ready = false;
  if (op->d_bvec) { CUDA_OK(cudaSetDevice(op->sp->mesh->ctx->device)); CUDA_OK(cudaStreamSynchronize(op->sp->mesh->ctx->stream)); CUDA_OK(cudaFree(op->d_bvec)); op->d_bvec = nullptr; }
  if (op->cg_graph) { cudaGraphExecDestroy(op->cg_graph); op->cg_graph = nullptr; }
  return B200FEM_OK;
}
extern "C" int b200fem_operator_linearize_dev(b200fem_operator* op, const double* u, double eps) {
  REQUIRE(op, B200FEM_ERR_INVALID, "linearize: null operator");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const long long n = s->size; const size_t bytes = sizeof(double) * (size_t)n;
  CUDA_OK(cudaSetDevice(c->device));
  if (op->cg_graph) { cudaGraphExecDestroy(op->cg_graph); op->cg_graph = nullptr; }        // the captured iteration applied another operator
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
extern "C" int b200fem_operator_dirichlet(b200fem_operator* op, uint8_t* mask, double* values) {
  REQUIRE(op && mask && values, B200FEM_ERR_INVALID, "null");
  if (op->h_dmask.empty()) { std::fill(mask, mask + op->sp->size, 0); return B200FEM_OK; }
  std::copy(op->h_dmask.begin(), op->h_dmask.end(), mask); std::copy(op->h_dvals.begin(), op->h_dvals.end(), values); return B200FEM_OK;
}
extern "C" int b200fem_operator_timing(b200fem_operator* op, b200fem_timing* out) {
  REQUIRE(op && out, B200FEM_ERR_INVALID, "null");
  if (!op->timing_enabled) { op->timing_enabled = true; }
  else if (op->timing.applies > 0) {
    CUDA_OK(cudaEventSynchronize(op->ev1)); float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, op->ev0, op->ev1)); op->timing.last_apply_ms = ms;
    if (op->communicate && op->sp->mesh->ctx->world > 1) { CUDA_OK(cudaEventElapsedTime(&ms, op->evx0, op->evx1)); op->timing.last_exchange_ms = ms; }
  }
  *out = op->timing; return B200FEM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// dot over primary dofs + global sum (function/common/scalarproducts.hh:115-127): two-stage device reduction,
// then ncclAllReduce on the scalar when there is more than one rank
static int reduce_sums(b200fem_operator* op, int count) {
  b200fem_ctx* c = op->sp->mesh->ctx;
  for (int i = 0; i < count; ++i) reduce_final_kernel<<<1, kRedThreads, 0, c->stream>>>(op->d_partial + (size_t)i * kRedBlocks, kRedBlocks, op->d_sums + i);
  CUDA_OK(cudaGetLastError());
  if (c->world > 1) { if (c->nccl.AllReduce(op->d_sums, op->d_sums, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, c->stream) != 0) return fail(B200FEM_ERR_COMM, "ncclAllReduce failed"); }
  return B200FEM_OK;
}
static int ensure_cg_buffers(b200fem_operator* op, int maxit) {
  const size_t bytes = sizeof(double) * (size_t)op->sp->size;
  if (!op->d_h) { CUDA_OK(cudaMalloc(&op->d_h, bytes)); CUDA_OK(cudaMalloc(&op->d_r, bytes)); CUDA_OK(cudaMalloc(&op->d_p, bytes)); }
  if (!op->d_partial) { CUDA_OK(cudaMalloc(&op->d_partial, sizeof(double) * 2 * kRedBlocks)); CUDA_OK(cudaMalloc(&op->d_sums, sizeof(double) * 4)); CUDA_OK(cudaMalloc(&op->d_cg, sizeof(CgState))); CUDA_OK(cudaMalloc(&op->d_counter, 2 * sizeof(unsigned int))); CUDA_OK(cudaMemset(op->d_counter, 0, 2 * sizeof(unsigned int))); }
  if (maxit > op->hist_cap) { if (op->d_hist) cudaFree(op->d_hist); CUDA_OK(cudaMalloc(&op->d_hist, sizeof(double) * (size_t)std::max(maxit, 1))); op->hist_cap = std::max(maxit, 1); }
  return B200FEM_OK;
}
extern "C" int b200fem_dot_dev(b200fem_operator* op, const double* x, const double* y, double* result) {
  REQUIRE(op && x && y && result, B200FEM_ERR_INVALID, "dot: null argument");
  b200fem_ctx* c = op->sp->mesh->ctx; CUDA_OK(cudaSetDevice(c->device));
  int rc = ensure_cg_buffers(op, 1); if (rc) return rc;
  dot_partial_kernel<<<kRedBlocks, kRedThreads, 0, c->stream>>>(x, y, op->d_aux, op->sp->size, op->d_partial);
  rc = reduce_sums(op, 1); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(result, op->d_sums, sizeof(double), cudaMemcpyDeviceToHost, c->stream)); CUDA_OK(cudaStreamSynchronize(c->stream));
  return B200FEM_OK;
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
  // one CG iteration, enqueued on the stream.  Single rank: 4 launches (the last block of a reduction kernel finishes the
  // reduction and updates the scalars); several ranks: the partial sums go through ncclAllReduce between two kernels.
  const bool single = c->world == 1;
  auto enqueue_iteration = [&]() -> int {
    cg_update_p_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_p, op->d_r, n, op->d_cg);                      // no-op in iteration 0
    op->want_dot = single; op->dot_parts = 0;
    int e = apply_dev_impl(op, op->d_p, op->d_h, true); op->want_dot = false; if (e) return e;                  // h = A q (+ <q,h> partials when the kernel can)
    if (single) {
      if (op->dot_parts > 0) cg_alpha_partials_kernel<<<1, kRedThreads, 0, st>>>(op->d_dot_partial, op->dot_parts, op->d_cg);
      else cg_dot_alpha_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(op->d_p, op->d_h, op->d_aux, n, op->d_partial, op->d_cg, op->d_counter);
      cg_update_xr_residual_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, op->d_r, op->d_p, op->d_h, op->d_aux, n, op->d_partial, op->d_cg, op->d_hist, op->d_counter + 1);
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
  // Single rank: a chunk of 16 iterations is captured once into a CUDA graph and replayed (launch-bound sizes such as the
  // 256^2 P1 grid spend their time in launch gaps otherwise).  Iterations past convergence / max_iterations are no-ops
  // on the device (every kernel checks the device-resident `done` flag), so whole chunks can always be replayed.
  static const bool no_graph = std::getenv("B200FEM_NO_CG_GRAPH") != nullptr;
  bool use_graph = single && !no_graph && maxit >= chunk;
  if (std::getenv("B200FEM_NO_COOP_CG") == nullptr && single && !op->jac_mode && s->kind == B200FEM_LAGRANGE && s->box.dim == 2 && op->model.gamma == 0.0 && n <= (1 << 20)) use_graph = false;   // cooperative path below
  if (use_graph && !(op->cg_graph && op->cg_graph_key[0] == (const void*)x && op->cg_graph_key[1] == (const void*)b && op->cg_graph_key[2] == (const void*)op->d_hist)) {
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
    op->cg_graph_key[0] = x; op->cg_graph_key[1] = b; op->cg_graph_key[2] = op->d_hist;
  }
  // Launch-bound sizes on a 2-D Lagrange lattice (BASELINE config 1): a chunk of iterations is ONE cooperative launch with
  // grid-wide barriers instead of kernel boundaries (cg_coop2d.cuh).  B200FEM_NO_COOP_CG disables it.
  static const bool no_coop = std::getenv("B200FEM_NO_COOP_CG") != nullptr;
  int coop_grid = 0;
  const bool use_coop = single && !no_coop && !op->jac_mode && s->kind == B200FEM_LAGRANGE && s->box.dim == 2 && op->model.gamma == 0.0 && !op->model.has_skeleton &&
                        default_quadrature(op) && n <= (1 << 20) && op->d_lag_rows != nullptr && (!op->model.strong_dirichlet || op->d_dmask);
  if (use_coop) {
    int per_sm = 0, sms = 0, coop_ok = 0;
    CUDA_OK(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, c->device));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if (s->order == 1) CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_coop2d_kernel<1>, kCoopThreads, 0));
    else CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_coop2d_kernel<2>, kCoopThreads, 0));
    const long long nodes = s->lay.lattice[0] * s->lay.lattice[1];
    coop_grid = coop_ok ? (int)std::min<long long>(std::min<long long>((long long)per_sm * sms, kRedBlocks), (nodes + kCoopThreads - 1) / kCoopThreads) : 0;
  }
  for (int it = 0; it < maxit;) {
    const int upto = std::min(maxit, it + chunk);
    if (coop_grid > 0) {
      int iters = chunk; const unsigned char* dm = op->model.strong_dirichlet ? op->d_dmask : nullptr; double* xx = x;
      void* args[] = {(void*)&s->lay, (void*)&op->lag_rows, (void*)&xx, (void*)&op->d_r, (void*)&op->d_p, (void*)&op->d_h, (void*)&dm, (void*)&op->d_partial,
                      (void*)&op->d_cg, (void*)&op->d_hist, (void*)&iters};
      if (s->order == 1) CUDA_OK(cudaLaunchCooperativeKernel((const void*)cg_coop2d_kernel<1>, dim3((unsigned)coop_grid), dim3(kCoopThreads), args, 0, st));
      else CUDA_OK(cudaLaunchCooperativeKernel((const void*)cg_coop2d_kernel<2>, dim3((unsigned)coop_grid), dim3(kCoopThreads), args, 0, st));
      it += chunk;
    }
    else if (use_graph) { CUDA_OK(cudaGraphLaunch(op->cg_graph, st)); it += chunk; }
    else for (; it < upto; ++it) { rc = enqueue_iteration(); if (rc) return rc; }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(&host, op->d_cg, sizeof(CgState), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
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

// diag(A) of the Kronecker form, on the host (setup cost O(N), once per operator)
static int host_diagonal(b200fem_operator* op, std::vector<double>& diag, bool dirichlet_rows = true) {
  b200fem_space* s = op->sp; const BoxDev& b = s->box;
  REQUIRE(op->model.gamma == 0.0 && default_quadrature(op), B200FEM_ERR_NOT_IMPLEMENTED, "diagonal: needs a linear model with the default quadrature (Kronecker form)");
  diag.assign((size_t)s->size, 0.0);
  if (s->kind == B200FEM_LAGRANGE) {
    REQUIRE(!op->model.has_skeleton, B200FEM_ERR_NOT_IMPLEMENTED, "skeleton integrands on continuous spaces");
    const int k = s->order;
    LagRowsHost rh = build_lagrange_rows(s->tab, op->model, b.dim, k, b.n, b.origin, b.gn, b.h);
    const int W = 2 * k + 1;
    LagrangeLayoutDev L = s->lay; L.lattice_map = s->lattice_map.empty() ? nullptr : s->lattice_map.data();
    for (long long g2 = 0; g2 < L.lattice[2]; ++g2) for (long long g1 = 0; g1 < L.lattice[1]; ++g1) for (long long g0 = 0; g0 < L.lattice[0]; ++g0) {
      const double m0 = rh.M[0][(size_t)g0 * W + k], m1 = rh.M[1][(size_t)g1 * W + k], m2 = rh.M[2][(size_t)g2 * W + k];
      const double t0 = rh.T[0][(size_t)g0 * W + k], t1 = rh.T[1][(size_t)g1 * W + k], t2 = rh.T[2][(size_t)g2 * W + k];
      const long long dof = lagrange_dof(L, g0, g1, g2);
      diag[(size_t)dof] = t0 * m1 * m2 + m0 * t1 * m2 + m0 * m1 * t2;
      if (dirichlet_rows && !op->h_dmask.empty() && op->h_dmask[(size_t)dof]) diag[(size_t)dof] = 1.0;      // DirichletWrapperOperator: identity rows
    }
    return B200FEM_OK;
  }
  const int N = s->n1, nb = s->nb;
  KronHost kh = build_kron_tables(s->tab, op->model, b.dim, b.h, mass_scale(op));
  for (long long e2 = 0; e2 < b.n[2]; ++e2) for (long long e1 = 0; e1 < b.n[1]; ++e1) for (long long e0 = 0; e0 < b.n[0]; ++e0) {
    const long long ec[3] = {e0, e1, e2}; const long long e = e0 + b.n[0] * (e1 + b.n[1] * e2);
    for (int t = 0; t < nb; ++t) {
      const int idx[3] = {t / (N * N), (t / N) % N, t % N}; double v = 0;
      for (int d = 0; d < 3; ++d) {
        const int ii = idx[d] * N + idx[d]; const long long g = b.origin[d] + ec[d];
        v += kh.S[d][ii] + (g == 0 ? kh.Dlo[d][ii] : 0.0) + (g == b.gn[d] - 1 ? kh.Dhi[d][ii] : 0.0);
      }
      diag[(size_t)(e * nb + s->perm[t])] = v;
    }
  }
  return B200FEM_OK;
}
extern "C" int b200fem_operator_diagonal(b200fem_operator* op, double* diag_host) {
  REQUIRE(op && diag_host, B200FEM_ERR_INVALID, "diagonal: null argument");
  std::vector<double> d; int rc = host_diagonal(op, d); if (rc) return rc;
  std::copy(d.begin(), d.end(), diag_host); return B200FEM_OK;
}
// LinearSolver::cg, preconditioned branch (solver/linear/cg.hh:52-56, 72-107) with the Jacobi preconditioner
extern "C" int b200fem_pcg_solve_dev(b200fem_operator* op, const double* b, double* x, double epsilon, int maxit, int tolcrit, int* iterations, double* history) {
  REQUIRE(op && b && x && iterations, B200FEM_ERR_INVALID, "pcg: null argument");
  REQUIRE(tolcrit >= 0 && tolcrit <= 2, B200FEM_ERR_INVALID, "pcg: unknown tolerance criterion");
  REQUIRE(!op->jac_mode, B200FEM_ERR_NOT_IMPLEMENTED, "pcg: no diagonal for a difference-quotient linearisation");
  b200fem_space* s = op->sp; b200fem_ctx* c = s->mesh->ctx; cudaStream_t st = c->stream; const long long n = s->size; const size_t bytes = sizeof(double) * (size_t)n;
  CUDA_OK(cudaSetDevice(c->device));
  int rc = ensure_cg_buffers(op, maxit); if (rc) return rc;
  if (!op->d_dinv || op->dinv_mass != op->inverse_mass) {
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
    CUDA_OK(cudaStreamSynchronize(st)); op->dinv_mass = op->inverse_mass;
  }
  double* p = op->d_p; double* q = op->d_pq; double* sv = op->d_ps; double* h = op->d_h;
  CgState init{}; init.epsilon = epsilon; init.max_iterations = maxit; init.tol_criteria = tolcrit;
  CUDA_OK(cudaMemcpyAsync(op->d_cg, &init, sizeof(CgState), cudaMemcpyHostToDevice, st));
  rc = apply_dev_impl(op, x, h, true); if (rc) return rc;                                                       // h = A x
  pcg_init_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(h, b, op->d_dinv, p, q, sv, op->d_aux, n, op->d_partial, op->d_partial + kRedBlocks);
  rc = reduce_sums(op, 2); if (rc) return rc;
  cg_init_final_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg);
  const bool single = c->world == 1;
  CgState host{}; const int chunk = 16;
  for (int it = 0; it < maxit;) {
    const int upto = std::min(maxit, it + chunk);
    for (; it < upto; ++it) {
      pcg_update_q_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(q, sv, n, op->d_cg);
      rc = apply_dev_impl(op, q, h, true); if (rc) return rc;                                                   // h = A q
      if (single) {
        cg_dot_alpha_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(q, h, op->d_aux, n, op->d_partial, op->d_cg, op->d_counter);
        pcg_update_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, p, sv, q, h, op->d_dinv, op->d_aux, n, op->d_partial, op->d_cg, op->d_hist, op->d_counter + 1);
      } else {
        cg_dot_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(q, h, op->d_aux, n, op->d_partial, op->d_cg);
        rc = reduce_sums(op, 1); if (rc) return rc;
        cg_alpha_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg);
        pcg_update_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, p, sv, q, h, op->d_dinv, op->d_aux, n, op->d_partial, op->d_cg, nullptr, nullptr);
        rc = reduce_sums(op, 1); if (rc) return rc;
        cg_residual_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_cg, op->d_hist);
      }
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(&host, op->d_cg, sizeof(CgState), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
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
  const size_t bytes = sizeof(double) * (size_t)n;
  while ((int)op->gmres_v.size() < m + 1) { double* q = nullptr; CUDA_OK(cudaMalloc(&q, bytes)); op->gmres_v.push_back(q); }
  if (op->gm_cap < m + 2) {
    if (op->d_gm_partial) cudaFree(op->d_gm_partial); if (op->d_gm_sums) cudaFree(op->d_gm_sums);
    CUDA_OK(cudaMalloc(&op->d_gm_partial, sizeof(double) * (size_t)(m + 2) * kRedBlocks)); CUDA_OK(cudaMalloc(&op->d_gm_sums, sizeof(double) * (size_t)(m + 2)));
    op->gm_cap = m + 2;
  }
  std::vector<double*>& v = op->gmres_v;
  // device scalar products of `count` (vector, v_l) pairs -> d_gm_sums[offset ..], globally reduced
  auto reduce = [&](int offset, int count) -> int {
    for (int q = 0; q < count; ++q) reduce_final_kernel<<<1, kRedThreads, 0, st>>>(op->d_gm_partial + (size_t)(offset + q) * kRedBlocks, kRedBlocks, op->d_gm_sums + offset + q);
    CUDA_OK(cudaGetLastError());
    if (c->world > 1 && c->nccl.AllReduce(op->d_gm_sums + offset, op->d_gm_sums + offset, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, st) != 0) return fail(B200FEM_ERR_COMM, "ncclAllReduce failed");
    return B200FEM_OK;
  };
  auto norm2 = [&](const double* x, double* out) -> int {            // <x,x> over primary dofs, on the host
    dot_partial_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, x, op->d_aux, n, op->d_gm_partial);
    int e = reduce(0, 1); if (e) return e;
    CUDA_OK(cudaMemcpyAsync(out, op->d_gm_sums, sizeof(double), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
    return B200FEM_OK;
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
  const bool single = c->world == 1;
  BicgState host{}; const int chunk = 8; int issued = 0;
  do {
    for (int k = 0; k < chunk; ++k, ++issued) {
      rc = apply_dev_impl(op, p, tmp, true); if (rc) return rc;                                                // tmp = A p
      if (single) bicg_dot_alpha_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(tmp, rstar, op->d_aux, n, op->d_partial, op->d_bicg, op->d_counter);
      else {
        bicg_dot_alpha_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(tmp, rstar, op->d_aux, n, op->d_partial, op->d_bicg, nullptr);
        rc = reduce_sums(op, 1); if (rc) return rc;
        bicg_alpha_kernel<<<1, 32, 0, st>>>(op->d_sums, op->d_bicg);
      }
      bicg_s_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(sv, r, tmp, n, op->d_bicg);
      rc = apply_dev_impl(op, sv, r, true); if (rc) return rc;                                                 // r = A s
      if (single) bicg_dots5_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(r, sv, rstar, op->d_aux, n, op->d_partial5, op->d_bicg, op->d_hist, op->d_counter + 1);
      else {
        bicg_dots5_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(r, sv, rstar, op->d_aux, n, op->d_partial5, op->d_bicg, op->d_hist, nullptr);
        for (int q = 0; q < 5; ++q) reduce_final_kernel<<<1, kRedThreads, 0, st>>>(op->d_partial5 + (size_t)q * kRedBlocks, kRedBlocks, op->d_sums5 + q);
        if (c->nccl.AllReduce(op->d_sums5, op->d_sums5, 5, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, st) != 0) return fail(B200FEM_ERR_COMM, "ncclAllReduce failed");
        bicg_scalars_kernel<<<1, 32, 0, st>>>(op->d_sums5, op->d_bicg, op->d_hist);
      }
      bicg_update_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(x, r, p, sv, tmp, n, op->d_bicg, op->d_counter);
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(&host, op->d_bicg, sizeof(BicgState), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
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

// ---------------------------------------------------------------------------------------------------------------
extern "C" int b200fem_ctx_set_nccl(b200fem_ctx* c, void* comm, int rank, int world) {
  REQUIRE(c && comm, B200FEM_ERR_INVALID, "set_nccl: null argument");
  if (!c->nccl.load()) return fail(B200FEM_ERR_COMM, "libnccl.so.2 not loadable");
  c->comm = comm; c->own_comm = false; c->rank = rank; c->world = world; return B200FEM_OK;
}
extern "C" int b200fem_nccl_unique_id(void* out128) {
  REQUIRE(out128, B200FEM_ERR_INVALID, "null"); NcclApi api; if (!api.load()) return fail(B200FEM_ERR_COMM, "libnccl.so.2 not loadable");
  if (api.GetUniqueId(out128) != 0) return fail(B200FEM_ERR_COMM, "ncclGetUniqueId failed"); return B200FEM_OK;
}
extern "C" int b200fem_nccl_init(b200fem_ctx* c, const void* id128, int rank, int world) {
  REQUIRE(c && id128, B200FEM_ERR_INVALID, "nccl_init: null argument");
  if (!c->nccl.load()) return fail(B200FEM_ERR_COMM, "libnccl.so.2 not loadable");
  CUDA_OK(cudaSetDevice(c->device));
  NcclUniqueId id; std::memcpy(&id, id128, 128);
  if (c->nccl.CommInitRank(&c->comm, world, id, rank) != 0) return fail(B200FEM_ERR_COMM, "ncclCommInitRank failed");
  c->own_comm = true; c->rank = rank; c->world = world; return B200FEM_OK;
}
extern "C" int b200fem_communicate_dev(b200fem_operator* op, double* v) {
  REQUIRE(op && v, B200FEM_ERR_INVALID, "communicate: null argument");
  b200fem_ctx* c = op->sp->mesh->ctx; if (c->world <= 1) return B200FEM_OK;
  return exchange(op, v, c->stream);
}
