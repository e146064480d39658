// capi.cu -- C ABI (include/b200fem.h): handles and host logic (contexts, meshes, spaces, operator life cycle, setters,
// Dirichlet marks, the matrix-free diagonal).  The kernels are launched from the other translation units (internal.hpp).
// No CPU compute fallback exists: every compute entry point needs a CUDA device and reports B200FEM_ERR_CUDA otherwise.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "internal.hpp"
#include "kron_tables.hpp"

using namespace b200fem;

static thread_local std::string g_error;
namespace b200fem {
int fail(int code, const std::string& msg) { g_error = msg; return code; }
int ensure_smem_attr(b200fem_ctx* c, const void* kernel, size_t bytes) {
  if (c->attr_set.count(kernel)) return B200FEM_OK;
  CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  c->attr_set.insert(kernel); return B200FEM_OK;
}
}  // namespace b200fem

extern "C" const char* b200fem_last_error(void) { return g_error.c_str(); }
extern "C" int b200fem_version(void) { return 200; }

// ---------------------------------------------------------------------------------------------------------------
extern "C" int b200fem_device_count(int* count) {
  REQUIRE(count, B200FEM_ERR_INVALID, "device_count: null argument");
  int n = 0; const cudaError_t e = cudaGetDeviceCount(&n); *count = e == cudaSuccess ? n : 0;
  if (e != cudaSuccess || n == 0) return fail(B200FEM_ERR_CUDA, "no CUDA device available: this library has no CPU fallback");
  return B200FEM_OK;
}
extern "C" int b200fem_ctx_create(int device, void* stream, b200fem_ctx** out) {
  REQUIRE(out, B200FEM_ERR_INVALID, "ctx_create: out is null");
  int count = 0; cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(B200FEM_ERR_CUDA, "no CUDA device available: this library has no CPU fallback");
  REQUIRE(device >= 0 && device < count, B200FEM_ERR_INVALID, "ctx_create: bad device index");
  CUDA_OK(cudaSetDevice(device));
  auto* c = new b200fem_ctx; c->device = device;
  if (stream) c->stream = (cudaStream_t)stream; else { CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  CUDA_OK(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
  // host-mapped word for communication time-outs reported by kernels
  CUDA_OK(cudaHostAlloc((void**)&c->h_comm_error, sizeof(int), cudaHostAllocMapped)); *c->h_comm_error = 0;
  CUDA_OK(cudaHostGetDevicePointer((void**)&c->d_comm_error, c->h_comm_error, 0));
  *out = c; return B200FEM_OK;
}
static void ctx_delete(b200fem_ctx* c) {
  cudaSetDevice(c->device);
  if (c->scalars.ok) peer_scalars_free(c->scalars);
  if (c->own_comm && c->comm && c->nccl.ok()) c->nccl.CommDestroy(c->comm);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  if (c->h_comm_error) cudaFreeHost(c->h_comm_error);
  delete c;
}
static void ctx_unref(b200fem_ctx* c) { if (--c->refs <= 0 && c->released) ctx_delete(c); }
extern "C" int b200fem_ctx_destroy(b200fem_ctx* c) {
  if (!c || c->released) return B200FEM_OK;
  c->released = true;
  if (c->refs <= 0) ctx_delete(c);
  return B200FEM_OK;
}
extern "C" int b200fem_ctx_synchronize(b200fem_ctx* c) { REQUIRE(c, B200FEM_ERR_INVALID, "null ctx"); CUDA_OK(cudaSetDevice(c->device)); CUDA_OK(cudaStreamSynchronize(c->stream)); return check_comm_error(c); }
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
  BoxDev& b = m->box; b.dim = dim; b.periodic = 0;
  for (int d = 0; d < 3; ++d) {
    // block distribution: the first (gn % proc) ranks along an axis get one extra cell
    const int q = m->gn[d] / m->proc[d], r = m->gn[d] % m->proc[d], c = m->pc[d];
    m->olo[d] = c * q + std::min(c, r); m->ohi[d] = m->olo[d] + q + (c < r ? 1 : 0);
    const int glo = m->olo[d] > 0 ? 1 : 0, ghi = m->ohi[d] < m->gn[d] ? 1 : 0;      // overlap 1 where a neighbour rank exists
    b.origin[d] = m->olo[d] - glo; b.n[d] = (m->ohi[d] - m->olo[d]) + glo + ghi;
    b.own_lo[d] = glo; b.own_hi[d] = glo + (m->ohi[d] - m->olo[d]);
    b.gn[d] = m->gn[d]; b.lo[d] = m->lo[d]; b.h[d] = m->h[d]; b.ih[d] = 1.0 / m->h[d];
  }
  ctx->refs += 1;
  *out = m; return B200FEM_OK;
}
extern "C" int b200fem_mesh_set_periodic(b200fem_mesh* m, int mask) {
  REQUIRE(m && mask >= 0 && mask < (1 << m->dim), B200FEM_ERR_INVALID, "mesh_set_periodic: bad argument");
  REQUIRE(!m->unstructured, B200FEM_ERR_NOT_IMPLEMENTED, "periodic grids: Cartesian meshes");
  REQUIRE(m->refs == 0, B200FEM_ERR_INVALID, "mesh_set_periodic: spaces already exist on this mesh");
  REQUIRE(mask == 0 || m->proc[0] * m->proc[1] * m->proc[2] == 1, B200FEM_ERR_NOT_IMPLEMENTED, "periodic grids: one rank");
  for (int d = 0; d < m->dim; ++d) REQUIRE(!((mask >> d) & 1) || m->gn[d] >= 2, B200FEM_ERR_INVALID, "periodic axis needs at least two cells");
  m->box.periodic = mask; return B200FEM_OK;
}
extern "C" int b200fem_mesh_cartesian(b200fem_ctx* ctx, int dim, const int32_t* n, const double* lo, const double* hi, b200fem_mesh** out) {
  return make_mesh(ctx, dim, n, lo, hi, nullptr, 0, out);
}
extern "C" int b200fem_mesh_cartesian_distributed(b200fem_ctx* ctx, int dim, const int32_t* n, const double* lo, const double* hi, const int32_t* proc, int rank, b200fem_mesh** out) {
  REQUIRE(proc, B200FEM_ERR_INVALID, "mesh: proc is null");
  return make_mesh(ctx, dim, n, lo, hi, proc, rank, out);
}
static void mesh_unref(b200fem_mesh* m) { if (--m->refs <= 0 && m->released) { ctx_unref(m->ctx); delete m; } }
extern "C" int b200fem_mesh_destroy(b200fem_mesh* m) {
  if (!m || m->released) return B200FEM_OK;
  m->released = true;
  if (m->refs <= 0) { ctx_unref(m->ctx); delete m; }
  return B200FEM_OK;
}
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
  REQUIRE(!m->unstructured, B200FEM_ERR_NOT_IMPLEMENTED, "mesh_local_box: Cartesian meshes");
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

extern "C" int b200fem_space_create_vector(b200fem_mesh* mesh, int kind, int order, int numbering, int dim_range, b200fem_space** out) {
  REQUIRE(dim_range >= 1 && dim_range <= 4, B200FEM_ERR_NOT_IMPLEMENTED, "space_create_vector: dimRange 1..4");
  REQUIRE(dim_range == 1 || order <= 3, B200FEM_ERR_NOT_IMPLEMENTED, "vector-valued spaces: orders 1..3");
  REQUIRE(dim_range == 1 || !(mesh && mesh->unstructured), B200FEM_ERR_NOT_IMPLEMENTED, "vector-valued spaces on unstructured meshes");
  int rc = b200fem_space_create(mesh, kind, order, numbering, out); if (rc) return rc;
  (*out)->dim_range = dim_range; (*out)->size *= dim_range;        // space.size() = blockMapper().size() * localBlockSize
  return B200FEM_OK;
}
extern "C" int b200fem_space_dim_range(b200fem_space* s, int32_t* dim_range) { REQUIRE(s && dim_range, B200FEM_ERR_INVALID, "null"); *dim_range = s->dim_range; return B200FEM_OK; }
extern "C" int b200fem_space_create(b200fem_mesh* mesh, int kind, int order, int numbering, b200fem_space** out) {
  REQUIRE(mesh && out, B200FEM_ERR_INVALID, "space_create: null argument");
  REQUIRE(kind >= 0 && kind <= 3, B200FEM_ERR_INVALID, "space_create: unknown space kind");
  try {
    auto s = std::unique_ptr<b200fem_space>(new b200fem_space);
    s->mesh = mesh; s->kind = kind; s->order = order; s->numbering = numbering; s->n1 = order + 1;
    const int dim = mesh->dim; s->nb = 1; for (int d = 0; d < dim; ++d) s->nb *= s->n1;
    s->box = mesh->box;
    if (mesh->unstructured) {
      REQUIRE(kind == B200FEM_LAGRANGE, B200FEM_ERR_NOT_IMPLEMENTED, "unstructured meshes carry continuous Lagrange spaces (DG spaces need the face connectivity of a Cartesian mesh)");
      REQUIRE(order == 1 || order == 2, B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange spaces: order 1 and 2 only");
      s->tab = tabulate_1d(Basis::Lagrange, order, gauss_points_for_order(2 * order));
      s->lay = LagrangeLayoutDev{}; s->lay.order = order;
      int rc = unstructured_space_setup(s.get()); if (rc) { unstructured_space_free(s.get()); return rc; }
      mesh->refs += 1;
      *out = s.release(); return B200FEM_OK;
    }
    if (kind == B200FEM_LAGRANGE) {
      REQUIRE(order >= 1 && order <= 3, B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange spaces: orders 1 to 3");
      REQUIRE(order <= 2 || numbering == B200FEM_NUMBERING_YASP, B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange order 3: YaspGrid numbering");
      REQUIRE(mesh->box.periodic == 0, B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange spaces on periodic grids (dof identification across the boundary)");
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
        for (int d = 0; d < dim; ++d) if ((sft >> d) & 1) c *= order - 1;          // (order - 1)^p nodes inside an entity of dimension p
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
      REQUIRE(order >= 1 && order <= (kind == B200FEM_DG_ONB ? 4 : 5), B200FEM_ERR_NOT_IMPLEMENTED, "DG spaces: Legendre orders 1..5, dgonb orders 1..4");
      const BoxDev& b = s->box;
      // the kernels work on the full tensor basis of the (3-D) cube; the space is a sub-basis of it (tables.hpp: dg_tensor_map)
      s->perm = dg_tensor_map(dim, order, kind, &s->nb);
      s->tensor_full = dim == 3 && s->nb == s->n1 * s->n1 * s->n1;
      s->elements = (long long)b.n[0] * b.n[1] * b.n[2]; s->size = s->elements * s->nb;
      s->tab = tabulate_1d(Basis::Legendre, order, gauss_points_for_order(2 * order));
    }
    mesh->refs += 1;
    *out = s.release(); return B200FEM_OK;
  } catch (const std::exception& ex) { return fail(B200FEM_ERR_INVALID, ex.what()); }
}
static void space_delete(b200fem_space* s) { unstructured_space_free(s); if (s->d_lattice_map) { cudaSetDevice(s->mesh->ctx->device); cudaFree(s->d_lattice_map); } mesh_unref(s->mesh); delete s; }
static void space_unref(b200fem_space* s) { if (--s->refs <= 0 && s->released) space_delete(s); }
extern "C" int b200fem_space_destroy(b200fem_space* s) {
  if (!s || s->released) return B200FEM_OK;
  s->released = true;
  if (s->refs <= 0) space_delete(s);
  return B200FEM_OK;
}
extern "C" int b200fem_space_size(b200fem_space* s, int64_t* size) { REQUIRE(s && size, B200FEM_ERR_INVALID, "null"); *size = s->size; return B200FEM_OK; }
extern "C" int b200fem_space_local_size(b200fem_space* s, int32_t* nb) { REQUIRE(s && nb, B200FEM_ERR_INVALID, "null"); *nb = s->nb; return B200FEM_OK; }
extern "C" int b200fem_space_elements(b200fem_space* s, int64_t* n) { REQUIRE(s && n, B200FEM_ERR_INVALID, "null"); *n = s->elements; return B200FEM_OK; }
extern "C" int b200fem_space_dofmap(b200fem_space* s, int64_t e, int64_t* out) {
  REQUIRE(s && out, B200FEM_ERR_INVALID, "null"); REQUIRE(e >= 0 && e < s->elements, B200FEM_ERR_INVALID, "dofmap: element out of range");
  if (s->kind != B200FEM_LAGRANGE) { for (int j = 0; j < s->nb; ++j) out[j] = e * s->nb + j; return B200FEM_OK; }
  if (s->unst) return unstructured_dofmap(s, e, out);
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
// ---------------------------------------------------------------------------------------------------------------
static void mark_dirichlet(b200fem_operator* op) {
  // DirichletConstraints::updateDirichletDofs (schemes/dirichletconstraints.hh:435-554): all Lagrange nodes on boundary
  // faces whose side is flagged; values g(x_node).  Host-side, closed form over the boundary lattice.
  b200fem_space* s = op->sp; const BoxDev& b = s->box; const int dim = b.dim, k = s->order;
  if (s->unst) { unstructured_mark_dirichlet(op); return; }
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
  REQUIRE(s->dim_range == 1, B200FEM_ERR_NOT_IMPLEMENTED, "the built-in advection-diffusion-reaction integrands are scalar: vector-valued spaces take run-time compiled integrands (b200fem_operator_create_jit)");
  return b200fem::operator_create_impl(s, model, out);
}
int b200fem::operator_create_impl(b200fem_space* s, const b200fem_model* model, b200fem_operator** out) {
  REQUIRE(!(model->strong_dirichlet && s->kind != B200FEM_LAGRANGE), B200FEM_ERR_INVALID, "strong Dirichlet constraints need a Lagrange space");
  REQUIRE(!(s->unst && s->mesh->ctx->world > 1), B200FEM_ERR_NOT_IMPLEMENTED, "unstructured meshes: one rank (the context joined a communicator after the mesh was made)");
  REQUIRE(!(s->unst && (model->has_skeleton || model->has_boundary)), B200FEM_ERR_NOT_IMPLEMENTED, "unstructured meshes: interior integrands and strong Dirichlet constraints (no skeleton / boundary terms)");
  b200fem_ctx* c = s->mesh->ctx;
  CUDA_OK(cudaSetDevice(c->device));
  auto* op = new b200fem_operator; op->sp = s; op->model = *model; s->refs += 1;
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
  if (c->world > 1) {
    // index lists of the NCCL transport (always built: they also define the auxiliary-dof mask and are the fallback)
    // doubles per exchanged block: an element of a DG space (n_b blocks of dimRange components), a node of a Lagrange space (dimRange)
    const int blk = s->kind == B200FEM_LAGRANGE ? s->dim_range : s->nb * s->dim_range;
    int rc = halo_plan_build(op->halo, s->mesh->proc, s->mesh->pc, s->box, s->kind == B200FEM_LAGRANGE, s->kind == B200FEM_LAGRANGE ? s->order : 0, blk, s->lay, s->size, &op->d_aux);
    if (!rc && s->kind != B200FEM_LAGRANGE) rc = halo_plan_dg_build(op->halo_dg, s->mesh->proc, s->mesh->pc, s->box, blk);
    if (rc) { b200fem_operator_destroy(op); return fail(B200FEM_ERR_COMM, "halo plan failed"); }
    // peer-memory transport (all ranks must take the same decision: it depends on the context's peer probe, made when the
    // communicator was attached, and on the collective outcome of the mailbox mapping)
    if (c->scalars.ok) {
      int ok;
      if (s->kind != B200FEM_LAGRANGE) {
        ok = halo_plan_p2p_build(op->halo_p2p, op->halo_dg, c->nccl, c->comm, c->rank, c->world, s->mesh->proc, s->mesh->pc, s->mesh->gn, s->box, s->nb * s->dim_range, c->d_comm_error, c->stream) == 0;
      } else {
        ok = halo_plan_add_build(op->halo_add, c->nccl, c->comm, c->rank, c->world, s->mesh->proc, s->mesh->pc, s->lay, s->lattice_map, s->box.dim, s->dim_range, c->d_comm_error, c->stream) == 0;
      }
      cudaGetLastError();
      // agree on the outcome: one failing rank switches everybody to NCCL
      int flag = ok ? 1 : 0; int* d_ok = nullptr; CUDA_OK(cudaMalloc(&d_ok, sizeof(int)));
      CUDA_OK(cudaMemcpy(d_ok, &flag, sizeof(int), cudaMemcpyHostToDevice));
      if (c->nccl.AllReduce(d_ok, d_ok, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, c->comm, c->stream) != 0) { cudaFree(d_ok); b200fem_operator_destroy(op); return fail(B200FEM_ERR_COMM, "ncclAllReduce failed"); }
      CUDA_OK(cudaStreamSynchronize(c->stream)); CUDA_OK(cudaMemcpy(&flag, d_ok, sizeof(int), cudaMemcpyDeviceToHost)); cudaFree(d_ok);
      if (!flag) { halo_plan_p2p_free(op->halo_p2p); halo_plan_add_free(op->halo_add); }
    }
  }
  *out = op; return B200FEM_OK;
}
extern "C" int b200fem_operator_destroy(b200fem_operator* op) {
  if (!op) return B200FEM_OK;
  b200fem_ctx* c = op->sp->mesh->ctx;
  cudaSetDevice(c->device); cudaStreamSynchronize(c->stream);
  for (void* p : {(void*)op->d_perm, (void*)op->d_bvec, (void*)op->d_dmask, (void*)op->d_dvals, (void*)op->d_aux, (void*)op->d_u, (void*)op->d_w, (void*)op->d_h, (void*)op->d_r,
                  (void*)op->d_p, (void*)op->d_x, (void*)op->d_b, (void*)op->d_partial, (void*)op->d_sums, (void*)op->d_hist, (void*)op->d_cg, (void*)op->d_lag_rows, (void*)op->d_counter, (void*)op->d_rstar, (void*)op->d_s, (void*)op->d_tmp, (void*)op->d_partial5, (void*)op->d_sums5, (void*)op->d_bicg, (void*)op->d_dot_partial, (void*)op->d_nw_res, (void*)op->d_nw_dw, (void*)op->d_nw_w, (void*)op->d_nw_u}) if (p) cudaFree(p);
  jit_free(op);
  if (op->cg_graph) cudaGraphExecDestroy(op->cg_graph);
  for (double* q : op->gmres_v) if (q) cudaFree(q);
  for (void* q : {(void*)op->d_dinv, (void*)op->d_pq, (void*)op->d_ps}) if (q) cudaFree(q);
  for (void* q : {(void*)op->d_jac_u, (void*)op->d_jac_opu, (void*)op->d_jac_b, (void*)op->d_fd}) if (q) cudaFree(q);
  if (op->d_gm_partial) cudaFree(op->d_gm_partial);
  if (op->d_gm_sums) cudaFree(op->d_gm_sums);
  // mailboxes are mapped by the peers: nobody may still be writing into them (every exchange has been received before its
  // kernel retired; the stream was drained above)
  halo_plan_p2p_free(op->halo_p2p); halo_plan_add_free(op->halo_add); halo_plan_free(op->halo); halo_plan_dg_free(op->halo_dg); free_march_cache(op);
  if (op->h2d_stream) cudaStreamDestroy(op->h2d_stream);
  if (op->d2h_stream) cudaStreamDestroy(op->d2h_stream);
  for (cudaEvent_t e : op->pipe_ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {op->ev0, op->ev1, op->evx0, op->evx1}) if (e) cudaEventDestroy(e);
  space_unref(op->sp);
  delete op; return B200FEM_OK;
}
extern "C" int b200fem_operator_apply(b200fem_operator* op, const double* u, double* w) { return apply_host(op, u, w, false); }
extern "C" int b200fem_operator_apply_linear(b200fem_operator* op, const double* u, double* w) { return apply_host(op, u, w, true); }
extern "C" int b200fem_operator_apply_dev(b200fem_operator* op, const double* u, double* w, int linear) {
  REQUIRE(op && u && w, B200FEM_ERR_INVALID, "apply_dev: null argument");
  b200fem_ctx* c = op->sp->mesh->ctx;
  int cur = -1; cudaGetDevice(&cur);
  if (cur != c->device) CUDA_OK(cudaSetDevice(c->device));
  if (c->world > 1) { int rc = check_comm_error(c); if (rc) return rc; }
  return apply_dev_impl(op, u, w, linear != 0);
}
extern "C" int b200fem_operator_load_vector(b200fem_operator* op, double* b_host) {
  REQUIRE(op && b_host, B200FEM_ERR_INVALID, "load_vector: null argument");
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const size_t bytes = sizeof(double) * (size_t)s->size;
  CUDA_OK(cudaSetDevice(s->mesh->ctx->device));
  if (!op->d_u) CUDA_OK(cudaMalloc(&op->d_u, bytes));
  if (!op->d_w) CUDA_OK(cudaMalloc(&op->d_w, bytes));
  CUDA_OK(cudaMemsetAsync(op->d_u, 0, bytes, st));
  const int saved = op->kernel_pref; op->kernel_pref = B200FEM_KERNEL_QUADRATURE;
  int rc = apply_dev_impl(op, op->d_u, op->d_w, false); op->kernel_pref = saved; if (rc) return rc;
  rc = negate_dev(op->d_w, s->size, st); if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(b_host, op->d_w, bytes, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  return check_comm_error(s->mesh->ctx);
}
// every setter that changes what an apply computes bumps the state version: cached Kronecker tables and captured Krylov
// graphs are rebuilt on the next use
extern "C" int b200fem_operator_set_communicate(b200fem_operator* op, int c) { REQUIRE(op, B200FEM_ERR_INVALID, "null"); if (op->communicate != (c != 0)) { op->communicate = c != 0; invalidate_cached_state(op); } return B200FEM_OK; }
extern "C" int b200fem_operator_set_quadrature_orders(b200fem_operator* op, unsigned qi, unsigned qs) {
  REQUIRE(op, B200FEM_ERR_INVALID, "null");
  if (op->q_interior != qi || op->q_surface != qs) {
    op->q_interior = qi; op->q_surface = qs; invalidate_cached_state(op);
    // the load vector was integrated with the old rules
    if (op->d_bvec) { CUDA_OK(cudaSetDevice(op->sp->mesh->ctx->device)); CUDA_OK(cudaStreamSynchronize(op->sp->mesh->ctx->stream)); CUDA_OK(cudaFree(op->d_bvec)); op->d_bvec = nullptr; }
  }
  return B200FEM_OK;
}
extern "C" int b200fem_operator_set_kernel(b200fem_operator* op, int k) {
  REQUIRE(op && k >= 0 && k <= 3, B200FEM_ERR_INVALID, "set_kernel: bad kernel id");
  if (op->kernel_pref != k) { op->kernel_pref = k; invalidate_cached_state(op); }
  return B200FEM_OK;
}
extern "C" int b200fem_operator_set_host_pipeline(b200fem_operator* op, int chunks) { REQUIRE(op && chunks >= 0, B200FEM_ERR_INVALID, "set_host_pipeline: bad argument"); op->host_pipeline_chunks = chunks; return B200FEM_OK; }
extern "C" int b200fem_operator_set_inverse_mass(b200fem_operator* op, int on) {
  REQUIRE(op, B200FEM_ERR_INVALID, "null");
  REQUIRE(op->sp->kind != B200FEM_LAGRANGE, B200FEM_ERR_NOT_IMPLEMENTED, "inverse mass (MOLGalerkinOperator): DG spaces only");
  if (op->inverse_mass == (on != 0)) return B200FEM_OK;
  op->inverse_mass = on != 0;
  // the scaled 1-D operators and the scaled load vector are rebuilt on the next apply
  invalidate_cached_state(op);
  if (op->d_bvec) { CUDA_OK(cudaSetDevice(op->sp->mesh->ctx->device)); CUDA_OK(cudaStreamSynchronize(op->sp->mesh->ctx->stream)); CUDA_OK(cudaFree(op->d_bvec)); op->d_bvec = nullptr; }
  return B200FEM_OK;
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

// diag(A) of the Kronecker form, on the host (setup cost O(N), once per operator)
int host_diagonal(b200fem_operator* op, std::vector<double>& diag, bool dirichlet_rows) {
  b200fem_space* s = op->sp; const BoxDev& b = s->box;
  if (s->unst) return unstructured_diagonal(op, diag, dirichlet_rows);
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
  std::vector<double> d; int rc = host_diagonal(op, d, true); if (rc) return rc;
  std::copy(d.begin(), d.end(), diag_host); return B200FEM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// peer-memory transport: probed once per context, collectively, when the communicator is attached
static int attach_peer_memory(b200fem_ctx* c) {
  if (c->world <= 1 || std::getenv("B200FEM_NO_P2P")) return B200FEM_OK;
  CUDA_OK(cudaSetDevice(c->device));
  if (peer_scalars_create(c->nccl, c->comm, c->rank, c->world, c->stream, c->d_comm_error, c->scalars) != 0) { cudaGetLastError(); c->scalars = PeerScalars(); }
  return B200FEM_OK;
}
extern "C" int b200fem_ctx_set_nccl(b200fem_ctx* c, void* comm, int rank, int world) {
  REQUIRE(c && comm, B200FEM_ERR_INVALID, "set_nccl: null argument");
  if (!c->nccl.load()) return fail(B200FEM_ERR_COMM, "libnccl.so.2 not loadable");
  c->comm = comm; c->own_comm = false; c->rank = rank; c->world = world;
  return attach_peer_memory(c);
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
  c->own_comm = true; c->rank = rank; c->world = world;
  return attach_peer_memory(c);
}
extern "C" int b200fem_ctx_transport(b200fem_ctx* c, int* peer_memory) { REQUIRE(c && peer_memory, B200FEM_ERR_INVALID, "null"); *peer_memory = c->scalars.ok ? 1 : 0; return B200FEM_OK; }
extern "C" int b200fem_communicate_dev(b200fem_operator* op, double* v) {
  REQUIRE(op && v, B200FEM_ERR_INVALID, "communicate: null argument");
  b200fem_ctx* c = op->sp->mesh->ctx; if (c->world <= 1) return B200FEM_OK;
  return exchange(op, v, c->stream);
}
