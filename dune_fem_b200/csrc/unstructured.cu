// unstructured.cu -- continuous Lagrange spaces (order 1, 2) on unstructured conforming cube meshes: the mesh handle, the dof
// numbering (index arrays), element colouring, Dirichlet marks, the matrix-free diagonal and the launcher of
// lagrange_unstructured.cuh.  One rank.
//
// What the reference gets from ALUGrid< dim, dim, cube, conforming > behind an AdaptiveLeafGridPart, restated on plain arrays:
//  * dof blocks by geometry type -- vertices, edges, faces, cells (space/mapper/indexsetdofmapper.hh:504-515) -- and inside a
//    block the AdaptiveLeafIndexSet's first-touch order: elements in index order, sub-entities of an element in reference-element
//    order (gridpart/adaptiveleafindexset.hh:884-906, 1015-1019); the element's local numbering has coordinate 0 fastest
//    (space/lagrange/genericlagrangepoints.hh:862-876).  The per-element index array is what DofMapperCode compiles to
//    (space/lagrange/dofmappercode.hh:56-104);
//  * geometry: the multilinear map of the element's 2^dim vertices.
#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <memory>

#include "internal.hpp"
#include "lagrange_unstructured.cuh"

namespace b200fem {

struct UnstructuredSpace {
  std::vector<int> dofs;                  // [nelem][nb]
  std::vector<double> node_x;             // [size][3]
  std::vector<uint8_t> boundary;          // [size]: node lies on a boundary face
  std::vector<int> order;                 // elements sorted by colour
  std::vector<int> colour_begin;          // [ncolours + 1]
  std::vector<double> tabB, tabG;         // host copies of the tabulation: B[q * nb + i], G[(q * nb + i) * 3 + d]
  std::vector<double> xq, wq;
  int* d_dofs = nullptr; int* d_order = nullptr; double* d_elem_x = nullptr; double* d_tab = nullptr;
  UnstructuredTabDev tab{};
};

namespace {

// sub-entities of the cube in reference-element order (dune-geometry), as lattice offsets in {0, 1, 2}: 1 = the entity extends
// along that axis
void sub_entity_order(int dim, std::vector<std::array<int, 3>> subs[4]) {
  for (int v = 0; v < (1 << dim); ++v) { std::array<int, 3> a = {0, 0, 0}; for (int d = 0; d < dim; ++d) a[d] = 2 * ((v >> d) & 1); subs[0].push_back(a); }
  if (dim == 2) { subs[1] = {{0, 1, 0}, {2, 1, 0}, {1, 0, 0}, {1, 2, 0}}; subs[2] = {{1, 1, 0}}; }
  if (dim == 3) {
    subs[1] = {{0, 0, 1}, {2, 0, 1}, {0, 2, 1}, {2, 2, 1}, {0, 1, 0}, {2, 1, 0}, {1, 0, 0}, {1, 2, 0}, {0, 1, 2}, {2, 1, 2}, {1, 0, 2}, {1, 2, 2}};
    subs[2] = {{0, 1, 1}, {2, 1, 1}, {1, 0, 1}, {1, 2, 1}, {1, 1, 0}, {1, 1, 2}};
    subs[3] = {{1, 1, 1}};
  }
}
using Key = std::array<long long, 4>;
Key entity_key(const b200fem_mesh* m, long long e, const std::array<int, 3>& a) {
  const int nv = 1 << m->dim; Key k = {-1, -1, -1, -1}; int n = 0;
  for (int v = 0; v < nv; ++v) {
    bool in = true;
    for (int d = 0; d < m->dim; ++d) if (a[d] != 1 && ((v >> d) & 1) != a[d] / 2) in = false;
    if (in && n < 4) k[n++] = m->uev[(size_t)e * nv + v];
  }
  std::sort(k.begin(), k.begin() + n); return k;
}
// multilinear map of element e at xi: position and Jacobian (J[i][d] = d x_i / d xi_d)
void element_map(const b200fem_mesh* m, long long e, const double* xi, double x[3], double J[3][3]) {
  const int dim = m->dim, nv = 1 << dim;
  for (int i = 0; i < 3; ++i) { x[i] = 0; for (int d = 0; d < 3; ++d) J[i][d] = (i == d && i >= dim) ? 1.0 : 0.0; }
  for (int v = 0; v < nv; ++v) {
    double N = 1, dN[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) {
      const double a = ((v >> d) & 1) ? xi[d] : 1.0 - xi[d], da = ((v >> d) & 1) ? 1.0 : -1.0;
      N *= a; for (int k = 0; k < dim; ++k) dN[k] *= k == d ? da : a;
    }
    const double* X = &m->ux[(size_t)m->uev[(size_t)e * nv + v] * dim];
    for (int i = 0; i < dim; ++i) { x[i] += N * X[i]; for (int d = 0; d < dim; ++d) J[i][d] += dN[d] * X[i]; }
  }
}
double invert3(const double J[3][3], double Ji[3][3]) {
  const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  const double id = 1.0 / det;
  Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
  Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
  Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  return det;
}

}  // namespace

// host-only part of the space setup: numbering, node positions, boundary marks, colouring (no CUDA; also behind
// b200fem_unstructured_numbering for CPU tests)
static int unstructured_number(const b200fem_mesh* m, int k, int nb, UnstructuredSpace* U, long long* size_out) {
  const int dim = m->dim, n1 = k + 1, nv = 1 << dim;
  REQUIRE(m->nelem < (1ll << 31), B200FEM_ERR_NOT_IMPLEMENTED, "unstructured meshes: 32-bit element indices");
  std::vector<std::array<int, 3>> subs[4]; sub_entity_order(dim, subs);
  auto local_index = [&](const std::array<int, 3>& a) { int l = 0, st = 1; for (int d = 0; d < dim; ++d) { l += st * (k == 1 ? a[d] / 2 : a[d]); st *= n1; } return l; };
  // ---- first-touch numbering per entity dimension; faces are counted for the boundary detection
  std::map<Key, long long> index[4]; std::map<Key, int> face_count;
  for (long long e = 0; e < m->nelem; ++e)
    for (int cd = 0; cd <= dim; ++cd) {
      const int pd = dim - cd;
      for (auto& a : subs[pd]) {
        const Key key = entity_key(m, e, a);
        if (pd == dim - 1) face_count[key] += 1;
        if (k == 1 && pd != 0) continue;
        if (!index[pd].count(key)) { const long long i = (long long)index[pd].size(); index[pd][key] = i; }
      }
    }
  long long off[5] = {0, 0, 0, 0, 0};
  for (int p = 0; p <= dim; ++p) off[p + 1] = off[p] + (long long)index[p].size();
  const long long size = off[dim + 1]; *size_out = size;
  REQUIRE(size < (1ll << 31), B200FEM_ERR_NOT_IMPLEMENTED, "unstructured meshes: 32-bit dof indices");
  U->dofs.assign((size_t)m->nelem * nb, -1); U->node_x.assign((size_t)size * 3, 0.0); U->boundary.assign((size_t)size, 0);
  for (long long e = 0; e < m->nelem; ++e)
    for (int pd = 0; pd <= dim; ++pd) {
      if (k == 1 && pd != 0) continue;
      for (auto& a : subs[pd]) {
        const int l = local_index(a); const long long g = off[pd] + index[pd][entity_key(m, e, a)];
        U->dofs[(size_t)e * nb + l] = (int)g;
        double xi[3] = {0, 0, 0}, J[3][3]; for (int d = 0; d < dim; ++d) xi[d] = a[d] / 2.0;
        element_map(m, e, xi, &U->node_x[(size_t)g * 3], J);
      }
    }
  for (long long e = 0; e < m->nelem; ++e)
    for (auto& f : subs[dim - 1]) {
      if (face_count[entity_key(m, e, f)] != 1) continue;
      for (int l = 0; l < nb; ++l) {
        int idx = l; bool on = true;
        for (int d = 0; d < dim; ++d) { const int a = (idx % n1) * 2 / k; idx /= n1; if (f[d] != 1 && a != f[d]) on = false; }
        if (on) U->boundary[(size_t)U->dofs[(size_t)e * nb + l]] = 1;
      }
    }
  // ---- element colouring: greedy over the vertex adjacency (elements sharing a dof share a vertex on a conforming mesh)
  {
    std::vector<std::vector<int>> at_vertex((size_t)m->nvert);
    for (long long e = 0; e < m->nelem; ++e) for (int v = 0; v < nv; ++v) at_vertex[(size_t)m->uev[(size_t)e * nv + v]].push_back((int)e);
    std::vector<int> colour((size_t)m->nelem, -1); int ncol = 0; std::vector<char> used;
    for (long long e = 0; e < m->nelem; ++e) {
      used.assign((size_t)ncol + 1, 0);
      for (int v = 0; v < nv; ++v) for (int o : at_vertex[(size_t)m->uev[(size_t)e * nv + v]]) if (colour[(size_t)o] >= 0) used[(size_t)colour[(size_t)o]] = 1;
      int c = 0; while (used[(size_t)c]) ++c;
      colour[(size_t)e] = c; ncol = std::max(ncol, c + 1);
    }
    U->colour_begin.assign((size_t)ncol + 1, 0);
    for (long long e = 0; e < m->nelem; ++e) U->colour_begin[(size_t)colour[(size_t)e] + 1] += 1;
    for (int c = 0; c < ncol; ++c) U->colour_begin[(size_t)c + 1] += U->colour_begin[(size_t)c];
    U->order.resize((size_t)m->nelem); std::vector<int> fill(U->colour_begin.begin(), U->colour_begin.end() - 1);
    for (long long e = 0; e < m->nelem; ++e) U->order[(size_t)fill[(size_t)colour[(size_t)e]]++] = (int)e;      // element order kept inside a colour
  }
  return B200FEM_OK;
}

int unstructured_space_setup(b200fem_space* s) {
  const b200fem_mesh* m = s->mesh; const int dim = m->dim, k = s->order, n1 = k + 1, nb = s->nb, nv = 1 << dim;
  UnstructuredSpace* U = new UnstructuredSpace; s->unst = U;        // (owned by the space from here on: a failing caller frees it, device arrays included)
  int rc = unstructured_number(m, k, nb, U, &s->size); if (rc) return rc;
  s->elements = m->nelem;
  // ---- tabulation of the tensor basis at the tensor Gauss rule (x0 fastest, quadrature/femquadratures_inline.hh:33-95)
  const Tab1D& t = s->tab; const int nq = nb;
  U->tabB.assign((size_t)nq * nb, 0.0); U->tabG.assign((size_t)nq * nb * 3, 0.0); U->xq.assign((size_t)nq * 3, 0.0); U->wq.assign((size_t)nq, 0.0);
  for (int q = 0; q < nq; ++q) {
    int qd[3] = {0, 0, 0}; { int z = q; for (int d = 0; d < dim; ++d) { qd[d] = z % n1; z /= n1; } }
    double wt = 1; for (int d = 0; d < dim; ++d) { U->xq[(size_t)q * 3 + d] = t.x[(size_t)qd[d]]; wt *= t.w[(size_t)qd[d]]; }
    U->wq[(size_t)q] = wt;
    for (int i = 0; i < nb; ++i) {
      int id[3] = {0, 0, 0}; { int z = i; for (int d = 0; d < dim; ++d) { id[d] = z % n1; z /= n1; } }
      double b = 1, g[3] = {1, 1, 1};
      for (int c = 0; c < dim; ++c) {
        const double p = t.B[(size_t)qd[c] * n1 + id[c]], dp = t.G[(size_t)qd[c] * n1 + id[c]];
        b *= p; for (int d = 0; d < dim; ++d) g[d] *= c == d ? dp : p;
      }
      U->tabB[(size_t)q * nb + i] = b;
      for (int d = 0; d < 3; ++d) U->tabG[((size_t)q * nb + i) * 3 + d] = d < dim ? g[d] : 0.0;
    }
  }
  // ---- device arrays
  CUDA_OK(cudaSetDevice(m->ctx->device));
  std::vector<double> ex((size_t)m->nelem * nv * 3, 0.0);
  for (long long e = 0; e < m->nelem; ++e) for (int v = 0; v < nv; ++v) for (int d = 0; d < dim; ++d) ex[((size_t)e * nv + v) * 3 + d] = m->ux[(size_t)m->uev[(size_t)e * nv + v] * dim + d];
  // tabulation with the point as the fast index, then points and weights
  std::vector<double> flat; const size_t nn = (size_t)nq * nb;
  flat.resize(nn * 4 + (size_t)nq * 4, 0.0);
  double* Bq = flat.data(); double* Gq = Bq + nn; double* xq = Gq + 3 * nn; double* wq = xq + 3 * (size_t)nq;
  for (int q = 0; q < nq; ++q) for (int i = 0; i < nb; ++i) {
    Bq[(size_t)i * nq + q] = U->tabB[(size_t)q * nb + i];
    for (int d = 0; d < 3; ++d) Gq[((size_t)d * nb + i) * nq + q] = U->tabG[((size_t)q * nb + i) * 3 + d];
  }
  std::copy(U->xq.begin(), U->xq.end(), xq); std::copy(U->wq.begin(), U->wq.end(), wq);
  CUDA_OK(cudaMalloc(&U->d_dofs, sizeof(int) * U->dofs.size())); CUDA_OK(cudaMemcpy(U->d_dofs, U->dofs.data(), sizeof(int) * U->dofs.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&U->d_order, sizeof(int) * U->order.size())); CUDA_OK(cudaMemcpy(U->d_order, U->order.data(), sizeof(int) * U->order.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&U->d_elem_x, sizeof(double) * ex.size())); CUDA_OK(cudaMemcpy(U->d_elem_x, ex.data(), sizeof(double) * ex.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&U->d_tab, sizeof(double) * flat.size())); CUDA_OK(cudaMemcpy(U->d_tab, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice));
  U->tab.B = U->d_tab; U->tab.G = U->d_tab + nn; U->tab.xq = U->d_tab + 4 * nn; U->tab.wq = U->d_tab + 4 * nn + 3 * (size_t)nq;
  return B200FEM_OK;
}

void unstructured_space_free(b200fem_space* s) {
  if (!s->unst) return;
  cudaSetDevice(s->mesh->ctx->device);
  for (void* p : {(void*)s->unst->d_dofs, (void*)s->unst->d_order, (void*)s->unst->d_elem_x, (void*)s->unst->d_tab}) if (p) cudaFree(p);
  delete s->unst; s->unst = nullptr;
}

int unstructured_dofmap(const b200fem_space* s, long long e, int64_t* out) {
  for (int l = 0; l < s->nb; ++l) out[l] = s->unst->dofs[(size_t)e * s->nb + l];
  return B200FEM_OK;
}

// DirichletConstraints::updateDirichletDofs (schemes/dirichletconstraints.hh:435-554) with every boundary intersection a
// Dirichlet intersection: all nodes on boundary faces, values g(x_node)
void unstructured_mark_dirichlet(b200fem_operator* op) {
  b200fem_space* s = op->sp; const int dim = s->mesh->dim;
  op->h_dmask.assign((size_t)s->size, 0); op->h_dvals.assign((size_t)s->size, 0.0);
  for (long long i = 0; i < s->size; ++i) {
    if (!s->unst->boundary[(size_t)i]) continue;
    const double* x = &s->unst->node_x[(size_t)i * 3]; double val = 0;
    if (op->model.data == 1) val = std::sin(x[0] * x[1]);
    else if (op->model.data == 2) { val = 1; for (int d = 0; d < dim; ++d) val *= std::sin(M_PI * x[d]); }
    op->h_dmask[(size_t)i] = 1; op->h_dvals[(size_t)i] = val;
  }
}

// diag(A) of the homogeneous linear part by element-local contractions (setup cost O(elements), once per operator):
// A_ii = sum_K sum_q w_q |det J| ( c phi_i^2 + (eps grad phi_i - b phi_i) . grad phi_i )
int unstructured_diagonal(b200fem_operator* op, std::vector<double>& diag, bool dirichlet_rows) {
  b200fem_space* s = op->sp; const b200fem_mesh* m = s->mesh; const UnstructuredSpace* U = s->unst; const int dim = m->dim, nb = s->nb, nq = nb;
  REQUIRE(op->model.gamma == 0.0 && default_quadrature(op), B200FEM_ERR_NOT_IMPLEMENTED, "diagonal: needs a linear model with the default quadrature");
  diag.assign((size_t)s->size, 0.0);
  for (long long e = 0; e < m->nelem; ++e)
    for (int q = 0; q < nq; ++q) {
      double x[3], J[3][3], Ji[3][3]; element_map(m, e, &U->xq[(size_t)q * 3], x, J);
      const double wt = U->wq[(size_t)q] * std::fabs(invert3(J, Ji));
      for (int i = 0; i < nb; ++i) {
        const double phi = U->tabB[(size_t)q * nb + i]; const double* gh = &U->tabG[((size_t)q * nb + i) * 3];
        double g[3] = {0, 0, 0}; for (int a = 0; a < dim; ++a) for (int d = 0; d < dim; ++d) g[a] += Ji[d][a] * gh[d];
        double v = op->model.c * phi * phi;
        for (int a = 0; a < dim; ++a) v += (op->model.eps * g[a] - op->model.b[a] * phi) * g[a];
        diag[(size_t)U->dofs[(size_t)e * nb + i]] += wt * v;
      }
    }
  if (dirichlet_rows && !op->h_dmask.empty()) for (long long i = 0; i < s->size; ++i) if (op->h_dmask[(size_t)i]) diag[(size_t)i] = 1.0;
  return B200FEM_OK;
}

constexpr int kUnstructuredCtasPerSm = 4;      // resident persistent CTAs per SM (shared memory: <= 34 KB each)

template <int DIM, int NB> static int launch_unst(b200fem_operator* op, const double* u, double* w, bool with_data) {
  using Cfg = UnstructuredCfg<DIM, NB>; b200fem_space* s = op->sp; const UnstructuredSpace* U = s->unst; cudaStream_t st = s->mesh->ctx->stream;
  CUDA_OK(cudaMemsetAsync(w, 0, sizeof(double) * (size_t)s->size, st));                       // w.clear() (galerkin.hh:1463)
  AdrIntegrands I; I.m = op->model; I.dim = DIM; I.with_data = with_data;
  auto kern = lagrange_unstructured_kernel<DIM, NB, AdrIntegrands>;
  int rc = ensure_smem_attr(s->mesh->ctx, (const void*)kern, Cfg::smem_bytes()); if (rc) return rc;
  int launches = 1;
  for (size_t c = 0; c + 1 < U->colour_begin.size(); ++c) {
    const int first = U->colour_begin[c], count = U->colour_begin[c + 1] - first; if (count <= 0) continue;
    const int nbatch = (count + Cfg::EB - 1) / Cfg::EB, grid = std::min(nbatch, kUnstructuredCtasPerSm * s->mesh->ctx->sms);     // persistent CTAs
    kern<<<(unsigned)grid, Cfg::kThreads, Cfg::smem_bytes(), st>>>(U->tab, I, U->d_order, U->d_dofs, U->d_elem_x, u, w, first, count);
    ++launches;
  }
  CUDA_OK(cudaGetLastError());
  op->timing.launches_per_apply = launches;
  return B200FEM_OK;
}
int unstructured_launch_info(const b200fem_space* s, UnstructuredLaunch* out) {
  const UnstructuredSpace* U = s->unst; const int dim = s->mesh->dim, k = s->order;
  out->tab = U->tab; out->order = U->d_order; out->dofs = U->d_dofs; out->elem_x = U->d_elem_x; out->colour_begin = &U->colour_begin;
  auto set = [&](int eb, int threads, size_t smem) { out->eb = eb; out->threads = threads; out->smem = smem; out->max_grid = kUnstructuredCtasPerSm * s->mesh->ctx->sms; };
  if (dim == 2 && k == 1) set(UnstructuredCfg<2, 4>::EB, UnstructuredCfg<2, 4>::kThreads, UnstructuredCfg<2, 4>::smem_bytes());
  else if (dim == 2) set(UnstructuredCfg<2, 9>::EB, UnstructuredCfg<2, 9>::kThreads, UnstructuredCfg<2, 9>::smem_bytes());
  else if (k == 1) set(UnstructuredCfg<3, 8>::EB, UnstructuredCfg<3, 8>::kThreads, UnstructuredCfg<3, 8>::smem_bytes());
  else set(UnstructuredCfg<3, 27>::EB, UnstructuredCfg<3, 27>::kThreads, UnstructuredCfg<3, 27>::smem_bytes());
  return B200FEM_OK;
}
int launch_lagrange_unstructured(b200fem_operator* op, const double* u, double* w, bool with_data) {
  const int dim = op->sp->mesh->dim, k = op->sp->order;
  if (dim == 2) return k == 1 ? launch_unst<2, 4>(op, u, w, with_data) : launch_unst<2, 9>(op, u, w, with_data);
  return k == 1 ? launch_unst<3, 8>(op, u, w, with_data) : launch_unst<3, 27>(op, u, w, with_data);
}

}  // namespace b200fem

using namespace b200fem;

extern "C" int b200fem_mesh_unstructured(b200fem_ctx* ctx, int dim, int64_t n_vertices, const double* coords, int64_t n_elements, const int64_t* elem_vertices, b200fem_mesh** out) {
  REQUIRE(ctx && coords && elem_vertices && out, B200FEM_ERR_INVALID, "mesh_unstructured: null argument");
  REQUIRE(dim == 2 || dim == 3, B200FEM_ERR_NOT_IMPLEMENTED, "mesh: dim must be 2 or 3");
  REQUIRE(n_vertices > 0 && n_elements > 0, B200FEM_ERR_INVALID, "mesh_unstructured: empty mesh");
  REQUIRE(ctx->world == 1, B200FEM_ERR_NOT_IMPLEMENTED, "unstructured meshes: one rank");
  const int nv = 1 << dim;
  for (int64_t i = 0; i < n_elements * nv; ++i) REQUIRE(elem_vertices[i] >= 0 && elem_vertices[i] < n_vertices, B200FEM_ERR_INVALID, "mesh_unstructured: vertex index out of range");
  auto* m = new b200fem_mesh; m->ctx = ctx; m->dim = dim; m->unstructured = true; m->nvert = n_vertices; m->nelem = n_elements;
  m->ux.assign(coords, coords + n_vertices * dim); m->uev.assign(elem_vertices, elem_vertices + n_elements * nv);
  for (int d = 0; d < 3; ++d) { m->gn[d] = 1; m->lo[d] = 0; m->hi[d] = 1; m->h[d] = 1; m->proc[d] = 1; m->pc[d] = 0; m->olo[d] = 0; m->ohi[d] = 1; }
  std::memset(&m->box, 0, sizeof(m->box)); m->box.dim = dim;
  // every element must be positively oriented somewhere sensible: det J > 0 at the centre (the kernels use |det J| like
  // integrationElement, but an inverted element is an input error worth reporting)
  for (int64_t e = 0; e < n_elements; ++e) {
    const double xi[3] = {0.5, 0.5, 0.5}; double x[3], J[3][3], Ji[3][3]; element_map(m, e, xi, x, J);
    if (!(invert3(J, Ji) > 0.0)) { delete m; return fail(B200FEM_ERR_INVALID, "mesh_unstructured: element with non-positive Jacobian determinant (vertex order must follow the cube reference element)"); }
  }
  ctx->refs += 1;
  *out = m; return B200FEM_OK;
}

/* host-only: dof numbering and element colouring of a Lagrange space on an unstructured cube mesh, as b200fem_space_create would
 * build them (no device needed: testable on a CPU-only box).  dofs_out[n_elements][(order+1)^dim], colour_out[n_elements]
 * (colour of every element: elements of a colour share no dof), boundary_out[size] may be NULL; *size_out = number of dofs. */
extern "C" int b200fem_unstructured_numbering(int dim, int64_t n_vertices, const double* coords, int64_t n_elements, const int64_t* elem_vertices, int order,
                                              int64_t* size_out, int32_t* dofs_out, int32_t* colour_out, uint8_t* boundary_out) {
  REQUIRE(coords && elem_vertices && size_out && (dim == 2 || dim == 3) && (order == 1 || order == 2) && n_vertices > 0 && n_elements > 0, B200FEM_ERR_INVALID, "unstructured_numbering: bad argument");
  const int nv = 1 << dim; int nb = 1; for (int d = 0; d < dim; ++d) nb *= order + 1;
  for (int64_t i = 0; i < n_elements * nv; ++i) REQUIRE(elem_vertices[i] >= 0 && elem_vertices[i] < n_vertices, B200FEM_ERR_INVALID, "unstructured_numbering: vertex index out of range");
  b200fem_mesh m; m.ctx = nullptr; m.dim = dim; m.unstructured = true; m.nvert = n_vertices; m.nelem = n_elements;
  m.ux.assign(coords, coords + n_vertices * dim); m.uev.assign(elem_vertices, elem_vertices + n_elements * nv);
  UnstructuredSpace U; long long size = 0;
  int rc = unstructured_number(&m, order, nb, &U, &size); if (rc) return rc;
  *size_out = size;
  if (dofs_out) std::copy(U.dofs.begin(), U.dofs.end(), dofs_out);
  if (boundary_out) std::copy(U.boundary.begin(), U.boundary.end(), boundary_out);
  if (colour_out) for (size_t c = 0; c + 1 < U.colour_begin.size(); ++c) for (int i = U.colour_begin[c]; i < U.colour_begin[c + 1]; ++i) colour_out[U.order[(size_t)i]] = (int32_t)c;
  return B200FEM_OK;
}
