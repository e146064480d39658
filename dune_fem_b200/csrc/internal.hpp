// internal.hpp -- handles behind the C ABI (include/b200fem.h) and the functions the translation units of the library share.
//
//   capi.cu             handles: context, mesh, space, operator life cycle; setters; Dirichlet marks; diagonal
//   apply.cu            GalerkinOperator::evaluate: kernel choice, halo exchange, Dirichlet wrapper, host-pointer pipeline
//   launch_march.cu     DG Q2 z-marching Kronecker kernel (+ fused halo send / receive)
//   launch_slab.cu      DG Q3..Q5 slab Kronecker kernel
//   launch_dg.cu        DG generic quadrature kernel, Q1/Q2 fallback Kronecker kernel
//   launch_lagrange.cu  Lagrange lattice kernel, generic quadrature kernels with colour-ordered scatter
//   solvers.cu          CG / Jacobi-CG / BiCGStab / GMRES drivers, BLAS-1
//   comm.cu             halo plans, peer-memory mailboxes, NCCL fallback, scalar all-reduce
//   jit.cu              run-time compiled (NVRTC) integrands in the generic quadrature kernel
//   newton.cu           NewtonInverseOperator over linearize + Krylov solve
// No CPU compute fallback exists anywhere: every compute entry point needs a CUDA device.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "../../include/b200fem.h"
#include "comm.cuh"
#include "dg_quadrature.cuh"
#include "lagrange_quadrature.cuh"
#include "lagrange_unstructured.cuh"
#include "tables.hpp"
#include "vec_types.hpp"

namespace b200fem {
int fail(int code, const std::string& msg);   // records the calling thread's error message, returns code
}
#define CUDA_OK(expr)                                                                                   \
  do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return ::b200fem::fail(B200FEM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)
#define REQUIRE(cond, code, msg) do { if (!(cond)) return ::b200fem::fail(code, msg); } while (0)

struct b200fem_ctx {
  int device = 0; cudaStream_t stream = nullptr; bool own_stream = false; int sms = 0;
  b200fem::NcclApi nccl; void* comm = nullptr; bool own_comm = false; int rank = 0, world = 1;
  b200fem::PeerScalars scalars;                 // peer-memory all-reduce of scalars (dot products), built with the communicator
  int* h_comm_error = nullptr; int* d_comm_error = nullptr;   // mapped host word: kernels report communication time-outs here
  std::set<const void*> attr_set;               // kernels whose dynamic shared memory limit has been raised ON THIS DEVICE
  int refs = 0; bool released = false;          // handles are reference counted: a parent destroyed before its children (garbage
                                                // collectors finalise in any order) lives on until the last child is gone
};
struct b200fem_mesh {
  b200fem_ctx* ctx; int dim; int gn[3]; double lo[3], hi[3], h[3];
  int proc[3], pc[3];                    // process grid and this rank's coordinates
  b200fem::BoxDev box;                   // local box incl. ghost layers (ghost layers only used by DG spaces)
  int olo[3], ohi[3];                    // owned range in global element coordinates
  // unstructured conforming cube meshes (b200fem_mesh_unstructured): vertex coordinates [nvert][dim] and element -> vertex arrays
  // [nelem][2^dim] in the cube reference element's vertex order; none of the Cartesian fields above is meaningful then
  bool unstructured = false; long long nvert = 0, nelem = 0; std::vector<double> ux; std::vector<long long> uev;
  int refs = 0; bool released = false;
};
namespace b200fem { struct UnstructuredSpace; }
struct b200fem_space {
  b200fem_mesh* mesh; int kind, order, numbering, n1, nb; long long size, elements;
  int dim_range = 1;                     // dimRange: dof blocks of dim_range components (size = blocks * dim_range); > 1: run-time compiled integrands only
  b200fem::BoxDev box;                   // DG: mesh box with ghosts; Lagrange: owned elements only
  b200fem::Tab1D tab; std::vector<int> perm;   // DG: tensor index -> stored local index over the full n1^3 tensor basis (-1: not in the space)
  bool tensor_full = false;              // DG: the space is the whole 3-D tensor basis (the Kronecker kernels apply)
  b200fem::LagrangeLayoutDev lay; long long* d_lattice_map = nullptr; std::vector<long long> lattice_map;
  b200fem::UnstructuredSpace* unst = nullptr;   // Lagrange space on an unstructured mesh: index arrays, geometry, colours (unstructured.cu)
  int refs = 0; bool released = false;
};
struct MarchMapCache;
namespace b200fem { struct JitState; }
struct b200fem_operator {
  b200fem_space* sp; b200fem_model model; int kernel_pref = B200FEM_KERNEL_AUTO; bool communicate = true;
  unsigned q_interior = 0, q_surface = 0; bool inverse_mass = false;
  int* d_perm = nullptr; double* d_bvec = nullptr; uint8_t* d_dmask = nullptr; double* d_dvals = nullptr; uint8_t* d_aux = nullptr;
  std::vector<uint8_t> h_dmask; std::vector<double> h_dvals;
  double *d_u = nullptr, *d_w = nullptr;                       // staging for the host-pointer API
  double *d_h = nullptr, *d_r = nullptr, *d_p = nullptr, *d_x = nullptr, *d_b = nullptr, *d_partial = nullptr, *d_sums = nullptr, *d_hist = nullptr;
  b200fem::CgState* d_cg = nullptr; int hist_cap = 0; unsigned int* d_counter = nullptr;
  bool jac_mode = false; double *d_jac_u = nullptr, *d_jac_opu = nullptr, *d_jac_b = nullptr; b200fem::FdState* d_fd = nullptr;   // AutomaticDifferenceLinearOperator
  double* d_dinv = nullptr; double *d_pq = nullptr, *d_ps = nullptr; bool dinv_mass = false; unsigned long long dinv_version = 0;   // Jacobi preconditioner: 1 / diag(A), PCG work vectors
  std::vector<double*> gmres_v; double* d_gm_partial = nullptr; double* d_gm_sums = nullptr; int gm_cap = 0;   // GMRES basis and reduction scratch
  double *d_rstar = nullptr, *d_s = nullptr, *d_tmp = nullptr, *d_partial5 = nullptr, *d_sums5 = nullptr; b200fem::BicgState* d_bicg = nullptr;   // BiCGStab work vectors
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr; cudaEvent_t pipe_ev[2 * 16 + 2] = {};   // host-pointer apply: copy/compute pipeline
  bool want_dot = false; int dot_parts = 0; double* d_dot_partial = nullptr; int dot_cap = 0;     // <u, A u> fused into the lattice kernel (CG)
  cudaGraphExec_t cg_graph = nullptr; const void* cg_graph_key[3] = {nullptr, nullptr, nullptr}; bool capturing = false;
  unsigned long long state_version = 0;        // bumped by every setter that changes what an apply computes (invalidates cached graphs)
  unsigned long long cg_graph_version = 0;
  bool kron_ready = false; int kron_chk = -1; bool fuse_dirichlet = false, fuse_linear = false, dirichlet_fused = false;
  double* d_lag_rows = nullptr; b200fem::LagKronRows lag_rows{}; std::vector<unsigned char> kron_tab; MarchMapCache* march_cache = nullptr;
  b200fem::HaloPlan halo; b200fem::HaloPlanDG halo_dg; b200fem::HaloPlanP2P halo_p2p; b200fem::HaloPlanAddP2P halo_add;
  const b200fem::BoxDev* active_box = nullptr;   // sub-box override (host-pointer pipeline)
  double *d_nw_res = nullptr, *d_nw_dw = nullptr, *d_nw_w = nullptr, *d_nw_u = nullptr;   // NewtonInverseOperator work vectors (newton.cu)
  b200fem::JitState* jit = nullptr;             // run-time compiled integrands (jit.cu); null: the built-in ADR family
  bool in_bvec = false;                         // the load vector is being computed (data terms on, no recursion into ensure_bvec)
  bool want_exchange = false;                   // apply_dev_impl -> launcher: the Copy exchange of w is due after this apply
  bool exchange_fused = false;                  // launcher -> apply_dev_impl: the kernel did the exchange itself
  int host_pipeline_chunks = 8;                 // host-pointer apply of DG spaces: z-slabs of the copy/compute pipeline (< 2: off)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evx0 = nullptr, evx1 = nullptr; b200fem_timing timing{};
  bool timing_enabled = false;     // event records cost ~1.5 us each on the host: only after b200fem_operator_timing was asked for
};

namespace b200fem {

// raises the dynamic shared memory limit of a kernel once per device (cudaFuncSetAttribute applies to the current device)
int ensure_smem_attr(b200fem_ctx* c, const void* kernel, size_t bytes);
inline double mass_scale(const b200fem_operator* op) {       // 1, or referenceVolume / volume when acting as MOLGalerkinOperator
  if (!op->inverse_mass) return 1.0;
  const BoxDev& b = op->sp->box; double vol = 1; for (int d = 0; d < b.dim; ++d) vol *= b.h[d];
  return 1.0 / vol;
}
inline bool default_quadrature(const b200fem_operator* op) {
  const int k = op->sp->order;
  const int mi = gauss_points_for_order(op->q_interior ? (int)op->q_interior : 2 * k), ms = gauss_points_for_order(op->q_surface ? (int)op->q_surface : 2 * k + 1);
  return mi == k + 1 && ms == k + 1;
}
inline void invalidate_cached_state(b200fem_operator* op) { op->state_version += 1; op->kron_ready = false; op->kron_chk = -1; }

// ---- launchers (one translation unit each; all asynchronous on the context's stream) ----
int launch_dg_march(b200fem_operator* op, const double* u, double* w, const double* bvec, bool fuse_exchange);   // Q2; sets op->exchange_fused
bool dg_march_ok(const b200fem_operator* op, const double* u, const double* w, const double* bvec);
int launch_dg_slab(b200fem_operator* op, const double* u, double* w, const double* bvec);                        // Q3..Q5
int launch_dg_kronecker_v1(b200fem_operator* op, const double* u, double* w, const double* bvec);                // Q1, Q2 fallback
int launch_dg_quadrature_any(b200fem_operator* op, const double* u, double* w, const double* bvec, bool with_data);   // Q1..Q5 generic
int launch_lagrange_quadrature(b200fem_operator* op, const double* u, double* w, bool with_data);
int launch_lagrange_kronecker(b200fem_operator* op, const double* u, double* w, const double* bvec);
void free_march_cache(b200fem_operator* op);

// ---- unstructured.cu: continuous Lagrange spaces on unstructured cube meshes ----
int unstructured_space_setup(b200fem_space* s);                     // numbering, colours, device arrays
void unstructured_space_free(b200fem_space* s);
int unstructured_dofmap(const b200fem_space* s, long long e, int64_t* out);
void unstructured_mark_dirichlet(b200fem_operator* op);             // all nodes on boundary faces
int unstructured_diagonal(b200fem_operator* op, std::vector<double>& diag, bool dirichlet_rows);
int launch_lagrange_unstructured(b200fem_operator* op, const double* u, double* w, bool with_data);
// what a launch of lagrange_unstructured_kernel needs (for the run-time compiled instantiations, jit.cu)
struct UnstructuredLaunch { UnstructuredTabDev tab; const int* order; const int* dofs; const double* elem_x; const std::vector<int>* colour_begin; int eb, threads, max_grid; size_t smem; };
int unstructured_launch_info(const b200fem_space* s, UnstructuredLaunch* out);

// ---- jit.cu ----
int operator_create_impl(b200fem_space* s, const b200fem_model* model, b200fem_operator** out);   // b200fem_operator_create without the scalar-space check
int apply_jit(b200fem_operator* op, const double* u, double* w, bool linear);   // w = L[u] or L[u] - L[0] with compiled integrands
void jit_free(b200fem_operator* op);

// ---- apply.cu ----
int apply_local(b200fem_operator* op, const double* u, double* w, bool linear);
int apply_dev_impl(b200fem_operator* op, const double* u, double* w, bool linear);
int apply_host(b200fem_operator* op, const double* u, double* w, bool linear);
int ensure_bvec(b200fem_operator* op);
int exchange(b200fem_operator* op, double* v, cudaStream_t st);
int check_comm_error(b200fem_ctx* c);

// ---- solvers.cu ----
int ensure_cg_buffers(b200fem_operator* op, int maxit);
int reduce_sums(b200fem_operator* op, int count);
int negate_dev(double* x, long long n, cudaStream_t st);
int dirichlet_sub_dev(const double* u, double* w, const uint8_t* mask, const double* vals, long long n, cudaStream_t st);
int coop_cg_chunk(b200fem_operator* op, double* x, int iters, int* coop_grid_inout);      // cg_coop2d.cuh launch (launch_lagrange.cu)

}  // namespace b200fem
