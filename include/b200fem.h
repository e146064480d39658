/* b200fem.h -- C ABI of the B200-native matrix-free Galerkin operator + CG.
 *
 * This is the drop-in boundary for DUNE-FEM's hot path.  Every entry
 * point names the reference interface it replaces (paths relative to
 * /root/reference/dune/fem).  Conventions: extern "C", opaque handles, caller-owned
 * arrays, return 0 on success and a negative b200fem_status on error (no exceptions
 * cross the ABI; b200fem_last_error() returns the message of the calling thread's last
 * failure).  A handle is not re-entrant (same contract as galerkin.hh:1498-1500).
 * There is no CPU fallback: every compute entry point fails with B200FEM_ERR_CUDA
 * when no CUDA device is usable.
 */
#ifndef B200FEM_H
#define B200FEM_H

#ifdef __CUDACC_RTC__   /* the device headers of the library are also compiled at run time (NVRTC has no host headers) */
typedef signed char int8_t; typedef unsigned char uint8_t; typedef int int32_t; typedef unsigned int uint32_t;
typedef long long int64_t; typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200fem_ctx b200fem_ctx;           /* device + stream + scratch            */
typedef struct b200fem_mesh b200fem_mesh;         /* Cartesian YaspGrid-equivalent (rank-local box of a global grid) */
typedef struct b200fem_space b200fem_space;       /* DiscreteFunctionSpace                */
typedef struct b200fem_operator b200fem_operator; /* GalerkinOperator (+DirichletWrapper) */

enum b200fem_status {
  B200FEM_OK = 0,
  B200FEM_ERR_INVALID = -1,        /* DUNE_THROW(InvalidStateException) analogue */
  B200FEM_ERR_NOT_IMPLEMENTED = -2,/* DUNE_THROW(NotImplemented) analogue        */
  B200FEM_ERR_CUDA = -3,           /* CUDA runtime / no device                   */
  B200FEM_ERR_COMM = -4            /* NCCL / peer access                         */
};

/* space kinds: python/dune/fem/space/_spaces.py:106 (lagrange), :183-229 (dglegendre, hierarchical flag), `dgonb`
 * (orthonormal P_k, space/shapefunctionset/orthonormal.hh:55-60 -- what pydemo/advectiondiffusion.py:9 imports).  The DG
 * spaces exist on 2-D and 3-D boxes; Q_k Legendre orders 1..5, dgonb orders 1..4 (PMAX3D of orthonormalbase_3d.hh). */
enum b200fem_space_kind { B200FEM_LAGRANGE = 0, B200FEM_DG_LEGENDRE = 1, B200FEM_DG_LEGENDRE_HIER = 2, B200FEM_DG_ONB = 3 };
/* sub-entity numbering of Lagrange dofs: YaspGrid leaf index set or AdaptiveLeafIndexSet first-touch order
 * (gridpart/adaptiveleafindexset.hh:884-906) */
enum b200fem_numbering { B200FEM_NUMBERING_YASP = 0, B200FEM_NUMBERING_ADAPTIVE_LEAF = 1 };
/* which device kernel evaluates the operator */
enum b200fem_kernel {
  B200FEM_KERNEL_AUTO = 0,       /* fastest kernel that is valid for the model                              */
  B200FEM_KERNEL_QUADRATURE = 1, /* generic: gather, basis evaluation at quadrature points, integrand, axpy */
  B200FEM_KERNEL_KRONECKER = 2,  /* linear constant-coefficient models on uniform boxes: 1-D operator form  */
  B200FEM_KERNEL_KRONECKER_TILE = 3 /* same form through the plain tile kernel (DG Q1/Q2; the fallback of the TMA marching
                                       kernel, selectable for A/B checks)                                        */
};
/* solver/parameter.hh:21-295 "fem.solver.errormeasure" */
enum b200fem_tolerance { B200FEM_TOL_ABSOLUTE = 0, B200FEM_TOL_RELATIVE = 1, B200FEM_TOL_RESIDUAL_REDUCTION = 2 };

/* Integrands (schemes/integrands.hh:152-375): the advection-diffusion-reaction family of the
 * BASELINE configs, i.e. the UFL form of pydemo/advectiondiffusion.py:33-60 plus c*u + gamma*u^3:
 *   interior  s = c u + gamma u^3 - f,   F = eps grad u - b u
 *   skeleton  eps beta/he [u][v] - eps {grad u}.n [v] - eps [u]{grad v}.n + [hatb u][v]
 *   boundary  -eps grad g.n v + dD (eps beta/hbnd (u-g) + hatb u + (b.n-hatb) g) v
 * data: 0 g=f=0, 1 g=sin(x0 x1), 2 g=prod sin(pi x_k); f is chosen so that g solves the PDE. */
typedef struct b200fem_model {
  double eps, b[3], c, gamma, beta;
  int32_t dirichlet_mask;   /* bit 2*axis+side: side carries Dirichlet data (weak for DG, strong for Lagrange) */
  int32_t data;
  int32_t has_skeleton, has_boundary;
  int32_t strong_dirichlet; /* wrap in DirichletWrapperOperator (schemes/dirichletwrapper.hh:101-105) */
} b200fem_model;

typedef struct b200fem_timing {
  double last_apply_ms;     /* device time of the last apply (CUDA events on the operator's stream) */
  double last_exchange_ms;  /* halo exchange share, cf. exchangeTime() space/common/communicationmanager.hh:203-209 */
  int64_t applies;          /* cf. gridSizeInterior()/call counters galerkin.hh:822,1479 */
  int32_t kernel;           /* b200fem_kernel actually used */
  int32_t launches_per_apply;
} b200fem_timing;

#ifndef __CUDACC_RTC__   /* (the run-time compiled device code only needs the types above) */
const char* b200fem_last_error(void);
int b200fem_version(void);

/* MPIManager / device selection (misc/mpimanager.hh:352-461).  `stream` may be NULL (own stream) or a cudaStream_t. */
int b200fem_device_count(int* count);     /* visible CUDA devices (0 and B200FEM_ERR_CUDA when there is none) */
int b200fem_ctx_create(int device, void* stream, b200fem_ctx** out);
int b200fem_ctx_destroy(b200fem_ctx* ctx);
int b200fem_ctx_synchronize(b200fem_ctx* ctx);
/* plain device buffers for callers without their own allocator */
int b200fem_malloc(b200fem_ctx* ctx, int64_t bytes, void** dev);
int b200fem_free(b200fem_ctx* ctx, void* dev);
int b200fem_memcpy_h2d(b200fem_ctx* ctx, void* dev, const void* host, int64_t bytes);
int b200fem_memcpy_d2h(b200fem_ctx* ctx, void* host, const void* dev, int64_t bytes);

/* YaspGrid< dim > on [lo,hi] with n cells per axis (replaces the GridPart handed to the space,
 * gridpart/common/gridpart.hh).  The distributed variant describes this rank's box of a global grid that is
 * split into proc[0] x proc[1] x proc[2] boxes (rank = c0 + proc0*(c1 + proc1*c2)); DG spaces then carry one
 * layer of ghost elements (overlap 1). */
int b200fem_mesh_cartesian(b200fem_ctx* ctx, int dim, const int32_t* n, const double* lo, const double* hi, b200fem_mesh** out);
/* Unstructured conforming cube meshes -- what ALUGrid< dim, dim, cube, conforming > behind an AdaptiveLeafGridPart hands to the
 * reference: coords[n_vertices][dim], elem_vertices[n_elements][2^dim] in the cube reference element's vertex order (vertex v:
 * bit d set <=> xi_d = 1; Jacobian determinant positive).  Continuous Lagrange spaces of order 1 and 2 on one rank:
 * the dof mapper becomes an index array (space/lagrange/dofmappercode.hh:56-104 -> space/mapper/indexsetdofmapper.hh:414-427;
 * dof blocks by geometry type, first-touch order inside a block, gridpart/adaptiveleafindexset.hh:884-906), the geometry the
 * element's multilinear map (integrationElement / jacobianInverseTransposed per quadrature point), the scatter colour-ordered.
 * Operators: interior integrands of the built-in family + strong Dirichlet constraints on the WHOLE boundary
 * (strong_dirichlet = 1, any dirichlet_mask), or run-time compiled interior() integrands (b200fem_operator_create_jit with
 * has_skeleton = has_boundary = 0); no skeleton / boundary terms, no Kronecker kernel.  Krylov solvers, the matrix-free diagonal
 * and Newton work as on Cartesian meshes. */
int b200fem_mesh_unstructured(b200fem_ctx* ctx, int dim, int64_t n_vertices, const double* coords, int64_t n_elements,
                              const int64_t* elem_vertices, b200fem_mesh** out);
/* host-only (no device): the dof numbering (dofs_out[n_elements][(order+1)^dim]), the element colouring of the colour-ordered scatter
 * (colour_out[n_elements]: elements of one colour share no dof) and the boundary marks (boundary_out[*size_out]) such a space gets;
 * any of the three outputs may be NULL. */
int b200fem_unstructured_numbering(int dim, int64_t n_vertices, const double* coords, int64_t n_elements, const int64_t* elem_vertices,
                                   int order, int64_t* size_out, int32_t* dofs_out, int32_t* colour_out, uint8_t* boundary_out);
int b200fem_mesh_cartesian_distributed(b200fem_ctx* ctx, int dim, const int32_t* n_global, const double* lo, const double* hi,
                                       const int32_t* proc, int rank, b200fem_mesh** out);
/* Periodic grid (YaspGrid's `periodic` bitset): bit d set = the faces on the two sides of axis d are interior faces whose
 * neighbour is the element on the far side; the operator treats them as skeleton faces first, as the reference does
 * (schemes/galerkin.hh:859-861).  DG spaces on one rank, evaluated by the generic quadrature kernel (Kronecker kernels step aside). */
int b200fem_mesh_set_periodic(b200fem_mesh* mesh, int mask);
int b200fem_mesh_destroy(b200fem_mesh* mesh);
/* The block partition used by the distributed mesh, without needing a device (host logic only): for `rank` of the
 * process grid, out[0..2] = global coordinates of local element (0,0,0) (ghost layer included), out[3..5] = local
 * extents, out[6..8] / out[9..11] = owned range [lo,hi) in local coordinates.  overlap = 1 for DG spaces, 0 for Lagrange.
 * Cells are dealt in blocks; the first (n % p) ranks along an axis hold one extra cell (YaspGrid's default load
 * balancer lives in dune-grid and is not restated; the process grid is explicit instead). */
int b200fem_partition_box(int dim, const int32_t* n_global, const int32_t* proc, int rank, int overlap, int32_t* out12);
int b200fem_mesh_local_box(b200fem_mesh* mesh, int overlap, int32_t* out12);
/* Host logic, no device needed: the static work schedule of the z-marching DG Q2 kernel for an owned box of on[0..2] elements on
 * `grid` CTAs.  flags: bit 0 / 1 = a rank interface below / above in z (their planes become single-plane runs that go first and are
 * flushed to the neighbour at once), bits 2..5 = the box touches the domain boundary at x-low, x-high, y-low, y-high, bits 6 / 7 = a rank interface
 * below / above in y (cost model).
 * runs_out (may be NULL; cap = runs that fit) receives (column, za, zb, flush) per run, begin_out[0..grid] the slice of each CTA. */
int b200fem_march_schedule(const int32_t* on, int grid, int flags, int32_t* runs_out, int32_t cap, int32_t* begin_out, int32_t* nruns_out);

/* DiscreteFunctionSpace (space/lagrange/space.hh:129-353, space/discontinuousgalerkin/legendre.hh).  Lagrange: orders 1..3 on
 * Cartesian meshes (order 3: YaspGrid numbering, quadrature kernels; several nodes inside an edge / face / cell are
 * numbered entity by entity, lower axes fastest, space/lagrange/genericlagrangepoints.hh:862-876), orders 1..2 on unstructured
 * meshes; DG: Legendre orders 1..5, dgonb orders 1..4 on Cartesian meshes. */
int b200fem_space_create(b200fem_mesh* mesh, int kind, int order, int numbering, b200fem_space** out);
/* Vector-valued spaces (FunctionSpace< ..., dimRange >; the reference builds them from the scalar shape functions,
 * space/shapefunctionset/vectorial.hh:508-526: local dof i * dimRange + c; DofVector blocks of dimRange components,
 * function/blockvectors/defaultblockvectors.hh:284-294): b200fem_space_size = blocks * dim_range, dof (block g, component c) =
 * g * dim_range + c; b200fem_space_dofmap / _local_size keep returning BLOCK indices (blockMapper()).  dim_range 1..4, orders 1..3;
 * also on distributed meshes (the Copy exchange of DG spaces moves element blocks of n_b * dim_range doubles, the Add exchange of
 * Lagrange spaces sums blocks of dim_range components on shared nodes); operators on them take run-time compiled integrands (b200fem_operator_create_jit) -- the built-in family is scalar. */
int b200fem_space_create_vector(b200fem_mesh* mesh, int kind, int order, int numbering, int dim_range, b200fem_space** out);
int b200fem_space_dim_range(b200fem_space* space, int32_t* dim_range);
int b200fem_space_destroy(b200fem_space* space);
int b200fem_space_size(b200fem_space* space, int64_t* size);            /* space.size()                               */
int b200fem_space_local_size(b200fem_space* space, int32_t* nb);        /* blockMapper().maxNumDofs()                 */
int b200fem_space_elements(b200fem_space* space, int64_t* n);
/* blockMapper().map(entity, out) (space/mapper/indexsetdofmapper.hh:414-427, codimensionmapper.hh:121-131) */
int b200fem_space_dofmap(b200fem_space* space, int64_t element, int64_t* out);

/* GalerkinOperator(dSpace, rSpace, integrands) (schemes/galerkin.hh:1383-1504) */
int b200fem_operator_create(b200fem_space* space, const b200fem_model* model, b200fem_operator** out);
int b200fem_operator_destroy(b200fem_operator* op);
/* operator()(u, w) (galerkin.hh:1430-1433; operator/common/operator.hh:55): w = L[u], host dof vectors in the
 * reference layout (function/blockvectors/defaultblockvectors.hh:284-294). Copies u to the device, runs the
 * kernels, copies w back. */
int b200fem_operator_apply(b200fem_operator* op, const double* u_host, double* w_host);
/* the homogeneous linear part A u = L[u] - L[0] (what a Krylov solver applies; SURVEY.md 8a row I') */
int b200fem_operator_apply_linear(b200fem_operator* op, const double* u_host, double* w_host);
/* device-pointer variant; linear != 0 selects A u.  Asynchronous on the context's stream. */
int b200fem_operator_apply_dev(b200fem_operator* op, const double* u_dev, double* w_dev, int linear);
/* b = -L[0] (python/dune/fem/operator/__init__.py:347-350) */
int b200fem_operator_load_vector(b200fem_operator* op, double* b_host);
int b200fem_operator_set_communicate(b200fem_operator* op, int communicate);                    /* galerkin.hh:1409 */
int b200fem_operator_set_quadrature_orders(b200fem_operator* op, unsigned interior, unsigned surface); /* :1418-1423 */
int b200fem_operator_set_kernel(b200fem_operator* op, int kernel);
/* host-pointer apply of DG spaces: number of z-slabs of the H2D / compute / D2H pipeline (default 8; 0 or 1: one copy each way) */
int b200fem_operator_set_host_pipeline(b200fem_operator* op, int chunks);
/* MOLGalerkinOperator (schemes/molgalerkin.hh:100-124, 162-197; python molGalerkin, operator/__init__.py:203-207): apply the
 * inverse of the local mass matrix after the evaluate, w = M^-1 L[u] -- the form explicit time stepping uses.  For the
 * orthonormal Legendre bases on affine cells this is the scalar referenceVolume / volume per element
 * (operator/1order/localmassmatrix.hh:304-311, 421-434), fused into the kernels (scaled 1-D operators and load vector).
 * DG spaces only (B200FEM_ERR_NOT_IMPLEMENTED otherwise). */
int b200fem_operator_set_inverse_mass(b200fem_operator* op, int on);
/* AutomaticDifferenceOperator::jacobian / AutomaticDifferenceLinearOperator (operator/common/automaticdifferenceoperator.hh:58-166):
 * Jacobian-free linearisation J(u) v = (L[u + eps v] - L[u]) / eps with the reference's dynamic eps = sqrt((1 + |u|) macheps / |v|^2)
 * when eps <= 0 ("fem.differenceoperator.eps").  linearize(u) is jOp.set(u, op, eps): it stores u and L[u] on the device; while a
 * linearisation is set, apply_linear / apply_dev(linear != 0) and the Krylov solvers act on J(u) instead of on the homogeneous
 * linear part, i.e. a Newton step is  linearize(u);  solve J(u) delta = -L[u];  u += delta.  u_host == NULL drops the
 * linearisation.  Works for every model, including the non-linear ones (gamma != 0) that have no Kronecker form. */
int b200fem_operator_linearize(b200fem_operator* op, const double* u_host, double eps);
int b200fem_operator_linearize_dev(b200fem_operator* op, const double* u_dev, double eps);
/* Generic integrands (schemes/integrands.hh:152-375).  The reference JIT-compiles a C++ Integrands class generated from the UFL
 * form (python/dune/models/integrands/model.py:10-106, python/dune/ufl/codegen.py) into its GalerkinOperator; this entry point
 * does the same for the device: `source` is CUDA C++ that defines the three integrand functions below, and it is compiled at
 * run time (NVRTC, sm_100a) INTO the generic quadrature kernel (dg_quadrature.cuh) -- one specialised kernel per operator,
 * exactly like the reference's one specialised operator class per form.  DG spaces (all kinds) and continuous Lagrange
 * spaces (no skeleton terms there; colour-ordered scatter, lagrange_quadrature.cuh), scalar or vector-valued range.
 *
 *   struct PointValue { double u; double du[3]; };   // DomainValueType  = (u, grad u) at the quadrature point
 *   struct PointRange { double s; double F[3]; };    // RangeValueType: tested as  s * phi + F . grad phi
 *   __device__ void interior(const double* x, const PointValue& u, PointRange& r, const double* c, int dim);
 *   __device__ void skeleton(const double* x, int axis, double sign, double ihe, const PointValue& in, const PointValue& out,
 *                            PointRange& rin, PointRange& rout, const double* c, int dim);
 *   __device__ void boundary(const double* x, int axis, int side, double ihbnd, const PointValue& u, PointRange& r,
 *                            const double* c, int dim);
 *
 * x: physical coordinates of the point; c: the `nconstants` (<= 32) values passed here (the reference's dune.ufl.Constant
 * coefficients; b200fem_operator_set_constants changes them without recompiling); sign * e_axis is the unit outer normal of
 * `in` (of the element, for boundary(): sign = side ? +1 : -1); ihe = 1 / he with he = avg(CellVolume) / FacetArea,
 * ihbnd = FacetArea / CellVolume.  r / rin / rout arrive zeroed.  The functions evaluate the WHOLE integrand, data terms
 * included: apply = L[u]; apply_linear = L[u] - L[0] (meaningful for integrands that are affine in u; non-linear ones are
 * solved through b200fem_operator_linearize).  Compile errors: B200FEM_ERR_INVALID with the NVRTC log as the message.
 *
 * Vector-valued spaces (b200fem_space_create_vector, dimRange = R > 1): the same three functions over
 *   template <int R> struct PointValueV { double u[R]; double du[R][3]; };   // (u_c, grad u_c)
 *   template <int R> struct PointRangeV { double s[R]; double F[R][3]; };    // tested as  s_c * phi + F_c . grad phi  per component
 * with `dimRange`, `VectorValue = PointValueV<dimRange>` and `VectorRange = PointRangeV<dimRange>` defined for the source, e.g.
 *   __device__ void interior(const double* x, const VectorValue& u, VectorRange& r, const double* c, int dim);
 * (the reference's own matrix-free check runs on such a space: dune/fempy/test/testoperator.py:14-36, Lagrange order 2, dimRange 2). */
int b200fem_operator_create_jit(b200fem_space* space, const char* source, const double* constants, int nconstants,
                                int has_skeleton, int has_boundary, b200fem_operator** out);
int b200fem_operator_set_constants(b200fem_operator* op, const double* constants, int nconstants);
/* Compiles `source` for a Q_order / dgonb space with the default quadrature orders without touching a device (NVRTC only):
 * 0 when it compiles; the log (NUL-terminated, truncated to log_len) is returned either way.  Host logic, testable on CPU. */
int b200fem_jit_compile_check(const char* source, int order, char* log, int log_len);
/* ... and for any space the integrands can run on: space kind, mesh dimension (2 / 3), order, dimRange */
int b200fem_jit_compile_check_space(const char* source, int kind, int dim, int order, int dim_range, int has_skeleton, int has_boundary,
                                    char* log, int log_len);
/* ... and for a Lagrange space on an unstructured cube mesh (interior() only; lagrange_unstructured.cuh) */
int b200fem_jit_compile_check_unstructured(const char* source, int dim, int order, char* log, int log_len);

/* strong Dirichlet marks and values (schemes/dirichletconstraints.hh:435-554) */
int b200fem_operator_dirichlet(b200fem_operator* op, uint8_t* mask_host, double* values_host);
int b200fem_operator_timing(b200fem_operator* op, b200fem_timing* out);

/* CgInverseOperator / KrylovInverseOperator<cg> (solver/krylovinverseoperators.hh:118-208 ->
 * solver/linear/cg.hh:18-117) on the homogeneous linear part.  x holds the initial guess on entry.
 * *iterations is negative if not converged (cg.hh:116).  history (may be NULL, maxit entries) receives
 * sqrt(residual) per iteration, the value the reference prints as "Fem::CG it: i : residual r". */
int b200fem_cg_solve(b200fem_operator* op, const double* b_host, double* x_host, double epsilon, int max_iterations,
                     int tolerance_criteria, int* iterations, double* history);
int b200fem_cg_solve_dev(b200fem_operator* op, const double* b_dev, double* x_dev, double epsilon, int max_iterations,
                         int tolerance_criteria, int* iterations, double* history);

/* Diagonal of the homogeneous linear part A (what DiagonalPreconditioner extracts from an ASSEMBLED operator,
 * solver/diagonalpreconditioner.hh:104-141; the reference throws NotImplemented for matrix-free operators, :37-57).  Available
 * when the operator has a Kronecker form (linear model, default quadrature): the diagonal is then a sum of products of 1-D
 * diagonal entries and costs O(N).  Strong-Dirichlet rows hold 1. */
int b200fem_operator_diagonal(b200fem_operator* op, double* diag_host);
/* CG with the Jacobi preconditioner B = diag(A)^-1: the preconditioned branch of LinearSolver::cg (solver/linear/cg.hh:52-56,
 * 72-107; residual = <r, B r>).  history receives sqrt(<r, B r>) per iteration. */
int b200fem_pcg_solve(b200fem_operator* op, const double* b_host, double* x_host, double epsilon, int max_iterations,
                      int tolerance_criteria, int* iterations, double* history);
int b200fem_pcg_solve_dev(b200fem_operator* op, const double* b_dev, double* x_dev, double epsilon, int max_iterations,
                          int tolerance_criteria, int* iterations, double* history);

/* KrylovInverseOperator<bicgstab> (solver/krylovinverseoperators.hh:46-281 -> solver/linear/bicgstab.hh:64-214, unpreconditioned)
 * on the homogeneous linear part -- the Krylov method of pydemo/advectiondiffusion.py (non-symmetric operators).  Same
 * conventions as the reference: no convergence test before the first iteration; `res` is compared with
 * epsilon * {1 | sqrt(b.b) | sqrt(r0.r0)}; *iterations is negative when max_iterations was reached (:208-211); history
 * receives `res` per iteration ("Fem::BiCGstab it: i : res"). */
int b200fem_bicgstab_solve(b200fem_operator* op, const double* b_host, double* x_host, double epsilon, int max_iterations,
                           int tolerance_criteria, int* iterations, double* history);
int b200fem_bicgstab_solve_dev(b200fem_operator* op, const double* b_dev, double* x_dev, double epsilon, int max_iterations,
                               int tolerance_criteria, int* iterations, double* history);

/* KrylovInverseOperator<gmres> (solver/krylovinverseoperators.hh:142-160 -> solver/linear/gmres.hh:117-301, unpreconditioned):
 * restarted GMRES(restart) ("fem.solver.gmres.restart", default 20, solver/parameter.hh:197-201) with the reference's classical
 * Gram-Schmidt sweep and Givens rotations; the basis vectors and all vector work stay on the device, the (restart+1) x restart
 * Hessenberg system on the host.  history receives |g[j+1]| per iteration ("Fem::GMRES it: i : ...").  *iterations is negative when
 * max_iterations was reached. */
int b200fem_gmres_solve(b200fem_operator* op, const double* b_host, double* x_host, int restart, double epsilon, int max_iterations,
                        int tolerance_criteria, int* iterations, double* history);
int b200fem_gmres_solve_dev(b200fem_operator* op, const double* b_dev, double* x_dev, int restart, double epsilon, int max_iterations,
                            int tolerance_criteria, int* iterations, double* history);

/* NewtonInverseOperator (solver/newtoninverseoperator.hh:690-803) -- the caller directly above the Krylov loop (FemScheme::solve):
 * solves L[w] = u (u == NULL: L[w] = 0) from the initial guess in w.  Jacobian = the difference quotient of
 * AutomaticDifferenceLinearOperator (b200fem_operator_linearize), linear solves by cg (0) / bicgstab (1) / gmres (2) with
 * "fem.solver.nonlinear.linear.*" tolerance, criterion and iteration budget (the budget is shared by all Newton steps, :745-757),
 * "fem.solver.nonlinear.{tolerance, maxiterations}", line search "none" (0) or "simple" (1, :588-629).  Linear operators take one step
 * (:761, 791-792).  *failure receives NewtonFailure (:389-400: 0 Success, 1 InvalidResidual, 4 LineSearchFailed, 5 TooManyIterations,
 * 6 TooManyLinearIterations, 7 LinearSolverFailed); *residual_norm = |L[w] - u| of the last iterate.  Everything runs on the device. */
int b200fem_newton_solve(b200fem_operator* op, const double* u_host, double* w_host, double tolerance, int max_iterations, int linear_method,
                         double linear_tolerance, int linear_max_iterations, int linear_tolerance_criteria, int gmres_restart, int line_search,
                         int* iterations, int* linear_iterations, double* residual_norm, int* failure);
int b200fem_newton_solve_dev(b200fem_operator* op, const double* u_dev, double* w_dev, double tolerance, int max_iterations, int linear_method,
                             double linear_tolerance, int linear_max_iterations, int linear_tolerance_criteria, int gmres_restart, int line_search,
                             int* iterations, int* linear_iterations, double* residual_norm, int* failure);

/* BLAS-1 on device dof vectors (function/blockvectors/defaultblockvectors.hh:39-150) and the dot product over
 * primary dofs followed by the global sum (function/common/scalarproducts.hh:115-127) */
int b200fem_dot_dev(b200fem_operator* op, const double* x_dev, const double* y_dev, double* result);
int b200fem_axpy_dev(b200fem_operator* op, double alpha, const double* x_dev, double* y_dev);

/* Multi-GPU: one process per GPU.  `nccl_comm` is an ncclComm_t created by the caller (e.g. from an id broadcast
 * through torch.distributed); halo exchange replaces DiscreteFunction::communicate
 * (function/common/discretefunction.hh:825-835, space/common/communicationmanager.hh:130-150): Copy for DG,
 * Add for Lagrange.  Default transport: peer-mapped mailboxes written directly over NVLink by the kernels (the DG marching
 * kernel sends and receives inside the apply kernel itself; scalar products are summed inside the reduction kernels), so that
 * Krylov iterations contain no library call and are replayed as CUDA graphs on every rank; ncclSend/Recv + ncclAllReduce when
 * peer mappings are unavailable (or B200FEM_NO_P2P is set when the communicator is attached).  A peer that does not arrive
 * within 20 s makes the next entry point return B200FEM_ERR_COMM instead of hanging. */
int b200fem_ctx_set_nccl(b200fem_ctx* ctx, void* nccl_comm, int rank, int world);
int b200fem_nccl_unique_id(void* out128);
int b200fem_nccl_init(b200fem_ctx* ctx, const void* id128, int rank, int world);
/* *peer_memory = 1 when halo exchange and scalar sums run over peer-mapped mailboxes (cudaIpc, NVLink), 0 when they use
 * ncclSend/ncclRecv/ncclAllReduce (the choice is collective: all ranks agree) */
int b200fem_ctx_transport(b200fem_ctx* ctx, int* peer_memory);
/* exchange the ghost layer of a device dof vector of `space` in place */
int b200fem_communicate_dev(b200fem_operator* op, double* v_dev);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* B200FEM_H */
