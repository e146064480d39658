// dg_quadrature.cuh -- generic matrix-free operator apply by quadrature on 3-D cubes (Legendre Q_k; the element kernel is
// shared with the continuous Lagrange spaces, lagrange_quadrature.cuh).
//
// One kernel does what GalerkinOperator::evaluate does per element (dune/fem/schemes/galerkin.hh:811-917):
//   gather u_K (getLocalDofs, function/common/discretefunction.hh:945-950)
//   interior integral  (galerkin.hh:332-360): evaluateAll/jacobianAll at the quadrature points
//                      (space/basisfunctionset/default.hh:276-372), integrand, axpy (default.hh:199-257)
//   skeleton/boundary integrals (galerkin.hh:414-435, 475-537) on the six faces
//   write w_K (addLocalDofs, discretefunction.hh:929-934)
// for ANY pointwise integrand (non-linear, variable coefficients) and ANY Gauss rule: the interior rule has MI and the
// surface rule MS points per axis (setQuadratureOrders, galerkin.hh:1418-1423; defaults MI = MS = k+1).
// B200-first differences from the reference loop:
//  * the dense tabulated contraction u_q = sum_i B[q][i] u_i (space/shapefunctionset/caching.hh:302-319) is
//    sum-factorised: B = B1 (x) B1 (x) B1 with the 1-D tables held in the constant bank;
//  * every element integrates all six of its faces and keeps only its own half of the skeleton integrand
//    (the reference integrates a face once, from the lower-index element, and scatters into both elements,
//    galerkin.hh:879-897).  Roles (inside = lower element index, or the owned element next to a ghost) follow
//    the reference, so the integrand sees identical arguments; the result is race-free without atomics or
//    colouring and w is written exactly once (the w.clear() of galerkin.hh:1463 is fused away);
//  * a CTA owns EB elements, P x P threads per element (P = max(k+1, MI, MS)), one thread per tensor line; intermediate
//    tensors live in shared memory with odd line strides (no bank conflicts for any k).
// Second generation (the first one executed ten instructions per FMA: 376-byte stack frames from face bookkeeping arrays
// and a gradient indexed by a run-time axis, 93 double divisions per thread, 370 KB of code): no run-time indexed
// register arrays, reciprocal cell sizes, the six faces share one code path whose axis enters through selects.
#pragma once
#ifndef __CUDACC_RTC__       // (also compiled at run time by NVRTC for user-supplied integrands, jit.cu: no host headers there)
#include <cuda_runtime.h>
#include <cstdint>
#endif
#include "integrands.cuh"

namespace b200fem {

// rank-local box of the global Cartesian grid (with ghost layers where a neighbouring rank exists)
struct BoxDev {
  int dim;
  int n[3];                 // local box extents in elements (owned + ghost)
  int own_lo[3], own_hi[3]; // owned sub-box, local coordinates
  int origin[3];            // global element coordinates of local element (0,0,0)
  int gn[3];                // global extents
  double lo[3], h[3], ih[3];   // ih = 1 / h (filled by the host: no division in the kernels)
  int periodic;             // bit d: periodic along axis d -- the faces on those sides have the element of the far side as neighbour
                            // (single rank; only the generic quadrature kernel wraps)
};

// 1-D tables: N basis functions tabulated at the MI-point interior rule and at the MS-point surface rule, traces at 0 / 1
template <int N, int MI, int MS>
struct QuadTabDev {
  double Bi[MI * N], Gi[MI * N], xi[MI], wi[MI];   // B[q*N+i] = phi_i(x_q), G = phi_i'(x_q); Gauss points / weights on [0,1]
  double Bs[MS * N], Gs[MS * N], xs[MS], ws[MS];
  double phi[2][N], dphi[2][N];
};

__host__ __device__ constexpr int quad_odd(int x) { return x | 1; }
__host__ __device__ constexpr int quad_max(int a, int b) { return a > b ? a : b; }

template <int N, int MI, int MS> struct DgQuadCfg {
  static constexpr int P = quad_max(N, quad_max(MI, MS)), T2 = P * P;          // threads per element
  static constexpr int LN = quad_odd(N), LM = quad_odd(MI), LS = quad_odd(MS); // padded line lengths
  static constexpr int kU = N * N * LN;                                        // dof tensor [i0][i1][i2 (padded)]
  static constexpr int kVol = 2 * N * N * LM + 3 * N * MI * LM;                // T1a, T1b; T2a, T2b, T2c
  static constexpr int kRegA = quad_max(24 * N * LN, 18 * N * LS);             // face coefficients C, later X
  static constexpr int kRegB = quad_max(36 * N * LS, 12 * N * LN);             // E, later R
  // stored-order dofs of the six face neighbours, fetched (coalesced) with the element's own dofs at kernel start and read by
  // the trace stage F0; they sit behind the volume scratch and the C region, inside or behind E (not yet in use at F0)
  static constexpr int kNb = quad_max(kVol, kRegA);
  static constexpr int kScratch = quad_max(kRegA + kRegB, kNb + 6 * N * N * N);
  static constexpr int kElemDoubles = quad_odd(2 * kU + kScratch);
  // Elements per CTA and the thread map.  The element slot is the FAST index of the thread id (slot = tid % EB, lane in the
  // element = tid / EB): neighbouring lanes of a warp do the same job on different elements, so every shared-memory access of
  // a (half-)warp is "same offset, element stride" -- conflict-free whatever the access pattern, because the element stride
  // is odd (kElemDoubles); and the job loops (trip count depends on the lane in the element) do not diverge inside a warp.
  // EB = 16 fills a half-warp; smaller powers of two where 16 elements do not fit ~100 KB / 640 threads.
  __host__ __device__ static constexpr int eb_fit() { int eb = 16; while (eb > 1 && (eb * T2 > 640 || (size_t)eb * kElemDoubles * 8 > 100 * 1024)) eb /= 2; return eb; }
  static constexpr int EB = eb_fit();
  static constexpr int kThreads = (EB * T2 + 31) / 32 * 32;
  __device__ static int slot(int tid) { return tid % EB; }
  __device__ static int lane(int tid) { return tid / EB; }     // >= T2 for the padding threads of the last warp
  __host__ __device__ static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)EB * kElemDoubles + sizeof(int) * N * N * N * 2 + sizeof(long long) * EB + sizeof(int) * 4 * EB; }
};

template <int N> __device__ __forceinline__ void quad_zero(double (&a)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) a[i] = 0;
}
// out[q] = sum_c M[q*N+c] in[c]   (M tabulated at Q points)
template <int Q, int N> __device__ __forceinline__ void quad_mv(const double* __restrict__ M, const double (&in)[N], double (&out)[Q]) {
#pragma unroll
  for (int q = 0; q < Q; ++q) { double s = 0;
#pragma unroll
    for (int c = 0; c < N; ++c) s = fma(M[q * N + c], in[c], s); out[q] = s; }
}
// out[c] += sum_q M[q*N+c] in[q]
template <int Q, int N> __device__ __forceinline__ void quad_mvt_add(const double* __restrict__ M, const double (&in)[Q], double (&out)[N]) {
#pragma unroll
  for (int c = 0; c < N; ++c) { double s = out[c];
#pragma unroll
    for (int q = 0; q < Q; ++q) s = fma(M[q * N + c], in[q], s); out[c] = s; }
}

// ---- vector-valued spaces (dimRange R > 1; DofVector blocks of R components, function/blockvectors/defaultblockvectors.hh:284-294,
// vectorial basis phi_i e_c with local index i * R + c, space/shapefunctionset/vectorial.hh:508-526).  The sum-factorised sweeps act on one component at a
// time, so the R components of an element occupy R neighbouring element SLOTS (= neighbouring lanes of a warp) and run the scalar
// code unchanged; only the integrand needs all components of (u, grad u) at a point: the lanes collect them with shuffles,
// every lane evaluates the integrand and keeps its own component of the result. ----
template <int R> struct PointValueV { double u[R]; double du[R][3]; };   // DomainValueType of a range-R space
template <int R> struct PointRangeV { double s[R]; double F[R][3]; };    // RangeValueType: tested as s_c phi + F_c . grad phi per component
__host__ __device__ constexpr int quad_pow2(int r) { int p = 1; while (p < r) p *= 2; return p; }
#ifdef __CUDACC__
template <int R> __device__ __forceinline__ PointValueV<R> quad_collect(const PointValue& pv, const unsigned mask) {
  constexpr int RS = quad_pow2(R); PointValueV<R> v;
#pragma unroll
  for (int c = 0; c < R; ++c) {
    v.u[c] = __shfl_sync(mask, pv.u, c, RS);
#pragma unroll
    for (int d = 0; d < 3; ++d) v.du[c][d] = __shfl_sync(mask, pv.du[d], c, RS);
  }
  return v;
}
template <int R> __device__ __forceinline__ PointRange quad_pick(const PointRangeV<R>& r, const int comp) {
  PointRange o; o.s = r.s[0]; o.F[0] = r.F[0][0]; o.F[1] = r.F[0][1]; o.F[2] = r.F[0][2];
#pragma unroll
  for (int c = 1; c < R; ++c) if (comp == c) { o.s = r.s[c]; o.F[0] = r.F[c][0]; o.F[1] = r.F[c][1]; o.F[2] = r.F[c][2]; }
  return o;
}
#endif

// All integrals of one element: U (tensor-ordered dofs, padded lines, shared memory) -> W.  Called by every thread of the CTA
// (it synchronises); `lt` is the thread's index inside its element group (P x P threads), `lc` the element's local
// coordinates, `e` its local index, u the global vector (only read for skeleton neighbours of DG spaces), perm the
// tensor -> stored permutation of the neighbours' dofs (null: identity).
// GEN = true: sub-basis spaces and periodic grids (run-time dof count, wrap-around neighbours); GEN = false compiles both out
// R > 1: `comp` is the thread's component and `cmask` the lanes holding the components of its element (all of them take the same
// path through this function: they share lt, lc and `active`).
template <int N, int MI, int MS, class Integrands, bool GEN = false, int R = 1>
__device__ __forceinline__ void element_integrals(const QuadTabDev<N, MI, MS>& T, const BoxDev& box, const Integrands& I, const int* perm,
                                                  const double* __restrict__ u, const bool active, const int (&lc)[3],
                                                  const long long e, const int lt, double* U, double* W, double* S,
                                                  const int comp = 0, const unsigned cmask = 0xffffffffu) {
  using Cfg = DgQuadCfg<N, MI, MS>;
  constexpr int P = Cfg::P, LN = Cfg::LN, LM = Cfg::LM, LS = Cfg::LS, N3 = N * N * N;
  const int la = lt / P, lb = lt % P;
  const double ih0 = box.ih[0], ih1 = box.ih[1], ih2 = box.ih[2];
  const double detJ = box.h[0] * box.h[1] * box.h[2];

  // =========================== interior integral ===========================
  double* T1a = S; double* T1b = S + N * N * LM;
  double* T2a = S + 2 * N * N * LM; double* T2b = T2a + N * MI * LM; double* T2c = T2b + N * MI * LM;
  if (active && la < N && lb < N) {   // V1: contract axis 2 (fastest); thread = line (i0, i1)
    double in[N], a[MI], b[MI];
#pragma unroll
    for (int c = 0; c < N; ++c) in[c] = U[(la * N + lb) * LN + c];
    quad_mv<MI, N>(T.Bi, in, a); quad_mv<MI, N>(T.Gi, in, b);
#pragma unroll
    for (int q = 0; q < MI; ++q) { T1a[(la * N + lb) * LM + q] = a[q]; T1b[(la * N + lb) * LM + q] = b[q]; }
  }
  __syncthreads();
  if (active && la < N && lb < MI) {   // V2: contract axis 1; thread = line (i0, q2)
    double ina[N], inb[N], oa[MI], ob[MI], oc[MI];
#pragma unroll
    for (int i = 0; i < N; ++i) { ina[i] = T1a[(la * N + i) * LM + lb]; inb[i] = T1b[(la * N + i) * LM + lb]; }
    quad_mv<MI, N>(T.Bi, ina, oa); quad_mv<MI, N>(T.Gi, ina, ob); quad_mv<MI, N>(T.Bi, inb, oc);
#pragma unroll
    for (int q = 0; q < MI; ++q) { T2a[(la * MI + q) * LM + lb] = oa[q]; T2b[(la * MI + q) * LM + lb] = ob[q]; T2c[(la * MI + q) * LM + lb] = oc[q]; }
  }
  __syncthreads();
  if (active && la < MI && lb < MI) {   // V3-5: contract axis 0, integrand at the MI points of line (q1, q2), test along axis 0
    double ina[N], inb[N], inc[N], v[MI], dx[MI], dy[MI], dz[MI];
#pragma unroll
    for (int i = 0; i < N; ++i) { ina[i] = T2a[(i * MI + la) * LM + lb]; inb[i] = T2b[(i * MI + la) * LM + lb]; inc[i] = T2c[(i * MI + la) * LM + lb]; }
    quad_mv<MI, N>(T.Bi, ina, v); quad_mv<MI, N>(T.Gi, ina, dx); quad_mv<MI, N>(T.Bi, inb, dy); quad_mv<MI, N>(T.Bi, inc, dz);
    double rs[MI], rx[MI], ry[MI], rz[MI];
    double xq[3];
    xq[1] = box.lo[1] + box.h[1] * ((box.origin[1] + lc[1]) + T.xi[la]);
    xq[2] = box.lo[2] + box.h[2] * ((box.origin[2] + lc[2]) + T.xi[lb]);
    const double w12 = T.wi[la] * T.wi[lb] * detJ;
#pragma unroll
    for (int q0 = 0; q0 < MI; ++q0) {
      xq[0] = box.lo[0] + box.h[0] * ((box.origin[0] + lc[0]) + T.xi[q0]);
      PointValue pv; pv.u = v[q0]; pv.du[0] = dx[q0] * ih0; pv.du[1] = dy[q0] * ih1; pv.du[2] = dz[q0] * ih2;
      PointRange r;
      if constexpr (R == 1) r = I.interior(xq, pv);
      else r = quad_pick<R>(I.interior(xq, quad_collect<R>(pv, cmask)), comp);
      const double wq = T.wi[q0] * w12;               // qp.weight() * integrationElement (galerkin.hh:353)
      rs[q0] = r.s * wq; rx[q0] = r.F[0] * (wq * ih0); ry[q0] = r.F[1] * (wq * ih1); rz[q0] = r.F[2] * (wq * ih2);
    }
    double oa[N], ob[N], oc[N]; quad_zero<N>(oa); quad_zero<N>(ob); quad_zero<N>(oc);
    quad_mvt_add<MI, N>(T.Bi, rs, oa); quad_mvt_add<MI, N>(T.Gi, rx, oa); quad_mvt_add<MI, N>(T.Bi, ry, ob); quad_mvt_add<MI, N>(T.Bi, rz, oc);
#pragma unroll
    for (int i = 0; i < N; ++i) { T2a[(i * MI + la) * LM + lb] = oa[i]; T2b[(i * MI + la) * LM + lb] = ob[i]; T2c[(i * MI + la) * LM + lb] = oc[i]; }
  }
  __syncthreads();
  if (active && la < N && lb < MI) {   // V6: test along axis 1; thread = line (i0, q2)
    double ina[MI], inb[MI], inc[MI], oa[N], ob[N]; quad_zero<N>(oa); quad_zero<N>(ob);
#pragma unroll
    for (int q = 0; q < MI; ++q) { ina[q] = T2a[(la * MI + q) * LM + lb]; inb[q] = T2b[(la * MI + q) * LM + lb]; inc[q] = T2c[(la * MI + q) * LM + lb]; }
    quad_mvt_add<MI, N>(T.Bi, ina, oa); quad_mvt_add<MI, N>(T.Gi, inb, oa); quad_mvt_add<MI, N>(T.Bi, inc, ob);
#pragma unroll
    for (int i = 0; i < N; ++i) { T1a[(la * N + i) * LM + lb] = oa[i]; T1b[(la * N + i) * LM + lb] = ob[i]; }
  }
  __syncthreads();
  if (active && la < N && lb < N) {   // V7: test along axis 2; thread = line (i0, i1)
    double ina[MI], inb[MI], o[N]; quad_zero<N>(o);
#pragma unroll
    for (int q = 0; q < MI; ++q) { ina[q] = T1a[(la * N + lb) * LM + q]; inb[q] = T1b[(la * N + lb) * LM + q]; }
    quad_mvt_add<MI, N>(T.Bi, ina, o); quad_mvt_add<MI, N>(T.Gi, inb, o);
#pragma unroll
    for (int c = 0; c < N; ++c) W[(la * N + lb) * LN + c] = o[c];
  }
  __syncthreads();

  // =========================== skeleton + boundary integrals ===========================
  if (!(I.m.has_skeleton || I.m.has_boundary)) return;
  // face f = 2*axis+side.  C[f][k][field] (k: 0 own / 1 neighbour; field: 0 trace, 1 normal-derivative trace), N x N coefficients each;
  // E[f][k][e] after the first tangential sweep; later X[f][x] overlays C and R[f][field] overlays E.
  double* const Cc = S; double* const Ee = S + Cfg::kRegA;
  constexpr int kC = N * LN, kE = N * LS;
  // F0: trace coefficients of u_K and of the neighbour on each face; thread = (ia, ib); one unrolled block per axis
  if (active && la < N && lb < N) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      constexpr int stp[3] = {N * LN, LN, 1}; constexpr int stt[3] = {N * N, N, 1};      // padded / tensor strides
      const int a = d == 0 ? 1 : 0, b = d == 2 ? 1 : 2;
      const int base = la * stp[a] + lb * stp[b], tbase = la * stt[a] + lb * stt[b];
      double uu[N];
#pragma unroll
      for (int c = 0; c < N; ++c) uu[c] = U[base + c * stp[d]];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int f = 2 * d + s;
        double tv = 0, td = 0;
#pragma unroll
        for (int c = 0; c < N; ++c) { tv = fma(T.phi[s][c], uu[c], tv); td = fma(T.dphi[s][c], uu[c], td); }
        Cc[((f * 2 + 0) * 2 + 0) * kC + la * LN + lb] = tv; Cc[((f * 2 + 0) * 2 + 1) * kC + la * LN + lb] = td;
        double nv = 0, nd = 0;
        const int cn = lc[d] + (s ? 1 : -1);
        if (I.m.has_skeleton && ((cn >= 0 && cn < box.n[d]) || (GEN && ((box.periodic >> d) & 1)))) {
          const double* un = S + Cfg::kNb + f * N3;     // staged by the kernel prologue (stored order)
#pragma unroll
          for (int c = 0; c < N; ++c) {
            const int t = tbase + c * stt[d];
            const int ps = perm ? perm[t] : t;                 // -1: the space holds no such function (sub-bases, see below)
            const double uv = ps >= 0 ? un[ps] : 0.0;
            nv = fma(T.phi[1 - s][c], uv, nv); nd = fma(T.dphi[1 - s][c], uv, nd);
          }
        }
        Cc[((f * 2 + 1) * 2 + 0) * kC + la * LN + lb] = nv; Cc[((f * 2 + 1) * 2 + 1) * kC + la * LN + lb] = nd;
      }
    }
  }
  __syncthreads();
  // F1: tangential sweep along ib; job = (f, k, ia): N coefficients of the trace and of the normal derivative -> MS points each
  if (active) {
    for (int job = lt; job < 12 * N; job += Cfg::T2) {
      const int fk = job / N, ia = job % N;
      const double* Cv = Cc + (fk * 2 + 0) * kC + ia * LN; const double* Cd = Cc + (fk * 2 + 1) * kC + ia * LN;
      double val[N], der[N], e1[MS], e2[MS], e3[MS];
#pragma unroll
      for (int i = 0; i < N; ++i) { val[i] = Cv[i]; der[i] = Cd[i]; }
      quad_mv<MS, N>(T.Bs, val, e1); quad_mv<MS, N>(T.Gs, val, e2); quad_mv<MS, N>(T.Bs, der, e3);
      double* E0 = Ee + (fk * 3) * kE + ia * LS;
#pragma unroll
      for (int q = 0; q < MS; ++q) { E0[q] = e1[q]; E0[kE + q] = e2[q]; E0[2 * kE + q] = e3[q]; }
    }
  }
  __syncthreads();
  // F2-4: sweep along ia, integrand at the MS points (., qb), test along qa; job = (f, qb).  The axis of the face is a run-time
  // value here (all six faces share the code): it enters through selects, never as a register-array index.
  if (active) {
    for (int job = lt; job < 6 * MS; job += Cfg::T2) {
      const int f = job / MS, qb = job % MS, d = f >> 1, s = f & 1;
      double val[2][MS], dta[2][MS], dtb[2][MS], dn[2][MS];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double* E0 = Ee + ((f * 2 + k) * 3) * kE + qb;
        double e1[N], e2[N], e3[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { e1[i] = E0[i * LS]; e2[i] = E0[kE + i * LS]; e3[i] = E0[2 * kE + i * LS]; }
        quad_mv<MS, N>(T.Bs, e1, val[k]); quad_mv<MS, N>(T.Gs, e1, dta[k]); quad_mv<MS, N>(T.Bs, e2, dtb[k]); quad_mv<MS, N>(T.Bs, e3, dn[k]);
      }
      // tangential axes a < b of the face; geometry of this axis
      const double ihd = d == 0 ? ih0 : d == 1 ? ih1 : ih2, iha = d == 0 ? ih1 : ih0, ihb = d == 2 ? ih1 : ih2;
      const double area = detJ * ihd, ihe = ihd;          // faceArea; 1 / he with he = avg(CellVolume)/FacetArea = h_d on a uniform box
      const int lcd = d == 0 ? lc[0] : d == 1 ? lc[1] : lc[2], nd_ = d == 0 ? box.n[0] : d == 1 ? box.n[1] : box.n[2];
      const int olo = d == 0 ? box.own_lo[0] : d == 1 ? box.own_lo[1] : box.own_lo[2], ohi = d == 0 ? box.own_hi[0] : d == 1 ? box.own_hi[1] : box.own_hi[2];
      // domain boundary in GLOBAL coordinates: the rank-local box of a continuous space ends at rank interfaces too (no ghost layers
      // there), and those faces carry no boundary integral
      const int gcd = (d == 0 ? box.origin[0] : d == 1 ? box.origin[1] : box.origin[2]) + lcd;
      const bool dom_bnd = s ? gcd == (d == 0 ? box.gn[0] : d == 1 ? box.gn[1] : box.gn[2]) - 1 : gcd == 0;
      int cn = lcd + (s ? 1 : -1);
      bool nb_exists = cn >= 0 && cn < nd_;
      if (GEN && !nb_exists && ((box.periodic >> d) & 1)) { cn = cn < 0 ? nd_ - 1 : 0; nb_exists = true; }     // periodic: the far side's element (galerkin.hh:859-861)
      const bool nb_owned = cn >= olo && cn < ohi;
      // inside = lower element index, or the owned element next to a ghost (one-sided from the owned side, galerkin.hh:866-878)
      const bool own_inside = !nb_owned || cn > lcd;      // (without wrap-around: the neighbour across the high side)
      // physical coordinates: the normal coordinate is fixed, (a, b) follow the face points.  The point is the INSIDE element's
      // (on a periodic face the two sides differ by the domain length)
      const int cface = own_inside ? lcd + s : cn + (s ? 0 : 1);
      const double xd = (d == 0 ? box.lo[0] + box.h[0] * (box.origin[0] + cface) : d == 1 ? box.lo[1] + box.h[1] * (box.origin[1] + cface) : box.lo[2] + box.h[2] * (box.origin[2] + cface));
      const double xb_ = d == 2 ? box.lo[1] + box.h[1] * ((box.origin[1] + lc[1]) + T.xs[qb]) : box.lo[2] + box.h[2] * ((box.origin[2] + lc[2]) + T.xs[qb]);
      double R0[MS], Ra[MS], Rb[MS], Rn[MS];
#pragma unroll
      for (int qa = 0; qa < MS; ++qa) {
        const double xa_ = d == 0 ? box.lo[1] + box.h[1] * ((box.origin[1] + lc[1]) + T.xs[qa]) : box.lo[0] + box.h[0] * ((box.origin[0] + lc[0]) + T.xs[qa]);
        double xq[3];
        xq[0] = d == 0 ? xd : xa_; xq[1] = d == 1 ? xd : (d == 0 ? xa_ : xb_); xq[2] = d == 2 ? xd : xb_;
        PointValue own, nb;
        {
          const double gn_ = dn[0][qa] * ihd, ga = dta[0][qa] * iha, gb = dtb[0][qa] * ihb;
          own.u = val[0][qa]; own.du[0] = d == 0 ? gn_ : ga; own.du[1] = d == 1 ? gn_ : (d == 0 ? ga : gb); own.du[2] = d == 2 ? gn_ : gb;
        }
        {
          const double gn_ = dn[1][qa] * ihd, ga = dta[1][qa] * iha, gb = dtb[1][qa] * ihb;
          nb.u = val[1][qa]; nb.du[0] = d == 0 ? gn_ : ga; nb.du[1] = d == 1 ? gn_ : (d == 0 ? ga : gb); nb.du[2] = d == 2 ? gn_ : gb;
        }
        PointRange r; r.s = 0; r.F[0] = r.F[1] = r.F[2] = 0;
        if constexpr (R == 1) {
          if (nb_exists) {
            if (I.m.has_skeleton) {
              PointRange rin, rout;
              if (own_inside) { I.skeleton(xq, d, s ? 1.0 : -1.0, ihe, own, nb, rin, rout); r = rin; }
              else            { I.skeleton(xq, d, s ? -1.0 : 1.0, ihe, nb, own, rin, rout); r = rout; }
            }
          } else if (I.m.has_boundary && d < box.dim && dom_bnd) {         // (a 2-D mesh is one layer of cells: its x2-faces are no boundary)
            r = I.boundary(d, s, ihe, xq, own);
          }
        } else {
          const PointValueV<R> ownv = quad_collect<R>(own, cmask), nbv = quad_collect<R>(nb, cmask);
          if (nb_exists) {
            if (I.m.has_skeleton) {
              PointRangeV<R> rin, rout;
              if (own_inside) { I.skeleton(xq, d, s ? 1.0 : -1.0, ihe, ownv, nbv, rin, rout); r = quad_pick<R>(rin, comp); }
              else            { I.skeleton(xq, d, s ? -1.0 : 1.0, ihe, nbv, ownv, rin, rout); r = quad_pick<R>(rout, comp); }
            }
          } else if (I.m.has_boundary && d < box.dim && dom_bnd) {
            r = quad_pick<R>(I.boundary(d, s, ihe, xq, ownv), comp);
          }
        }
        const double wq = T.ws[qa] * T.ws[qb] * area;
        const double Fn = d == 0 ? r.F[0] : d == 1 ? r.F[1] : r.F[2], Fa = d == 0 ? r.F[1] : r.F[0], Fb = d == 2 ? r.F[1] : r.F[2];
        R0[qa] = r.s * wq; Ra[qa] = Fa * (wq * iha); Rb[qa] = Fb * (wq * ihb); Rn[qa] = Fn * (wq * ihd);
      }
      double x1[N], x2[N], x3[N]; quad_zero<N>(x1); quad_zero<N>(x2); quad_zero<N>(x3);
      quad_mvt_add<MS, N>(T.Bs, R0, x1); quad_mvt_add<MS, N>(T.Gs, Ra, x1); quad_mvt_add<MS, N>(T.Bs, Rb, x2); quad_mvt_add<MS, N>(T.Bs, Rn, x3);
      double* X0 = Cc + (f * 3) * kE + qb;               // X overlays C (dead since F1)
#pragma unroll
      for (int i = 0; i < N; ++i) { X0[i * LS] = x1[i]; X0[kE + i * LS] = x2[i]; X0[2 * kE + i * LS] = x3[i]; }
    }
  }
  __syncthreads();
  // F5: test along qb; job = (f, ia)
  if (active) {
    for (int job = lt; job < 6 * N; job += Cfg::T2) {
      const int f = job / N, ia = job % N;
      const double* X0 = Cc + (f * 3) * kE + ia * LS;
      double x1[MS], x2[MS], x3[MS], rv[N], rd[N]; quad_zero<N>(rv); quad_zero<N>(rd);
#pragma unroll
      for (int q = 0; q < MS; ++q) { x1[q] = X0[q]; x2[q] = X0[kE + q]; x3[q] = X0[2 * kE + q]; }
      quad_mvt_add<MS, N>(T.Bs, x1, rv); quad_mvt_add<MS, N>(T.Gs, x2, rv); quad_mvt_add<MS, N>(T.Bs, x3, rd);
      double* R0 = Ee + (f * 2) * kC + ia * LN;          // R overlays E (dead since F2)
#pragma unroll
      for (int i = 0; i < N; ++i) { R0[i] = rv[i]; R0[kC + i] = rd[i]; }
    }
  }
  __syncthreads();
  // F6: lift the face residuals into the element; one axis at a time (both sides of an axis hit the same lines)
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (active && la < N && lb < N) {
      constexpr int stp[3] = {N * LN, LN, 1};
      const int a = d == 0 ? 1 : 0, b = d == 2 ? 1 : 2;
      const int base = la * stp[a] + lb * stp[b];
      const double rv0 = Ee[((2 * d) * 2) * kC + la * LN + lb], rd0 = Ee[((2 * d) * 2 + 1) * kC + la * LN + lb];
      const double rv1 = Ee[((2 * d + 1) * 2) * kC + la * LN + lb], rd1 = Ee[((2 * d + 1) * 2 + 1) * kC + la * LN + lb];
#pragma unroll
      for (int c = 0; c < N; ++c) {
        double wv = W[base + c * stp[d]];
        wv = fma(T.phi[0][c], rv0, wv); wv = fma(T.dphi[0][c], rd0, wv); wv = fma(T.phi[1][c], rv1, wv); wv = fma(T.dphi[1][c], rd1, wv);
        W[base + c * stp[d]] = wv;
      }
    }
    __syncthreads();
  }
}

template <int N, int MI, int MS, class Integrands, bool GEN, int R = 1>
__global__ void __launch_bounds__(DgQuadCfg<N, MI, MS>::kThreads)
dg_quadrature_kernel(const __grid_constant__ QuadTabDev<N, MI, MS> T, const __grid_constant__ BoxDev box,
                     const __grid_constant__ Integrands I, const int* __restrict__ perm_g, const int nbs_arg,
                     const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                     long long n_owned, const double out_scale) {
  using Cfg = DgQuadCfg<N, MI, MS>;
  constexpr int N2 = N * N, N3 = N * N * N, EB = Cfg::EB, ELEM = Cfg::kElemDoubles, LN = Cfg::LN;
  const int nbs = GEN ? nbs_arg : N3;          // (a compile-time constant in the common case: the index divisions below fold)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  long long* elem_of = reinterpret_cast<long long*>(smem + (size_t)EB * ELEM);   // local element index per slot
  int* perm = reinterpret_cast<int*>(elem_of + EB);                               // tensor index -> stored index
  int* tinv = perm + N3;                                                          // stored index -> tensor index (padded offset)
  int* ecs = tinv + N3;                                                           // element coordinates per slot (4 ints)

  // range-R spaces: RS = R rounded up to a power of two slots per element (slot % RS = component; the slots c >= R idle along)
  constexpr int RS = quad_pow2(R);
  static_assert(R == 1 || RS <= EB, "dimRange: the components of an element must fit the element slots of a CTA");
  const int tid = threadIdx.x, es = Cfg::slot(tid), lt = Cfg::lane(tid);
  const int comp = es % RS;
  const unsigned cmask = R == 1 ? 0xffffffffu : (((1u << RS) - 1u) << ((tid & 31) / RS * RS));
  const int on0 = box.own_hi[0] - box.own_lo[0], on1 = box.own_hi[1] - box.own_lo[1];
  const long long oe = (long long)blockIdx.x * (EB / RS) + es / RS;
  const bool active = lt < Cfg::T2 && oe < n_owned;
  int lc[3] = {0, 0, 0};
  if (active) {
    lc[0] = box.own_lo[0] + (int)(oe % on0);
    lc[1] = box.own_lo[1] + (int)((oe / on0) % on1);
    lc[2] = box.own_lo[2] + (int)(oe / ((long long)on0 * on1));
  }
  const long long e = lc[0] + (long long)box.n[0] * (lc[1] + (long long)box.n[1] * lc[2]);
  if (lt == 0) { elem_of[es] = active && comp < R ? e : -1; ecs[4 * es] = lc[0]; ecs[4 * es + 1] = lc[1]; ecs[4 * es + 2] = lc[2]; }
  // The space may be a SUB-BASIS of the full tensor basis the kernel works on: nbs <= N^3 stored dofs per element, perm_g[t] = -1
  // for tensor functions it does not hold (2-D Q_k: functions constant in x2; dgonb P_k: total degree <= k).  Their coefficients
  // are zero on input and their residuals are dropped on output -- the Galerkin operator of the sub-space, exactly.
  for (int i = tid; i < N3; i += blockDim.x) { const int p = perm_g[i]; perm[i] = p; if (p >= 0) tinv[p] = (i / N2 * N + (i / N) % N) * LN + i % N; }
  if (nbs < N3)
    for (int idx = tid; idx < EB * Cfg::kU; idx += blockDim.x) smem[(size_t)(idx / Cfg::kU) * ELEM + idx % Cfg::kU] = 0.0;
  __syncthreads();

  // ---- gather: the elements' own dofs (stored order -> tensor order, padded lines) and, for skeleton terms, the stored-order
  //      dofs of their six face neighbours.  Coalesced: consecutive slots are consecutive elements, and so are their
  //      neighbours across one face.  All loads of a thread are issued before the first store (one latency, not seven). ----
  {
    constexpr int KI = (EB * N3 + Cfg::kThreads - 1) / Cfg::kThreads;
    const long long estep1 = box.n[0], estep2 = (long long)box.n[0] * box.n[1];
    const bool skel = I.m.has_skeleton;
    double own[KI], nbv[KI][6]; bool have[KI][6]; int base[KI], jj[KI];
#pragma unroll
    for (int k = 0; k < KI; ++k) {
      const int idx = tid + k * Cfg::kThreads, s2 = idx / nbs, j = idx % nbs;
      const long long e2 = idx < EB * nbs ? elem_of[s2] : -1;
      base[k] = e2 >= 0 ? s2 * ELEM : -1; jj[k] = j;
      const int cc = R == 1 ? 0 : s2 % RS;                                // dof (element, j, component) = (e * nbs + j) * R + c
      if (e2 >= 0) own[k] = u[(e2 * nbs + j) * R + cc];
#pragma unroll
      for (int f = 0; f < 6; ++f) {
        const int d = f >> 1;
        have[k][f] = false;
        if (skel && e2 >= 0) {
          const int c0 = ecs[4 * s2 + d]; int cn = c0 + ((f & 1) ? 1 : -1);
          const long long step = d == 0 ? 1 : d == 1 ? estep1 : estep2;
          if (GEN && (cn < 0 || cn >= box.n[d]) && ((box.periodic >> d) & 1)) cn = cn < 0 ? box.n[d] - 1 : 0;
          if (cn >= 0 && cn < box.n[d]) { have[k][f] = true; nbv[k][f] = u[((e2 + (long long)(cn - c0) * step) * nbs + j) * R + cc]; }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KI; ++k) {
      if (base[k] >= 0) {
        smem[base[k] + tinv[jj[k]]] = own[k];
#pragma unroll
        for (int f = 0; f < 6; ++f)
          if (have[k][f]) smem[base[k] + 2 * Cfg::kU + Cfg::kNb + f * N3 + jj[k]] = nbv[k][f];
      }
    }
  }
  __syncthreads();

  double* U = smem + (size_t)es * ELEM;
  element_integrals<N, MI, MS, Integrands, GEN, R>(T, box, I, perm, u, active, lc, e, lt, U, U + Cfg::kU, U + 2 * Cfg::kU, comp, cmask);

  // ---- write w_K once (tensor order -> stored order), optionally w = A u - b ----
  for (int idx = tid; idx < EB * nbs; idx += blockDim.x) {
    const int s2 = idx / nbs, j = idx % nbs; const long long e2 = elem_of[s2];
    if (e2 >= 0) {
      double val = smem[(size_t)s2 * ELEM + Cfg::kU + tinv[j]] * out_scale;      // out_scale: inverse mass of MOLGalerkinOperator (1 otherwise)
      const long long g = (e2 * nbs + j) * R + (R == 1 ? 0 : s2 % RS);
      if (bvec) val -= bvec[g];
      w[g] = val;
    }
  }
}

}  // namespace b200fem
