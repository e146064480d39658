// dg_quadrature.cuh -- generic matrix-free DG operator apply by quadrature (3-D cubes, Legendre Q_k).
//
// One kernel does what GalerkinOperator::evaluate does per element (dune/fem/schemes/galerkin.hh:811-917):
//   gather u_K (getLocalDofs, function/common/discretefunction.hh:945-950)
//   interior integral  (galerkin.hh:332-360): evaluateAll/jacobianAll at the quadrature points
//                      (space/basisfunctionset/default.hh:276-372), integrand, axpy (default.hh:199-257)
//   skeleton/boundary integrals (galerkin.hh:414-435, 475-537) on the six faces
//   write w_K (addLocalDofs, discretefunction.hh:929-934)
// B200-first differences from the reference loop:
//  * the dense tabulated contraction u_q = sum_i B[q][i] u_i (space/shapefunctionset/caching.hh:302-319) is
//    sum-factorised: B = B1 (x) B1 (x) B1 with the 1-D tables held in the constant bank;
//  * every element integrates all six of its faces and keeps only its own half of the skeleton integrand
//    (the reference integrates a face once, from the lower-index element, and scatters into both elements,
//    galerkin.hh:879-897).  Roles (inside = lower element index, or the owned element next to a ghost) follow
//    the reference, so the integrand sees identical arguments; the result is race-free without atomics or
//    colouring and w is written exactly once (the w.clear() of galerkin.hh:1463 is fused away);
//  * a CTA owns EB consecutive elements, N*N threads per element, one thread per tensor line; intermediate
//    tensors live in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "integrands.cuh"

namespace b200fem {

// rank-local box of the global Cartesian grid (with ghost layers where a neighbouring rank exists)
struct BoxDev {
  int dim;
  int n[3];                 // local box extents in elements (owned + ghost)
  int own_lo[3], own_hi[3]; // owned sub-box, local coordinates
  int origin[3];            // global element coordinates of local element (0,0,0)
  int gn[3];                // global extents
  double lo[3], h[3];
};

template <int N>
struct DgTabDev {
  double B[N * N], G[N * N];     // B[q*N+i] = phi_i(x_q), G = phi_i'(x_q)   (volume and face rules coincide)
  double x[N], w[N];             // Gauss points / weights on [0,1]
  double phi[2][N], dphi[2][N];  // traces at 0 and 1
};

template <int N> struct DgQuadCfg {
  static constexpr int N2 = N * N, N3 = N * N * N;
  static constexpr int kFaceScratch = 60 * N2;
  static constexpr int kVolScratch = 9 * N3;
  static constexpr int kScratch = kFaceScratch > kVolScratch ? kFaceScratch : kVolScratch;
  static constexpr int kElemDoubles = 2 * N3 + kScratch;
  static constexpr int EB = N == 2 ? 32 : N == 3 ? 16 : N == 4 ? 8 : N == 5 ? 6 : 4;   // elements per CTA
  static constexpr int kThreads = EB * N2;
  static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)EB * kElemDoubles + sizeof(int) * N3 * 2 + sizeof(long long) * EB + sizeof(int) * 4 * EB; }
};

// out[q] = sum_c M[q*N+c] in[c]
template <int N> __device__ __forceinline__ void mv(const double* __restrict__ M, const double (&in)[N], double (&out)[N]) {
#pragma unroll
  for (int q = 0; q < N; ++q) { double s = 0;
#pragma unroll
    for (int c = 0; c < N; ++c) s = fma(M[q * N + c], in[c], s); out[q] = s; }
}
// out[c] += sum_q M[q*N+c] in[q]
template <int N> __device__ __forceinline__ void mvt_add(const double* __restrict__ M, const double (&in)[N], double (&out)[N]) {
#pragma unroll
  for (int c = 0; c < N; ++c) { double s = out[c];
#pragma unroll
    for (int q = 0; q < N; ++q) s = fma(M[q * N + c], in[q], s); out[c] = s; }
}
template <int N> __device__ __forceinline__ void zero(double (&a)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) a[i] = 0;
}

// All integrals of one element: U (tensor-ordered dofs, shared memory) -> W.  Called by every thread of the CTA
// (it synchronises); `lt` is the thread's line index inside its element group, `lc` the element's local coordinates,
// `e` its local index, u the global vector (only read for skeleton neighbours of DG spaces).
template <int N, class Integrands>
__device__ __forceinline__ void element_integrals(const DgTabDev<N>& T, const BoxDev& box, const Integrands& I, const int* perm,
                                                  const double* __restrict__ u, const bool active, const int (&lc)[3],
                                                  const long long e, const int lt, double* U, double* W, double* S) {
  constexpr int N2 = N * N, N3 = N * N * N;
  const double h0 = box.h[0], h1 = box.h[1], h2 = box.h[2];
  const double detJ = h0 * h1 * h2;

  // =========================== interior integral ===========================
  double* T1a = S; double* T1b = S + N3;
  double* T2a = S + 2 * N3; double* T2b = S + 3 * N3; double* T2c = S + 4 * N3;
  if (active) {   // V1: contract axis 2 (fastest); thread = line (i0,i1)
    double in[N], a[N], b[N];
#pragma unroll
    for (int c = 0; c < N; ++c) in[c] = U[lt * N + c];
    mv<N>(T.B, in, a); mv<N>(T.G, in, b);
#pragma unroll
    for (int q = 0; q < N; ++q) { T1a[lt * N + q] = a[q]; T1b[lt * N + q] = b[q]; }
  }
  __syncthreads();
  if (active) {   // V2: contract axis 1; thread = line (i0,q2)
    const int base = (lt / N) * N2 + (lt % N);
    double ina[N], inb[N], oa[N], ob[N], oc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { ina[i] = T1a[base + i * N]; inb[i] = T1b[base + i * N]; }
    mv<N>(T.B, ina, oa); mv<N>(T.G, ina, ob); mv<N>(T.B, inb, oc);
#pragma unroll
    for (int q = 0; q < N; ++q) { T2a[base + q * N] = oa[q]; T2b[base + q * N] = ob[q]; T2c[base + q * N] = oc[q]; }
  }
  __syncthreads();
  if (active) {   // V3-5: contract axis 0, integrand at the N points of line (q1,q2), test along axis 0
    const int q1 = lt / N, q2 = lt % N;
    double ina[N], inb[N], inc[N], v[N], dx[N], dy[N], dz[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { ina[i] = T2a[lt + i * N2]; inb[i] = T2b[lt + i * N2]; inc[i] = T2c[lt + i * N2]; }
    mv<N>(T.B, ina, v); mv<N>(T.G, ina, dx); mv<N>(T.B, inb, dy); mv<N>(T.B, inc, dz);
    double rs[N], rx[N], ry[N], rz[N];
    double xq[3];
    xq[1] = box.lo[1] + h1 * ((box.origin[1] + lc[1]) + T.x[q1]);
    xq[2] = box.lo[2] + h2 * ((box.origin[2] + lc[2]) + T.x[q2]);
    const double w12 = T.w[q1] * T.w[q2] * detJ;
#pragma unroll
    for (int q0 = 0; q0 < N; ++q0) {
      xq[0] = box.lo[0] + h0 * ((box.origin[0] + lc[0]) + T.x[q0]);
      PointValue pv; pv.u = v[q0]; pv.du[0] = dx[q0] / h0; pv.du[1] = dy[q0] / h1; pv.du[2] = dz[q0] / h2;
      PointRange r = I.interior(xq, pv);
      const double wq = T.w[q0] * w12;               // qp.weight() * integrationElement (galerkin.hh:353)
      rs[q0] = r.s * wq; rx[q0] = r.F[0] * wq / h0; ry[q0] = r.F[1] * wq / h1; rz[q0] = r.F[2] * wq / h2;
    }
    double oa[N], ob[N], oc[N]; zero<N>(oa); zero<N>(ob); zero<N>(oc);
    mvt_add<N>(T.B, rs, oa); mvt_add<N>(T.G, rx, oa); mvt_add<N>(T.B, ry, ob); mvt_add<N>(T.B, rz, oc);
#pragma unroll
    for (int i = 0; i < N; ++i) { T2a[lt + i * N2] = oa[i]; T2b[lt + i * N2] = ob[i]; T2c[lt + i * N2] = oc[i]; }
  }
  __syncthreads();
  if (active) {   // V6: test along axis 1; thread = line (i0,q2)
    const int base = (lt / N) * N2 + (lt % N);
    double ina[N], inb[N], inc[N], oa[N], ob[N]; zero<N>(oa); zero<N>(ob);
#pragma unroll
    for (int q = 0; q < N; ++q) { ina[q] = T2a[base + q * N]; inb[q] = T2b[base + q * N]; inc[q] = T2c[base + q * N]; }
    mvt_add<N>(T.B, ina, oa); mvt_add<N>(T.G, inb, oa); mvt_add<N>(T.B, inc, ob);
#pragma unroll
    for (int i = 0; i < N; ++i) { T1a[base + i * N] = oa[i]; T1b[base + i * N] = ob[i]; }
  }
  __syncthreads();
  if (active) {   // V7: test along axis 2; thread = line (i0,i1)
    double ina[N], inb[N], o[N]; zero<N>(o);
#pragma unroll
    for (int q = 0; q < N; ++q) { ina[q] = T1a[lt * N + q]; inb[q] = T1b[lt * N + q]; }
    mvt_add<N>(T.B, ina, o); mvt_add<N>(T.G, inb, o);
#pragma unroll
    for (int c = 0; c < N; ++c) W[lt * N + c] = o[c];
  }
  __syncthreads();

  // =========================== skeleton + boundary integrals ===========================
  const bool do_faces = I.m.has_skeleton || I.m.has_boundary;
  if (do_faces) {
    // per face f = 2*axis+side: C = [TV,TD,NV,ND] (4 N2), E (6 N2); later X overlays C, R overlays E
    constexpr int FS = 10 * N2;
    const int st[3] = {N2, N, 1};
    const double hh[3] = {h0, h1, h2};
    // neighbour bookkeeping for this element
    long long enb[6]; bool nb_exists[6], own_inside[6];
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const int d = f >> 1, s = f & 1; int c = lc[d] + (s ? 1 : -1);
      nb_exists[f] = active && c >= 0 && c < box.n[d];
      const long long step = d == 0 ? 1 : d == 1 ? box.n[0] : (long long)box.n[0] * box.n[1];
      enb[f] = e + (s ? step : -step);
      const bool nb_owned = c >= box.own_lo[d] && c < box.own_hi[d];
      own_inside[f] = !nb_owned || e < enb[f];      // ghost neighbour: one-sided from the owned element (galerkin.hh:866-878)
    }
    if (active) {   // F0: trace coefficients of u_K and of the neighbour on each face; thread = (ia,ib)
      const int ia = lt / N, ib = lt % N;
#pragma unroll
      for (int f = 0; f < 6; ++f) {
        const int d = f >> 1, s = f & 1, a = d == 0 ? 1 : 0, b = d == 2 ? 1 : 2;
        const int base = ia * st[a] + ib * st[b];
        double* C = S + f * FS;
        double tv = 0, td = 0;
#pragma unroll
        for (int c = 0; c < N; ++c) { const double uu = U[base + c * st[d]]; tv = fma(T.phi[s][c], uu, tv); td = fma(T.dphi[s][c], uu, td); }
        C[lt] = tv; C[N2 + lt] = td;
        double nv = 0, nd = 0;
        if (nb_exists[f] && I.m.has_skeleton) {
          const double* un = u + enb[f] * N3;
#pragma unroll
          for (int c = 0; c < N; ++c) { const double uu = un[perm[base + c * st[d]]]; nv = fma(T.phi[1 - s][c], uu, nv); nd = fma(T.dphi[1 - s][c], uu, nd); }
        }
        C[2 * N2 + lt] = nv; C[3 * N2 + lt] = nd;
      }
    }
    __syncthreads();
    if (active) {   // F1: tangential sweep along ib; job = (f, k, ia)
      for (int job = lt; job < 12 * N; job += N2) {
        const int f = job / (2 * N), k = (job / N) & 1, ia = job % N;
        double* C = S + f * FS; double* E = C + 4 * N2;
        double val[N], der[N], e1[N], e2[N], e3[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { val[i] = C[(2 * k) * N2 + ia * N + i]; der[i] = C[(2 * k + 1) * N2 + ia * N + i]; }
        mv<N>(T.B, val, e1); mv<N>(T.G, val, e2); mv<N>(T.B, der, e3);
#pragma unroll
        for (int q = 0; q < N; ++q) { E[(3 * k) * N2 + ia * N + q] = e1[q]; E[(3 * k + 1) * N2 + ia * N + q] = e2[q]; E[(3 * k + 2) * N2 + ia * N + q] = e3[q]; }
      }
    }
    __syncthreads();
    if (active) {   // F2-4: sweep along ia, integrand at the N points (.,qb), test along qa; job = (f, qb)
      for (int job = lt; job < 6 * N; job += N2) {
        const int f = job / N, qb = job % N;
        const int d = f >> 1, s = f & 1, a = d == 0 ? 1 : 0, b = d == 2 ? 1 : 2;
        double* C = S + f * FS; double* E = C + 4 * N2;
        double val[2][N], dta[2][N], dtb[2][N], dn[2][N];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          double e1[N], e2[N], e3[N];
#pragma unroll
          for (int i = 0; i < N; ++i) { e1[i] = E[(3 * k) * N2 + i * N + qb]; e2[i] = E[(3 * k + 1) * N2 + i * N + qb]; e3[i] = E[(3 * k + 2) * N2 + i * N + qb]; }
          mv<N>(T.B, e1, val[k]); mv<N>(T.G, e1, dta[k]); mv<N>(T.B, e2, dtb[k]); mv<N>(T.B, e3, dn[k]);
        }
        const double area = detJ / hh[d], he = hh[d];      // faceArea, avg(CellVolume)/FacetArea on a uniform box
        double xq[3];
        xq[d] = box.lo[d] + hh[d] * ((box.origin[d] + lc[d]) + s);
        xq[b] = box.lo[b] + hh[b] * ((box.origin[b] + lc[b]) + T.x[qb]);
        double R0[N], Ra[N], Rb[N], Rn[N];
#pragma unroll
        for (int qa = 0; qa < N; ++qa) {
          xq[a] = box.lo[a] + hh[a] * ((box.origin[a] + lc[a]) + T.x[qa]);
          PointValue own, nb;
          own.u = val[0][qa]; own.du[d] = dn[0][qa] / hh[d]; own.du[a] = dta[0][qa] / hh[a]; own.du[b] = dtb[0][qa] / hh[b];
          nb.u = val[1][qa];  nb.du[d] = dn[1][qa] / hh[d];  nb.du[a] = dta[1][qa] / hh[a];  nb.du[b] = dtb[1][qa] / hh[b];
          PointRange r; r.s = 0; r.F[0] = r.F[1] = r.F[2] = 0;
          if (nb_exists[f]) {
            if (I.m.has_skeleton) {
              PointRange rin, rout;
              if (own_inside[f]) { I.skeleton(d, s ? 1.0 : -1.0, he, own, nb, rin, rout); r = rin; }
              else               { I.skeleton(d, s ? -1.0 : 1.0, he, nb, own, rin, rout); r = rout; }
            }
          } else if (I.m.has_boundary) {
            r = I.boundary(d, s, he, xq, own);
          }
          const double wq = T.w[qa] * T.w[qb] * area;
          R0[qa] = r.s * wq; Ra[qa] = r.F[a] * wq / hh[a]; Rb[qa] = r.F[b] * wq / hh[b]; Rn[qa] = r.F[d] * wq / hh[d];
        }
        double x1[N], x2[N], x3[N]; zero<N>(x1); zero<N>(x2); zero<N>(x3);
        mvt_add<N>(T.B, R0, x1); mvt_add<N>(T.G, Ra, x1); mvt_add<N>(T.B, Rb, x2); mvt_add<N>(T.B, Rn, x3);
#pragma unroll
        for (int i = 0; i < N; ++i) { C[i * N + qb] = x1[i]; C[N2 + i * N + qb] = x2[i]; C[2 * N2 + i * N + qb] = x3[i]; }
      }
    }
    __syncthreads();
    if (active) {   // F5: test along qb; job = (f, ia)
      for (int job = lt; job < 6 * N; job += N2) {
        const int f = job / N, ia = job % N;
        double* C = S + f * FS; double* E = C + 4 * N2;
        double x1[N], x2[N], x3[N], rv[N], rd[N]; zero<N>(rv); zero<N>(rd);
#pragma unroll
        for (int q = 0; q < N; ++q) { x1[q] = C[ia * N + q]; x2[q] = C[N2 + ia * N + q]; x3[q] = C[2 * N2 + ia * N + q]; }
        mvt_add<N>(T.B, x1, rv); mvt_add<N>(T.G, x2, rv); mvt_add<N>(T.B, x3, rd);
#pragma unroll
        for (int i = 0; i < N; ++i) { E[ia * N + i] = rv[i]; E[N2 + ia * N + i] = rd[i]; }
      }
    }
    __syncthreads();
    // F6: lift the face residuals into the element; one axis at a time (both sides of an axis hit the same lines)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (active) {
        const int a = d == 0 ? 1 : 0, b = d == 2 ? 1 : 2;
        const int ia = lt / N, ib = lt % N, base = ia * st[a] + ib * st[b];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const double* E = S + (2 * d + s) * FS + 4 * N2;
          const double rv = E[lt], rd = E[N2 + lt];
#pragma unroll
          for (int c = 0; c < N; ++c) W[base + c * st[d]] += T.phi[s][c] * rv + T.dphi[s][c] * rd;
        }
      }
      __syncthreads();
    }
  }

}

template <int N, class Integrands>
__global__ void __launch_bounds__(DgQuadCfg<N>::kThreads)
dg_quadrature_kernel(const __grid_constant__ DgTabDev<N> T, const __grid_constant__ BoxDev box,
                     const __grid_constant__ Integrands I, const int* __restrict__ perm_g,
                     const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                     long long n_owned, const double out_scale) {
  using Cfg = DgQuadCfg<N>;
  constexpr int N2 = Cfg::N2, N3 = Cfg::N3, EB = Cfg::EB, ELEM = Cfg::kElemDoubles;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  long long* elem_of = reinterpret_cast<long long*>(smem + (size_t)EB * ELEM);   // local element index per slot
  int* perm = reinterpret_cast<int*>(elem_of + EB);                               // tensor index -> stored index
  int* tinv = perm + N3;                                                          // stored index -> tensor index

  const int tid = threadIdx.x, es = tid / N2, lt = tid % N2;
  const int on0 = box.own_hi[0] - box.own_lo[0], on1 = box.own_hi[1] - box.own_lo[1];
  const long long oe = (long long)blockIdx.x * EB + es;
  const bool active = oe < n_owned;
  int lc[3] = {0, 0, 0};
  if (active) {
    lc[0] = box.own_lo[0] + (int)(oe % on0);
    lc[1] = box.own_lo[1] + (int)((oe / on0) % on1);
    lc[2] = box.own_lo[2] + (int)(oe / ((long long)on0 * on1));
  }
  const long long e = lc[0] + (long long)box.n[0] * (lc[1] + (long long)box.n[1] * lc[2]);
  if (lt == 0) elem_of[es] = active ? e : -1;
  for (int i = tid; i < N3; i += blockDim.x) { int p = perm_g[i]; perm[i] = p; tinv[p] = i; }
  __syncthreads();

  // ---- gather (coalesced over the CTA's elements), stored order -> tensor order ----
  for (int idx = tid; idx < EB * N3; idx += blockDim.x) {
    const int s2 = idx / N3, j = idx % N3; const long long e2 = elem_of[s2];
    if (e2 >= 0) smem[(size_t)s2 * ELEM + tinv[j]] = u[e2 * N3 + j];
  }
  __syncthreads();

  double* U = smem + (size_t)es * ELEM;
  element_integrals<N, Integrands>(T, box, I, perm, u, active, lc, e, lt, U, U + N3, U + 2 * N3);

  // ---- write w_K once (tensor order -> stored order), optionally w = A u - b ----
  for (int idx = tid; idx < EB * N3; idx += blockDim.x) {
    const int s2 = idx / N3, j = idx % N3; const long long e2 = elem_of[s2];
    if (e2 >= 0) {
      double val = smem[(size_t)s2 * ELEM + N3 + tinv[j]] * out_scale;      // out_scale: inverse mass of MOLGalerkinOperator (1 otherwise)
      if (bvec) val -= bvec[e2 * N3 + j];
      w[e2 * N3 + j] = val;
    }
  }
}

}  // namespace b200fem
