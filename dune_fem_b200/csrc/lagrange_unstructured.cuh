// lagrange_unstructured.cuh -- GalerkinOperator::evaluate for continuous Lagrange spaces (order 1, 2) on UNSTRUCTURED conforming
// cube meshes: element -> dof index arrays and per-element geometry instead of the closed forms of the Cartesian kernels.
//
// Per element, what the reference loop does (dune/fem/schemes/galerkin.hh:811-917 without intersections):
//   getLocalDofs through the mapper's index array (space/mapper/indexsetdofmapper.hh:414-427; the array is what
//     DofMapperCode compiles to, space/lagrange/dofmappercode.hh:56-104)
//   interior integral (galerkin.hh:332-360): evaluateAll / jacobianAll with the TABULATED basis
//     (space/shapefunctionset/caching.hh:302-319), gradients through jacobianInverseTransposed of the element's multilinear map
//     (basisfunctionset/default.hh:239, 266; transformation.hh:35-45), weight * integrationElement (galerkin.hh:353), axpy
//   addLocalDofs (discretefunction.hh:929-934)
// Layout: a CTA takes EB elements of ONE colour (elements of a colour share no dof: plain read-modify-write, deterministic --
// the colour order is the summation order); NB threads per element, thread t is quadrature point t in the evaluation phase and
// basis function t in the axpy phase (NB = NQ for the (k+1)-point Gauss rule).  The tabulation lives in global memory in both
// orientations so that either phase reads it coalesced; element dofs, vertex coordinates and the weighted integrand values are
// staged in shared memory.  Algorithmic traffic per element: 4 NB B of indices + 24 * 2^dim B of coordinates on top of the
// 16 B/dof of the structured kernels (SURVEY.md 8d reports this separately from the headline).
#pragma once
#ifndef __CUDACC_RTC__       // (also compiled at run time by NVRTC for user-supplied integrands, jit.cu)
#include <cuda_runtime.h>
#include <cstdint>
#endif
#include "integrands.cuh"

namespace b200fem {

struct UnstructuredTabDev {
  const double* Bq;    // [i * NQ + q]       phi_i(x_q)            (threads = points)
  const double* Gq;    // [(d * NB + i) * NQ + q]  d phi_i / d xi_d
  const double* Bi;    // [q * NB + i]                              (threads = basis functions)
  const double* Gi;    // [(d * NQ + q) * NB + i]
  const double* xq;    // [q * 3 + d]  Gauss points on the reference cube
  const double* wq;    // [q]
};

template <int DIM, int NB> struct UnstructuredCfg {
  static constexpr int NV = 1 << DIM, EB = (128 / NB) > 0 ? 128 / NB : 1, kThreads = EB * NB;
  __host__ __device__ static constexpr unsigned long long smem_bytes() { return sizeof(double) * (unsigned long long)EB * (NB + 3 * NV + 4 * NB); }
};

template <int DIM, int NB, class Integrands>
__global__ void __launch_bounds__(UnstructuredCfg<DIM, NB>::kThreads)
lagrange_unstructured_kernel(const UnstructuredTabDev T, const __grid_constant__ Integrands I, const int* __restrict__ elem_order, const int* __restrict__ elem_dofs,
                             const double* __restrict__ elem_x, const double* __restrict__ u, double* __restrict__ w, const int first, const int count) {
  using Cfg = UnstructuredCfg<DIM, NB>;
  constexpr int NV = Cfg::NV, EB = Cfg::EB, NQ = NB;
  extern __shared__ __align__(16) unsigned char ust_smem[];
  double* U = reinterpret_cast<double*>(ust_smem);        // [EB][NB]
  double* X = U + EB * NB;                                 // [EB][NV][3]
  double* R = X + EB * NV * 3;                             // [EB][NQ][4]: weighted s, J^-1 F
  const int tid = threadIdx.x, es = tid / NB, t = tid % NB;
  const int slot = blockIdx.x * EB + es;
  const bool active = slot < count;
  const int e = active ? elem_order[first + slot] : 0;
  int dof = 0;
  if (active) {
    dof = elem_dofs[(size_t)e * NB + t];
    U[es * NB + t] = u[dof];
    for (int i = t; i < NV * 3; i += NB) X[es * NV * 3 + i] = elem_x[(size_t)e * NV * 3 + i];
  }
  __syncthreads();
  if (active) {
    // ---- thread = quadrature point t: geometry, evaluateAll / jacobianAll, integrand
    const double xi[3] = {T.xq[3 * t], T.xq[3 * t + 1], T.xq[3 * t + 2]};
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, x[3] = {0, 0, 0};
    const double* Xe = X + es * NV * 3;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double N = 1, dN[3] = {1, 1, 1};
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const double a = ((v >> d) & 1) ? xi[d] : 1.0 - xi[d], da = ((v >> d) & 1) ? 1.0 : -1.0;
        N *= a;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dN[k] *= k == d ? da : a;
      }
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        const double xv = Xe[3 * v + i];
        x[i] = fma(N, xv, x[i]);
#pragma unroll
        for (int d = 0; d < DIM; ++d) J[i][d] = fma(dN[d], xv, J[i][d]);
      }
    }
    double Ji[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, det;
    if (DIM == 2) {
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      const double id = 1.0 / det;
      Ji[0][0] = J[1][1] * id; Ji[0][1] = -J[0][1] * id; Ji[1][0] = -J[1][0] * id; Ji[1][1] = J[0][0] * id;
    } else {
      det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      const double id = 1.0 / det;
      Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    }
    double uq = 0, gh[3] = {0, 0, 0};
    const double* Ue = U + es * NB;
#pragma unroll 3
    for (int i = 0; i < NB; ++i) {
      const double ui = Ue[i];
      uq = fma(T.Bq[i * NQ + t], ui, uq);
#pragma unroll
      for (int d = 0; d < DIM; ++d) gh[d] = fma(T.Gq[(d * NB + i) * NQ + t], ui, gh[d]);
    }
    PointValue pv; pv.u = uq;
#pragma unroll
    for (int i = 0; i < 3; ++i) { pv.du[i] = 0; if (i < DIM) { for (int d = 0; d < DIM; ++d) pv.du[i] = fma(Ji[d][i], gh[d], pv.du[i]); } }   // J^-T gradhat u
    const PointRange r = I.interior(x, pv);
    const double weight = T.wq[t] * fabs(det);                                                                                   // qp.weight() * integrationElement
    double* Rq = R + (es * NQ + t) * 4;
    Rq[0] = r.s * weight;
#pragma unroll
    for (int d = 0; d < 3; ++d) { double f = 0; if (d < DIM) { for (int i = 0; i < DIM; ++i) f = fma(Ji[d][i], r.F[i], f); } Rq[1 + d] = f * weight; }
  }
  __syncthreads();
  if (active) {
    // ---- thread = basis function t: axpy over the points, then addLocalDofs (the colour guarantees exclusive ownership of the dof)
    double acc = 0;
    const double* Re = R + es * NQ * 4;
#pragma unroll 3
    for (int q = 0; q < NQ; ++q) {
      acc = fma(T.Bi[q * NB + t], Re[4 * q], acc);
#pragma unroll
      for (int d = 0; d < DIM; ++d) acc = fma(T.Gi[(d * NQ + q) * NB + t], Re[4 * q + 1 + d], acc);
    }
    w[dof] += acc;
  }
}

}  // namespace b200fem
