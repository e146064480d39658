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
// basis function t in the axpy phase (NB = NQ for the (k+1)-point Gauss rule).  The tabulation, element dofs, vertex coordinates
// and the weighted integrand values are staged in shared memory.  Algorithmic traffic per element: 4 NB B of indices + 24 * 2^dim B of coordinates on top of the
// 16 B/dof of the structured kernels (SURVEY.md 8d reports this separately from the headline).
#pragma once
#ifndef __CUDACC_RTC__       // (also compiled at run time by NVRTC for user-supplied integrands, jit.cu)
#include <cuda_runtime.h>
#include <cstdint>
#endif
#include "integrands.cuh"

namespace b200fem {

struct UnstructuredTabDev {
  const double* B;     // [i * NQ + q]              phi_i(x_q)
  const double* G;     // [(d * NB + i) * NQ + q]   d phi_i / d xi_d
  const double* xq;    // [q * 3 + d]  Gauss points on the reference cube
  const double* wq;    // [q]
};

template <int DIM, int NB> struct UnstructuredCfg {
  static constexpr int NV = 1 << DIM, NQ = NB;
  static constexpr int EB = NB >= 27 ? 8 : NB >= 8 ? 16 : 32, kThreads = EB * NB;       // elements per batch
  static constexpr int LQ = NQ | 1;                                                     // odd table stride: conflict-free for both phases
  static constexpr int kTab = (1 + DIM) * NB * LQ + 4 * NQ;                             // B, G[DIM], xq[3], wq
  static constexpr int kBatch = EB * (NB + 3 * NV + 4 * NQ);                            // U, X, R of one batch
  __host__ __device__ static constexpr unsigned long long smem_bytes() { return sizeof(double) * (unsigned long long)(kTab + kBatch); }
};

// Second generation: PERSISTENT CTAs.  The tabulation is copied into shared memory once per CTA (odd row stride: the evaluation
// phase reads it with the point as the fast index, the axpy phase with the basis function as the fast index -- both conflict-free)
// and the CTA then loops over batches of EB elements of the colour.  (First generation: one batch per CTA, tables read from
// global memory in every inner-loop step: 63 % of the warp samples waited on those loads, profiles/r02_unstructured.md.)
template <int DIM, int NB, class Integrands>
__global__ void __launch_bounds__(UnstructuredCfg<DIM, NB>::kThreads)
lagrange_unstructured_kernel(const UnstructuredTabDev T, const __grid_constant__ Integrands I, const int* __restrict__ elem_order, const int* __restrict__ elem_dofs,
                             const double* __restrict__ elem_x, const double* __restrict__ u, double* __restrict__ w, const int first, const int count) {
  using Cfg = UnstructuredCfg<DIM, NB>;
  constexpr int NV = Cfg::NV, EB = Cfg::EB, NQ = NB, LQ = Cfg::LQ;
  extern __shared__ __align__(16) unsigned char ust_smem[];
  double* Bs = reinterpret_cast<double*>(ust_smem);       // [NB][LQ]
  double* Gs = Bs + NB * LQ;                               // [DIM][NB][LQ]
  double* Xq = Gs + DIM * NB * LQ;                         // [NQ][3]
  double* Wq = Xq + 3 * NQ;                                // [NQ]
  double* U = Wq + NQ;                                     // [EB][NB]
  double* X = U + EB * NB;                                 // [EB][NV][3]
  double* R = X + EB * NV * 3;                             // [EB][NQ][4]: weighted s, J^-1 F
  const int tid = threadIdx.x, es = tid / NB, t = tid % NB;
  for (int idx = tid; idx < NB * NQ; idx += Cfg::kThreads) {
    const int i = idx / NQ, q = idx % NQ;
    Bs[i * LQ + q] = T.B[idx];
#pragma unroll
    for (int d = 0; d < DIM; ++d) Gs[(d * NB + i) * LQ + q] = T.G[d * NB * NQ + idx];
  }
  for (int idx = tid; idx < 3 * NQ; idx += Cfg::kThreads) Xq[idx] = T.xq[idx];
  for (int idx = tid; idx < NQ; idx += Cfg::kThreads) Wq[idx] = T.wq[idx];
  const int nbatch = (count + EB - 1) / EB;
  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    const int slot = batch * EB + es;
    const bool active = slot < count;
    const int e = active ? elem_order[first + slot] : 0;
    int dof = 0;
    if (active) {
      dof = elem_dofs[(size_t)e * NB + t];
      U[es * NB + t] = u[dof];
      for (int i = t; i < NV * 3; i += NB) X[es * NV * 3 + i] = elem_x[(size_t)e * NV * 3 + i];
    }
    __syncthreads();                                       // (also covers the table copy before the first batch)
    if (active) {
      // ---- thread = quadrature point t: geometry, evaluateAll / jacobianAll, integrand
      const double xi[3] = {Xq[3 * t], Xq[3 * t + 1], Xq[3 * t + 2]};
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
#pragma unroll 9
      for (int i = 0; i < NB; ++i) {
        const double ui = Ue[i];
        uq = fma(Bs[i * LQ + t], ui, uq);
#pragma unroll
        for (int d = 0; d < DIM; ++d) gh[d] = fma(Gs[(d * NB + i) * LQ + t], ui, gh[d]);
      }
      PointValue pv; pv.u = uq;
#pragma unroll
      for (int i = 0; i < 3; ++i) { pv.du[i] = 0; if (i < DIM) { for (int d = 0; d < DIM; ++d) pv.du[i] = fma(Ji[d][i], gh[d], pv.du[i]); } }   // J^-T gradhat u
      const PointRange r = I.interior(x, pv);
      const double weight = Wq[t] * fabs(det);                                                                                     // qp.weight() * integrationElement
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
#pragma unroll 9
      for (int q = 0; q < NQ; ++q) {
        acc = fma(Bs[t * LQ + q], Re[4 * q], acc);
#pragma unroll
        for (int d = 0; d < DIM; ++d) acc = fma(Gs[(d * NB + t) * LQ + q], Re[4 * q + 1 + d], acc);
      }
      w[dof] += acc;
    }
  }
}

}  // namespace b200fem
