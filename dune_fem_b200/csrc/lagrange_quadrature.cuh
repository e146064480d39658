// lagrange_quadrature.cuh -- generic matrix-free apply for continuous Lagrange Q_k spaces (k = 1, 2) on cubes.
//
// Same per-element work as the DG kernel (element_integrals), but local->global maps follow
// IndexSetDofMapper::mapEach (dune/fem/space/mapper/indexsetdofmapper.hh:414-427): dof blocks ordered by geometry
// type (vertices, edges, faces, cells; :504-515), index inside a block = indexSet.subIndex.  On a Cartesian box the
// YaspGrid index is closed-form, so no index arrays are read; the AdaptiveLeafIndexSet numbering
// (gridpart/adaptiveleafindexset.hh:884-906) uses a precomputed lattice->dof table instead.
//
// Scatter: the reference guards addLocalDofs with a shared_mutex (galerkin.hh:963-991).  Here elements are split
// into 2^dim colours (parity of the element coordinates); elements of one colour share no dof, so each colour is
// one launch doing plain read-modify-write -- deterministic, no atomics (colour order = summation order).
#pragma once
#ifndef __CUDACC_RTC__       // (also compiled at run time by NVRTC for user-supplied integrands, jit.cu)
#include <cuda_runtime.h>
#include <cstdint>
#endif
#include "dg_quadrature.cuh"

namespace b200fem {

// 1-D tables of the (k+1)-point rule for the register-only 2-D kernel
template <int N>
struct DgTabDev {
  double B[N * N], G[N * N];     // B[q*N+i] = phi_i(x_q), G = phi_i'(x_q)   (volume and face rules coincide)
  double x[N], w[N];             // Gauss points / weights on [0,1]
  double phi[2][N], dphi[2][N];  // traces at 0 and 1
};

struct LagrangeLayoutDev {
  int order;                        // 1, 2 or 3
  long long group_offset[8];        // per "shift" bit set (directions the sub-entity extends in)
  long long group_dims[8][3];
  long long lattice[3];             // lattice extents k*n+1
  const long long* lattice_map;     // optional: lattice index -> dof (adaptive-leaf numbering), else null
};

__host__ __device__ inline long long lagrange_dof(const LagrangeLayoutDev& L, const long long g0, const long long g1, const long long g2) {
  if (L.lattice_map) return L.lattice_map[g0 + L.lattice[0] * (g1 + L.lattice[1] * g2)];
  int s = 0; long long c0 = g0, c1 = g1, c2 = g2;
  if (L.order == 2) { s = (int)(g0 & 1) | ((int)(g1 & 1) << 1) | ((int)(g2 & 1) << 2); c0 >>= 1; c1 >>= 1; c2 >>= 1; }
  if (L.order >= 3) {
    // k - 1 nodes inside an edge, (k-1)^2 inside a face, (k-1)^3 inside a cell: block = offset[type] + numDofs(entity) * entity index
    // + j (space/mapper/indexsetdofmapper.hh:414-427), j = the node's position inside its entity, lower axes fastest (the local
    // numbering of the Lagrange points, genericlagrangepoints.hh:862-876; Cartesian grids: no twists, lagrange/space.hh:68-71)
    const int k = L.order; const int r0 = (int)(g0 % k), r1 = (int)(g1 % k), r2 = (int)(g2 % k);
    c0 = g0 / k; c1 = g1 / k; c2 = g2 / k;
    s = (r0 != 0 ? 1 : 0) | (r1 != 0 ? 2 : 0) | (r2 != 0 ? 4 : 0);
    int j = 0, nd = 1;
    if (r0) { j += nd * (r0 - 1); nd *= k - 1; }
    if (r1) { j += nd * (r1 - 1); nd *= k - 1; }
    if (r2) { j += nd * (r2 - 1); nd *= k - 1; }
    return L.group_offset[s] + nd * (c0 + L.group_dims[s][0] * (c1 + L.group_dims[s][1] * c2)) + j;
  }
  return L.group_offset[s] + c0 + L.group_dims[s][0] * (c1 + L.group_dims[s][1] * c2);
}

// ------------------------------------------------------------------ 3-D: N*N threads per element, shared-memory tensors
// (R > 1: range-R spaces, dof (node, component) = node * R + c; the components sit in neighbouring element slots, dg_quadrature.cuh)
template <int N, class Integrands, int R = 1>
__global__ void __launch_bounds__(DgQuadCfg<N, N, N>::kThreads)
lagrange3d_quadrature_kernel(const __grid_constant__ QuadTabDev<N, N, N> T, const __grid_constant__ BoxDev box,
                             const __grid_constant__ Integrands I, const __grid_constant__ LagrangeLayoutDev L,
                             const double* __restrict__ u, double* __restrict__ w,
                             int c0, int c1, int c2, int m0, int m1, long long n_colour) {
  using Cfg = DgQuadCfg<N, N, N>;
  constexpr int N2 = N * N, N3 = N * N * N, EB = Cfg::EB, ELEM = Cfg::kElemDoubles, LN = Cfg::LN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  int* ecs = reinterpret_cast<int*>(smem + (size_t)EB * ELEM);       // 4 ints per slot: coords + active flag

  constexpr int RS = quad_pow2(R);
  static_assert(R == 1 || RS <= EB, "dimRange: the components of an element must fit the element slots of a CTA");
  const int tid = threadIdx.x, es = Cfg::slot(tid), lt = Cfg::lane(tid);
  const int comp = es % RS;
  const unsigned cmask = R == 1 ? 0xffffffffu : (((1u << RS) - 1u) << ((tid & 31) / RS * RS));
  const long long oe = (long long)blockIdx.x * (EB / RS) + es / RS;
  const bool active = lt < Cfg::T2 && oe < n_colour;
  int lc[3] = {0, 0, 0};
  if (active) {
    lc[0] = box.own_lo[0] + 2 * (int)(oe % m0) + c0;
    lc[1] = box.own_lo[1] + 2 * (int)((oe / m0) % m1) + c1;
    lc[2] = box.own_lo[2] + 2 * (int)(oe / ((long long)m0 * m1)) + c2;
  }
  const long long e = lc[0] + (long long)box.n[0] * (lc[1] + (long long)box.n[1] * lc[2]);
  if (lt == 0) { ecs[4 * es] = lc[0]; ecs[4 * es + 1] = lc[1]; ecs[4 * es + 2] = lc[2]; ecs[4 * es + 3] = active && comp < R; }
  __syncthreads();
  const int k = N - 1;
  for (int idx = tid; idx < EB * N3; idx += blockDim.x) {            // gather (getLocalDofs)
    const int s2 = idx / N3, t = idx % N3;
    if (ecs[4 * s2 + 3]) {
      const int i0 = t / N2, i1 = (t / N) % N, i2 = t % N;
      smem[(size_t)s2 * ELEM + (i0 * N + i1) * LN + i2] = u[lagrange_dof(L, (long long)k * ecs[4 * s2] + i0, (long long)k * ecs[4 * s2 + 1] + i1, (long long)k * ecs[4 * s2 + 2] + i2) * R + (R == 1 ? 0 : s2 % RS)];
    }
  }
  __syncthreads();
  double* U = smem + (size_t)es * ELEM;
  element_integrals<N, N, N, Integrands, false, R>(T, box, I, nullptr, u, active, lc, e, lt, U, U + Cfg::kU, U + 2 * Cfg::kU, comp, cmask);
  for (int idx = tid; idx < EB * N3; idx += blockDim.x) {            // coloured scatter-add (addLocalDofs)
    const int s2 = idx / N3, t = idx % N3;
    if (ecs[4 * s2 + 3]) {
      const int i0 = t / N2, i1 = (t / N) % N, i2 = t % N;
      const long long g = lagrange_dof(L, (long long)k * ecs[4 * s2] + i0, (long long)k * ecs[4 * s2 + 1] + i1, (long long)k * ecs[4 * s2 + 2] + i2);
      w[g * R + (R == 1 ? 0 : s2 % RS)] += smem[(size_t)s2 * ELEM + Cfg::kU + (i0 * N + i1) * LN + i2];
    }
  }
}

// ------------------------------------------------------------------ 2-D: one thread per element, registers only
// (range-R spaces: the thread carries all R components of its element; scalar integrands go through the same code with R = 1)
template <int R, class Integrands> __device__ __forceinline__ PointRangeV<R> lag2d_interior(const Integrands& I, const double* x, const PointValueV<R>& v) {
  if constexpr (R == 1) {
    PointValue pv; pv.u = v.u[0]; pv.du[0] = v.du[0][0]; pv.du[1] = v.du[0][1]; pv.du[2] = v.du[0][2];
    const PointRange r = I.interior(x, pv);
    PointRangeV<1> o; o.s[0] = r.s; o.F[0][0] = r.F[0]; o.F[0][1] = r.F[1]; o.F[0][2] = r.F[2]; return o;
  } else return I.interior(x, v);
}
template <int R, class Integrands> __device__ __forceinline__ PointRangeV<R> lag2d_boundary(const Integrands& I, int d, int s, double ih, const double* x, const PointValueV<R>& v) {
  if constexpr (R == 1) {
    PointValue pv; pv.u = v.u[0]; pv.du[0] = v.du[0][0]; pv.du[1] = v.du[0][1]; pv.du[2] = v.du[0][2];
    const PointRange r = I.boundary(d, s, ih, x, pv);
    PointRangeV<1> o; o.s[0] = r.s; o.F[0][0] = r.F[0]; o.F[0][1] = r.F[1]; o.F[0][2] = r.F[2]; return o;
  } else return I.boundary(d, s, ih, x, v);
}

template <int N, class Integrands, int R = 1>
__global__ void __launch_bounds__(128)
lagrange2d_quadrature_kernel(const __grid_constant__ DgTabDev<N> T, const __grid_constant__ BoxDev box,
                             const __grid_constant__ Integrands I, const __grid_constant__ LagrangeLayoutDev L,
                             const double* __restrict__ u, double* __restrict__ w, int c0, int c1, int m0, long long n_colour) {
  const long long oe = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (oe >= n_colour) return;
  constexpr int k = N - 1;
  const int lc[2] = {box.own_lo[0] + 2 * (int)(oe % m0) + c0, box.own_lo[1] + 2 * (int)(oe / m0) + c1};
  const double hh[2] = {box.h[0], box.h[1]};
  const double detJ = hh[0] * hh[1];
  double ul[R][N][N], wl[R][N][N];            // [c][i1][i0]
  long long dof[N][N];
#pragma unroll
  for (int i1 = 0; i1 < N; ++i1)
#pragma unroll
    for (int i0 = 0; i0 < N; ++i0) {
      dof[i1][i0] = lagrange_dof(L, (long long)k * lc[0] + i0, (long long)k * lc[1] + i1, 0);
#pragma unroll
      for (int c = 0; c < R; ++c) { ul[c][i1][i0] = u[dof[i1][i0] * R + c]; wl[c][i1][i0] = 0; }
    }

  // interior integral (galerkin.hh:332-360), sum-factorised in registers
  double tb[R][N][N], tg[R][N][N];            // [c][q0][i1]
#pragma unroll
  for (int c = 0; c < R; ++c)
#pragma unroll
    for (int q0 = 0; q0 < N; ++q0)
#pragma unroll
      for (int i1 = 0; i1 < N; ++i1) { double a = 0, b = 0;
#pragma unroll
        for (int i0 = 0; i0 < N; ++i0) { a = fma(T.B[q0 * N + i0], ul[c][i1][i0], a); b = fma(T.G[q0 * N + i0], ul[c][i1][i0], b); }
        tb[c][q0][i1] = a; tg[c][q0][i1] = b; }
  double zs[R][N][N], zx[R][N][N];            // [c][q0][i1] after testing along axis 1
#pragma unroll
  for (int q0 = 0; q0 < N; ++q0) {
    double rs[R][N], rx[R][N], ry[R][N];
#pragma unroll
    for (int q1 = 0; q1 < N; ++q1) {
      PointValueV<R> pv;
#pragma unroll
      for (int c = 0; c < R; ++c) {
        pv.u[c] = 0; pv.du[c][0] = pv.du[c][1] = pv.du[c][2] = 0;
#pragma unroll
        for (int i1 = 0; i1 < N; ++i1) { pv.u[c] = fma(T.B[q1 * N + i1], tb[c][q0][i1], pv.u[c]); pv.du[c][0] = fma(T.B[q1 * N + i1], tg[c][q0][i1], pv.du[c][0]); pv.du[c][1] = fma(T.G[q1 * N + i1], tb[c][q0][i1], pv.du[c][1]); }
        pv.du[c][0] /= hh[0]; pv.du[c][1] /= hh[1];
      }
      double xq[3] = {box.lo[0] + hh[0] * ((box.origin[0] + lc[0]) + T.x[q0]), box.lo[1] + hh[1] * ((box.origin[1] + lc[1]) + T.x[q1]), 0.0};
      const PointRangeV<R> r = lag2d_interior<R>(I, xq, pv);
      const double wq = T.w[q0] * T.w[q1] * detJ;
#pragma unroll
      for (int c = 0; c < R; ++c) { rs[c][q1] = r.s[c] * wq; rx[c][q1] = r.F[c][0] * wq / hh[0]; ry[c][q1] = r.F[c][1] * wq / hh[1]; }
    }
#pragma unroll
    for (int c = 0; c < R; ++c)
#pragma unroll
      for (int i1 = 0; i1 < N; ++i1) { double a = 0, b = 0;
#pragma unroll
        for (int q1 = 0; q1 < N; ++q1) { a = fma(T.B[q1 * N + i1], rs[c][q1], a); a = fma(T.G[q1 * N + i1], ry[c][q1], a); b = fma(T.B[q1 * N + i1], rx[c][q1], b); }
        zs[c][q0][i1] = a; zx[c][q0][i1] = b; }
  }
#pragma unroll
  for (int c = 0; c < R; ++c)
#pragma unroll
    for (int i1 = 0; i1 < N; ++i1)
#pragma unroll
      for (int i0 = 0; i0 < N; ++i0) { double a = 0;
#pragma unroll
        for (int q0 = 0; q0 < N; ++q0) { a = fma(T.B[q0 * N + i0], zs[c][q0][i1], a); a = fma(T.G[q0 * N + i0], zx[c][q0][i1], a); }
        wl[c][i1][i0] += a; }

  // boundary integrals (galerkin.hh:414-435) on domain-boundary edges
  if (I.m.has_boundary) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int d = f >> 1, s = f & 1, a = 1 - d;
      const int gc = box.origin[d] + lc[d];
      if ((s == 0 && gc != 0) || (s == 1 && gc != box.gn[d] - 1)) continue;
      double tv[R][N], td[R][N];               // trace coefficients along the edge (index along axis a)
#pragma unroll
      for (int c = 0; c < R; ++c)
#pragma unroll
        for (int ia = 0; ia < N; ++ia) { double v = 0, dv = 0;
#pragma unroll
          for (int id = 0; id < N; ++id) { const double uu = d == 0 ? ul[c][ia][id] : ul[c][id][ia]; v = fma(T.phi[s][id], uu, v); dv = fma(T.dphi[s][id], uu, dv); }
          tv[c][ia] = v; td[c][ia] = dv; }
      const double area = detJ / hh[d];
      double rv[R][N], rd[R][N];
#pragma unroll
      for (int c = 0; c < R; ++c)
#pragma unroll
        for (int ia = 0; ia < N; ++ia) rv[c][ia] = rd[c][ia] = 0;
#pragma unroll
      for (int q = 0; q < N; ++q) {
        PointValueV<R> pv;
#pragma unroll
        for (int c = 0; c < R; ++c) {
          pv.u[c] = 0; pv.du[c][0] = pv.du[c][1] = pv.du[c][2] = 0; double dn = 0, dt = 0;
#pragma unroll
          for (int ia = 0; ia < N; ++ia) { pv.u[c] = fma(T.B[q * N + ia], tv[c][ia], pv.u[c]); dn = fma(T.B[q * N + ia], td[c][ia], dn); dt = fma(T.G[q * N + ia], tv[c][ia], dt); }
          pv.du[c][d] = dn / hh[d]; pv.du[c][a] = dt / hh[a];
        }
        double xq[3] = {0, 0, 0};
        xq[d] = box.lo[d] + hh[d] * (gc + s); xq[a] = box.lo[a] + hh[a] * ((box.origin[a] + lc[a]) + T.x[q]);
        const PointRangeV<R> r = lag2d_boundary<R>(I, d, s, 1.0 / hh[d], xq, pv);
        const double wq = T.w[q] * area;
#pragma unroll
        for (int c = 0; c < R; ++c)
#pragma unroll
          for (int ia = 0; ia < N; ++ia) {
            rv[c][ia] = fma(T.B[q * N + ia], r.s[c] * wq, rv[c][ia]); rv[c][ia] = fma(T.G[q * N + ia], r.F[c][a] * wq / hh[a], rv[c][ia]);
            rd[c][ia] = fma(T.B[q * N + ia], r.F[c][d] * wq / hh[d], rd[c][ia]);
          }
      }
#pragma unroll
      for (int c = 0; c < R; ++c)
#pragma unroll
        for (int ia = 0; ia < N; ++ia)
#pragma unroll
          for (int id = 0; id < N; ++id) {
            const double add = T.phi[s][id] * rv[c][ia] + T.dphi[s][id] * rd[c][ia];
            if (d == 0) wl[c][ia][id] += add; else wl[c][id][ia] += add;
          }
    }
  }
#pragma unroll
  for (int i1 = 0; i1 < N; ++i1)
#pragma unroll
    for (int i0 = 0; i0 < N; ++i0)
#pragma unroll
      for (int c = 0; c < R; ++c) w[dof[i1][i0] * R + c] += wl[c][i1][i0];
}

}  // namespace b200fem
