// lagrange_kronecker.cuh -- matrix-free apply for continuous Lagrange Q_k spaces (k = 1, 2) with LINEAR,
// CONSTANT-COEFFICIENT integrands on a uniform Cartesian box, as a sum-factorised lattice stencil.
//
// For such a model the quadrature loop of the reference (dune/fem/schemes/galerkin.hh:332-360, 414-435, element loop
// :811-917, scatter :963-991) computes, exactly up to summation order,
//     A = T_0 (x) M_1 (x) M_2  +  M_0 (x) T_1 (x) M_2  +  M_0 (x) M_1 (x) T_2
// on the lattice of Lagrange nodes, because the basis and the Gauss rule are tensor products and geometry factors are
// constants: M_d is the assembled 1-D mass matrix of axis d and T_d = eps K_d - b_d C_d (+ c M_0 for d = 0, + the
// u-dependent boundary term on the two end nodes), each built from the SAME 1-D tabulations and weights the quadrature
// kernel uses (host: build_lagrange_rows).  They are banded with half-width k.  Rank-local boxes assemble over their own
// elements only, i.e. interface planes hold partial sums exactly like the element loop would leave them, and the
// existing Add exchange (halo.cuh) completes them.
//
// Kernel: a CTA owns a (32-2k) x (16-2k) patch of lattice columns and marches through z.
//   z-pass  every thread of the halo'd 32 x 16 patch keeps the 2k+1 z-neighbours of its column in REGISTERS:
//           a = M_z u,  b = T_z u                                   (no shared memory, u is read exactly once per column)
//   y-pass  c = M_y a,  s = T_y a + M_y b     out of shared memory  (row coefficients are warp-uniform)
//   x-pass  w = T_x c + M_x s                 out of shared memory  (row coefficients per lane, in registers)
// 2 __syncthreads per lattice plane, double-buffered planes.  No atomics, no colouring, one launch: the colour-ordered
// scatter of the generic kernel (8 launches, read-modify-write of w) disappears because every lattice node is written
// exactly once.  Dofs are addressed through the closed-form YaspGrid numbering (8 parity classes, each a dense array)
// or through the lattice->dof table of the adaptive-leaf numbering.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "lagrange_quadrature.cuh"

namespace b200fem {

struct LagKronRows {
  const double* M[3];      // [L_d][2k+1]: row g holds the coefficients of columns g-k .. g+k
  const double* T[3];
};

template <int K> struct LagKronCfg {
  static constexpr int W = 2 * K + 1, HX = 32, HY = 16, TX = HX - 2 * K, TY = HY - 2 * K, kThreads = HX * HY;
};

template <int K, bool MAPPED>
__global__ void __launch_bounds__(LagKronCfg<K>::kThreads)
lagrange_kronecker_kernel(const __grid_constant__ LagrangeLayoutDev L, const __grid_constant__ LagKronRows R,
                          const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                          const int tiles_x, const int tiles_y, const int zseg) {
  using Cfg = LagKronCfg<K>;
  constexpr int W = Cfg::W, HX = Cfg::HX, HY = Cfg::HY, TX = Cfg::TX, TY = Cfg::TY;
  __shared__ double Sa[2][HY][HX], Sb[2][HY][HX], Sc[2][HY][HX], Ss[2][HY][HX];

  const int tid = threadIdx.x, hx = tid % HX, hy = tid / HX;
  const int tile = blockIdx.x % (tiles_x * tiles_y), seg = blockIdx.x / (tiles_x * tiles_y);
  const int L0 = (int)L.lattice[0], L1 = (int)L.lattice[1], L2 = (int)L.lattice[2];
  const int gx = (tile % tiles_x) * TX - K + hx, gy = (tile / tiles_x) * TY - K + hy;
  const bool in_xy = gx >= 0 && gx < L0 && gy >= 0 && gy < L1;
  const int z0 = seg * zseg, z1 = min(L2, z0 + zseg);

  // dof address of lattice node (gx, gy, gz) = base[p] + stride[p] * (gz >> zshift), p = parity class of gz
  long long base[2] = {0, 0}, stride[2] = {0, 0}; int zshift = 0;
  if (!MAPPED && in_xy) {
    if (L.order == 2) {
      const int sxy = (gx & 1) | ((gy & 1) << 1);
      for (int p = 0; p < 2; ++p) {
        const int s = sxy | (p << 2);
        base[p] = L.group_offset[s] + (gx >> 1) + L.group_dims[s][0] * (long long)(gy >> 1);
        stride[p] = L.group_dims[s][0] * L.group_dims[s][1];
      }
      zshift = 1;
    } else {
      base[0] = base[1] = L.group_offset[0] + gx + L.group_dims[0][0] * (long long)gy;
      stride[0] = stride[1] = L.group_dims[0][0] * L.group_dims[0][1];
    }
  }
  auto dof = [&](int gz) -> long long {
    if (MAPPED) return L.lattice_map[gx + (long long)L0 * (gy + (long long)L1 * gz)];
    const int p = zshift ? (gz & 1) : 0;
    return base[p] + stride[p] * (long long)(gz >> zshift);
  };
  auto load_u = [&](int gz) -> double { return (in_xy && gz >= 0 && gz < L2) ? u[dof(gz)] : 0.0; };

  // y-rows are the same for the whole warp (one warp = one hy)
  double ym[W], yt[W];
  const bool y_owned = hy >= K && hy < HY - K && gy < L1;    // (gy >= 0 follows from hy >= K)
#pragma unroll
  for (int j = 0; j < W; ++j) { ym[j] = y_owned ? R.M[1][(size_t)gy * W + j] : 0.0; yt[j] = y_owned ? R.T[1][(size_t)gy * W + j] : 0.0; }
  const bool x_owned = y_owned && hx >= K && hx < HX - K && gx < L0;
  double xm[W], xt[W];                                       // x-rows differ per lane
#pragma unroll
  for (int j = 0; j < W; ++j) { xm[j] = x_owned ? R.M[0][(size_t)gx * W + j] : 0.0; xt[j] = x_owned ? R.T[0][(size_t)gx * W + j] : 0.0; }

  double uw[W];                                              // u(gx, gy, z-K .. z+K)
#pragma unroll
  for (int j = 1; j < W; ++j) uw[j] = load_u(z0 - K + j - 1);
  double pre = load_u(z0 + K);                               // one plane of prefetch distance

  for (int z = z0; z < z1; ++z) {
    const int buf = (z - z0) & 1;
#pragma unroll
    for (int j = 0; j < W - 1; ++j) uw[j] = uw[j + 1];
    uw[W - 1] = pre; pre = load_u(z + K + 1);
    double a = 0, b = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) { a = fma(R.M[2][(size_t)z * W + j], uw[j], a); b = fma(R.T[2][(size_t)z * W + j], uw[j], b); }
    Sa[buf][hy][hx] = a; Sb[buf][hy][hx] = b;
    __syncthreads();
    if (y_owned) {
      double c = 0, s = 0;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const double aj = Sa[buf][hy - K + j][hx], bj = Sb[buf][hy - K + j][hx];
        c = fma(ym[j], aj, c); s = fma(yt[j], aj, s); s = fma(ym[j], bj, s);
      }
      Sc[buf][hy][hx] = c; Ss[buf][hy][hx] = s;
    }
    __syncthreads();
    if (x_owned) {
      double r = 0;
#pragma unroll
      for (int j = 0; j < W; ++j) { r = fma(xt[j], Sc[buf][hy][hx - K + j], r); r = fma(xm[j], Ss[buf][hy][hx - K + j], r); }
      const long long g = dof(z);
      w[g] = bvec ? r - bvec[g] : r;
    }
  }
}

}  // namespace b200fem
