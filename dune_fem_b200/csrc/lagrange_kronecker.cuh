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
#include "vec_types.hpp"
#include <type_traits>
#include <utility>

namespace b200fem {

template <class F, int... I> __device__ __forceinline__ void lagk_static_for_impl(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F> __device__ __forceinline__ void lagk_static_for(F&& f) { lagk_static_for_impl(f, std::make_integer_sequence<int, N>{}); }

#ifndef LAGK_PF
#define LAGK_PF 1      // prefetch distance in planes: 1 needs no spills at 64 registers and measured fastest (284 vs 306 us at 3, C3)
#endif
template <int K, int HYP = 16> struct LagKronCfg {
  static constexpr int W = 2 * K + 1, HX = 32, HY = HYP, TX = HX - 2 * K, TY = HY - 2 * K, kThreads = HX * HY;
  static constexpr int kCtasPerSm = HYP <= 16 ? 2 : 1;
  static constexpr int kMaxSeg = 128;                        // planes per z-segment (their z-rows are staged in shared memory)
};

// dmask / dvals (optional): DirichletWrapperOperator fused into the store, w_d = u_d - g_d on constrained nodes
// (schemes/dirichletwrapper.hh:101-105); only valid when no halo exchange follows (single rank).
//
// Schedule per lattice plane (ONE __syncthreads): the z-pass of plane z+1 writes the other half of the double-buffered
// (a, b) planes while the y-pass of plane z reads this half; the x-pass exchanges (c, s) between the lanes of a warp with
// shuffles (a warp is one lattice row), so it needs neither shared memory nor a barrier.
template <int K, bool MAPPED, int HYP>
__global__ void __launch_bounds__(LagKronCfg<K, HYP>::kThreads, LagKronCfg<K, HYP>::kCtasPerSm)
lagrange_kronecker_kernel(const __grid_constant__ LagrangeLayoutDev L, const __grid_constant__ LagKronRows R,
                          const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                          const unsigned char* __restrict__ dmask, const double* __restrict__ dvals,
                          const int tiles_x, const int tiles_y, const int zseg, double* __restrict__ dot_partial) {
  using Cfg = LagKronCfg<K, HYP>;
  constexpr int W = Cfg::W, HX = Cfg::HX, HY = Cfg::HY, TX = Cfg::TX, TY = Cfg::TY;
  __shared__ double Sa[2][HY][HX], Sb[2][HY][HX];
  __shared__ double Zr[Cfg::kMaxSeg][2 * W];                  // z-rows of this segment: M then T (uniform per plane)
  __shared__ double Yr[HY][2 * W];                            // y-rows of this tile (uniform per warp)
  __shared__ double Xr[HX][2 * W + 1];                        // x-rows of this tile (per lane; odd stride: conflict-free)

  const int tid = threadIdx.x, hx = tid % HX, hy = tid / HX;
  const int tile = blockIdx.x % (tiles_x * tiles_y), seg = blockIdx.x / (tiles_x * tiles_y);
  const int L0 = (int)L.lattice[0], L1 = (int)L.lattice[1], L2 = (int)L.lattice[2];
  const int gx = (tile % tiles_x) * TX - K + hx, gy0 = (tile / tiles_x) * TY - K, gy = gy0 + hy;
  const bool in_xy = gx >= 0 && gx < L0 && gy >= 0 && gy < L1;
  const int z0 = seg * zseg, z1 = min(L2, z0 + zseg);

  for (int i = tid; i < (z1 - z0) * W; i += Cfg::kThreads) {
    const int p = i / W, j = i % W;
    Zr[p][j] = R.M[2][(size_t)(z0 + p) * W + j]; Zr[p][W + j] = R.T[2][(size_t)(z0 + p) * W + j];
  }
  if (tid < HY * W) {
    const int r = tid / W, j = tid % W, g = gy0 + r; const bool ok = g >= 0 && g < L1;
    Yr[r][j] = ok ? R.M[1][(size_t)g * W + j] : 0.0; Yr[r][W + j] = ok ? R.T[1][(size_t)g * W + j] : 0.0;
  } else if (tid >= 256 && tid < 256 + HX * W) {   // (HY * W <= 120 < 256 <= kThreads - HX * W)
    const int r = (tid - 256) / W, j = (tid - 256) % W, g = gx - hx + r; const bool ok = g >= 0 && g < L0;
    Xr[r][j] = ok ? R.M[0][(size_t)g * W + j] : 0.0; Xr[r][W + j] = ok ? R.T[0][(size_t)g * W + j] : 0.0;
  }

  // dof address of lattice node (gx, gy, gz): base_p + stride_p * (gz >> zshift), p = parity class of gz (closed-form
  // YaspGrid numbering: 8 parity classes, each a dense array), or lattice_map[linear lattice index].  Addresses are not
  // recomputed per plane: the load stream (planes z0-K, z0-K+1, ...) and the store stream (planes z0, z0+1, ...) each keep
  // two 32-bit cursors, one per parity of the sequence index, advanced by the class stride after use -- the NW-fold
  // unrolled march (NW even) knows the parity at compile time.  (Address arithmetic was > 50 % of all instructions.)
  int base0 = 0, base1 = 0, stride0 = 0, stride1 = 0, zshift = 0;
  if (MAPPED) { base0 = base1 = gx + L0 * gy; stride0 = stride1 = L0 * L1; }
  else if (in_xy) {
    if (L.order == 2) {
      const int s0 = (gx & 1) | ((gy & 1) << 1), s1 = s0 | 4;
      base0 = (int)(L.group_offset[s0] + (gx >> 1) + L.group_dims[s0][0] * (long long)(gy >> 1)); stride0 = (int)(L.group_dims[s0][0] * L.group_dims[s0][1]);
      base1 = (int)(L.group_offset[s1] + (gx >> 1) + L.group_dims[s1][0] * (long long)(gy >> 1)); stride1 = (int)(L.group_dims[s1][0] * L.group_dims[s1][1]);
      zshift = 1;
    } else {
      base0 = base1 = (int)(L.group_offset[0] + gx + L.group_dims[0][0] * (long long)gy);
      stride0 = stride1 = (int)(L.group_dims[0][0] * L.group_dims[0][1]);
    }
  }
  auto slot_of = [&](int gz) -> int { return (zshift & gz) ? base1 + stride1 * (gz >> zshift) : base0 + stride0 * (gz >> zshift); };
  auto step_of = [&](int gz) -> int { return (zshift & gz) ? stride1 : stride0; };      // advance of a cursor over two planes (one plane if zshift == 0)
  // cursors: [q] serves the planes whose sequence index has parity q; with zshift == 0 consecutive planes are one stride
  // apart, so a cursor that is used every other plane advances by two strides
  const int mul = zshift ? 1 : 2;
  int ld[2] = {slot_of(z0 - K), slot_of(z0 - K + 1)}, ldstep[2] = {mul * step_of(z0 - K), mul * step_of(z0 - K + 1)};
  int sc[2] = {slot_of(z0), slot_of(z0 + 1)}, scstep[2] = {mul * step_of(z0), mul * step_of(z0 + 1)};
  auto dof_at = [&](int slot) -> long long { return MAPPED ? L.lattice_map[slot] : (long long)slot; };

  const bool y_owned = hy >= K && hy < HY - K && gy < L1;    // warp-uniform (gy >= 0 follows from hy >= K)
  const bool x_owned = y_owned && hx >= K && hx < HX - K && gx < L0;

  // Register RING of z-neighbours: slot (p - (z0 - K)) mod NW holds u(gx, gy, p) for the W planes of the current window and
  // the PF planes that are prefetched ahead.  The march is unrolled NW-fold so that every slot index is a compile-time
  // constant: no register moves.  (The first version shifted a window through registers; ncu showed the compiler placing
  // the move of the freshly loaded value at the end of the SAME plane step, i.e. every step waited for its own global
  // load -- long_scoreboard was the top stall and the prefetch distance had no effect.)  The Dirichlet mask byte of a
  // node travels with its u value.
  constexpr int PF = LAGK_PF, NW = W + PF;
  static_assert(NW % 2 == 0, "the parity of a plane's sequence index must be a compile-time constant of the unrolled march");
  double ring[NW]; int mring[NW];
  // loads u (and the mask byte) of plane gz through load cursor q and advances the cursor
  auto load_plane = [&](int q, int gz, double& uv, int& mv) {
    const int slot = ld[q]; ld[q] += ldstep[q];
    uv = 0.0; mv = 0;
    if (in_xy && (unsigned)gz < (unsigned)L2) { const long long g = dof_at(slot); uv = u[g]; if (dmask) mv = (int)dmask[g]; }
  };
#pragma unroll
  for (int i = 0; i < NW; ++i) load_plane(i & 1, z0 - K + i, ring[i], mring[i]);
  __syncthreads();                                           // rows staged

  // z-pass of plane zc (ring phase r = (zc - z0) mod NW) into buffer nb; afterwards slot r is dead and takes plane zc - K + NW
  auto z_pass = [&](auto rc, int zc, int nb) {
    constexpr int r = decltype(rc)::value;
    const double* zr = Zr[zc - z0];
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const double v = ring[(r + j) % NW];
      if (j & 1) { a1 = fma(zr[j], v, a1); b1 = fma(zr[W + j], v, b1); } else { a0 = fma(zr[j], v, a0); b0 = fma(zr[W + j], v, b0); }
    }
    Sa[nb][hy][hx] = a0 + a1; Sb[nb][hy][hx] = b0 + b1;
    load_plane(r & 1, zc - K + NW, ring[r], mring[r]);      // sequence index (zc - z0) + NW has the parity of r
  };
  z_pass(std::integral_constant<int, 0>{}, z0, 0);
  __syncthreads();

  double dacc = 0.0;                                         // <u, w> over the nodes this thread stores (CG: <p, A p> without a second sweep)
  auto step = [&](auto rc, int z) {
    constexpr int r = decltype(rc)::value;
    const int cb = (z - z0) & 1;
    // operands of the store, requested now, used at the end of the step
    long long g = 0; double bq = 0, dq = 0;
    const double uc = ring[(r + K) % NW];                    // u(gx, gy, z)
    const bool constrained = mring[(r + K) % NW] != 0;
    { const int slot = sc[r & 1]; sc[r & 1] += scstep[r & 1]; if (x_owned) g = dof_at(slot); }
    if (x_owned) {
      if (bvec) bq = bvec[g];
      if (constrained && dvals) dq = dvals[g];
    }
    if (z + 1 < z1) z_pass(std::integral_constant<int, (r + 1) % NW>{}, z + 1, cb ^ 1);
    if (y_owned) {
      const double* yr = Yr[hy];
      double c = 0, s0 = 0, s1 = 0;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const double aj = Sa[cb][hy - K + j][hx], bj = Sb[cb][hy - K + j][hx];
        c = fma(yr[j], aj, c); s0 = fma(yr[W + j], aj, s0); s1 = fma(yr[j], bj, s1);
      }
      const double s = s0 + s1;
      double r0 = 0, r1 = 0;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const int src = (hx - K + j) & 31;
        r0 = fma(Xr[hx][W + j], __shfl_sync(0xffffffffu, c, src), r0); r1 = fma(Xr[hx][j], __shfl_sync(0xffffffffu, s, src), r1);
      }
      if (x_owned) { const double val = constrained ? uc - dq : (r0 + r1) - bq; w[g] = val; dacc = fma(uc, val, dacc); }
    }
    __syncthreads();
  };
  for (int zb = z0; zb < z1; zb += NW) lagk_static_for<NW>([&](auto rc) { const int z = zb + decltype(rc)::value; if (z < z1) step(rc, z); });
  // optional fused scalar product <u, w>: one partial per CTA, summed in block order by cg_alpha_partials_kernel -- deterministic
  if (dot_partial) { const double t = block_sum(dacc); if (tid == 0) dot_partial[blockIdx.x] = t; }
}

}  // namespace b200fem
