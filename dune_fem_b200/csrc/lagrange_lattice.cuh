// lagrange_lattice.cuh -- matrix-free apply for continuous Lagrange Q_k spaces (k = 1, 2) with LINEAR,
// CONSTANT-COEFFICIENT integrands on a uniform Cartesian box, as a sum-factorised lattice stencil (second generation).
//
// For such a model the quadrature loop of the reference (dune/fem/schemes/galerkin.hh:332-360, 414-435, element loop
// :811-917, scatter :963-991) computes, exactly up to summation order,
//     A = T_0 (x) M_1 (x) M_2  +  M_0 (x) T_1 (x) M_2  +  M_0 (x) M_1 (x) T_2
// on the lattice of Lagrange nodes: M_d is the assembled 1-D mass matrix of axis d and T_d = eps K_d - b_d C_d (+ c M_0 for
// d = 0, + the u-dependent boundary term on the two end nodes), banded with half-width k, built from the SAME 1-D
// tabulations and weights the quadrature kernel uses.  Rank-local boxes assemble over their own elements only, i.e.
// interface planes hold partial sums exactly like the element loop would leave them, and the Add exchange completes them.
//
// What changed against the first generation (lagrange_kronecker.cuh, 9.5 warp instructions per dof, 0.15 of the HBM
// roofline: per-row coefficient tables in shared memory, one node per thread, 34 % of the threads pure halo):
//  * ROWS ARE OF TWO TYPES ONLY.  An assembled row depends on the node type (element vertex / element-interior node) and,
//    on the first and last lattice plane of the box, lacks one element contribution.  Nodes outside the box read as zero,
//    so an end row is the interior row plus a correction of its DIAGONAL entry (M_lo = -Me[k][k], M_hi = -Me[0][0], same for
//    T plus the boundary term).  All coefficients are therefore 2 x (2k+1) numbers per axis in the constant bank: no
//    coefficient tables, no shared-memory loads for coefficients, FMAs with constant operands.
//  * FOUR NODES PER THREAD along x.  The x-pass exchanges only two neighbour values per side and field by shuffles (8 SHFL.32
//    per node instead of 20), the node type along x is a compile-time constant (tiles start at even lattice coordinates),
//    global accesses are pairs of neighbouring dofs of the same parity class, and a thread carries four independent FMA chains.
//  * A row of the tile is LX lanes (4 LX nodes) wide, a warp holds 32 / LX rows: 64 x 32 node tiles (LX = 16) have 82 % useful
//    threads instead of 66 %, and fit lattices like 257^3 with 9 % padding.
//  * Dirichlet marks are computed from the lattice coordinates (a node is constrained iff it lies on a masked side of the
//    global box: dirichletconstraints.hh:435-554 in closed form) -- no mask bytes travel with u.
//
// Kernel: a CTA owns a tile of lattice columns and marches through z.
//   z-pass  every thread keeps the 2k+1 z-neighbours of its four columns in registers (a window shifted once per plane):
//           a = M_z u,  b = T_z u                                   (no shared memory, u is read once per column)
//   y-pass  c = M_y a,  s = T_y a + M_y b     neighbours' (a, b) out of shared memory, 16-byte accesses
//   x-pass  w = T_x c + M_x s                 neighbours' (c, s) by warp shuffles
// ONE __syncthreads per lattice plane (double-buffered (a, b) planes).  No atomics, no colouring, one launch, every lattice
// node written exactly once; the Dirichlet wrapper (w_d = u_d - g_d, schemes/dirichletwrapper.hh:101-105), the load vector
// and the scalar product <u, w> of CG ride along in the store.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include <utility>
#include "lagrange_quadrature.cuh"
#include "vec_types.hpp"

namespace b200fem {

template <class F, int... I> __device__ __forceinline__ void lat_static_for_impl(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F> __device__ __forceinline__ void lat_static_for(F&& f) { lat_static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// coefficients of the three 1-D operators; offsets -k..k in slots 0..2k; type 0 = element vertex, 1 = element-interior node
template <int K> struct LagStencilDev {
  static constexpr int W = 2 * K + 1;
  double M[3][2][W], T[3][2][W];
  double Mlo[3], Mhi[3], Tlo[3], Thi[3];   // diagonal corrections on the first / last lattice plane of the LOCAL box
  int glo[3], gend[3];                     // global lattice coordinate of local node 0; last global lattice coordinate (k * gn)
  int dirichlet_bits;                      // sides whose nodes are constrained (bit 2*axis+side), 0: no fused Dirichlet wrapper
  int affine;                              // 1: constrained rows store u - g (g from dvals), 0: they store u (homogeneous part)
};

template <int K, int LX, int WARPS = 16> struct LagLatCfg {
  static constexpr int W = 2 * K + 1, R = 4, kWarps = WARPS, kThreads = 32 * kWarps, kCtasPerSm = 16 / WARPS;
  static constexpr int RPW = 32 / LX;                        // tile rows per warp
  static constexpr int HY = kWarps * RPW, NX = R * LX;       // tile extents incl. halo (rows, nodes per row)
  static constexpr int TXO = NX - 2 * K, TYO = HY - 2 * K;   // nodes a tile produces
  static constexpr size_t smem_bytes() { return sizeof(double) * 2 * (size_t)HY * 2 * NX; }
  static_assert(32 % LX == 0 && TXO % 2 == 0 && TYO % 2 == 0, "tiles start on even lattice coordinates: node types are fixed per register slot / warp");
};

// DATA = false: neither a load vector nor Dirichlet values are streamed (the homogeneous apply A u of the Krylov loops): their
// eight operand registers are compiled out, which is what keeps this instantiation free of spills at the 128-register cap
template <int K, int LX, bool MAPPED, int WARPS, bool DATA>
__global__ void __launch_bounds__(LagLatCfg<K, LX, WARPS>::kThreads, LagLatCfg<K, LX, WARPS>::kCtasPerSm)
lagrange_lattice_kernel(const __grid_constant__ LagrangeLayoutDev L, const __grid_constant__ LagStencilDev<K> S,
                        const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec, const double* __restrict__ dvals,
                        const int tiles_x, const int tiles_y, const int zseg, double* __restrict__ dot_partial) {
  using Cfg = LagLatCfg<K, LX, WARPS>;
  constexpr int W = Cfg::W, R = Cfg::R, HY = Cfg::HY, NX = Cfg::NX;
  extern __shared__ __align__(16) unsigned char lat_smem[];
  double* const AB = reinterpret_cast<double*>(lat_smem);      // [buffer][row][field][NX]
  auto plane_ptr = [&](int buf, int row, int field) { return AB + (((size_t)buf * HY + row) * 2 + field) * NX; };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lx = lane % LX, row = warp + Cfg::kWarps * (lane / LX);      // rows of a warp are kWarps apart: same parity
  const int tile = blockIdx.x % (tiles_x * tiles_y), seg = blockIdx.x / (tiles_x * tiles_y);
  const int L0 = (int)L.lattice[0], L1 = (int)L.lattice[1], L2 = (int)L.lattice[2];
  const int gx0 = (tile % tiles_x) * Cfg::TXO - K + R * lx;             // lattice x of this thread's first node
  const int gy = (tile / tiles_x) * Cfg::TYO - K + row;
  const bool row_ok = gy >= 0 && gy < L1;
  const int z0 = seg * zseg, z1 = min(L2, z0 + zseg);
  const int ty = K == 2 ? (gy & 1) : 0;                                 // node type of this row (warp-uniform)

  // which of the four nodes exist / are produced by this thread
  unsigned xin = 0, xout = 0;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int gx = gx0 + j;
    if (gx >= 0 && gx < L0 && row_ok) xin |= 1u << j;
    if (gx >= 0 && gx < L0 && row_ok && R * lx + j >= K && R * lx + j < NX - K && row >= K && row < HY - K) xout |= 1u << j;
  }
  // Dirichlet marks in closed form: a node is constrained iff it lies on a masked side of the GLOBAL box
  const int db = S.dirichlet_bits;
  unsigned xdir = 0;                                                    // per node: constrained through its x or y coordinate
  {
    const int gyg = S.glo[1] + gy;
    const bool ydir = (gyg == 0 && (db & 4)) || (gyg == S.gend[1] && (db & 8));
#pragma unroll
    for (int j = 0; j < R; ++j) { const int gxg = S.glo[0] + gx0 + j; if (ydir || (gxg == 0 && (db & 1)) || (gxg == S.gend[0] && (db & 2))) xdir |= 1u << j; }
  }

  // dof addresses.  Closed-form YaspGrid numbering: 8 parity classes (k = 2), each a dense array; the thread's nodes 0, 2 (even x)
  // are neighbours in one class, nodes 1, 3 (odd x) in another: dof = base[class] + stride[class] * (gz >> 1) (+1 for the second
  // node).  k = 1: one class, four consecutive dofs.  Adaptive-leaf numbering: lattice -> dof table.  Addresses are not
  // recomputed per plane: the load stream and the store stream each keep a CURSOR for the plane they touch next and one for the
  // plane after it (the other z-parity class); a cursor is advanced by its class stride after use and the two swap roles.
  struct LatCursor { int e, o, se, so; };
  auto make_cursor = [&](const int gz) -> LatCursor {
    LatCursor c;
    if (MAPPED) { c.e = gx0 + L0 * (gy + L1 * gz); c.o = 0; c.se = 2 * L0 * L1; c.so = 0; }
    else if (K == 2) {
      const int xe = gx0 >> 1, pz = gz & 1;                             // (gx0 is even; arithmetic shifts keep the halo consistent)
      const int s0 = ((gy & 1) << 1) | (pz << 2), s1 = s0 | 1;
      c.se = (int)(L.group_dims[s0][0] * L.group_dims[s0][1]); c.so = (int)(L.group_dims[s1][0] * L.group_dims[s1][1]);
      c.e = (int)(L.group_offset[s0] + xe + L.group_dims[s0][0] * (long long)(gy >> 1)) + c.se * (gz >> 1);
      c.o = (int)(L.group_offset[s1] + xe + L.group_dims[s1][0] * (long long)(gy >> 1)) + c.so * (gz >> 1);
    } else {
      const int st = (int)(L.group_dims[0][0] * L.group_dims[0][1]);
      c.e = (int)(L.group_offset[0] + gx0 + L.group_dims[0][0] * (long long)gy) + st * gz; c.o = 0; c.se = 2 * st; c.so = 0;
    }
    return c;
  };
  auto node_dof = [&](const LatCursor& c, const int j) -> int {
    if (MAPPED) return (int)L.lattice_map[c.e + j];
    return K == 2 ? ((j & 1) ? c.o : c.e) + (j >> 1) : c.e + j;
  };
  auto advance = [&](LatCursor& a, LatCursor& b) { a.e += a.se; a.o += a.so; const LatCursor t = a; a = b; b = t; };
  LatCursor ld_a = make_cursor(z0 - K), ld_b = make_cursor(z0 - K + 1);   // load stream: planes z0-K, z0-K+1, ...
  LatCursor st_a = make_cursor(z0), st_b = make_cursor(z0 + 1);           // store stream: planes z0, z0+1, ...

  // z-neighbours of the thread's four columns in registers: win[j][t] = u(., ., z - K + t), t = 0 .. 2K, for the plane z whose
  // z-pass comes next, plus the plane after the window, which is loaded one whole plane step ahead of its use (nxt).  The window
  // is shifted by register moves once per plane (the march loop is NOT unrolled: the first generation's ring-fold unrolling
  // made 190 KB of code; the moves are 8 per node against ~70 other instructions).
  double win[R][W], nxt[R];
  auto load_plane = [&](const int gz, double (&dst)[R]) {          // (planes are requested in ascending order, one per call)
    const bool zok = (unsigned)gz < (unsigned)L2;
#pragma unroll
    for (int j = 0; j < R; ++j) dst[j] = (zok && ((xin >> j) & 1u)) ? u[node_dof(ld_a, j)] : 0.0;
    advance(ld_a, ld_b);
  };
#pragma unroll
  for (int t = 0; t < W; ++t) {
    double tmp[R]; load_plane(z0 - K + t, tmp);
#pragma unroll
    for (int j = 0; j < R; ++j) win[j][t] = tmp[j];
  }
  load_plane(z0 + K + 1, nxt);

  // z-pass of plane zc (the window is centred on it) into buffer nb
  auto z_pass = [&](const int zc, const int nb) {
    double a[R], b[R];
    const int tz = K == 2 ? (zc & 1) : 0;
#pragma unroll
    for (int j = 0; j < R; ++j) { a[j] = 0.0; b[j] = 0.0; }
    if (K == 2 && tz) {                                                  // element-interior plane: three non-zero coefficients
#pragma unroll
      for (int t = 1; t < W - 1; ++t) {
        const double cm = S.M[2][1][t], ct = S.T[2][1][t];
#pragma unroll
        for (int j = 0; j < R; ++j) { a[j] = fma(cm, win[j][t], a[j]); b[j] = fma(ct, win[j][t], b[j]); }
      }
    } else {
#pragma unroll
      for (int t = 0; t < W; ++t) {
        const double cm = S.M[2][0][t], ct = S.T[2][0][t];
#pragma unroll
        for (int j = 0; j < R; ++j) { a[j] = fma(cm, win[j][t], a[j]); b[j] = fma(ct, win[j][t], b[j]); }
      }
    }
    if (zc == 0 || zc == L2 - 1) {                                       // first / last plane of the box: diagonal corrections
      const double cm = (zc == 0 ? S.Mlo[2] : 0.0) + (zc == L2 - 1 ? S.Mhi[2] : 0.0), ct = (zc == 0 ? S.Tlo[2] : 0.0) + (zc == L2 - 1 ? S.Thi[2] : 0.0);
#pragma unroll
      for (int j = 0; j < R; ++j) { a[j] = fma(cm, win[j][K], a[j]); b[j] = fma(ct, win[j][K], b[j]); }
    }
    // A row of a plane is stored as two halves: the lanes' node pairs (0, 1) side by side, then their pairs (2, 3).  16-byte
    // accesses of neighbouring lanes are then 16 bytes apart (a lane's four nodes in one 32-byte piece put every other lane of a
    // quarter-warp on the same banks: half of the y-pass wavefronts were conflicts, profiles/r02_lattice.md)
    double* pa = plane_ptr(nb, row, 0) + 2 * lx; double* pb = plane_ptr(nb, row, 1) + 2 * lx;
    *reinterpret_cast<double2*>(pa) = make_double2(a[0], a[1]); *reinterpret_cast<double2*>(pa + NX / 2) = make_double2(a[2], a[3]);
    *reinterpret_cast<double2*>(pb) = make_double2(b[0], b[1]); *reinterpret_cast<double2*>(pb + NX / 2) = make_double2(b[2], b[3]);
  };
  z_pass(z0, 0);
  __syncthreads();

  const bool y_owned = row >= K && row < HY - K && gy < L1;              // (gy >= 0 follows from row >= K)
  const unsigned ymask = __ballot_sync(0xffffffffu, y_owned);            // lanes that take part in the x-pass shuffles (whole tile rows)
  const bool ylo = gy == 0, yhi = gy == L1 - 1;
  double dacc = 0.0;                                                     // <u, w> over the nodes this thread stores (CG: <p, A p> without a second sweep)
#pragma unroll 1
  for (int z = z0; z < z1; ++z) {
    const int cb = (z - z0) & 1;
    // operands of the store, requested now, used at the end of the step
    double uc[R], bq[R], dq[R]; int g[R];                                 // (dof indices fit 32 bits: checked by the launcher)
    const int gzg = S.glo[2] + z;
    const bool zdir = (gzg == 0 && (db & 16)) || (gzg == S.gend[2] && (db & 32));
    const unsigned cons = zdir ? 0xfu : xdir;                            // constrained nodes of this thread on this plane
#pragma unroll
    for (int j = 0; j < R; ++j) {                                        // (straight-line, predicated: no divergent blocks)
      const bool out = (xout >> j) & 1u;
      uc[j] = win[j][K]; g[j] = out ? node_dof(st_a, j) : 0;
      if (DATA) {
        bq[j] = (out && bvec != nullptr) ? bvec[g[j]] : 0.0;
        dq[j] = (out && ((cons >> j) & 1u) && S.affine) ? dvals[g[j]] : 0.0;
      } else { bq[j] = 0.0; dq[j] = 0.0; }
    }
    advance(st_a, st_b);
    if (z + 1 < z1) {
      // shift the window to plane z+1 (the plane that enters was requested a whole step ago), request the plane after it,
      // and do the z-pass of plane z+1 into the other buffer
#pragma unroll
      for (int j = 0; j < R; ++j) {
#pragma unroll
        for (int t = 0; t < W - 1; ++t) win[j][t] = win[j][t + 1];
        win[j][W - 1] = nxt[j];
      }
      load_plane(z + K + 2, nxt);
      z_pass(z + 1, cb ^ 1);
    }
    if (y_owned) {
      // ---- y-pass: own row from shared memory too (the registers of the z-pass are long gone), neighbours by 16-byte loads
      double c[R], s[R];
#pragma unroll
      for (int j = 0; j < R; ++j) { c[j] = 0.0; s[j] = 0.0; }
      auto y_term = [&](const int t, const double cm, const double ct) {
        const double2 a01 = *reinterpret_cast<const double2*>(plane_ptr(cb, row - K + t, 0) + 2 * lx), a23 = *reinterpret_cast<const double2*>(plane_ptr(cb, row - K + t, 0) + 2 * lx + NX / 2);
        const double2 b01 = *reinterpret_cast<const double2*>(plane_ptr(cb, row - K + t, 1) + 2 * lx), b23 = *reinterpret_cast<const double2*>(plane_ptr(cb, row - K + t, 1) + 2 * lx + NX / 2);
        const double av[R] = {a01.x, a01.y, a23.x, a23.y}, bv[R] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int j = 0; j < R; ++j) { c[j] = fma(cm, av[j], c[j]); s[j] = fma(ct, av[j], s[j]); s[j] = fma(cm, bv[j], s[j]); }
      };
      if (K == 2 && ty) {
#pragma unroll
        for (int t = 1; t < W - 1; ++t) y_term(t, S.M[1][1][t], S.T[1][1][t]);
      } else {
#pragma unroll
        for (int t = 0; t < W; ++t) y_term(t, S.M[1][0][t], S.T[1][0][t]);
      }
      if (ylo || yhi) y_term(K, (ylo ? S.Mlo[1] : 0.0) + (yhi ? S.Mhi[1] : 0.0), (ylo ? S.Tlo[1] : 0.0) + (yhi ? S.Thi[1] : 0.0));
      // ---- x-pass: two neighbour nodes per side and field by shuffles (the lanes of a tile row are consecutive)
      double ce[R + 2 * K], se[R + 2 * K];
#pragma unroll
      for (int j = 0; j < R; ++j) { ce[K + j] = c[j]; se[K + j] = s[j]; }
#pragma unroll
      for (int q = 0; q < K; ++q) {
        ce[q] = __shfl_up_sync(ymask, c[R - K + q], 1); se[q] = __shfl_up_sync(ymask, s[R - K + q], 1);
        ce[K + R + q] = __shfl_down_sync(ymask, c[q], 1); se[K + R + q] = __shfl_down_sync(ymask, s[q], 1);
      }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int tx = K == 2 ? (j & 1) : 0;                               // compile-time node type (tiles start at even x)
        double r0 = 0.0, r1 = 0.0;
#pragma unroll
        for (int t = 0; t < W; ++t) {
          if (K == 2 && tx && (t == 0 || t == W - 1)) continue;
          r0 = fma(S.T[0][tx][t], ce[j + t], r0); r1 = fma(S.M[0][tx][t], se[j + t], r1);
        }
        const int gx = gx0 + j;
        if (gx == 0 || gx == L0 - 1) {
          r0 = fma((gx == 0 ? S.Tlo[0] : 0.0) + (gx == L0 - 1 ? S.Thi[0] : 0.0), c[j], r0);
          r1 = fma((gx == 0 ? S.Mlo[0] : 0.0) + (gx == L0 - 1 ? S.Mhi[0] : 0.0), s[j], r1);
        }
        const double val = ((cons >> j) & 1u) ? uc[j] - dq[j] : (r0 + r1) - bq[j];
        if ((xout >> j) & 1u) w[g[j]] = val;
        dacc = fma(((xout >> j) & 1u) ? uc[j] : 0.0, val, dacc);
      }
    }
    __syncthreads();
  }
  // optional fused scalar product <u, w>: one partial per CTA, summed in block order by cg_alpha_partials_kernel -- deterministic
  if (dot_partial) { const double t = block_sum(dacc); if (tid == 0) dot_partial[blockIdx.x] = t; }
}

}  // namespace b200fem
