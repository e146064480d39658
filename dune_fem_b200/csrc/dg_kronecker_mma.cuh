// dg_kronecker_mma.cuh -- Kronecker-form DG apply for Q3 (n = 4 Legendre modes per axis) on the FP64 tensor cores.
//
// Same operator as dg_kronecker_slab.cuh (linear constant-coefficient models on uniform boxes):
//     w_K = sum_d [ S_d u_K + L_d u_{K-e_d} + R_d u_{K+e_d} ] - b_K          (1-D operators from kron_tables.hpp)
// restated as small matrix products that fit mma.sync.m8n8k4.f64 EXACTLY when n = 4 (no padding of the contraction index):
//   * a warp owns a 2 x 2 patch of elements in (x, y); for a fixed z-mode c the patch's 64 outputs form an 8 x 8 accumulator
//     tile  D[(ex, a)][(ey, b)]   (a, b: x- and y-mode; ex, ey: element inside the patch),
//   * x-axis:  D += T_x (8 x 16) * U (16 x 8): the block-Toeplitz operator [L S R 0; 0 L S R] as the A operand (four k-steps =
//     the four source elements x-1 .. x+2), the data as the B operand (k = x-mode of the source element, n = (ey, b)),
//   * y-axis:  D += U (8 x 16) * T_y^T (16 x 8): the data as the A operand (rows (ex, a), k = y-mode of the source element), the
//     operator as B -- so both axes accumulate into the SAME fragment with no transposition in between,
//   * z-axis:  the z-mode c is the fragment index: 4 x 4 scalar coefficients applied fragment-wise with plain DFMAs (constant-bank
//     operands).  The kernel marches through z: the data of plane z is used once for plane z (S_z), for plane z-1 (R_z) and for
//     plane z+1 (L_z), whose accumulators wait in registers -- z-neighbours are never re-read and only ONE plane of u lives in
//     shared memory.
// 8 DMMA + 24 DFMA per element and lane replace the slab kernel's 72 DFMA: the math needs a third of the issue slots.
// Per plane step a CTA (8 warps, 16 patches = an 8 x 8 column of elements) stages the 10 x 10 elements of the plane incl. the x/y halo
// into a swizzled tensor-order layout (element stride 66, row stride 11 elements, (a, b) index XOR-swizzled: every fragment load is a
// conflict-free 16-byte access), computes, and leaves the finished plane z-1 through a staging buffer with coalesced 16-byte stores
// (hierarchical dof order and the load vector b are applied there).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "dg_kronecker.cuh"
#include "kron_common.cuh"

namespace b200fem {

// TY_ rows of elements per column tile, two 2 x 2 patches per warp: TY_ = 8 -> 8 warps, one CTA per SM; TY_ = 4 -> 4 warps, TWO CTAs
// per SM that run out of phase (one CTA's shared-memory-bound re-layout passes under the other's tensor-core phase)
template <int TY_> struct KronMmaCfgT {
  static constexpr int N = 4, N3 = 64, TX = 8, TY = TY_, kWarps = TX * TY_ / 8, kThreads = 32 * kWarps, kCtasPerSm = 8 / kWarps;
  static constexpr int PX = TX + 2, PY = TY + 2;          // plane incl. halo
  static constexpr int ES = 66, RS = 11;                  // element stride (doubles), row stride (elements): see the header
  static constexpr int kLand = PX * PY * N3;              // TMA landing buffer: the plane in stored order, dense
  static constexpr int kOut = TX * TY * N3;
  static constexpr int kPlane = (PY * RS * ES + 15) / 16 * 16;   // doubles (128-byte multiple)
  static constexpr int kQ = ((TY - 1) * RS + TX) * ES;           // finished plane in the fragment layout (owned elements only)
  // [landing | out (the load-vector tile lands here first) | plane | q | 2 mbarriers]
  static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)(kLand + kOut + kPlane + kQ) + 16 + 128; }
  // offset of (a, b, c = 0 | 2) inside an element: 16-byte chunks, chunk = (c >> 1) * 16 + ((a * 4 + b) ^ ((a >> 1) << 1))
  __host__ __device__ static constexpr int eoff(int a, int b, int chalf) { return 2 * (chalf * 16 + ((a * 4 + b) ^ ((a >> 1) << 1))); }
};

// tensor maps of one launch: u over the local box [z][y][x][64] (box 64 x 10 x 10 x 1, out-of-bounds = zero = missing neighbour),
// w and b over the OWNED sub-box (box 64 x 8 x 8 x 1: stores are clipped to the owned range by the TMA unit)
using KronMmaCfg = KronMmaCfgT<4>;

struct KronMmaMaps { CUtensorMap u_plane, w_tile, b_tile; };
// The stored (hierarchical) dof order is a permutation of the tensor order the fragments use.  Both re-layout passes (landing buffer
// -> fragment layout, finished plane -> stored order) move one 8-byte word per lane between a dense element (bank = dof % 16) and
// the swizzled element (bank = offset % 16): `dof[i]` assigns dofs to lanes such that every half-warp touches 16 different banks on
// BOTH sides (the 64 (source bank, destination bank) pairs form a 4-regular bipartite multigraph; it splits into four perfect
// matchings, one per half-warp of the two passes -- computed on the host).  off[i] = offset of dof[i] inside a swizzled element.
struct KronMmaOrder { int dof[64], off[64]; };

__device__ __forceinline__ void dmma884(double (&d)[2], const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <bool HAS_B, int TY_>
__global__ void __launch_bounds__(KronMmaCfgT<TY_>::kThreads, KronMmaCfgT<TY_>::kCtasPerSm)
dg_kronecker_mma_kernel(const __grid_constant__ KronTabDev<4> K, const __grid_constant__ BoxDev box, const __grid_constant__ KronMmaMaps M,
                        const __grid_constant__ KronMmaOrder O, const int tx, const int ty, long long* __restrict__ dbg) {
  using Cfg = KronMmaCfgT<TY_>;
  constexpr int N = 4, N3 = 64, ES = Cfg::ES, RS = Cfg::RS;
  extern __shared__ unsigned char mma_smem_raw[];
  // TMA wants 128-byte alignment.  (An offset added to the shared symbol keeps the accesses LDS / STS; a round trip of the pointer
  // through an integer would turn every one of them into a generic LD / ST.)
  unsigned char* const mma_smem = mma_smem_raw + ((128u - (ptx::smem_addr(mma_smem_raw) & 127u)) & 127u);
  double* const LND = reinterpret_cast<double*>(mma_smem);          // landing buffer of the TMA load: plane in stored order
  double* const OUT = LND + Cfg::kLand;                              // finished plane: [oy][ox][stored order], leaves by TMA store;
                                                                     // the load-vector tile of that plane is loaded into it beforehand
  double* const P = OUT + Cfg::kOut;                                 // plane z of u: [ey'][ex' (RS)][swizzled tensor order (ES)]
  double* const Q = P + Cfg::kPlane;                                 // finished plane, fragment layout: [oy][ox (RS)][swizzled (ES)]
  const uint32_t bar_l = ptx::smem_addr(Q + Cfg::kQ), bar_b = bar_l + 8;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  // the two dofs this lane moves in the re-layout passes (see KronMmaOrder)
  const int rl_j0 = O.dof[lane], rl_j1 = O.dof[32 + lane], rl_o0 = O.off[lane], rl_o1 = O.off[32 + lane];

  // ---- operator fragments (constant over the kernel) ----
  // x-axis, A operand: row g = (ex, a), column t = a'  of  [L S R 0; 0 L S R]  for the four source elements s = 0 .. 3
  // y-axis, B operand: row t = b', column g = (ey, b)  of the transposed Toeplitz operator
  double ax[4], by[4];
  {
    const int e = g >> 2, m = g & 3;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      ax[s] = s == e ? K.L[0][m * N + t] : s == e + 1 ? K.S[0][m * N + t] : s == e + 2 ? K.R[0][m * N + t] : 0.0;
      by[s] = s == e ? K.L[1][m * N + t] : s == e + 1 ? K.S[1][m * N + t] : s == e + 2 ? K.R[1][m * N + t] : 0.0;
    }
  }
  const double dlo_x = K.Dlo[0][(g & 3) * N + t], dhi_x = K.Dhi[0][(g & 3) * N + t];     // boundary corrections of S (same fragment roles)
  const double dlo_y = K.Dlo[1][(g & 3) * N + t], dhi_y = K.Dhi[1][(g & 3) * N + t];

  const int on0 = box.own_hi[0] - box.own_lo[0], on1 = box.own_hi[1] - box.own_lo[1], on2 = box.own_hi[2] - box.own_lo[2];
  // Work split.  Enough columns for every CTA: whole columns, dealt round-robin -- the CTAs of a wave then march through z in
  // lockstep over NEIGHBOURING columns, so the x/y halo rows two of them share are read from DRAM once and hit the L2 the second
  // time.  Fewer columns than CTAs: the (column, z) plane steps are split evenly and a column is shared by several CTAs.
  const int ncols = tx * ty;
  const bool whole_columns = ncols >= (int)gridDim.x;
  const long long total = whole_columns ? (long long)((ncols - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * on2 : (long long)ncols * on2;
  long long s0 = whole_columns ? 0 : total * blockIdx.x / gridDim.x;
  const long long s1 = whole_columns ? total : total * (blockIdx.x + 1) / gridDim.x;
  if (tid == 0) { ptx::mbar_init(bar_l, 1); ptx::mbar_init(bar_b, 1); ptx::fence_barrier_init(); ptx::prefetch_tensormap(&M.u_plane); ptx::prefetch_tensormap(&M.w_tile); if (HAS_B) ptx::prefetch_tensormap(&M.b_tile); }
  unsigned n_l = 0, n_b = 0;                       // completed waits on the two barriers (their phase parities)
  long long tph[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();   // diagnostics (dbg != nullptr): cycles per phase seen by one thread
  auto stamp = [&](int i) { if (dbg) { const long long now = clock64(); tph[i] += now - tlast; tlast = now; } };
  __syncthreads();

  while (s0 < s1) {
    const int col = whole_columns ? (int)blockIdx.x + (int)(s0 / on2) * (int)gridDim.x : (int)(s0 / on2);
    const int za = (int)(s0 % on2), zb = (int)min((long long)on2, za + (s1 - s0));
    s0 += zb - za;
    const int x0 = box.own_lo[0] + (col % tx) * Cfg::TX, y0 = box.own_lo[1] + (col / tx) * Cfg::TY;   // local coordinates of the column's first element
    // accumulators of the planes z-1 (m), z (0), z+1 (p) for the two patches of this warp: [patch][c][2]
    double accm[2][4][2], acc0[2][4][2], accp[2][4][2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int c = 0; c < 4; ++c) { accm[q][c][0] = accm[q][c][1] = acc0[q][c][0] = acc0[q][c][1] = accp[q][c][0] = accp[q][c][1] = 0.0; }

    const int zfirst = box.own_lo[2] + za - 1, zlast = box.own_lo[2] + zb;
    const int ox0 = x0 - box.own_lo[0], oy0 = y0 - box.own_lo[1];          // tile origin in the owned sub-box (w / b tensor maps)
    if (tid == 0) {
      ptx::mbar_expect_tx(bar_l, 8u * Cfg::kLand); ptx::tma_load_4d(ptx::smem_addr(LND), &M.u_plane, 0, x0 - 1, y0 - 1, zfirst, bar_l);
      if (HAS_B) { ptx::mbar_expect_tx(bar_b, 8u * Cfg::kOut); ptx::tma_load_4d(ptx::smem_addr(OUT), &M.b_tile, 0, ox0, oy0, za, bar_b); }
    }
    for (int zl = zfirst; zl <= zlast; ++zl) {
      // ---- plane zl (with its x/y halo; elements outside the local box arrive as zeros = missing neighbours) has landed in stored
      //      order: re-lay it into the swizzled tensor order the fragment loads want ----
      stamp(5);
      ptx::mbar_wait(bar_l, n_l & 1u); ++n_l;
      stamp(0);
      {
        // warp w re-lays the elements pe = w, w + 8, ...: two 8-byte words per lane and element, conflict-free on both sides
        constexpr int kItems = (Cfg::PX * Cfg::PY + Cfg::kWarps - 1) / Cfg::kWarps;
        double v0[kItems], v1[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) { const int pe = warp + Cfg::kWarps * k; if (pe < Cfg::PX * Cfg::PY) { v0[k] = LND[pe * N3 + rl_j0]; v1[k] = LND[pe * N3 + rl_j1]; } }
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
          const int pe = warp + Cfg::kWarps * k, ey = pe / Cfg::PX, ex = pe - ey * Cfg::PX;
          if (pe < Cfg::PX * Cfg::PY) { double* const pel = P + (ey * RS + ex) * ES; pel[rl_o0] = v0[k]; pel[rl_o1] = v1[k]; }
        }
      }
      stamp(1);
      if (tid == 0) ptx::bulk_wait_read();            // the previous plane's TMA store has read OUT
      __syncthreads();
      stamp(2);
      if (tid == 0 && zl < zlast) {                   // next plane of u, under this plane's arithmetic
        ptx::fence_proxy_async();
        ptx::mbar_expect_tx(bar_l, 8u * Cfg::kLand); ptx::tma_load_4d(ptx::smem_addr(LND), &M.u_plane, 0, x0 - 1, y0 - 1, zl + 1, bar_l);
      }

      const bool own_plane = zl >= box.own_lo[2] + za && zl < box.own_lo[2] + zb;
      const int gz = box.origin[2] + zl;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int pidx = warp * 2 + q, pxi = pidx & 3, pyi = pidx >> 2;
        if (own_plane) {
          // ---- x-axis: data as B operand.  lane (g, t): source element (2 pxi + s, 2 pyi + ey + 1), a' = t, b = g & 3, ey = g >> 2
          {
            double bx[4][4];
            const double* base = P + ((2 * pyi + (g >> 2) + 1) * RS + 2 * pxi) * ES + Cfg::eoff(t, g & 3, 0);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const double2 v0 = *reinterpret_cast<const double2*>(base + s * ES), v1 = *reinterpret_cast<const double2*>(base + s * ES + 32);
              bx[s][0] = v0.x; bx[s][1] = v0.y; bx[s][2] = v1.x; bx[s][3] = v1.y;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int s = 0; s < 4; ++s) dmma884(acc0[q][c], ax[s], bx[s][c]);
            // elements on the domain boundary in x: S -> S + Dlo / Dhi (rows of element ex, source = the element itself)
            const int gx0 = box.origin[0] + x0 + 2 * pxi;
            if (gx0 == 0 || gx0 + 1 >= box.gn[0] - 1) {
              const int gxe = gx0 + (g >> 2);
              const double cx = (gxe == 0 ? dlo_x : 0.0) + (gxe == box.gn[0] - 1 ? dhi_x : 0.0);
              const double c0 = (g >> 2) == 0 ? cx : 0.0, c1 = (g >> 2) == 1 ? cx : 0.0;
#pragma unroll
              for (int c = 0; c < 4; ++c) { dmma884(acc0[q][c], c0, bx[1][c]); dmma884(acc0[q][c], c1, bx[2][c]); }
            }
          }
          // ---- y-axis: data as A operand.  lane (g, t): source element (2 pxi + ex + 1, 2 pyi + s), a = g & 3, ex = g >> 2, b' = t
          {
            double ay[4][4];
            const double* base = P + ((2 * pyi) * RS + 2 * pxi + (g >> 2) + 1) * ES + Cfg::eoff(g & 3, t, 0);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const double2 v0 = *reinterpret_cast<const double2*>(base + s * RS * ES), v1 = *reinterpret_cast<const double2*>(base + s * RS * ES + 32);
              ay[s][0] = v0.x; ay[s][1] = v0.y; ay[s][2] = v1.x; ay[s][3] = v1.y;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int s = 0; s < 4; ++s) dmma884(acc0[q][c], ay[s][c], by[s]);
            const int gy0 = box.origin[1] + y0 + 2 * pyi;
            if (gy0 == 0 || gy0 + 1 >= box.gn[1] - 1) {
              const int gye = gy0 + (g >> 2);                       // (B operand: column g = (ey, b))
              const double cy = (gye == 0 ? dlo_y : 0.0) + (gye == box.gn[1] - 1 ? dhi_y : 0.0);
              const double c0 = (g >> 2) == 0 ? cy : 0.0, c1 = (g >> 2) == 1 ? cy : 0.0;
#pragma unroll
              for (int c = 0; c < 4; ++c) { dmma884(acc0[q][c], ay[1][c], c0); dmma884(acc0[q][c], ay[2][c], c1); }
            }
          }
        }
        // ---- z-axis: the plane's data in accumulator layout.  lane (g, t): element (2 pxi + ex + 1, 2 pyi + ey + 1), a = g & 3,
        //      ex = g >> 2, columns 2t, 2t+1 -> ey = t >> 1, b = 2 (t & 1) + r
        {
          double uz[2][4];
          const double* base = P + ((2 * pyi + (t >> 1) + 1) * RS + 2 * pxi + (g >> 2) + 1) * ES;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int o = Cfg::eoff(g & 3, 2 * (t & 1) + r, 0);
            const double2 v0 = *reinterpret_cast<const double2*>(base + o), v1 = *reinterpret_cast<const double2*>(base + o + 32);
            uz[r][0] = v0.x; uz[r][1] = v0.y; uz[r][2] = v1.x; uz[r][3] = v1.y;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              double sm = accm[q][c][r], sp = accp[q][c][r];
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) { sm = fma(K.R[2][c * N + cc], uz[r][cc], sm); sp = fma(K.L[2][c * N + cc], uz[r][cc], sp); }
              accm[q][c][r] = sm; accp[q][c][r] = sp;
            }
          if (own_plane) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                double s = acc0[q][c][r];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) s = fma(K.S[2][c * N + cc], uz[r][cc], s);
                acc0[q][c][r] = s;
              }
            if (gz == 0 || gz == box.gn[2] - 1) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                  double s = acc0[q][c][r];
#pragma unroll
                  for (int cc = 0; cc < 4; ++cc) s = fma((gz == 0 ? K.Dlo[2][c * N + cc] : 0.0) + (gz == box.gn[2] - 1 ? K.Dhi[2][c * N + cc] : 0.0), uz[r][cc], s);
                  acc0[q][c][r] = s;
                }
            }
          }
        }
        // ---- plane zl-1 is complete: into Q in the fragment layout (16-byte stores, conflict-free); rotate the accumulators ----
        {
          if (zl - 1 >= box.own_lo[2] + za) {
            double* const qel = Q + ((2 * pyi + (t >> 1)) * RS + 2 * pxi + (g >> 2)) * ES;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int o = Cfg::eoff(g & 3, 2 * (t & 1) + r, 0);
              *reinterpret_cast<double2*>(qel + o) = make_double2(accm[q][0][r], accm[q][1][r]);
              *reinterpret_cast<double2*>(qel + o + 32) = make_double2(accm[q][2][r], accm[q][3][r]);
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) { accm[q][c][r] = acc0[q][c][r]; acc0[q][c][r] = accp[q][c][r]; accp[q][c][r] = 0.0; }
        }
      }
      stamp(3);
      const bool out_plane = zl - 1 >= box.own_lo[2] + za;
      __syncthreads();
      // ---- fragment layout -> stored order (dense, what the TMA store wants), minus the load vector ----
      if (out_plane) {
        if (HAS_B) { ptx::mbar_wait(bar_b, n_b & 1u); ++n_b; }
#pragma unroll
        for (int k = 0; k < Cfg::TX * Cfg::TY / Cfg::kWarps; ++k) {
          const int eo = warp + Cfg::kWarps * k, oy = eo >> 3, ox = eo & 7;
          const double* const qel = Q + (oy * RS + ox) * ES;
          double a0 = qel[rl_o0], a1 = qel[rl_o1];
          if (HAS_B) { a0 -= OUT[eo * N3 + rl_j0]; a1 -= OUT[eo * N3 + rl_j1]; }
          OUT[eo * N3 + rl_j0] = a0; OUT[eo * N3 + rl_j1] = a1;
        }
        ptx::fence_proxy_async();                     // generic writes of OUT -> visible to the TMA store
      }
      __syncthreads();
      stamp(4);
      // ---- plane zl-1 leaves by one TMA store (clipped to the owned range); its successor's load-vector tile is requested ----
      if (tid == 0 && out_plane) {
        const int zo = zl - 1 - box.own_lo[2];
        ptx::tma_store_4d(&M.w_tile, 0, ox0, oy0, zo, ptx::smem_addr(OUT)); ptx::bulk_commit();
        if (HAS_B && zl < zlast) {                    // the next plane's load-vector tile, into OUT once the store has read it
          ptx::bulk_wait_read();
          ptx::mbar_expect_tx(bar_b, 8u * Cfg::kOut); ptx::tma_load_4d(ptx::smem_addr(OUT), &M.b_tile, 0, ox0, oy0, zo + 1, bar_b);
        }
      }
    }
    if (tid == 0) ptx::bulk_wait_read();
    __syncthreads();
  }
  if (tid == 0) ptx::bulk_wait_all();
  if (dbg && blockIdx.x == 1 && (tid == 0 || tid == Cfg::kThreads - 1)) { for (int i = 0; i < 6; ++i) dbg[(tid ? 8 : 0) + i] = tph[i]; dbg[(tid ? 8 : 0) + 6] = n_l; }
  (void)on0; (void)on1;
}

}  // namespace b200fem
