// dg_kronecker.cuh -- DG operator apply for LINEAR, CONSTANT-COEFFICIENT integrands on a uniform Cartesian box.
//
// For such a model the Galerkin operator of dune/fem/schemes/galerkin.hh:811-917 is, exactly (up to rounding),
//     w_K = sum_d [ S_d^{cls} u_K + L_d u_{K-e_d} + R_d u_{K+e_d} ]        (each an n x n matrix acting on axis d)
// because (i) the reference basis is a tensor product, (ii) the Gauss rules integrate the 1-D mass matrices of
// the orthonormal Legendre basis exactly (they are the identity), (iii) geometry factors are constants.
// The 1-D matrices are built on the host from the same 1-D tabulations and quadrature weights the quadrature
// kernel uses (kron_tables.hpp), so both kernels agree to rounding; this one needs 9 n^4 FMA per element instead
// of ~50 n^4 and is bounded by HBM (16 B/dof) rather than by the FP64 pipe.
//
// v1 mapping: one CTA per tile of TX x TY x TZ elements; tile + face halo staged in shared memory by coalesced
// loads; one thread per element keeps u_K and w_K in registers (n = 2, 3).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "dg_quadrature.cuh"

namespace b200fem {

template <int N>
struct KronTabDev {
  double S[3][N * N];          // interior self matrix per axis (volume + both interior faces)
  double Dlo[3][N * N];        // correction when the low face is a domain boundary:  S_lo-bnd - S_int
  double Dhi[3][N * N];        // same for the high face
  double L[3][N * N];          // coupling to the low neighbour  (w_K += L_d u_{K-e_d})
  double R[3][N * N];          // coupling to the high neighbour
};

template <int N, int TX, int TY, int TZ> struct KronCfg {
  static constexpr int N3 = N * N * N;
  static constexpr int HX = TX + 2, HY = TY + 2, HZ = TZ + 2;
  static constexpr int kThreads = TX * TY * TZ;
  static constexpr int kSlots = HX * HY * HZ;
  static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)kSlots * N3 + sizeof(int) * N3; }
};

// acc[.. i ..] += sum_j M[i*N+j] v[.. j ..] along tensor axis AX (tensor index (i0*N+i1)*N+i2)
template <int N, int AX>
__device__ __forceinline__ void apply_axis(const double* __restrict__ M, const double (&v)[N * N * N], double (&acc)[N * N * N]) {
  constexpr int st = AX == 0 ? N * N : AX == 1 ? N : 1;
#pragma unroll
  for (int t = 0; t < N * N * N; ++t) {
    const int i = (t / st) % N, base = t - i * st;
    double s = acc[t];
#pragma unroll
    for (int j = 0; j < N; ++j) s = fma(M[i * N + j], v[base + j * st], s);
    acc[t] = s;
  }
}

template <int N, int TX, int TY, int TZ>
__global__ void __launch_bounds__(KronCfg<N, TX, TY, TZ>::kThreads)
dg_kronecker_kernel(const __grid_constant__ KronTabDev<N> K, const __grid_constant__ BoxDev box, const int* __restrict__ perm_g,
                    const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                    int tiles_x, int tiles_y) {
  using Cfg = KronCfg<N, TX, TY, TZ>;
  constexpr int N3 = Cfg::N3, HX = Cfg::HX, HY = Cfg::HY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  int* perm = reinterpret_cast<int*>(sm + (size_t)Cfg::kSlots * N3);
  const int tid = threadIdx.x;
  for (int i = tid; i < N3; i += Cfg::kThreads) perm[i] = perm_g[i];

  // tile origin in local coordinates (owned box is tiled)
  const int bx = blockIdx.x % tiles_x, by = (blockIdx.x / tiles_x) % tiles_y, bz = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = box.own_lo[0] + bx * TX, y0 = box.own_lo[1] + by * TY, z0 = box.own_lo[2] + bz * TZ;

  // ---- stage tile + face halo: rows along x are contiguous in the dof vector ----
  constexpr int ROW = HX * N3;
  for (int r = 0; r < HY * Cfg::HZ; ++r) {
    const int hy = r % HY, hz = r / HY;
    const bool yh = hy == 0 || hy == HY - 1, zh = hz == 0 || hz == Cfg::HZ - 1;
    if (yh && zh) continue;                                   // edges/corners are never read
    const int ly = y0 + hy - 1, lz = z0 + hz - 1;
    if (ly < 0 || ly >= box.n[1] || lz < 0 || lz >= box.n[2]) continue;
    const long long row_e = (long long)box.n[0] * (ly + (long long)box.n[1] * lz);
    for (int k = tid; k < ROW; k += Cfg::kThreads) {
      const int lx = x0 - 1 + k / N3;
      if (lx >= 0 && lx < box.n[0]) sm[(size_t)r * ROW + k] = u[(row_e + lx) * N3 + (k % N3)];
    }
  }
  __syncthreads();

  const int tx = tid % TX, ty = (tid / TX) % TY, tz = tid / (TX * TY);
  const int lx = x0 + tx, ly = y0 + ty, lz = z0 + tz;
  const bool active = lx < box.own_hi[0] && ly < box.own_hi[1] && lz < box.own_hi[2];
  const int slot = (tx + 1) + HX * ((ty + 1) + HY * (tz + 1));
  double acc[N3];
  if (active) {
    double v[N3];
#pragma unroll
    for (int t = 0; t < N3; ++t) { v[t] = sm[(size_t)slot * N3 + perm[t]]; acc[t] = 0; }
    apply_axis<N, 0>(K.S[0], v, acc); apply_axis<N, 1>(K.S[1], v, acc); apply_axis<N, 2>(K.S[2], v, acc);
    const int lc[3] = {lx, ly, lz};
    // domain-boundary corrections (rare, divergent only in boundary warps)
    if (box.origin[0] + lx == 0) apply_axis<N, 0>(K.Dlo[0], v, acc);
    if (box.origin[0] + lx == box.gn[0] - 1) apply_axis<N, 0>(K.Dhi[0], v, acc);
    if (box.origin[1] + ly == 0) apply_axis<N, 1>(K.Dlo[1], v, acc);
    if (box.origin[1] + ly == box.gn[1] - 1) apply_axis<N, 1>(K.Dhi[1], v, acc);
    if (box.origin[2] + lz == 0) apply_axis<N, 2>(K.Dlo[2], v, acc);
    if (box.origin[2] + lz == box.gn[2] - 1) apply_axis<N, 2>(K.Dhi[2], v, acc);
    // neighbours
    if (lc[0] > 0)            { const double* p = sm + (size_t)(slot - 1) * N3;
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[perm[t]]; apply_axis<N, 0>(K.L[0], v, acc); }
    if (lc[0] < box.n[0] - 1) { const double* p = sm + (size_t)(slot + 1) * N3;
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[perm[t]]; apply_axis<N, 0>(K.R[0], v, acc); }
    if (lc[1] > 0)            { const double* p = sm + (size_t)(slot - HX) * N3;
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[perm[t]]; apply_axis<N, 1>(K.L[1], v, acc); }
    if (lc[1] < box.n[1] - 1) { const double* p = sm + (size_t)(slot + HX) * N3;
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[perm[t]]; apply_axis<N, 1>(K.R[1], v, acc); }
    if (lc[2] > 0)            { const double* p = sm + (size_t)(slot - HX * HY) * N3;
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[perm[t]]; apply_axis<N, 2>(K.L[2], v, acc); }
    if (lc[2] < box.n[2] - 1) { const double* p = sm + (size_t)(slot + HX * HY) * N3;
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[perm[t]]; apply_axis<N, 2>(K.R[2], v, acc); }
  }
  __syncthreads();                     // everyone is done reading neighbours: reuse own slot for the result
  if (active) {
#pragma unroll
    for (int t = 0; t < N3; ++t) sm[(size_t)slot * N3 + perm[t]] = acc[t];
  }
  __syncthreads();
  // ---- coalesced store of the tile rows ----
  constexpr int OROW = TX * N3;
  for (int r = 0; r < TY * TZ; ++r) {
    const int ty2 = r % TY, tz2 = r / TY, ly2 = y0 + ty2, lz2 = z0 + tz2;
    if (ly2 >= box.own_hi[1] || lz2 >= box.own_hi[2]) continue;
    const long long row_e = (long long)box.n[0] * (ly2 + (long long)box.n[1] * lz2);
    const size_t srow = ((size_t)(1 + HX * ((ty2 + 1) + HY * (tz2 + 1)))) * N3;
    for (int k = tid; k < OROW; k += Cfg::kThreads) {
      const int lx2 = x0 + k / N3;
      if (lx2 < box.own_hi[0]) {
        const long long g = (row_e + lx2) * N3 + (k % N3);
        double val = sm[srow + k];
        if (bvec) val -= bvec[g];
        w[g] = val;
      }
    }
  }
}

}  // namespace b200fem
