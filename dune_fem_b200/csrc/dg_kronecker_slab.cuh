// dg_kronecker_slab.cuh -- Kronecker-form DG apply for the higher orders (Q3 .. Q5, n = 4 .. 6 modes per axis).
//
// Same arithmetic as dg_kronecker.cuh:  w_K = sum_d [ S_d u_K + L_d u_{K-e_d} + R_d u_{K+e_d} ] - b_K  with n x n
// matrices acting along one tensor axis (9 n^4 FMA per element).  One thread per element (the Q1/Q2 kernels) is not
// possible any more -- an element is n^3 = 64 .. 216 doubles -- so an element is shared by n threads and the work is
// organised around SLABS, the unit that gives n FMAs per shared-memory load (profiles/micro/dmma_bench_b200.txt: the
// FP64 pipe needs >= 4 FMA per LDS.64 to stay fed; mma.sync.m8n8k4.f64 has the same 37 TFLOP/s peak as DFMA on sm_100a,
// so padding 6 -> 8 rows and 18 -> 20 columns for the tensor pipe could only lose):
//
//   phase A  thread (K, j) owns the slab u_K[m0 = j][.][.]: y- and z-contractions of the element itself happen in
//            registers (2 n^3 FMA for n^2 loads); the y/z face neighbours are streamed line by line (n loads ->
//            n^2 FMA).  Result: n^2 accumulators acc[m1][m2] in registers.
//   phase B  the same thread takes the slab u[.][.][m2 = j] of K and of its two x-neighbours, does the x-contraction
//            (n loads -> n^2 FMA per line) and writes the partial result into an exchange buffer;
//   combine  after a barrier the thread reads slab m0 = j of the exchange buffer, adds it to acc and leaves the sum
//            there; warps then write whole elements to global memory in stored order (coalesced), minus b_K.
//
// A CTA owns TX x TY x TZ elements; the tile and its six face halos are staged by 8-byte cp.async (LDGSTS, no
// register round trip; missing neighbours are zero-filled = "no coupling"), in TENSOR order with a padded layout:
// slab stride S0 and element stride ES are chosen so that both access patterns -- lanes (K, j) reading slab m0 = j
// (bank = K ES + j S0) and slab m2 = j (bank = K ES + j) -- hit 16 distinct 8-byte banks per half-warp.  The dof
// permutation of the hierarchical ordering (legendre.hh:236-250) is folded into the staging addresses, so the compute
// phases use immediates only.  The exchange buffer aliases the y/z halo slots, which are dead after phase A.  Two to four
// CTAs per SM overlap one tile's staging with the others' FMAs.
//
// Tried and dropped (profiles/r01_dg_kronecker_slab_q5.md): a variant with three warp-uniform thread roles per slab (one per
// axis, 3 n threads per element, n^2 accumulators each).  It triples the warps but 9 warps per CTA put 5 warps on one SM
// sub-partition, which caps the kernel at 96 registers -> spills, and the three partial results cost two more shared
// round trips: 54 instead of 99 GDoF/s for Q5.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "dg_kronecker.cuh"

namespace b200fem {

template <int N, int TX, int TY, int TZ, int SPLIT = 1> struct KronSlabCfg {
  static constexpr int N2 = N * N, N3 = N * N * N;
  static constexpr int NR = (N + SPLIT - 1) / SPLIT;         // slab rows per thread: SPLIT = 2 shares a slab between two threads
  static constexpr int slab_stride() {
    int s = N2;
    while ((N * (s - 1)) % 16 != 0 && s < N2 + 6) ++s;       // bank(K, j) = N K + S0 j distinct over 16 consecutive lanes
    return (N * (s - 1)) % 16 == 0 ? s : (N2 | 1);
  }
  static constexpr int S0 = slab_stride();
  static constexpr int elem_stride() { int e = N * S0; while (e % 16 != N % 16) ++e; return e; }
  static constexpr int ES = elem_stride();
  static constexpr int NO = TX * TY * TZ, NX = 2 * TY * TZ, NY = 2 * TX * TZ, NZ = 2 * TX * TY, kSlots = NO + NX + NY + NZ;
  static constexpr int kWork = NO * N, kHalf = (kWork + 31) / 32 * 32, kThreads = SPLIT * kHalf, kWarps = kThreads / 32;   // halves are whole warps
  static constexpr int KS = (N3 + 31) / 32;                  // doubles per lane when a warp moves one element
  static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)kSlots * ES + sizeof(int) * N3; }
  static_assert(NY + NZ >= NO, "the exchange buffer aliases the y/z halo slots");
};

namespace slab {
__device__ __forceinline__ void cp_async8(uint32_t dst, const double* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// acc[m1][m2] += sum_k M[i][k] v[..k..] inside one N x N slab held in registers; AX = 0 contracts the slow index, AX = 1
// the fast one.  k is the OUTER loop: consecutive FMAs hit different accumulators (no dependent chains back to back).
template <int N, int AX>
__device__ __forceinline__ void self(const double* __restrict__ M, const double (&v)[N * N], double (&acc)[N * N]) {
  constexpr int st = AX == 0 ? N : 1;
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int t = 0; t < N * N; ++t) {
      const int i = AX == 0 ? t / N : t % N, base = t - i * st;
      acc[t] = fma(M[i * N + k], v[base + k * st], acc[t]);
    }
  }
}
// the same with the operand streamed from shared memory line by line: src points at the slab, whose slow index has
// stride SLOW and whose fast index has stride FAST (doubles)
template <int N, int AX, int SLOW, int FAST>
__device__ __forceinline__ void streamed(const double* __restrict__ M, const double* __restrict__ src, double (&acc)[N * N]) {
#pragma unroll
  for (int l = 0; l < N; ++l) {                              // l = the index that is NOT contracted
    double line[N];
#pragma unroll
    for (int k = 0; k < N; ++k) line[k] = AX == 0 ? src[k * SLOW + l * FAST] : src[l * SLOW + k * FAST];
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int t = AX == 0 ? i * N + l : l * N + i;
        acc[t] = fma(M[i * N + k], line[k], acc[t]);
      }
    }
  }
}


// ---- row-restricted variants for SPLIT = 2: a thread owns the slab rows [R0, R1) (slow index of its slab); acc is indexed
// [(row - R0) * N + fast].  Which half a thread takes is warp-uniform, so the matrix indices below stay compile-time
// constants (constant-bank operands) in both instantiations.
template <int N, int AX, int R0, int R1, int NA>
__device__ __forceinline__ void self_rows(const double* __restrict__ M, const double (&v)[N * N], double (&acc)[NA]) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int r = R0; r < R1; ++r) {
#pragma unroll
      for (int f = 0; f < N; ++f)
        acc[(r - R0) * N + f] = AX == 0 ? fma(M[r * N + k], v[k * N + f], acc[(r - R0) * N + f]) : fma(M[f * N + k], v[r * N + k], acc[(r - R0) * N + f]);
    }
  }
}
// contraction over the SLOW index of the streamed slab, output rows restricted to [R0, R1): lines run along the slow index
template <int N, int R0, int R1, int SLOW, int FAST, int NA>
__device__ __forceinline__ void streamed_slow_rows(const double* __restrict__ M, const double* __restrict__ src, double (&acc)[NA]) {
#pragma unroll
  for (int f = 0; f < N; ++f) {
    double line[N];
#pragma unroll
    for (int k = 0; k < N; ++k) line[k] = src[k * SLOW + f * FAST];
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int r = R0; r < R1; ++r) acc[(r - R0) * N + f] = fma(M[r * N + k], line[k], acc[(r - R0) * N + f]);
    }
  }
}
// contraction over the FAST index, rows [R0, R1) only: lines run along the fast index of those rows
template <int N, int R0, int R1, int SLOW, int FAST, int NA>
__device__ __forceinline__ void streamed_fast_rows(const double* __restrict__ M, const double* __restrict__ src, double (&acc)[NA]) {
#pragma unroll
  for (int r = R0; r < R1; ++r) {
    double line[N];
#pragma unroll
    for (int k = 0; k < N; ++k) line[k] = src[r * SLOW + k * FAST];
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int f = 0; f < N; ++f) acc[(r - R0) * N + f] = fma(M[f * N + k], line[k], acc[(r - R0) * N + f]);
    }
  }
}
// x-contraction of phase B for the slab m2 = j: slow index m0 (contracted, stride S0), fast index m1 restricted to [R0, R1);
// pb is indexed [m0 * (R1 - R0) + (m1 - R0)]
template <int N, int R0, int R1, int S0, int NA>
__device__ __forceinline__ void streamed_x_rows(const double* __restrict__ M, const double* __restrict__ src, double (&pb)[NA]) {
#pragma unroll
  for (int r = R0; r < R1; ++r) {
    double line[N];
#pragma unroll
    for (int k = 0; k < N; ++k) line[k] = src[k * S0 + r * N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int i = 0; i < N; ++i) pb[i * (R1 - R0) + (r - R0)] = fma(M[i * N + k], line[k], pb[i * (R1 - R0) + (r - R0)]);
    }
  }
}

// Stages the tile and its six face halos.  Slot sl of the CTA is handled by warp sl % kWarps; the decode of a slot (local
// element coordinates, existence, global offset) is done ONCE per slot by one lane and broadcast by shuffle -- done per
// warp iteration it was 45 % of all instructions of the first version (profiles/r01_dg_kronecker_slab_q3.md).
template <class Cfg, int TX, int TY, int TZ>
__device__ __forceinline__ void stage_tile(const BoxDev& box, const double* __restrict__ u, double* U, const int (&doff)[Cfg::KS],
                                           const int x0, const int y0, const int z0, const int warp, const int lane) {
  constexpr int N3 = Cfg::N3, NO = Cfg::NO, NX = Cfg::NX, NY = Cfg::NY, KS = Cfg::KS, SPW = (Cfg::kSlots + Cfg::kWarps - 1) / Cfg::kWarps;
  static_assert(SPW <= 32, "one lane decodes one slot of its warp");
  long long my_src = -1;                                     // offset (doubles) of the element in slot warp + lane * kWarps, -1: none
  {
    const int sl = warp + lane * Cfg::kWarps;
    if (sl < Cfg::kSlots) {
      int ex, ey, ez;
      if (sl < NO) { ex = sl % TX; ey = (sl / TX) % TY; ez = sl / (TX * TY); }
      else if (sl < NO + NX) { int r = sl - NO; const int side = r / (TY * TZ); r -= side * TY * TZ; ex = side ? TX : -1; ey = r % TY; ez = r / TY; }
      else if (sl < NO + NX + NY) { int r = sl - NO - NX; const int side = r / (TX * TZ); r -= side * TX * TZ; ey = side ? TY : -1; ex = r % TX; ez = r / TX; }
      else { int r = sl - NO - NX - NY; const int side = r / (TX * TY); r -= side * TX * TY; ez = side ? TZ : -1; ex = r % TX; ey = r / TX; }
      const int lx = x0 + ex, ly = y0 + ey, lz = z0 + ez;
      if (lx >= 0 && lx < box.n[0] && ly >= 0 && ly < box.n[1] && lz >= 0 && lz < box.n[2])
        my_src = (lx + (long long)box.n[0] * (ly + (long long)box.n[1] * lz)) * N3;
    }
  }
  const uint32_t ubase = (uint32_t)__cvta_generic_to_shared(U);
#pragma unroll 1
  for (int i = 0; i < SPW; ++i) {
    const int sl = warp + i * Cfg::kWarps;
    if (sl >= Cfg::kSlots) break;
    const long long so = __shfl_sync(0xffffffffu, my_src, i);
    const bool ok = so >= 0;
    const double* src = u + (ok ? so : 0) + lane;            // (a missing element reads nothing: src-size 0 zero-fills; the address only has to be valid)
    const int bytes = ok ? 8 : 0;
    const uint32_t dst = ubase + 8u * (uint32_t)(sl * Cfg::ES);
#pragma unroll
    for (int k = 0; k < KS; ++k)
      if (lane + 32 * k < N3) cp_async8(dst + 8u * doff[k], src + 32 * k, bytes);
  }
}
}  // namespace slab

// phase A of thread (e, j) for the slab rows [R0, R1): slab m0 = j, axes y (slow index m1) and z (fast index m2)
template <int N, int TX, int TY, int TZ, int R0, int R1, class Cfg, int NA>
__device__ __forceinline__ void slab_phase_a(const KronTabDev<N>& K, const BoxDev& box, const double* __restrict__ U, const int e, const int j,
                                             const int tx, const int ty, const int tz, const int gy, const int gz, double (&acc)[NA]) {
  constexpr int N2 = Cfg::N2, S0 = Cfg::S0, ES = Cfg::ES, NO = Cfg::NO, NX = Cfg::NX, NY = Cfg::NY;
  const double* own = U + (size_t)e * ES + j * S0;
  {
    double v[N2];
#pragma unroll
    for (int t = 0; t < N2; ++t) v[t] = own[t];
#pragma unroll
    for (int t = 0; t < NA; ++t) acc[t] = 0.0;
    slab::self_rows<N, 0, R0, R1>(K.S[1], v, acc);
    slab::self_rows<N, 1, R0, R1>(K.S[2], v, acc);
    if (gy == 0) slab::self_rows<N, 0, R0, R1>(K.Dlo[1], v, acc);
    if (gy == box.gn[1] - 1) slab::self_rows<N, 0, R0, R1>(K.Dhi[1], v, acc);
    if (gz == 0) slab::self_rows<N, 1, R0, R1>(K.Dlo[2], v, acc);
    if (gz == box.gn[2] - 1) slab::self_rows<N, 1, R0, R1>(K.Dhi[2], v, acc);
  }
  const int s_ym = ty > 0 ? e - TX : NO + NX + tz * TX + tx, s_yp = ty < TY - 1 ? e + TX : NO + NX + TX * TZ + tz * TX + tx;
  const int s_zm = tz > 0 ? e - TX * TY : NO + NX + NY + ty * TX + tx, s_zp = tz < TZ - 1 ? e + TX * TY : NO + NX + NY + TX * TY + ty * TX + tx;
  slab::streamed_slow_rows<N, R0, R1, N, 1>(K.L[1], U + (size_t)s_ym * ES + j * S0, acc);
  slab::streamed_slow_rows<N, R0, R1, N, 1>(K.R[1], U + (size_t)s_yp * ES + j * S0, acc);
  slab::streamed_fast_rows<N, R0, R1, N, 1>(K.L[2], U + (size_t)s_zm * ES + j * S0, acc);
  slab::streamed_fast_rows<N, R0, R1, N, 1>(K.R[2], U + (size_t)s_zp * ES + j * S0, acc);
}
// phase B: slab m2 = j, axis x, the fast index m1 restricted to [R0, R1); the partial result goes to the exchange slot
template <int N, int TX, int TY, int TZ, int R0, int R1, class Cfg>
__device__ __forceinline__ void slab_phase_b(const KronTabDev<N>& K, const BoxDev& box, double* __restrict__ U, const int e, const int j,
                                             const int tx, const int ty, const int tz, const int gx) {
  constexpr int S0 = Cfg::S0, ES = Cfg::ES, NO = Cfg::NO, NX = Cfg::NX, NRR = R1 - R0;
  if (NRR <= 0) return;
  double pb[N * (NRR > 0 ? NRR : 1)];
#pragma unroll
  for (int t = 0; t < N * NRR; ++t) pb[t] = 0.0;
  const int s_xm = tx > 0 ? e - 1 : NO + tz * TY + ty, s_xp = tx < TX - 1 ? e + 1 : NO + TY * TZ + tz * TY + ty;
  const double* own = U + (size_t)e * ES + j;
  slab::streamed_x_rows<N, R0, R1, S0>(K.S[0], own, pb);
  if (gx == 0) slab::streamed_x_rows<N, R0, R1, S0>(K.Dlo[0], own, pb);
  if (gx == box.gn[0] - 1) slab::streamed_x_rows<N, R0, R1, S0>(K.Dhi[0], own, pb);
  slab::streamed_x_rows<N, R0, R1, S0>(K.L[0], U + (size_t)s_xm * ES + j, pb);
  slab::streamed_x_rows<N, R0, R1, S0>(K.R[0], U + (size_t)s_xp * ES + j, pb);
  double* Xe = U + (size_t)(NO + NX + e) * ES;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int r = R0; r < R1; ++r) Xe[i * S0 + r * N + j] = pb[i * NRR + (r - R0)];
  }
}
// combine: rows [R0, R1) of slab m0 = j of the exchange slot += acc, the sum stays there for the store loop
template <int N, int R0, int R1, class Cfg, int NA>
__device__ __forceinline__ void slab_combine(double* __restrict__ U, const int e, const int j, const double (&acc)[NA]) {
  double* X = U + (size_t)(Cfg::NO + Cfg::NX + e) * Cfg::ES + j * Cfg::S0;
#pragma unroll
  for (int r = R0; r < R1; ++r) {
#pragma unroll
    for (int f = 0; f < N; ++f) X[r * N + f] += acc[(r - R0) * N + f];
  }
}

template <int N, int TX, int TY, int TZ, int MINB, int SPLIT>
__global__ void __launch_bounds__(KronSlabCfg<N, TX, TY, TZ, SPLIT>::kThreads, MINB)
dg_kronecker_slab_kernel(const __grid_constant__ KronTabDev<N> K, const __grid_constant__ BoxDev box, const int* __restrict__ perm_g,
                         const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                         const int tiles_x, const int tiles_y, const int ntiles, long long* __restrict__ timeline) {
  using Cfg = KronSlabCfg<N, TX, TY, TZ, SPLIT>;
  constexpr int N2 = Cfg::N2, N3 = Cfg::N3, S0 = Cfg::S0, ES = Cfg::ES, NO = Cfg::NO, NX = Cfg::NX, KS = Cfg::KS, NR = Cfg::NR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* const U = reinterpret_cast<double*>(smem_raw);
  int* const tinv = reinterpret_cast<int*>(U + (size_t)Cfg::kSlots * ES);     // stored index -> tensor index
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int t = tid; t < N3; t += Cfg::kThreads) tinv[perm_g[t]] = t;
  __syncthreads();

  // shared-memory offset (inside an element) of the doubles this lane moves when its warp stages / stores an element
  int doff[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) { const int s = lane + 32 * k; const int t = s < N3 ? tinv[s] : 0; doff[k] = (t / N2) * S0 + (t % N2); }

  // persistent over tiles: the per-CTA prologue above and the CTA launch are paid once per resident slot, not once per tile
  // (a CTA's lifetime was ~10 us per tile of which only ~3 us were FMAs)
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
  const int bx = tile % tiles_x, by = (tile / tiles_x) % tiles_y, bz = tile / (tiles_x * tiles_y);
  const int x0 = box.own_lo[0] + bx * TX, y0 = box.own_lo[1] + by * TY, z0 = box.own_lo[2] + bz * TZ;

  // optional timeline of one CTA (B200FEM_SLAB_TIMELINE): clock64 stamps of the phases of its second tile
  const bool stamp = timeline != nullptr && blockIdx.x == 77 && tile == blockIdx.x + (int)gridDim.x && tid == 0;
  if (stamp) timeline[0] = clock64();
  slab::stage_tile<Cfg, TX, TY, TZ>(box, u, U, doff, x0, y0, z0, warp, lane);
  if (stamp) timeline[1] = clock64();
  slab::cp_async_wait_all();
  __syncthreads();
  if (stamp) timeline[2] = clock64();

  // ------------------------------ compute: thread (e, j[, half]) ------------------------------
  // SPLIT = 2: the two halves of a slab (rows [0, NR) and [NR, N)) belong to different WARPS (half = warp-uniform), so both
  // instantiations keep compile-time matrix indices and consecutive lanes still are consecutive (e, j): no bank conflicts
  const int half = SPLIT == 1 ? 0 : tid / Cfg::kHalf, idx = tid - half * Cfg::kHalf;
  const bool worker = idx < Cfg::kWork;
  const int e = worker ? idx / N : 0, j = idx % N;
  const int tx = e % TX, ty = (e / TX) % TY, tz = e / (TX * TY);
  const int gx = box.origin[0] + x0 + tx, gy = box.origin[1] + y0 + ty, gz = box.origin[2] + z0 + tz;
  double acc[NR * N];
  if (worker) {
    if (half == 0) slab_phase_a<N, TX, TY, TZ, 0, NR, Cfg>(K, box, U, e, j, tx, ty, tz, gy, gz, acc);
    else slab_phase_a<N, TX, TY, TZ, (SPLIT == 1 ? N : NR), N, Cfg>(K, box, U, e, j, tx, ty, tz, gy, gz, acc);
  }
  __syncthreads();                                            // y/z halo slots are dead from here on: they become the exchange buffer
  if (stamp) timeline[3] = clock64();
  if (worker) {
    if (half == 0) slab_phase_b<N, TX, TY, TZ, 0, NR, Cfg>(K, box, U, e, j, tx, ty, tz, gx);
    else slab_phase_b<N, TX, TY, TZ, (SPLIT == 1 ? N : NR), N, Cfg>(K, box, U, e, j, tx, ty, tz, gx);
  }
  // the load-vector rows of the elements this warp will store are requested now: their latency hides behind the
  // barrier and the combine step instead of sitting in front of every store
  constexpr int ELW = (NO + Cfg::kWarps - 1) / Cfg::kWarps;   // elements per warp in the store loop
  double breg[ELW][KS];
  long long gbase[ELW];
#pragma unroll
  for (int i = 0; i < ELW; ++i) {
    const int el = warp + i * Cfg::kWarps;
    const int lx = x0 + el % TX, ly = y0 + (el / TX) % TY, lz = z0 + el / (TX * TY);
    const bool owned = el < NO && lx < box.own_hi[0] && ly < box.own_hi[1] && lz < box.own_hi[2];
    gbase[i] = owned ? (lx + (long long)box.n[0] * (ly + (long long)box.n[1] * lz)) * N3 : -1;
#pragma unroll
    for (int k = 0; k < KS; ++k) { const int s = lane + 32 * k; breg[i][k] = (bvec && owned && s < N3) ? bvec[gbase[i] + s] : 0.0; }
  }
  if (stamp) timeline[4] = clock64();
  __syncthreads();
  if (worker) {
    if (half == 0) slab_combine<N, 0, NR, Cfg>(U, e, j, acc);
    else slab_combine<N, (SPLIT == 1 ? N : NR), N, Cfg>(U, e, j, acc);
  }
  __syncthreads();

  if (stamp) timeline[5] = clock64();
  // ------------------------------ store: one warp per element, stored order, coalesced ------------------------------
#pragma unroll
  for (int i = 0; i < ELW; ++i) {
    if (gbase[i] < 0) continue;
    const double* X = U + (size_t)(NO + NX + warp + i * Cfg::kWarps) * ES;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int s = lane + 32 * k;
      if (s < N3) w[gbase[i] + s] = X[doff[k]] - breg[i][k];
    }
  }
  __syncthreads();                                            // the exchange slots have been read: the next tile may be staged
  if (stamp) timeline[6] = clock64();
  }
}

}  // namespace b200fem
