// dg_kronecker_tma.cuh -- Kronecker-form DG apply, v2: tile + face halo staged by the bulk-copy engine.
//
// Same arithmetic as dg_kronecker.cuh (w_K = sum_d [S_d u_K + L_d u_{K-e_d} + R_d u_{K+e_d}] - b_K), different data
// movement:
//  * every x-row segment of the tile and of its y/z face halo is one `cp.async.bulk` (TMA unit, 1-D bulk copy,
//    SASS UBLKCP) global -> shared, completion counted on an mbarrier; the two x-halo elements of interior rows and
//    odd leftovers go through 8-byte cp.async (LDGSTS).  No register staging: a CTA has its whole 60-110 KB working set
//    in flight at once, two CTAs per SM overlap one CTA's loads with the other's FMAs;
//  * the load vector tile b_K lands in the output staging area, the thread overwrites it with (A u - b)_K and rows go
//    back with `cp.async.bulk` shared -> global;
//  * element rows are 27*8 = 216 B, i.e. only 8-byte aligned for odd elements: each shared row carries a one-double pad
//    chosen so that shared and global addresses of an element have the same 16-byte phase; bulk copies cover the largest
//    16-byte-aligned even sub-range, leftovers (at most one element per end) use the 8-byte path;
//  * the local (hierarchical / lexicographic) dof permutation is a compile-time table: shared-memory offsets are
//    immediates.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "dg_kronecker.cuh"

namespace b200fem {

// perm[tensor index] = stored local index, computed at compile time (legendre.hh:169-194, 236-250)
template <int N, bool HIER> struct PermTable {
  int p[N * N * N];
  constexpr PermTable() : p{} {
    for (int t = 0; t < N * N * N; ++t) {
      if (!HIER) { p[t] = t; continue; }
      const int a0 = t / (N * N), a1 = (t / N) % N, a2 = t % N;
      const int ma = a0 > a1 ? (a0 > a2 ? a0 : a2) : (a1 > a2 ? a1 : a2);
      int rank = 0;
      for (int s = 0; s < N * N * N; ++s) {
        const int b0 = s / (N * N), b1 = (s / N) % N, b2 = s % N;
        const int mb = b0 > b1 ? (b0 > b2 ? b0 : b2) : (b1 > b2 ? b1 : b2);
        const bool before = mb != ma ? mb < ma : (b0 != a0 ? b0 < a0 : (b1 != a1 ? b1 < a1 : b2 < a2));
        if (before) ++rank;
      }
      p[t] = rank;
    }
  }
};

namespace ptx {
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
}  // namespace ptx

template <int N, int TX, int TY, int TZ> struct KronTmaCfg {
  static constexpr int N3 = N * N * N;
  static constexpr int HY = TY + 2, HZ = TZ + 2;
  static constexpr int kThreads = TX * TY * TZ;
  static constexpr int RS = ((TX + 2) * N3 + 2 + 1) / 2 * 2;     // doubles per staged u row (x-halo + phase pad), even
  static constexpr int RSO = (TX * N3 + 2 + 1) / 2 * 2;          // doubles per output row
  static constexpr int kRows = HY * HZ, kOutRows = TY * TZ;
  static constexpr size_t smem_bytes() { return sizeof(double) * ((size_t)kRows * RS + (size_t)kOutRows * RSO) + 16; }
};

template <int N, bool HIER, int TX, int TY, int TZ>
__global__ void __launch_bounds__(KronTmaCfg<N, TX, TY, TZ>::kThreads, 2)
dg_kronecker_tma_kernel(const __grid_constant__ KronTabDev<N> K, const __grid_constant__ BoxDev box,
                        const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                        int tiles_x, int tiles_y) {
  using Cfg = KronTmaCfg<N, TX, TY, TZ>;
  constexpr int N3 = Cfg::N3, HY = Cfg::HY, HZ = Cfg::HZ, RS = Cfg::RS, RSO = Cfg::RSO;
  static_assert(N3 % 2 == 1, "the 16-byte phase logic assumes an odd number of doubles per element");
  constexpr PermTable<N, HIER> P{};
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* su = reinterpret_cast<double*>(smem_raw);                 // staged u rows
  double* so = su + (size_t)Cfg::kRows * RS;                        // output / load-vector rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(so + (size_t)Cfg::kOutRows * RSO);
  const uint32_t bar_a = ptx::smem_addr(bar);
  const int tid = threadIdx.x;

  const int bx = blockIdx.x % tiles_x, by = (blockIdx.x / tiles_x) % tiles_y, bz = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = box.own_lo[0] + bx * TX, y0 = box.own_lo[1] + by * TY, z0 = box.own_lo[2] + bz * TZ;
  const int xe = min(x0 + TX, box.own_hi[0]);                       // end of the tile's owned x-range
  const int ub8 = (int)((reinterpret_cast<uintptr_t>(u) >> 3) & 1), wb8 = (int)((reinterpret_cast<uintptr_t>(w) >> 3) & 1);

  if (tid == 0) { ptx::mbar_init(bar_a, 1); ptx::fence_barrier_init(); ptx::fence_proxy_async(); }
  __syncthreads();

  // ---------------- issue loads ----------------
  // row r = (hy, hz) of the halo'd tile; element x of that row lives at su[r*RS + pad_r + (x - x0 + 1)*N3]
  auto row_info = [&](int r, long long& row_e, bool& needed, bool& interior) {
    const int hy = r % HY, hz = r / HY;
    const bool yh = hy == 0 || hy == HY - 1, zh = hz == 0 || hz == HZ - 1;
    const int ly = y0 + hy - 1, lz = z0 + hz - 1;
    interior = !yh && !zh;
    needed = !(yh && zh) && ly >= 0 && ly < box.n[1] && lz >= 0 && lz < box.n[2];
    row_e = (long long)box.n[0] * (ly + (long long)box.n[1] * lz);
  };
  if (tid < 32) {
    // pass 1: bytes this lane will request through the bulk engine; pass 2: issue
    uint32_t my_bytes = 0;
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) {
        uint32_t total = my_bytes;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (tid == 0) ptx::mbar_expect_tx(bar_a, total);
        __syncwarp();
      }
      for (int r = tid; r < Cfg::kRows + Cfg::kOutRows; r += 32) {
        if (r < Cfg::kRows) {
          long long row_e; bool needed, interior; row_info(r, row_e, needed, interior);
          if (!needed) continue;
          const int par = (int)((ub8 + row_e + x0) & 1);            // 16-byte phase of element x0 in global memory
          const int xs = x0 + par;                                   // first element on a 16-byte boundary
          const int cnt = ((xe - xs) > 0 ? (xe - xs) : 0) & ~1;      // even number of elements
          if (cnt <= 0) continue;
          const int pad = (par + 1) & 1;                             // same phase in shared memory (RS even, N3 odd)
          const uint32_t bytes = (uint32_t)cnt * N3 * 8;
          if (pass == 0) my_bytes += bytes;
          else ptx::bulk_g2s(ptx::smem_addr(su + (size_t)r * RS + pad + (xs - x0 + 1) * N3), u + (row_e + xs) * N3, bytes, bar_a);
        } else if (bvec) {
          const int ro = r - Cfg::kRows, ly = y0 + ro % TY, lz = z0 + ro / TY;
          if (ly >= box.own_hi[1] || lz >= box.own_hi[2]) continue;
          const long long row_e = (long long)box.n[0] * (ly + (long long)box.n[1] * lz);
          const int par = (int)((wb8 + row_e + x0) & 1), xs = x0 + par, cnt = ((xe - xs) > 0 ? (xe - xs) : 0) & ~1;
          if (cnt <= 0) continue;
          const uint32_t bytes = (uint32_t)cnt * N3 * 8;
          if (pass == 0) my_bytes += bytes;
          else ptx::bulk_g2s(ptx::smem_addr(so + (size_t)ro * RSO + par + (xs - x0) * N3), bvec + (row_e + xs) * N3, bytes, bar_a);
        }
      }
    }
  }
  // 8-byte path: x-halo elements of interior rows and odd leftovers of every row
  for (int r = 0; r < Cfg::kRows; ++r) {
    long long row_e; bool needed, interior; row_info(r, row_e, needed, interior);
    if (!needed) continue;
    const int par = (int)((ub8 + row_e + x0) & 1), xs = x0 + par, cnt = ((xe - xs) > 0 ? (xe - xs) : 0) & ~1, pad = (par + 1) & 1;
    double* rowp = su + (size_t)r * RS + pad + N3;                  // element x0
    // candidates: x0-1 (interior rows), x0 (if par), xs+cnt (tail leftover), xe (interior rows)
    const int cand[4] = {interior && x0 - 1 >= 0 ? x0 - 1 : -1, par ? x0 : -1, (xs + cnt < xe) ? xs + cnt : -1, interior && xe < box.n[0] ? xe : -1};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int x = cand[c];
      if (x < 0 || (c == 2 && x == cand[1])) continue;
      for (int j = tid; j < N3; j += Cfg::kThreads) ptx::cp_async8(ptx::smem_addr(rowp + (x - x0) * N3 + j), u + (row_e + x) * N3 + j);
    }
  }
  if (bvec) {
    for (int ro = 0; ro < Cfg::kOutRows; ++ro) {
      const int ly = y0 + ro % TY, lz = z0 + ro / TY;
      if (ly >= box.own_hi[1] || lz >= box.own_hi[2]) continue;
      const long long row_e = (long long)box.n[0] * (ly + (long long)box.n[1] * lz);
      const int par = (int)((wb8 + row_e + x0) & 1), xs = x0 + par, cnt = ((xe - xs) > 0 ? (xe - xs) : 0) & ~1;
      double* rowp = so + (size_t)ro * RSO + par;
      const int cand[2] = {par ? x0 : -1, (xs + cnt < xe) ? xs + cnt : -1};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int x = cand[c];
        if (x < 0 || (c == 1 && x == cand[0])) continue;
        for (int j = tid; j < N3; j += Cfg::kThreads) ptx::cp_async8(ptx::smem_addr(rowp + (x - x0) * N3 + j), bvec + (row_e + x) * N3 + j);
      }
    }
  }
  ptx::cp_async_wait_all();
  __syncthreads();
  ptx::mbar_wait(bar_a, 0);

  // ---------------- compute: one thread per element ----------------
  const int tx = tid % TX, ty = (tid / TX) % TY, tz = tid / (TX * TY);
  const int lx = x0 + tx, ly = y0 + ty, lz = z0 + tz;
  const bool active = lx < box.own_hi[0] && ly < box.own_hi[1] && lz < box.own_hi[2];
  auto elem_ptr = [&](int hy, int hz) -> const double* {          // element (lx, row hy/hz) in the staged tile
    const int yy = y0 + hy - 1, zz = z0 + hz - 1;
    const long long row_e = (long long)box.n[0] * (yy + (long long)box.n[1] * zz);
    const int pad = (int)((ub8 + row_e + x0 + 1) & 1);
    return su + (size_t)(hy + HY * hz) * RS + pad + (tx + 1) * N3;
  };
  double acc[N3];
  if (active) {
    double v[N3];
    const double* own = elem_ptr(ty + 1, tz + 1);
#pragma unroll
    for (int t = 0; t < N3; ++t) { v[t] = own[P.p[t]]; acc[t] = 0; }
    apply_axis<N, 0>(K.S[0], v, acc); apply_axis<N, 1>(K.S[1], v, acc); apply_axis<N, 2>(K.S[2], v, acc);
    if (box.origin[0] + lx == 0) apply_axis<N, 0>(K.Dlo[0], v, acc);
    if (box.origin[0] + lx == box.gn[0] - 1) apply_axis<N, 0>(K.Dhi[0], v, acc);
    if (box.origin[1] + ly == 0) apply_axis<N, 1>(K.Dlo[1], v, acc);
    if (box.origin[1] + ly == box.gn[1] - 1) apply_axis<N, 1>(K.Dhi[1], v, acc);
    if (box.origin[2] + lz == 0) apply_axis<N, 2>(K.Dlo[2], v, acc);
    if (box.origin[2] + lz == box.gn[2] - 1) apply_axis<N, 2>(K.Dhi[2], v, acc);
    if (lx > 0)            { const double* p = own - N3;
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[P.p[t]]; apply_axis<N, 0>(K.L[0], v, acc); }
    if (lx < box.n[0] - 1) { const double* p = own + N3;
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[P.p[t]]; apply_axis<N, 0>(K.R[0], v, acc); }
    if (ly > 0)            { const double* p = elem_ptr(ty, tz + 1);
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[P.p[t]]; apply_axis<N, 1>(K.L[1], v, acc); }
    if (ly < box.n[1] - 1) { const double* p = elem_ptr(ty + 2, tz + 1);
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[P.p[t]]; apply_axis<N, 1>(K.R[1], v, acc); }
    if (lz > 0)            { const double* p = elem_ptr(ty + 1, tz);
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[P.p[t]]; apply_axis<N, 2>(K.L[2], v, acc); }
    if (lz < box.n[2] - 1) { const double* p = elem_ptr(ty + 1, tz + 2);
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = p[P.p[t]]; apply_axis<N, 2>(K.R[2], v, acc); }
    // (A u - b)_K into the output staging row
    const long long row_e = (long long)box.n[0] * (ly + (long long)box.n[1] * lz);
    double* o = so + (size_t)(ty + TY * tz) * RSO + (int)((wb8 + row_e + x0) & 1) + tx * N3;
    if (bvec) {
#pragma unroll
      for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t] - o[P.p[t]];
    } else {
#pragma unroll
      for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t];
    }
  }
  ptx::fence_proxy_async();
  __syncthreads();

  // ---------------- store rows ----------------
  if (tid < Cfg::kOutRows) {
    const int ro = tid, ly2 = y0 + ro % TY, lz2 = z0 + ro / TY;
    if (ly2 < box.own_hi[1] && lz2 < box.own_hi[2]) {
      const long long row_e = (long long)box.n[0] * (ly2 + (long long)box.n[1] * lz2);
      const int par = (int)((wb8 + row_e + x0) & 1), xs = x0 + par, cnt = ((xe - xs) > 0 ? (xe - xs) : 0) & ~1;
      if (cnt > 0) ptx::bulk_s2g(w + (row_e + xs) * N3, ptx::smem_addr(so + (size_t)ro * RSO + par + (xs - x0) * N3), (uint32_t)cnt * N3 * 8);
    }
    ptx::bulk_commit();
  }
  for (int ro = 0; ro < Cfg::kOutRows; ++ro) {                      // odd leftovers: plain stores
    const int ly2 = y0 + ro % TY, lz2 = z0 + ro / TY;
    if (ly2 >= box.own_hi[1] || lz2 >= box.own_hi[2]) continue;
    const long long row_e = (long long)box.n[0] * (ly2 + (long long)box.n[1] * lz2);
    const int par = (int)((wb8 + row_e + x0) & 1), xs = x0 + par, cnt = ((xe - xs) > 0 ? (xe - xs) : 0) & ~1;
    const double* rowp = so + (size_t)ro * RSO + par;
    const int cand[2] = {par && x0 < xe ? x0 : -1, (xs + cnt < xe) ? xs + cnt : -1};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int x = cand[c];
      if (x < 0 || (c == 1 && x == cand[0])) continue;
      for (int j = tid; j < N3; j += Cfg::kThreads) w[(row_e + x) * N3 + j] = rowp[(x - x0) * N3 + j];
    }
  }
  if (tid < Cfg::kOutRows) ptx::bulk_wait_read();
}

}  // namespace b200fem
