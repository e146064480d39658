// dg_kronecker_pipe.cuh -- Kronecker-form DG apply, v3: persistent, warp-specialised, double-buffered.
//
// One CTA per SM loops over tiles of TX x TY x TZ elements.  Warp 4 (producer) drives the bulk-copy engine:
//   tile i+1:  cp.async.bulk  global -> shared   (u rows of tile + y/z face halo, load-vector rows)   -> full[s]
//   tile i-1:  cp.async.bulk  shared -> global   (finished output rows)                               <- done[s]
// while warps 0-3 (consumers, one thread per element) compute tile i out of the other stage.  All traffic of the
// aligned case goes through UBLKCP; there is no register staging and no per-thread address loop in the consumers.
// Interior rows are loaded with two extra elements on each side (16-byte phase, see dg_kronecker_tma.cuh), which also
// provides the x-halo.  Elements that cannot be reached by a 16-byte aligned, even-sized bulk copy (odd row lengths,
// 8-byte aligned user vectors) are copied by the producer lanes with plain loads/stores before it arrives on the
// barrier, so the kernel is correct for every box size.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "dg_kronecker_tma.cuh"

namespace b200fem {

template <int N, int TX, int TY, int TZ> struct KronPipeCfg {
  static constexpr int N3 = N * N * N;
  static constexpr int kConsumers = TX * TY * TZ, kThreads = kConsumers + 32;
  static constexpr int RSI = ((TX + 4) * N3 + 2 + 1) / 2 * 2;      // interior row: x0-2 .. x0+TX+1, + phase pad
  static constexpr int RSH = (TX * N3 + 2 + 1) / 2 * 2;            // halo / output row: x0 .. x0+TX-1
  static constexpr int kIntRows = TY * TZ, kHaloRows = 2 * TZ + 2 * TY, kOutRows = TY * TZ;
  static constexpr int kStageU = kIntRows * RSI + kHaloRows * RSH; // doubles
  static constexpr int kStageO = kOutRows * RSH;
  static constexpr int kStages = 2;
  static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)kStages * (kStageU + kStageO) + 64; }
};

namespace ptx {
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}  // namespace ptx

template <int N, bool HIER, int TX, int TY, int TZ>
__global__ void __launch_bounds__(KronPipeCfg<N, TX, TY, TZ>::kThreads, 1)
dg_kronecker_pipe_kernel(const __grid_constant__ KronTabDev<N> K, const __grid_constant__ BoxDev box,
                         const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                         int tiles_x, int tiles_y, int ntiles) {
  using Cfg = KronPipeCfg<N, TX, TY, TZ>;
  constexpr int N3 = Cfg::N3, RSI = Cfg::RSI, RSH = Cfg::RSH;
  static_assert(N3 % 2 == 1, "16-byte phase logic assumes an odd number of doubles per element");
  static_assert(Cfg::kIntRows + Cfg::kHaloRows <= 32 && Cfg::kOutRows <= 32, "one producer lane per row");
  constexpr PermTable<N, HIER> P{};
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sbase = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbase + (size_t)Cfg::kStages * (Cfg::kStageU + Cfg::kStageO));
  const uint32_t full_a = ptx::smem_addr(bars), done_a = ptx::smem_addr(bars + 2);
  const int tid = threadIdx.x;
  const int ub8 = (int)((reinterpret_cast<uintptr_t>(u) >> 3) & 1), wb8 = (int)((reinterpret_cast<uintptr_t>(w) >> 3) & 1);

  if (tid == 0) {
    ptx::mbar_init(full_a, 1); ptx::mbar_init(full_a + 8, 1);
    ptx::mbar_init(done_a, Cfg::kConsumers); ptx::mbar_init(done_a + 8, Cfg::kConsumers);
    ptx::fence_barrier_init(); ptx::fence_proxy_async();
  }
  __syncthreads();

  auto tile_origin = [&](int tile, int& x0, int& y0, int& z0) {
    const int bx = tile % tiles_x, by = (tile / tiles_x) % tiles_y, bz = tile / (tiles_x * tiles_y);
    x0 = box.own_lo[0] + bx * TX; y0 = box.own_lo[1] + by * TY; z0 = box.own_lo[2] + bz * TZ;
  };

  if (tid >= Cfg::kConsumers) {
    // ============================== producer warp ==============================
    const int lane = tid - Cfg::kConsumers;
    for (int it = 0;; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const bool has = tile < ntiles;
      if (has) {
        const int s = it & 1;
        double* su = sbase + (size_t)s * (Cfg::kStageU + Cfg::kStageO);
        double* so = su + Cfg::kStageU;
        if (it >= 2) ptx::bulk_wait_read();                         // stores of tile it-2 have left so[s]
        __syncwarp();
        int x0, y0, z0; tile_origin(tile, x0, y0, z0);
        const int xe = min(x0 + TX, box.own_hi[0]);
        uint32_t bytes_u = 0, bytes_b = 0;
        // ---- u row handled by this lane ----
        {
          int ly, lz, off; bool interior;
          if (lane < Cfg::kIntRows) { interior = true; ly = y0 + lane % TY; lz = z0 + lane / TY; off = lane * RSI; }
          else {
            interior = false; const int h = lane - Cfg::kIntRows; off = Cfg::kIntRows * RSI + h * RSH;
            if (h < TZ) { ly = y0 - 1; lz = z0 + h; } else if (h < 2 * TZ) { ly = y0 + TY; lz = z0 + h - TZ; }
            else if (h < 2 * TZ + TY) { ly = y0 + h - 2 * TZ; lz = z0 - 1; } else { ly = y0 + h - 2 * TZ - TY; lz = z0 + TZ; }
          }
          if (lane < Cfg::kIntRows + Cfg::kHaloRows && ly >= 0 && ly < box.n[1] && lz >= 0 && lz < box.n[2]) {
            const long long row_e = (long long)box.n[0] * (ly + (long long)box.n[1] * lz);
            const int par = (int)((ub8 + row_e + x0) & 1);
            const int first = interior ? max(x0 - 1, 0) : x0;                       // needed range [first, last)
            const int last = interior ? min(xe + 1, box.n[0]) : xe;
            int xs = interior ? x0 - 2 + par : x0 + par;                            // 16-byte aligned start candidates
            if (xs < 0) xs += 2;
            const int xlim = interior ? min(x0 + TX + 2, box.n[0]) : xe;
            const int cnt = max(xlim - xs, 0) & ~1;
            double* rowp = su + off + par + (interior ? 2 * N3 : 0);                // element x0
            if (cnt > 0) { bytes_u = (uint32_t)cnt * N3 * 8; ptx::bulk_g2s(ptx::smem_addr(rowp + (xs - x0) * N3), u + (row_e + xs) * N3, bytes_u, full_a + 8 * s); }
            for (int x = first; x < last; ++x) {                                     // leftovers not covered by the bulk copy
              if (cnt > 0 && x >= xs && x < xs + cnt) continue;
              for (int j = 0; j < N3; ++j) rowp[(x - x0) * N3 + j] = u[(row_e + x) * N3 + j];
            }
          }
        }
        // ---- load-vector row handled by this lane ----
        if (bvec && lane < Cfg::kOutRows) {
          const int ly = y0 + lane % TY, lz = z0 + lane / TY;
          if (ly < box.own_hi[1] && lz < box.own_hi[2]) {
            const long long row_e = (long long)box.n[0] * (ly + (long long)box.n[1] * lz);
            const int par = (int)((wb8 + row_e + x0) & 1), xs = x0 + par, cnt = max(xe - xs, 0) & ~1;
            double* rowp = so + lane * RSH + par;
            if (cnt > 0) { bytes_b = (uint32_t)cnt * N3 * 8; ptx::bulk_g2s(ptx::smem_addr(rowp + (xs - x0) * N3), bvec + (row_e + xs) * N3, bytes_b, full_a + 8 * s); }
            for (int x = x0; x < xe; ++x) {
              if (cnt > 0 && x >= xs && x < xs + cnt) continue;
              for (int j = 0; j < N3; ++j) rowp[(x - x0) * N3 + j] = bvec[(row_e + x) * N3 + j];
            }
          }
        }
        uint32_t total = bytes_u + bytes_b;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        __syncwarp();
        if (lane == 0) ptx::mbar_expect_tx(full_a + 8 * s, total);                  // release: plain leftover writes are ordered before
      }
      if (it >= 1) {
        const int ptile = blockIdx.x + (it - 1) * gridDim.x;
        if (ptile < ntiles) {
          const int sp = (it - 1) & 1;
          double* so = sbase + (size_t)sp * (Cfg::kStageU + Cfg::kStageO) + Cfg::kStageU;
          ptx::mbar_wait(done_a + 8 * sp, ((it - 1) >> 1) & 1);
          int x0, y0, z0; tile_origin(ptile, x0, y0, z0);
          const int xe = min(x0 + TX, box.own_hi[0]);
          if (lane < Cfg::kOutRows) {
            const int ly = y0 + lane % TY, lz = z0 + lane / TY;
            if (ly < box.own_hi[1] && lz < box.own_hi[2]) {
              const long long row_e = (long long)box.n[0] * (ly + (long long)box.n[1] * lz);
              const int par = (int)((wb8 + row_e + x0) & 1), xs = x0 + par, cnt = max(xe - xs, 0) & ~1;
              const double* rowp = so + lane * RSH + par;
              if (cnt > 0) ptx::bulk_s2g(w + (row_e + xs) * N3, ptx::smem_addr(rowp + (xs - x0) * N3), (uint32_t)cnt * N3 * 8);
              for (int x = x0; x < xe; ++x) {
                if (cnt > 0 && x >= xs && x < xs + cnt) continue;
                for (int j = 0; j < N3; ++j) w[(row_e + x) * N3 + j] = rowp[(x - x0) * N3 + j];
              }
            }
          }
          ptx::bulk_commit();
        }
      }
      if (!has) break;
    }
    ptx::bulk_wait_all();
    return;
  }

  // ============================== consumer warps: one thread per element ==============================
  const int tx = tid % TX, ty = (tid / TX) % TY, tz = tid / (TX * TY);
  for (int it = 0;; ++it) {
    const int tile = blockIdx.x + it * gridDim.x;
    if (tile >= ntiles) break;
    const int s = it & 1;
    const double* su = sbase + (size_t)s * (Cfg::kStageU + Cfg::kStageO);
    double* so = const_cast<double*>(su) + Cfg::kStageU;
    int x0, y0, z0; tile_origin(tile, x0, y0, z0);
    const int lx = x0 + tx, ly = y0 + ty, lz = z0 + tz;
    const bool active = lx < box.own_hi[0] && ly < box.own_hi[1] && lz < box.own_hi[2];
    // phase pads of the five rows this thread reads (all zero when the row length is even and vectors are 16-byte aligned)
    const long long n0 = box.n[0], n1 = box.n[1];
    auto rowpar = [&](int yy, int zz, int base8) { return (int)((base8 + n0 * (yy + n1 * zz) + x0) & 1); };
    const double* own = su + (ty + TY * tz) * RSI + rowpar(ly, lz, ub8) + (tx + 2) * N3;
    const double* ylo = ty > 0 ? own - RSI - rowpar(ly, lz, ub8) + rowpar(ly - 1, lz, ub8)
                               : su + Cfg::kIntRows * RSI + tz * RSH + rowpar(ly - 1, lz, ub8) + tx * N3;
    const double* yhi = ty < TY - 1 ? own + RSI - rowpar(ly, lz, ub8) + rowpar(ly + 1, lz, ub8)
                                    : su + Cfg::kIntRows * RSI + (TZ + tz) * RSH + rowpar(ly + 1, lz, ub8) + tx * N3;
    const double* zlo = tz > 0 ? own - TY * RSI - rowpar(ly, lz, ub8) + rowpar(ly, lz - 1, ub8)
                               : su + Cfg::kIntRows * RSI + (2 * TZ + ty) * RSH + rowpar(ly, lz - 1, ub8) + tx * N3;
    const double* zhi = tz < TZ - 1 ? own + TY * RSI - rowpar(ly, lz, ub8) + rowpar(ly, lz + 1, ub8)
                                    : su + Cfg::kIntRows * RSI + (2 * TZ + TY + ty) * RSH + rowpar(ly, lz + 1, ub8) + tx * N3;
    double* o = so + (ty + TY * tz) * RSH + rowpar(ly, lz, wb8) + tx * N3;

    ptx::mbar_wait(full_a + 8 * s, (it >> 1) & 1);
    if (active) {
      double acc[N3], v[N3];
#pragma unroll
      for (int t = 0; t < N3; ++t) { v[t] = own[P.p[t]]; acc[t] = 0; }
      apply_axis<N, 0>(K.S[0], v, acc); apply_axis<N, 1>(K.S[1], v, acc); apply_axis<N, 2>(K.S[2], v, acc);
      if (box.origin[0] + lx == 0) apply_axis<N, 0>(K.Dlo[0], v, acc);
      if (box.origin[0] + lx == box.gn[0] - 1) apply_axis<N, 0>(K.Dhi[0], v, acc);
      if (box.origin[1] + ly == 0) apply_axis<N, 1>(K.Dlo[1], v, acc);
      if (box.origin[1] + ly == box.gn[1] - 1) apply_axis<N, 1>(K.Dhi[1], v, acc);
      if (box.origin[2] + lz == 0) apply_axis<N, 2>(K.Dlo[2], v, acc);
      if (box.origin[2] + lz == box.gn[2] - 1) apply_axis<N, 2>(K.Dhi[2], v, acc);
      if (lx > 0) {
#pragma unroll
        for (int t = 0; t < N3; ++t) v[t] = own[P.p[t] - N3];
        apply_axis<N, 0>(K.L[0], v, acc); }
      if (lx < box.n[0] - 1) {
#pragma unroll
        for (int t = 0; t < N3; ++t) v[t] = own[P.p[t] + N3];
        apply_axis<N, 0>(K.R[0], v, acc); }
      if (ly > 0) {
#pragma unroll
        for (int t = 0; t < N3; ++t) v[t] = ylo[P.p[t]];
        apply_axis<N, 1>(K.L[1], v, acc); }
      if (ly < box.n[1] - 1) {
#pragma unroll
        for (int t = 0; t < N3; ++t) v[t] = yhi[P.p[t]];
        apply_axis<N, 1>(K.R[1], v, acc); }
      if (lz > 0) {
#pragma unroll
        for (int t = 0; t < N3; ++t) v[t] = zlo[P.p[t]];
        apply_axis<N, 2>(K.L[2], v, acc); }
      if (lz < box.n[2] - 1) {
#pragma unroll
        for (int t = 0; t < N3; ++t) v[t] = zhi[P.p[t]];
        apply_axis<N, 2>(K.R[2], v, acc); }
      if (bvec) {
#pragma unroll
        for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t] - o[P.p[t]];
      } else {
#pragma unroll
        for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t];
      }
    }
    ptx::fence_proxy_async();
    ptx::mbar_arrive(done_a + 8 * s);
  }
}

}  // namespace b200fem
