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

// shared-memory row strides are chosen == 8 (mod 16) doubles: the 4 tile rows a warp (8 x 4 lanes) touches with one
// 64-bit load then fall on complementary bank sets (lane stride 27 doubles covers the 16 even banks per row pair)
constexpr int bank_friendly(int doubles) { return doubles + ((8 - doubles % 16) + 16) % 16; }

template <int N, int TX, int TY, int TZ, bool SPLIT> struct KronPipeCfg {
  static constexpr int N3 = N * N * N;
  static constexpr int kElems = TX * TY * TZ;                       // elements per tile = threads per axis group
  static constexpr int kConsumers = (SPLIT ? 3 : 1) * kElems, kThreads = kConsumers + 32;
  static constexpr int RSI = bank_friendly((TX + 4) * N3 + 1);     // interior row: x0-2 .. x0+TX+1, + phase pad
  static constexpr int RSH = bank_friendly(TX * N3 + 1);           // halo / output row: x0 .. x0+TX-1
  static constexpr int kIntRows = TY * TZ, kHaloRows = 2 * TZ + 2 * TY, kOutRows = TY * TZ;
  static constexpr int kStageU = kIntRows * RSI + kHaloRows * RSH; // doubles
  static constexpr int kStageO = kOutRows * RSH;
  static constexpr int kStages = 2;
  static constexpr size_t smem_bytes() { return sizeof(double) * ((size_t)kStages * (kStageU + kStageO) + N3 + 1) + 64; }
};

namespace ptx {
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}  // namespace ptx

// acc[.. i ..] += sum_j M[i*N+j] src[perm(.. j ..)] along tensor axis AX, reading the source element line by line from
// shared memory (keeps only one line of N values live besides the accumulators)
template <int N, int AX, bool HIER>
__device__ __forceinline__ void apply_axis_smem(const double* __restrict__ M, const double* __restrict__ src, double (&acc)[N * N * N]) {
  constexpr PermTable<N, HIER> P{};
  constexpr int st = AX == 0 ? N * N : AX == 1 ? N : 1;
#pragma unroll
  for (int l = 0; l < N * N; ++l) {
    // base tensor index of line l (all indices except the one along AX)
    const int base = AX == 0 ? l : AX == 1 ? (l / N) * N * N + (l % N) : l * N;
    double line[N];
#pragma unroll
    for (int j = 0; j < N; ++j) line[j] = src[P.p[base + j * st]];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double a = acc[base + i * st];
#pragma unroll
      for (int j = 0; j < N; ++j) a = fma(M[i * N + j], line[j], a);
      acc[base + i * st] = a;
    }
  }
}

template <int N, bool HIER, int TX, int TY, int TZ, bool SPLIT>
__global__ void __launch_bounds__(KronPipeCfg<N, TX, TY, TZ, SPLIT>::kThreads, 1)
dg_kronecker_pipe_kernel(const __grid_constant__ KronTabDev<N> K, const __grid_constant__ BoxDev box,
                         const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ bvec,
                         int tiles_x, int tiles_y, int ntiles, int debug_skip, long long* dbg) {
  using Cfg = KronPipeCfg<N, TX, TY, TZ, SPLIT>;
  constexpr int N3 = Cfg::N3, RSI = Cfg::RSI, RSH = Cfg::RSH;
  static_assert(N3 % 2 == 1, "16-byte phase logic assumes an odd number of doubles per element");
  static_assert(Cfg::kIntRows + Cfg::kHaloRows <= 32 && Cfg::kOutRows <= 32, "one producer lane per row");
  constexpr PermTable<N, HIER> P{};
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sbase = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbase + (size_t)Cfg::kStages * (Cfg::kStageU + Cfg::kStageO));
  double* zeros = reinterpret_cast<double*>(bars + 8);              // a zero element standing in for missing neighbours
  const uint32_t full_a = ptx::smem_addr(bars), done_a = ptx::smem_addr(bars + 2);
  const int tid = threadIdx.x;
  const int ub8 = (int)((reinterpret_cast<uintptr_t>(u) >> 3) & 1), wb8 = (int)((reinterpret_cast<uintptr_t>(w) >> 3) & 1);

  if (tid < N3) zeros[tid] = 0.0;
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
    // Steady state for tile `it` being computed:  wait done(it-1)  ->  u rows of tile it+1 into the stage it-1 just
    // vacated (issued first: they are the long pole)  ->  bulk store of tile it-1  ->  once that store has read its
    // rows, the load-vector rows of tile it+1 go into the same output rows  ->  arrive on full(it+1).
    const int lane = tid - Cfg::kConsumers;
    auto load_u = [&](int tile, int s) -> uint32_t {
      double* su = sbase + (size_t)s * (Cfg::kStageU + Cfg::kStageO);
      int x0, y0, z0; tile_origin(tile, x0, y0, z0);
      const int xe = min(x0 + TX, box.own_hi[0]);
      uint32_t bytes_u = 0;
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
      return bytes_u;
    };
    auto load_b = [&](int tile, int s) -> uint32_t {
      uint32_t bytes_b = 0;
      if (!bvec || lane >= Cfg::kOutRows) return 0;
      double* so = sbase + (size_t)s * (Cfg::kStageU + Cfg::kStageO) + Cfg::kStageU;
      int x0, y0, z0; tile_origin(tile, x0, y0, z0);
      const int xe = min(x0 + TX, box.own_hi[0]);
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
      return bytes_b;
    };
    auto arrive_full = [&](int s, uint32_t bytes) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
      __syncwarp();
      if (lane == 0) ptx::mbar_expect_tx(full_a + 8 * s, bytes);                // release: plain leftover writes are ordered before
    };
    auto store_w = [&](int tile, int s) {
      const double* so = sbase + (size_t)s * (Cfg::kStageU + Cfg::kStageO) + Cfg::kStageU;
      int x0, y0, z0; tile_origin(tile, x0, y0, z0);
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
    };
    // prologue: the first two tiles
    for (int it = 0; it < 2; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      if (tile < ntiles) { uint32_t b = load_u(tile, it); b += load_b(tile, it); arrive_full(it, b); }
    }
    for (int it = 1;; ++it) {                                                   // tile `it` is (about to be) computed
      const int ptile = blockIdx.x + (it - 1) * gridDim.x;                      // the tile whose stage frees up next
      if (ptile >= ntiles) break;
      const int sp = (it - 1) & 1;
      ptx::mbar_wait(done_a + 8 * sp, ((it - 1) >> 1) & 1);
      if (dbg && blockIdx.x == 0 && lane == 0) dbg[8 * it + 0] = clock64();
      const int ntile = blockIdx.x + (it + 1) * gridDim.x;
      uint32_t bytes = 0;
      if (ntile < ntiles) bytes = load_u(ntile, sp);
      if (dbg && blockIdx.x == 0 && lane == 0) dbg[8 * it + 1] = clock64();
      store_w(ptile, sp);
      if (dbg && blockIdx.x == 0 && lane == 0) dbg[8 * it + 2] = clock64();
      if (ntile < ntiles) {
        ptx::bulk_wait_read();                                                  // tile it-1's rows have left the output stage
        __syncwarp();
        if (dbg && blockIdx.x == 0 && lane == 0) dbg[8 * it + 3] = clock64();
        bytes += load_b(ntile, sp);
        arrive_full(sp, bytes);
        if (dbg && blockIdx.x == 0 && lane == 0) dbg[8 * it + 4] = clock64();
      }
    }
    ptx::bulk_wait_all();
    return;
  }

  // ============================== consumer warps ==============================
  // axis split: group g = tid / kElems (4 warps) applies the three 1-D operators of axis g to element e = tid % kElems:
  //   acc_g = S_g u_K + L_g u_{K-e_g} + R_g u_{K+e_g}     (243 FMA, 81 shared loads per thread)
  // the groups then combine in the output row: g=0 writes acc_0 - b, g=1 and g=2 add theirs (two named barriers).
  const int g = tid / Cfg::kElems, e = tid % Cfg::kElems;
  const int tx = e % TX, ty = (e / TX) % TY, tz = e / (TX * TY);
  for (int it = 0;; ++it) {
    const int tile = blockIdx.x + it * gridDim.x;
    if (tile >= ntiles) break;
    const int s = it & 1;
    const double* su = sbase + (size_t)s * (Cfg::kStageU + Cfg::kStageO);
    double* so = const_cast<double*>(su) + Cfg::kStageU;
    int x0, y0, z0; tile_origin(tile, x0, y0, z0);
    const int lx = x0 + tx, ly = y0 + ty, lz = z0 + tz;
    const bool active = lx < box.own_hi[0] && ly < box.own_hi[1] && lz < box.own_hi[2];
    const long long n0 = box.n[0], n1 = box.n[1];
    auto rowpar = [&](int yy, int zz, int base8) { return (int)((base8 + n0 * (yy + n1 * zz) + x0) & 1); };
    const int pown = rowpar(ly, lz, ub8);
    const double* own = su + (ty + TY * tz) * RSI + pown + (tx + 2) * N3;
    // neighbours along axis a; a missing neighbour (box boundary) reads the zero element, so the FMA section is
    // branch-free
    auto neighbours = [&](int a, const double*& lo, const double*& hi) {
      int lc, nax;
      if (a == 0) { lo = own - N3; hi = own + N3; lc = lx; nax = box.n[0]; }
      else if (a == 1) {
        lo = ty > 0 ? own - RSI - pown + rowpar(ly - 1, lz, ub8) : su + Cfg::kIntRows * RSI + tz * RSH + rowpar(ly - 1, lz, ub8) + tx * N3;
        hi = ty < TY - 1 ? own + RSI - pown + rowpar(ly + 1, lz, ub8) : su + Cfg::kIntRows * RSI + (TZ + tz) * RSH + rowpar(ly + 1, lz, ub8) + tx * N3;
        lc = ly; nax = box.n[1];
      } else {
        lo = tz > 0 ? own - TY * RSI - pown + rowpar(ly, lz - 1, ub8) : su + Cfg::kIntRows * RSI + (2 * TZ + ty) * RSH + rowpar(ly, lz - 1, ub8) + tx * N3;
        hi = tz < TZ - 1 ? own + TY * RSI - pown + rowpar(ly, lz + 1, ub8) : su + Cfg::kIntRows * RSI + (2 * TZ + TY + ty) * RSH + rowpar(ly, lz + 1, ub8) + tx * N3;
        lc = lz; nax = box.n[2];
      }
      if (!(lc > 0)) lo = zeros;
      if (!(lc < nax - 1)) hi = zeros;
    };
    const int lcs[3] = {lx, ly, lz};
    double* o = so + (ty + TY * tz) * RSH + rowpar(ly, lz, wb8) + tx * N3;
    double acc[N3];
#pragma unroll
    for (int t = 0; t < N3; ++t) acc[t] = 0;

    if (dbg && blockIdx.x == 0 && tid == 0) dbg[8 * it + 5] = clock64();
    if constexpr (SPLIT) {
      const double *lo, *hi; neighbours(g, lo, hi);
      const int gc = box.origin[g] + lcs[g];                            // global coordinate along the axis
      const bool bnd_lo = gc == 0, bnd_hi = gc == box.gn[g] - 1;
      ptx::mbar_wait(full_a + 8 * s, (it >> 1) & 1);
      if (active) {
        if (g == 0) {
          apply_axis_smem<N, 0, HIER>(K.S[0], own, acc); apply_axis_smem<N, 0, HIER>(K.L[0], lo, acc); apply_axis_smem<N, 0, HIER>(K.R[0], hi, acc);
          if (bnd_lo) apply_axis_smem<N, 0, HIER>(K.Dlo[0], own, acc);
          if (bnd_hi) apply_axis_smem<N, 0, HIER>(K.Dhi[0], own, acc);
        } else if (g == 1) {
          apply_axis_smem<N, 1, HIER>(K.S[1], own, acc); apply_axis_smem<N, 1, HIER>(K.L[1], lo, acc); apply_axis_smem<N, 1, HIER>(K.R[1], hi, acc);
          if (bnd_lo) apply_axis_smem<N, 1, HIER>(K.Dlo[1], own, acc);
          if (bnd_hi) apply_axis_smem<N, 1, HIER>(K.Dhi[1], own, acc);
        } else {
          apply_axis_smem<N, 2, HIER>(K.S[2], own, acc); apply_axis_smem<N, 2, HIER>(K.L[2], lo, acc); apply_axis_smem<N, 2, HIER>(K.R[2], hi, acc);
          if (bnd_lo) apply_axis_smem<N, 2, HIER>(K.Dlo[2], own, acc);
          if (bnd_hi) apply_axis_smem<N, 2, HIER>(K.Dhi[2], own, acc);
        }
      }
      if (g == 0 && active) {
        if (bvec) {
#pragma unroll
          for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t] - o[P.p[t]];
        } else {
#pragma unroll
          for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kConsumers) : "memory");
      if (g == 1 && active) {
#pragma unroll
        for (int t = 0; t < N3; ++t) o[P.p[t]] += acc[t];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kConsumers) : "memory");
      if (g == 2 && active) {
#pragma unroll
        for (int t = 0; t < N3; ++t) o[P.p[t]] += acc[t];
      }
    } else {
      const double *xlo, *xhi, *ylo, *yhi, *zlo, *zhi;
      neighbours(0, xlo, xhi); neighbours(1, ylo, yhi); neighbours(2, zlo, zhi);
      ptx::mbar_wait(full_a + 8 * s, (it >> 1) & 1);
      if (dbg && blockIdx.x == 0 && tid == 0) dbg[8 * it + 6] = clock64();
      if (active) {
        double v[N3];
#pragma unroll
        for (int t = 0; t < N3; ++t) v[t] = own[P.p[t]];
        if (debug_skip) {
#pragma unroll
          for (int t = 0; t < N3; ++t) acc[t] = v[t];
        } else {
        apply_axis<N, 0>(K.S[0], v, acc); apply_axis<N, 1>(K.S[1], v, acc); apply_axis<N, 2>(K.S[2], v, acc);
        apply_axis_smem<N, 0, HIER>(K.L[0], xlo, acc); apply_axis_smem<N, 0, HIER>(K.R[0], xhi, acc);
        apply_axis_smem<N, 1, HIER>(K.L[1], ylo, acc); apply_axis_smem<N, 1, HIER>(K.R[1], yhi, acc);
        apply_axis_smem<N, 2, HIER>(K.L[2], zlo, acc); apply_axis_smem<N, 2, HIER>(K.R[2], zhi, acc);
        // domain-boundary corrections of the self matrix (only in boundary tiles)
        if (box.origin[0] + lx == 0) apply_axis<N, 0>(K.Dlo[0], v, acc);
        if (box.origin[0] + lx == box.gn[0] - 1) apply_axis<N, 0>(K.Dhi[0], v, acc);
        if (box.origin[1] + ly == 0) apply_axis<N, 1>(K.Dlo[1], v, acc);
        if (box.origin[1] + ly == box.gn[1] - 1) apply_axis<N, 1>(K.Dhi[1], v, acc);
        if (box.origin[2] + lz == 0) apply_axis<N, 2>(K.Dlo[2], v, acc);
        if (box.origin[2] + lz == box.gn[2] - 1) apply_axis<N, 2>(K.Dhi[2], v, acc);
        }
        if (bvec) {
#pragma unroll
          for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t] - o[P.p[t]];
        } else {
#pragma unroll
          for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t];
        }
      }
    }
    ptx::fence_proxy_async();
    ptx::mbar_arrive(done_a + 8 * s);
    if (dbg && blockIdx.x == 0 && tid == 0) dbg[8 * it + 7] = clock64();
  }
}

}  // namespace b200fem
