// dg_kronecker_tensor.cuh -- Kronecker-form DG apply, v4: persistent, warp-specialised, TMA tensor tiles.
//
// The dof vector of a DG space on a Cartesian box is a dense 3-D tensor [z][y][x*n^3] of doubles (element blocks are
// contiguous along x: function/blockvectors/defaultblockvectors.hh:284-294 + lexicographic element order).  A tile of
// TX x TY x TZ elements and each of its six face-halo slabs is therefore one TMA box:
//     cp.async.bulk.tensor.3d  global -> shared   (UTMALDG)    7 boxes of u  (+1 box of the load vector b)
//     cp.async.bulk.tensor.3d  shared -> global   (UTMASTG)    1 box of w
// i.e. 8-9 TMA instructions per tile, issued by one elected thread, instead of 64 per-row bulk copies (the per-row
// variant, dg_kronecker_pipe.cuh, is issue-bound: ~110 cycles per copy from one warp).  Boxes that stick out of the
// rank-local tensor are zero-filled by the TMA unit, which is exactly "no neighbour": the consumer code has no
// boundary branches for the neighbour terms.  The w / b maps cover only the owned sub-box, so partial tiles are
// clipped by the hardware on store.
//
// Pipeline (one CTA per SM, persistent over tiles, 2 stages): warp 4 lane 0 = producer, warps 0-3 = consumers
// (one thread per element, u_K and w_K in registers, neighbours streamed from shared memory).
// Requirements: n0 even (global strides must be multiples of 16 bytes) and 16-byte aligned vectors; otherwise the
// per-row kernel is used.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "dg_kronecker_pipe.cuh"
#include "halo.cuh"

namespace b200fem {

struct KronTensorMaps {
  CUtensorMap u_tile, u_xhalo, u_yhalo, u_zhalo;   // boxes {TX*N3,TY,TZ}, {2*N3,TY,TZ}, {TX*N3,1,TZ}, {TX*N3,TY,1} over the local box
  CUtensorMap b_tile, w_tile;                      // box {TX*N3,TY,TZ} over the owned sub-box
};

// Fused halo send: the eight neighbours in the y-z plane (faces and edges), direction code d = (dy+1) + 3*(dz+1).
// remote[d][b] is the neighbour's mailbox buffer b for my message: a dense [z-part][y-part][on0*N3] array where a part
// is the whole owned range for a zero direction component and the single interface layer otherwise.
struct KronSendDev {
  int any; int enabled[9];
  double* remote[9][2]; unsigned long long* remote_ready[9]; const unsigned long long* local_ack[9];
  unsigned int* dir_counter; unsigned int expected[9];     // tiles that contribute to each message
  unsigned long long seq; int* error;
};

namespace ptx {
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
}  // namespace ptx

template <int N, int TX, int TY, int TZ, bool SPLIT> struct KronTensorCfg {
  static constexpr int N3 = N * N * N;
  static constexpr int kElems = TX * TY * TZ, kConsumers = (SPLIT ? 3 : 1) * kElems, kThreads = kConsumers + 32;
  static constexpr int RX = TX * N3;                                  // doubles per tile row (== 8 mod 16 for TX=8, N=3: conflict-free)
  static constexpr int kA = TZ * TY * RX, kXH = TZ * TY * 2 * N3, kYH = TZ * RX, kZH = TY * RX;   // doubles
  static constexpr int kStage = 2 * kA + 2 * kXH + 2 * kYH + 2 * kZH; // u tile, halos, output tile
  static constexpr int kStages = 2;
  static constexpr uint32_t kBytesU = 8u * (kA + 2 * kXH + 2 * kYH + 2 * kZH), kBytesB = 8u * kA;
  static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)kStages * kStage + 64 + 128; }
  static_assert((kA * 8) % 128 == 0 && (kXH * 8) % 128 == 0 && (kYH * 8) % 128 == 0 && (kZH * 8) % 128 == 0, "TMA destinations must stay 128-byte aligned");
};

template <int N, bool HIER, int TX, int TY, int TZ, bool SPLIT>
__global__ void __launch_bounds__(KronTensorCfg<N, TX, TY, TZ, SPLIT>::kThreads, 1)
dg_kronecker_tensor_kernel(const __grid_constant__ KronTabDev<N> K, const __grid_constant__ BoxDev box,
                           const __grid_constant__ KronTensorMaps M, const __grid_constant__ KronSendDev SND, const int has_b,
                           int tiles_x, int tiles_y, int ntiles) {
  using Cfg = KronTensorCfg<N, TX, TY, TZ, SPLIT>;
  constexpr int N3 = Cfg::N3, RX = Cfg::RX;
  constexpr PermTable<N, HIER> P{};
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  // 128-byte aligned base, computed as an offset so that the pointer keeps its shared-memory address space (a
  // uintptr_t round trip would turn every access into a generic LD/ST)
  double* sbase = reinterpret_cast<double*>(smem_dyn + ((128u - (ptx::smem_addr(smem_dyn) & 127u)) & 127u));
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbase + (size_t)Cfg::kStages * Cfg::kStage);
  const uint32_t full_a = ptx::smem_addr(bars), done_a = ptx::smem_addr(bars + 2);
  const int tid = threadIdx.x;

  if (tid == 0) {
    ptx::mbar_init(full_a, 1); ptx::mbar_init(full_a + 8, 1);
    ptx::mbar_init(done_a, Cfg::kConsumers); ptx::mbar_init(done_a + 8, Cfg::kConsumers);
    ptx::fence_barrier_init(); ptx::fence_proxy_async();
  }
  __syncthreads();

  // tile order: the first and the last tile layer of y and z come first, so that tiles on rank interfaces are computed
  // (and their halo rows sent) at the very beginning of the kernel
  const int tiles_z = ntiles / (tiles_x * tiles_y);
  auto front = [](int i, int n) { return i == 0 ? 0 : (i == 1 ? n - 1 : i - 1); };
  auto tile_origin = [&](int tile, int& x0, int& y0, int& z0) {
    const int bx = tile % tiles_x, by = front((tile / tiles_x) % tiles_y, tiles_y), bz = front(tile / (tiles_x * tiles_y), tiles_z);
    x0 = box.own_lo[0] + bx * TX; y0 = box.own_lo[1] + by * TY; z0 = box.own_lo[2] + bz * TZ;
  };
  // stage layout: [A | XL | XH | YL | YH | ZL | ZH | O]
  auto stage = [&](int s) { return sbase + (size_t)s * Cfg::kStage; };
  constexpr int oXL = Cfg::kA, oXH = oXL + Cfg::kXH, oYL = oXH + Cfg::kXH, oYH = oYL + Cfg::kYH, oZL = oYH + Cfg::kYH, oZH = oZL + Cfg::kZH, oO = oZH + Cfg::kZH;

  if (tid >= Cfg::kConsumers) {
    // ============================== producer: one elected thread ==============================
    if (tid != Cfg::kConsumers) return;
    ptx::prefetch_tensormap(&M.u_tile); ptx::prefetch_tensormap(&M.u_xhalo); ptx::prefetch_tensormap(&M.u_yhalo);
    ptx::prefetch_tensormap(&M.u_zhalo); ptx::prefetch_tensormap(&M.b_tile); ptx::prefetch_tensormap(&M.w_tile);
    auto load_u = [&](int tile, int s) {
      int x0, y0, z0; tile_origin(tile, x0, y0, z0);
      double* st = stage(s); const uint32_t bar = full_a + 8 * s;
      ptx::tma_load_3d(ptx::smem_addr(st), &M.u_tile, x0 * N3, y0, z0, bar);
      ptx::tma_load_3d(ptx::smem_addr(st + oXL), &M.u_xhalo, (x0 - 2) * N3, y0, z0, bar);
      ptx::tma_load_3d(ptx::smem_addr(st + oXH), &M.u_xhalo, (x0 + TX) * N3, y0, z0, bar);
      ptx::tma_load_3d(ptx::smem_addr(st + oYL), &M.u_yhalo, x0 * N3, y0 - 1, z0, bar);
      ptx::tma_load_3d(ptx::smem_addr(st + oYH), &M.u_yhalo, x0 * N3, y0 + TY, z0, bar);
      ptx::tma_load_3d(ptx::smem_addr(st + oZL), &M.u_zhalo, x0 * N3, y0, z0 - 1, bar);
      ptx::tma_load_3d(ptx::smem_addr(st + oZH), &M.u_zhalo, x0 * N3, y0, z0 + TZ, bar);
    };
    auto load_b = [&](int tile, int s) {
      int x0, y0, z0; tile_origin(tile, x0, y0, z0);
      ptx::tma_load_3d(ptx::smem_addr(stage(s) + oO), &M.b_tile, (x0 - box.own_lo[0]) * N3, y0 - box.own_lo[1], z0 - box.own_lo[2], full_a + 8 * s);
    };
    bool ack_checked = false;
    auto store_w = [&](int tile, int s) -> int {
      int x0, y0, z0; tile_origin(tile, x0, y0, z0);
      ptx::tma_store_3d(&M.w_tile, (x0 - box.own_lo[0]) * N3, y0 - box.own_lo[1], z0 - box.own_lo[2], ptx::smem_addr(stage(s) + oO));
      int mask = 0;
      if (SND.any) {
        // rows of this tile that lie on a rank interface go straight into the neighbour's mailbox (1-D bulk copies over
        // NVLink; row starts are 16-byte aligned because on0 and TX are even)
        const int on0 = box.own_hi[0] - box.own_lo[0], on1 = box.own_hi[1] - box.own_lo[1], xr = x0 - box.own_lo[0];
        const uint32_t bytes = (uint32_t)(min(TX, box.own_hi[0] - x0)) * N3 * 8;
        const int buf = (int)(SND.seq & 1);
        // is this tile at the low / high interface layer of y, z?  (index 0: low, 2: high)
        const bool aty[3] = {y0 == box.own_lo[1], true, y0 + TY >= box.own_hi[1]};
        const bool atz[3] = {z0 == box.own_lo[2], true, z0 + TZ >= box.own_hi[2]};
        for (int d = 0; d < 9; ++d) if (SND.enabled[d] && aty[d % 3] && atz[d / 3]) mask |= 1 << d;
        if (mask && !ack_checked) {                                  // the mailbox buffer was last used by message seq-2
          for (int d = 0; d < 9; ++d) if (SND.enabled[d] && SND.seq > 2) {
            const long long t0 = clock64();
            while (ld_acquire_sys(SND.local_ack[d]) < SND.seq - 2) if (clock64() - t0 > kP2PSpinLimit) { *SND.error = 3; break; }
          }
          ack_checked = true;
        }
        const double* O = stage(s) + oO;
        for (int d = 0; d < 9; ++d) if ((mask >> d) & 1) {
          const int dy = d % 3 - 1, dz = d / 3 - 1;
          const int ty_lo = dy < 0 ? 0 : dy > 0 ? box.own_hi[1] - 1 - y0 : 0, ty_hi = dy == 0 ? TY : ty_lo + 1;
          const int tz_lo = dz < 0 ? 0 : dz > 0 ? box.own_hi[2] - 1 - z0 : 0, tz_hi = dz == 0 ? TZ : tz_lo + 1;
          const int ny = dy == 0 ? on1 : 1;                            // rows per z-part of the message
          for (int tz_ = tz_lo; tz_ < tz_hi && z0 + tz_ < box.own_hi[2]; ++tz_)
            for (int ty_ = ty_lo; ty_ < ty_hi && y0 + ty_ < box.own_hi[1]; ++ty_) {
              const long long zi = dz == 0 ? z0 + tz_ - box.own_lo[2] : 0, yi = dy == 0 ? y0 + ty_ - box.own_lo[1] : 0;
              ptx::bulk_s2g(SND.remote[d][buf] + ((zi * ny + yi) * on0 + xr) * N3, ptx::smem_addr(O + (tz_ * TY + ty_) * RX), bytes);
            }
        }
      }
      ptx::bulk_commit();
      return mask;
    };
    // A tile's halo rows may only be counted once they have landed in the neighbour's memory.  Waiting right after the
    // store would stall the load pipeline for an NVLink round trip, so the accounting is deferred by one tile: by then the
    // bulk group is (almost always) complete already.  The message is published by whoever completes its last tile.
    auto publish = [&](int mask) {
      __threadfence_system();
      for (int d = 0; d < 9; ++d) if (((mask >> d) & 1) && atomicAdd(&SND.dir_counter[d], 1u) == SND.expected[d] - 1) {
        SND.dir_counter[d] = 0; __threadfence_system(); st_release_sys(SND.remote_ready[d] + (SND.seq & 1), SND.seq);
      }
    };
    int pending = 0;
    const uint32_t bytes = Cfg::kBytesU + (has_b ? Cfg::kBytesB : 0u);
    for (int it = 0; it < 2; ++it) {                                            // prologue: the first two tiles
      const int tile = blockIdx.x + it * gridDim.x;
      if (tile < ntiles) { load_u(tile, it); if (has_b) load_b(tile, it); ptx::mbar_expect_tx(full_a + 8 * it, bytes); }
    }
    for (int it = 1;; ++it) {                                                   // tile `it` is (about to be) computed
      const int ptile = blockIdx.x + (it - 1) * gridDim.x;                      // the tile whose stage frees up next
      if (ptile >= ntiles) break;
      const int sp = (it - 1) & 1;
      ptx::mbar_wait(done_a + 8 * sp, ((it - 1) >> 1) & 1);
      const int ntile = blockIdx.x + (it + 1) * gridDim.x;
      if (ntile < ntiles) load_u(ntile, sp);                                    // long pole first
      const int mask = store_w(ptile, sp);
      if (pending) { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); publish(pending); }   // the previous tile's group is complete
      pending = mask;
      if (ntile < ntiles) {
        ptx::bulk_wait_read();                                                  // tile it-1's output has left the stage: its slot
        if (has_b) load_b(ntile, sp);                                           // may receive the next load-vector tile
        ptx::mbar_expect_tx(full_a + 8 * sp, bytes);                            // only now may the phase complete (consumers write O)
      }
    }
    ptx::bulk_wait_all();
    if (pending) publish(pending);
    return;
  }

  // ============================== consumers ==============================
  // SPLIT = false: one thread per element applies all nine 1-D operators (u_K, w_K in registers).
  // SPLIT = true : three groups of kElems threads; group g applies S_g, L_g, R_g (243 FMA) and the groups combine in the
  //                output tile (g=0 writes acc - b, g=1 and g=2 add) -- 12 consumer warps hide FMA / shared-load latency.
  const int g = SPLIT ? tid / Cfg::kElems : 0, e = tid % Cfg::kElems;
  const int tx = e % TX, ty = (e / TX) % TY, tz = e / (TX * TY);
  const int row = tz * TY + ty;
  for (int it = 0;; ++it) {
    const int tile = blockIdx.x + it * gridDim.x;
    if (tile >= ntiles) break;
    const int s = it & 1;
    const double* st = stage(s);
    int x0, y0, z0; tile_origin(tile, x0, y0, z0);
    const int lx = x0 + tx, ly = y0 + ty, lz = z0 + tz;
    const double* own = st + row * RX + tx * N3;
    const double* xlo = tx > 0 ? own - N3 : st + oXL + row * 2 * N3 + N3;       // x0-1 is the second element of the low x-halo
    const double* xhi = tx < TX - 1 ? own + N3 : st + oXH + row * 2 * N3;
    const double* ylo = ty > 0 ? own - RX : st + oYL + tz * RX + tx * N3;
    const double* yhi = ty < TY - 1 ? own + RX : st + oYH + tz * RX + tx * N3;
    const double* zlo = tz > 0 ? own - TY * RX : st + oZL + ty * RX + tx * N3;
    const double* zhi = tz < TZ - 1 ? own + TY * RX : st + oZH + ty * RX + tx * N3;
    double* o = const_cast<double*>(st) + oO + row * RX + tx * N3;
    const int gcx = box.origin[0] + lx, gcy = box.origin[1] + ly, gcz = box.origin[2] + lz;
    double acc[N3];
#pragma unroll
    for (int t = 0; t < N3; ++t) acc[t] = 0;

    ptx::mbar_wait(full_a + 8 * s, (it >> 1) & 1);
    if constexpr (SPLIT) {
      if (g == 0) {
        apply_axis_smem<N, 0, HIER>(K.S[0], own, acc); apply_axis_smem<N, 0, HIER>(K.L[0], xlo, acc); apply_axis_smem<N, 0, HIER>(K.R[0], xhi, acc);
        if (gcx == 0) apply_axis_smem<N, 0, HIER>(K.Dlo[0], own, acc);
        if (gcx == box.gn[0] - 1) apply_axis_smem<N, 0, HIER>(K.Dhi[0], own, acc);
      } else if (g == 1) {
        apply_axis_smem<N, 1, HIER>(K.S[1], own, acc); apply_axis_smem<N, 1, HIER>(K.L[1], ylo, acc); apply_axis_smem<N, 1, HIER>(K.R[1], yhi, acc);
        if (gcy == 0) apply_axis_smem<N, 1, HIER>(K.Dlo[1], own, acc);
        if (gcy == box.gn[1] - 1) apply_axis_smem<N, 1, HIER>(K.Dhi[1], own, acc);
      } else {
        apply_axis_smem<N, 2, HIER>(K.S[2], own, acc); apply_axis_smem<N, 2, HIER>(K.L[2], zlo, acc); apply_axis_smem<N, 2, HIER>(K.R[2], zhi, acc);
        if (gcz == 0) apply_axis_smem<N, 2, HIER>(K.Dlo[2], own, acc);
        if (gcz == box.gn[2] - 1) apply_axis_smem<N, 2, HIER>(K.Dhi[2], own, acc);
      }
      if (g == 0) {
        if (has_b) {
#pragma unroll
          for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t] - o[P.p[t]];
        } else {
#pragma unroll
          for (int t = 0; t < N3; ++t) o[P.p[t]] = acc[t];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kConsumers) : "memory");
      if (g == 1) {
#pragma unroll
        for (int t = 0; t < N3; ++t) o[P.p[t]] += acc[t];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kConsumers) : "memory");
      if (g == 2) {
#pragma unroll
        for (int t = 0; t < N3; ++t) o[P.p[t]] += acc[t];
      }
    } else {
      double v[N3];
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = own[P.p[t]];
      apply_axis<N, 0>(K.S[0], v, acc); apply_axis<N, 1>(K.S[1], v, acc); apply_axis<N, 2>(K.S[2], v, acc);
      apply_axis_smem<N, 0, HIER>(K.L[0], xlo, acc); apply_axis_smem<N, 0, HIER>(K.R[0], xhi, acc);
      apply_axis_smem<N, 1, HIER>(K.L[1], ylo, acc); apply_axis_smem<N, 1, HIER>(K.R[1], yhi, acc);
      apply_axis_smem<N, 2, HIER>(K.L[2], zlo, acc); apply_axis_smem<N, 2, HIER>(K.R[2], zhi, acc);
      // domain-boundary corrections of the self matrix (only in boundary tiles)
      if (gcx == 0) apply_axis<N, 0>(K.Dlo[0], v, acc);
      if (gcx == box.gn[0] - 1) apply_axis<N, 0>(K.Dhi[0], v, acc);
      if (gcy == 0) apply_axis<N, 1>(K.Dlo[1], v, acc);
      if (gcy == box.gn[1] - 1) apply_axis<N, 1>(K.Dhi[1], v, acc);
      if (gcz == 0) apply_axis<N, 2>(K.Dlo[2], v, acc);
      if (gcz == box.gn[2] - 1) apply_axis<N, 2>(K.Dhi[2], v, acc);
      if (has_b) {
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
