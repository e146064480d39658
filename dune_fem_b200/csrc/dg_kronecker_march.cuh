// dg_kronecker_march.cuh -- Kronecker-form DG apply, v5: persistent z-marching with a register pipeline.
//
// Same arithmetic as dg_kronecker.cuh:  w_K = sum_d [ S_d u_K + L_d u_{K-e_d} + R_d u_{K+e_d} ] - b_K.
// What changed against the tile kernel (dg_kronecker_tensor.cuh) is the data movement.  Measured there (profiles/
// r01_dg_kronecker_tensor.md): all global->SM traffic, L2 hits included, is capped near the DRAM rate, so the 2.5x halo
// re-read of 8x4x4 tiles -- not DRAM -- bounded the kernel, and 4 consumer warps could not hide shared-load latency.
//
//  * A CTA owns a column of TX x TY elements and marches through z.  One thread per element keeps THREE things in
//    registers: u_K of the current plane and the two partial results w(z-1), w(z).  When plane z arrives,
//        w(z-1) += R_z u(z)   -> complete, written out;      w(z+1)  = L_z u(z);      w(z) += S u(z) + x/y neighbours
//    so z-neighbours are never read from shared memory and never re-read from L2: only the two end planes of a run
//    are loaded twice (as own-only boxes without x/y halo).
//  * One plane incl. its x/y halo is ONE 4-D TMA box (cp.async.bulk.tensor.4d) over the view [z][y][x/2][2*n^3] of the
//    dof vector (element pairs make the innermost extent a multiple of 16 bytes); x/y neighbours are at constant
//    offsets (+-n^3, +-row) inside that box.  Out-of-range parts are zero-filled by the TMA unit = "no neighbour".
//  * 8 consumer warps (16x16 elements per plane) + 1 producer warp (setmaxnreg moves its registers to the consumers),
//    2-stage plane ring.  Output is warp-autonomous: every warp has its own staging slab (its two tile rows) that
//    receives the rows of the load vector b (TMA, per-warp mbarrier), is overwritten with A u - b and leaves through
//    a TMA store issued by lane 0 -- no CTA-wide barrier and no producer round trip on the output path.
//  * The plane-tiles (column, z) are split evenly over the persistent grid: every CTA gets the same number of planes.
//
// Traffic per dof (C2, 148 CTAs, runs of ~7 planes): u is read 1.7x instead of 2.5x; shared-memory loads per element
// drop from 7 to 5 element blocks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "comm.cuh"
#include "kron_common.cuh"
#include "march_schedule.hpp"

namespace b200fem {

struct KronMarchMaps {
  CUtensorMap u_plane;   // box {2*N3, (TX+4)/2, TY+2, 1} over the local box (owned + ghost)
  CUtensorMap u_edge;    // box {2*N3, TX/2, TY, 1}       over the local box
  CUtensorMap b_tile;    // box {2*N3, TX/2, 32/TX, 1}    over the owned sub-box of the load vector: the rows of one warp
  CUtensorMap w_tile;    // same over the owned sub-box of w
};


template <int N, int TX, int TY> struct KronMarchCfg {
  static constexpr int N3 = N * N * N;
  static constexpr int kConsumers = TX * TY, kThreads = kConsumers + 128;   // 2 consumer warpgroups + 1 producer warpgroup
  static constexpr int RS = (TX + 4) * N3;                 // doubles per staged row (x0-2 .. x0+TX+1)
  static constexpr int RO = TX * N3;                       // doubles per output / edge-plane row
  static constexpr int kU = ((TY + 2) * RS + 15) / 16 * 16;  // doubles per plane stage, 128-byte multiple
  static constexpr int kO = TY * RO;
  static constexpr uint32_t kBytesU = 8u * (TY + 2) * RS, kBytesE = 8u * kO, kBytesB = 8u * kO;
  static constexpr size_t smem_bytes() { return sizeof(double) * (2 * (size_t)kU + kO + 6 * N * N + 2) + 8 * (4 + kConsumers / 32) + 4 * 9 * (kConsumers / 32) + 32 + 128; }
  static_assert(TX % 2 == 0 && (kO * 8) % 128 == 0, "TMA destinations must stay 128-byte aligned");
};

// The work of a launch is a list of RUNS (one column of TX x TY elements, planes z in [za, zb)), built on the host
// (launch_march.cu: march_schedule) and handed to the CTAs as contiguous slices run_begin[b] .. run_begin[b+1]: a run is
// processed as the steps z = za-1 .. zb (the first and the last step only touch the neighbouring plane's own elements).
// The list balances the CTAs with a cost model (planes, run overheads, boundary columns) and, on several ranks, puts the
// single-plane runs of the rank-interface planes FIRST (flush = 1: their rows are on their way to the neighbours' mailboxes,
// and their sequence number published, while the rest of the box is still being computed).
struct MarchCursor {
  const MarchRun* runs; int i, iend, col, za, zb, z, flush; bool valid;
  __device__ __forceinline__ void start_run() {
    valid = i < iend;
    if (valid) { const int4 r = __ldg(reinterpret_cast<const int4*>(runs + i)); col = r.x; za = r.y; zb = r.z; flush = r.w; z = za - 1; }
  }
  __device__ __forceinline__ void init(const MarchRun* r, int b, int e) { runs = r; i = b; iend = e; start_run(); }
  __device__ __forceinline__ void advance() { if (++z > zb) { ++i; start_run(); } }
  __device__ __forceinline__ bool has_prev() const { return z > za; }          // plane z-1 is owned by this run: it completes now
  __device__ __forceinline__ bool edge() const { return z == za - 1 || z == zb; }
};

// acc += M v along axis AX where M has the checkerboard pattern M[i][j] = 0 for i + j odd (pure-diffusion SIPG axes of
// the Legendre basis: even and odd polynomials decouple) -- the zero products are not issued
template <int N, int AX, bool CHK>
__device__ __forceinline__ void apply_axis_chk(const double* __restrict__ M, const double (&v)[N * N * N], double (&acc)[N * N * N]) {
  constexpr int st = AX == 0 ? N * N : AX == 1 ? N : 1;
#pragma unroll
  for (int t = 0; t < N * N * N; ++t) {
    const int i = (t / st) % N, base = t - i * st;
    double s = acc[t];
#pragma unroll
    for (int j = 0; j < N; ++j) if (!CHK || ((i + j) & 1) == 0) s = fma(M[i * N + j], v[base + j * st], s);
    acc[t] = s;
  }
}

// One line (N values along axis AX, line index l in [0, N*N)) of acc += M v: the unit of the interleaved schedule below.
template <int N, int AX> __device__ __forceinline__ constexpr int line_base(int l) { return AX == 0 ? l : AX == 1 ? (l / N) * N * N + (l % N) : l * N; }
template <int N, int AX, bool CHK>
__device__ __forceinline__ void unit_reg(const double* __restrict__ M, const double (&v)[N * N * N], double (&acc)[N * N * N], const int l) {
  constexpr int st = AX == 0 ? N * N : AX == 1 ? N : 1;
  const int base = line_base<N, AX>(l);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double a = acc[base + i * st];
#pragma unroll
    for (int j = 0; j < N; ++j) if (!CHK || ((i + j) & 1) == 0) a = fma(M[i * N + j], v[base + j * st], a);
    acc[base + i * st] = a;
  }
}
template <int N, int AX, bool HIER>
__device__ __forceinline__ void unit_smem(const double* __restrict__ M, const double* __restrict__ src, double (&acc)[N * N * N], const int l) {
  constexpr PermTable<N, HIER> P{};
  constexpr int st = AX == 0 ? N * N : AX == 1 ? N : 1;
  const int base = line_base<N, AX>(l);
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

template <int N, bool HIER, int TX, int TY, bool HAS_B, bool CHK>
__global__ void __launch_bounds__(KronMarchCfg<N, TX, TY>::kThreads, 1)
dg_kronecker_march_kernel(const __grid_constant__ KronTabDev<N> K, const __grid_constant__ BoxDev box,
                          const __grid_constant__ KronMarchMaps M, const __grid_constant__ MarchCommDev C,
                          const MarchRun* __restrict__ runs, const int* __restrict__ run_begin, const int tiles_x) {
  using Cfg = KronMarchCfg<N, TX, TY>;
  constexpr int N3 = Cfg::N3, RS = Cfg::RS, RO = Cfg::RO, NN = N * N;
  constexpr int kWarps = Cfg::kConsumers / 32, kRowsPerWarp = 32 / TX, kSlab = kRowsPerWarp * RO;   // doubles per warp slab
  static_assert(32 % TX == 0 && (kSlab * 8) % 128 == 0, "a warp owns whole tile rows; its slab is a TMA box");
  constexpr PermTable<N, HIER> P{};
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  double* sbase = reinterpret_cast<double*>(smem_dyn + ((128u - (ptx::smem_addr(smem_dyn) & 127u)) & 127u));
  double* const O = sbase + 2 * (size_t)Cfg::kU;                    // kWarps output slabs
  double* const Dsm = O + Cfg::kO;                                  // boundary corrections [axis][lo|hi][N*N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Dsm + 6 * NN + (6 * NN) % 2);
  const uint32_t ufull = ptx::smem_addr(bars), ufree = ptx::smem_addr(bars + 2), wbar0 = ptx::smem_addr(bars + 4);
  unsigned int* const s_cnt = reinterpret_cast<unsigned int*>(bars + 4 + kWarps);   // [warp][9] row segments sent per direction, not yet accounted
  const int tid = threadIdx.x;
  if (tid < 9 * kWarps) s_cnt[tid] = 0u;

  const int nz = box.own_hi[2] - box.own_lo[2];
  const int t0 = __ldg(run_begin + blockIdx.x), t1 = __ldg(run_begin + blockIdx.x + 1);       // this CTA's slice of the run list (static:
                                                                                               // read before the dependency wait)
  auto plane_exists = [&](const MarchCursor& c) { const int lz = box.own_lo[2] + c.z; return lz >= 0 && lz < box.n[2]; };
  auto issue_u = [&](const MarchCursor& c, int s) {
    const int x0 = box.own_lo[0] + (c.col % tiles_x) * TX, y0 = box.own_lo[1] + (c.col / tiles_x) * TY, lz = box.own_lo[2] + c.z;
    const uint32_t dst = ptx::smem_addr(sbase + (size_t)s * Cfg::kU), bar = ufull + 8 * s;
    if (c.edge()) { ptx::mbar_expect_tx(bar, Cfg::kBytesE); ptx::tma_load_4d(dst, &M.u_edge, 0, x0 / 2, y0, lz, bar); }
    else          { ptx::mbar_expect_tx(bar, Cfg::kBytesU); ptx::tma_load_4d(dst, &M.u_plane, 0, (x0 - 2) / 2, y0 - 1, lz, bar); }
  };
  MarchCursor ld;                                                    // producer thread: next plane to load
  if (tid == Cfg::kConsumers) {
    // the producer thread sets the barriers up and gets the first two planes moving before anything else happens
    ptx::mbar_init(ufull, 1); ptx::mbar_init(ufull + 8, 1);
    ptx::mbar_init(ufree, Cfg::kConsumers); ptx::mbar_init(ufree + 8, Cfg::kConsumers);
    for (int w = 0; w < kWarps; ++w) ptx::mbar_init(wbar0 + 8 * w, 1);
    ptx::fence_barrier_init(); ptx::fence_proxy_async();
    // programmatic dependent launch: this CTA may have started while the previous kernel of the stream was still draining;
    // nothing of u / b / w is touched before that kernel has completed (no-op for an ordinary launch)
    ld.init(runs, t0, t1);                                            // (the run list is static: fetched before the wait)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = 0; i < 2; ++i) { while (ld.valid && !plane_exists(ld)) ld.advance(); if (ld.valid) { issue_u(ld, i); ld.advance(); } }
  }
  if (tid < 3 * NN) { Dsm[(tid / NN) * 2 * NN + tid % NN] = K.Dlo[tid / NN][tid % NN]; Dsm[(tid / NN) * 2 * NN + NN + tid % NN] = K.Dhi[tid / NN][tid % NN]; }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;");               // the next kernel may take over SMs as soon as CTAs of this one retire
  if (tid != Cfg::kConsumers) asm volatile("griddepcontrol.wait;" ::: "memory");

  if (tid >= Cfg::kConsumers) {
    // ============================== producer: one elected thread streams the u planes ==============================
    // register reallocation between warpgroups (as in warp-specialised GEMMs): the producer group gives its registers
    // to the consumers, which need 3 x n^3 doubles each
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (tid != Cfg::kConsumers) return;
    MarchCursor cur; cur.init(runs, t0, t1);
    auto next_load = [&]() { while (ld.valid && !plane_exists(ld)) ld.advance(); };
    int ku = 0;
    for (; cur.valid; cur.advance()) {
      if (!plane_exists(cur)) continue;
      next_load(); if (!ld.valid) break;
      ptx::mbar_wait(ufree + 8 * (ku & 1), (ku >> 1) & 1);           // plane consumed: its stage takes the load after next
      issue_u(ld, ku & 1); ld.advance(); ++ku;
    }
    return;
  }

  // ============================== consumers: one thread per element of the plane ==============================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
  const int tx = tid % TX, ty = tid / TX, warp = tid / 32, lane = tid % 32;
  double* const slab = O + (size_t)warp * kSlab;                     // this warp's rows of the output tile
  double* const o = slab + (ty % kRowsPerWarp) * RO + tx * N3;
  const uint32_t wbar = wbar0 + 8 * warp, slab_a = ptx::smem_addr(slab);
  const int wrow = warp * kRowsPerWarp;                              // first tile row of this warp
  const int on1 = box.own_hi[1] - box.own_lo[1];
  const int gz0 = box.origin[2] + box.own_lo[2];
  double A[N3], B[N3];                                               // A: plane z-1 (waits for R_z u(z)),  B: plane z
  int ku = 0, eb = 0;
  MarchCursor cur; cur.init(runs, t0, t1);
  int gcx = 0, gcy = 0, cx = 0, cy = 0;                              // global element coordinates; TMA coordinates of the warp's rows
  bool rows_owned = false, xb = false, yb = false;
  if (lane == 0) { ptx::prefetch_tensormap(&M.w_tile); if (HAS_B) ptx::prefetch_tensormap(&M.b_tile); }
  // fused Copy exchange of w (several ranks, comm.cuh): sequence number of this exchange (device-resident, so that the
  // launch can sit inside a captured graph); read after griddepcontrol.wait, i.e. after the previous exchange has retired
  const unsigned long long seq = C.any ? *C.seq + 1 : 0ull;
  const bool stamp = C.ts != nullptr && blockIdx.x == 0 && tid == 0;
  if (stamp) { C.ts[9] = C.ts[7]; C.ts[10] = C.ts[0]; C.ts[8] = gtimer_ns(); }     // (previous launch: end of tail, end of main loop)
  if (C.ts != nullptr && tid == 0) { C.ts[16 + 4 * blockIdx.x + 2] = C.ts[16 + 4 * blockIdx.x + 3]; C.ts[16 + 4 * blockIdx.x] = gtimer_ns(); }   // per CTA: start, loop end, previous tail end, tail end
  const int on0 = box.own_hi[0] - box.own_lo[0];
  // lane 0: the rows of the completed plane zc (owned coordinates) that lie on a rank interface go straight from the
  // warp's slab into the neighbours' mailboxes (1-D bulk copies over NVLink; 16-byte aligned because on0 and TX are even)
  int zflush = 0, commits = 0;                                       // (lane 0) steps until the z-interface rows are accounted; store groups committed since
  auto send_rows = [&](const int zc, const int col) {
    const int xoff = (col % tiles_x) * TX;
    const uint32_t bytes = (uint32_t)min(TX, on0 - xoff) * N3 * 8u;
    const bool zlo = zc == 0, zhi = zc == nz - 1;
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int yc = cy + r;
      if (yc >= on1) break;
      const bool ylo = yc == 0, yhi = yc == on1 - 1;
      if (!(ylo | yhi | zlo | zhi)) continue;
      for (int dzc = 0; dzc < 3; ++dzc) {
        if ((dzc == 0 && !zlo) || (dzc == 2 && !zhi)) continue;
        for (int dyc = 0; dyc < 3; ++dyc) {
          if ((dyc == 0 && !ylo) || (dyc == 2 && !yhi) || (dyc == 1 && dzc == 1)) continue;
          const int d = dyc + 3 * dzc;
          if (!C.enabled[d]) continue;
          const long long zi = dzc == 1 ? zc : 0, yi = dyc == 1 ? yc : 0, ny = dyc == 1 ? on1 : 1;
          double* const mbox = C.remote[d][seq & 1];
          ptx::bulk_s2g(mbox + ((zi * ny + yi) * on0 + xoff) * N3, slab_a + (uint32_t)(r * RO) * 8u, bytes);
          s_cnt[9 * warp + d] += 1u;
          if (dzc != 1) zflush = 2;                                  // a z-interface plane is on its way: account it two steps from now
        }
      }
    }
  };

  // lane 0: accounts the row segments this warp has sent so far.  All its bulk stores are complete (wait_group), a gpu-scope
  // fence orders them before the counter update; whoever completes a message publishes its sequence number with ONE
  // system-scope release, which is cumulative over everything ordered before it.
  auto flush_sends = [&](const bool all) {
    if (all) ptx::bulk_wait_all(); else asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");   // (all but the newest store group)
    bool any = false;
    for (int d = 0; d < 9; ++d) any = any || s_cnt[9 * warp + d] != 0u;
    if (!any) return;
    __threadfence();
    for (int d = 0; d < 9; ++d) {
      const unsigned int c = s_cnt[9 * warp + d];
      if (!c) continue;
      s_cnt[9 * warp + d] = 0u;
      if (atomicAdd(&C.counters[d], c) + c == C.expected[d]) {
        C.counters[d] = 0; __threadfence(); st_release_sys(C.remote_ready[d] + (seq & 1), seq);
      }
    }
  };
  for (; cur.valid; cur.advance()) {
    // rows of a z-interface plane were sent two steps ago: by now their bulk stores have (almost always) completed, so the
    // accounting -- and, for whoever completes the message, the publication of its sequence number -- costs no round trip.
    // The peers thus hold the interface planes, which the schedule puts first, long before they finish their own box.
    if (lane == 0 && zflush && --zflush == 0) { flush_sends(commits == 0); }
    if (cur.z == cur.za - 1) {                                       // a new run starts
      const int bx = cur.col % tiles_x, by = cur.col / tiles_x;
      gcx = box.origin[0] + box.own_lo[0] + bx * TX + tx; gcy = box.origin[1] + box.own_lo[1] + by * TY + ty;
      xb = gcx == 0 || gcx == box.gn[0] - 1; yb = gcy == 0 || gcy == box.gn[1] - 1;
      cx = bx * (TX / 2); cy = by * TY + wrow; rows_owned = cy < on1;
#pragma unroll
      for (int t = 0; t < N3; ++t) { A[t] = 0.0; B[t] = 0.0; }
    }
    const bool exists = plane_exists(cur), edge = cur.edge(), hp = cur.has_prev();
    const double* st = sbase + (size_t)(ku & 1) * Cfg::kU;
    const double* own = edge ? st + ty * RO + tx * N3 : st + (ty + 1) * RS + (tx + 2) * N3;
    double v[N3];
    if (exists) {
      ptx::mbar_wait(ufull + 8 * (ku & 1), (ku >> 1) & 1);
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = own[P.p[t]];
    } else {
#pragma unroll
      for (int t = 0; t < N3; ++t) v[t] = 0.0;
    }
    // The step is a hand-interleaved stream: every register-only line unit (R, S: FP64 pipe only) is paired with a
    // neighbour line unit (3 shared loads + 9 FMA), so that shared-memory traffic is spread evenly over the FMA work
    // instead of being bunched in "neighbour phases" where the load pipe, not the FP64 pipe, would set the pace.
    if (hp && !edge) {
#pragma unroll
      for (int l = 0; l < NN; ++l) { if (exists) unit_reg<N, 2, false>(K.R[2], v, A, l); unit_smem<N, 0, HIER>(K.L[0], own - N3, B, l); }
    } else {
      if (hp && exists) apply_axis<N, 2>(K.R[2], v, A);              // plane z-1 is complete after R_z u(z)
      if (!edge) apply_axis_smem<N, 0, HIER>(K.L[0], own - N3, B);
    }
    if (hp) {
      if (HAS_B) { ptx::mbar_wait(wbar, eb & 1); ++eb; }              // b rows have landed (and the previous store has read the slab)
      else { if (lane == 0) ptx::bulk_wait_read(); __syncwarp(); }   // the previous store has read the slab
#pragma unroll
      for (int t = 0; t < N3; ++t) o[P.p[t]] = HAS_B ? A[t] - o[P.p[t]] : A[t];
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0 && rows_owned) { ptx::tma_store_4d(&M.w_tile, 0, cx, cy, cur.z - 1, slab_a); if (C.any) { commits += 1; send_rows(cur.z - 1, cur.col); if (zflush == 2) commits = 0; } ptx::bulk_commit(); }
    }
    if (!edge) {
#pragma unroll
      for (int l = 0; l < NN; ++l) { unit_reg<N, 0, false>(K.S[0], v, B, l); unit_smem<N, 0, HIER>(K.R[0], own + N3, B, l); }
#pragma unroll
      for (int l = 0; l < NN; ++l) { unit_reg<N, 1, CHK>(K.S[1], v, B, l); unit_smem<N, 1, HIER>(K.L[1], own - RS, B, l); }
    }
    if (HAS_B && lane == 0 && cur.z >= cur.za && cur.z < cur.zb) {    // plane z completes in the next step: the slab takes its b rows
      ptx::bulk_wait_read();                                         // (the store issued above has read the slab by now)
      if (rows_owned) { ptx::mbar_expect_tx(wbar, 8u * kSlab); ptx::tma_load_4d(slab_a, &M.b_tile, 0, cx, cy, cur.z, wbar); }
      else ptx::mbar_arrive(wbar);
    }
    if (!edge) {
#pragma unroll
      for (int l = 0; l < NN; ++l) { unit_reg<N, 2, CHK>(K.S[2], v, B, l); unit_smem<N, 1, HIER>(K.R[1], own + RS, B, l); }
      // domain-boundary corrections of the self matrices (tables in shared memory: the choice is per thread)
      const int gcz = gz0 + cur.z;
      const bool zb = gcz == 0 || gcz == box.gn[2] - 1;
      if (xb | yb | zb) {
        const int gc[3] = {gcx, gcy, gcz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const bool lo = gc[a] == 0, hi = gc[a] == box.gn[a] - 1;
          if (lo || hi) {
            double m[NN];
#pragma unroll
            for (int i = 0; i < NN; ++i) m[i] = (lo ? Dsm[a * 2 * NN + i] : 0.0) + (hi ? Dsm[a * 2 * NN + NN + i] : 0.0);
            if (a == 0) apply_axis<N, 0>(m, v, B); else if (a == 1) apply_axis<N, 1>(m, v, B); else apply_axis<N, 2>(m, v, B);
          }
        }
      }
    }
    if (exists) { ptx::mbar_arrive(ufree + 8 * (ku & 1)); ++ku; }
    // rotate: plane z now waits for R_{z+1}; plane z+1 starts with L_z u(z)
#pragma unroll
    for (int t = 0; t < N3; ++t) { A[t] = B[t]; B[t] = 0.0; }
    if (exists && cur.z + 1 < cur.zb) apply_axis<N, 2>(K.L[2], v, B);
  }
  if (stamp) C.ts[0] = gtimer_ns();
  if (C.ts != nullptr && tid == 0) C.ts[16 + 4 * blockIdx.x + 1] = gtimer_ns();
  if (lane == 0) ptx::bulk_wait_all();
  if (stamp) C.ts[1] = gtimer_ns();
  if (!C.any) { if (stamp) C.ts[7] = gtimer_ns(); if (C.ts != nullptr && tid == 0) C.ts[16 + 4 * blockIdx.x + 3] = gtimer_ns(); return; }

  // ============================== fused Copy exchange: publish, receive, retire ==============================
  // (1) the rest of this warp's row segments (no CTA-wide barrier: warps account on their own)
  if (lane == 0) flush_sends(true);
  if (stamp) C.ts[2] = gtimer_ns();
  if (stamp) C.ts[3] = gtimer_ns();
  if (stamp) C.ts[4] = gtimer_ns();
  // (2) receive: the row segments of the incoming messages are dealt to all consumer warps of the (fully resident)
  // grid; a warp waits for the flag of a message the first time it needs it.  Peers publish from their own compute
  // kernels, which never wait for anything of this exchange: no cycle.
  {
    constexpr int kSplit = 1;                                          // an item is one row segment (TX elements), all its loads in flight at once
    const int on2 = nz, n0 = box.n[0], n1 = box.n[1];
    // Items (quarter row segments) are handed out dynamically, per message: CTAs that finish early (the schedule lets those
    // with rank-interface rows finish first) do the receiving, the CTAs on the critical path find nothing left.  First pass:
    // only messages whose sequence number has already arrived (the z-interface planes, published early in the peers' kernels);
    // second pass: wait for the rest.  Peers publish from their own compute kernels, which never wait for anything of this
    // exchange: no cycle.
    for (int pass = 0; pass < 2; ++pass) {
      for (int d = 0; d < 9; ++d) {
        if (!C.enabled[d]) continue;
        const int dyc = d % 3, dzc = d / 3, ny = dyc == 1 ? on1 : 1;
        const long long nitems = (long long)kSplit * tiles_x * ny * (dzc == 1 ? on2 : 1);
        int ok = 0;
        if (lane == 0) {
          if (*reinterpret_cast<volatile unsigned int*>(&C.counters[10 + d]) >= nitems) ok = 2;          // nothing left of this message
          else if (pass == 0) ok = ld_acquire_sys(C.local_ready[d] + (seq & 1)) >= seq ? 1 : 0;
          else ok = wait_flag_ge(C.local_ready[d] + (seq & 1), seq, C.err, kCommTimeoutFused) ? 1 : 2;
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok != 1) continue;
        if (stamp && C.ts[5] < C.ts[8]) C.ts[5] = gtimer_ns();
        while (true) {
          long long it = 0;
          if (lane == 0) it = (long long)atomicAdd(&C.counters[10 + d], 1u);
          it = __shfl_sync(0xffffffffu, it, 0);
          if (it >= nitems) break;
          const int part = (int)(it % kSplit); const long long q = it / kSplit;
          const int seg = (int)(q % tiles_x); const long long row = q / tiles_x;
          const int yi = (int)(row % ny), zi = (int)(row / ny);
          const int gy = dyc == 1 ? box.own_lo[1] + yi : (dyc == 0 ? box.own_lo[1] - 1 : box.own_hi[1]);
          const int gz = dzc == 1 ? box.own_lo[2] + zi : (dzc == 0 ? box.own_lo[2] - 1 : box.own_hi[2]);
          const int xoff = seg * TX, cnt2 = min(TX, on0 - xoff) * N3 / 2;                 // double2 items of this segment
          const double2* src = reinterpret_cast<const double2*>(C.local[d][seq & 1] + (row * on0 + xoff) * N3);
          double2* dst = reinterpret_cast<double2*>(C.w + (((long long)gz * n1 + gy) * n0 + box.own_lo[0] + xoff) * N3);
          (void)part;
          constexpr int kPer = (TX * N3 / 2 + 31) / 32;                                    // double2 per lane
          double2 v[kPer];
#pragma unroll
          for (int k = 0; k < kPer; ++k) { const int i = lane + 32 * k; v[k] = i < cnt2 ? __ldcg(src + i) : make_double2(0.0, 0.0); }
#pragma unroll
          for (int k = 0; k < kPer; ++k) { const int i = lane + 32 * k; if (i < cnt2) dst[i] = v[k]; }
        }
      }
    }
  }
  // (3) the CTA that finishes last advances the sequence number (the next exchange on this stream reads it)
  if (stamp) C.ts[6] = gtimer_ns();
  asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kConsumers) : "memory");
  if (stamp) C.ts[7] = gtimer_ns();
  if (C.ts != nullptr && tid == 0) C.ts[16 + 4 * blockIdx.x + 3] = gtimer_ns();
  if (tid == 0) { __threadfence(); if (atomicAdd(&C.counters[9], 1u) == gridDim.x - 1) { C.counters[9] = 0; for (int d = 0; d < 9; ++d) C.counters[10 + d] = 0; *C.seq = seq; } }
}

}  // namespace b200fem
