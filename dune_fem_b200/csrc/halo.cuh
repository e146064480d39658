// halo.cuh -- halo exchange of dof vectors between the boxes of a Cartesian domain decomposition.
//
// Replaces DiscreteFunction::communicate() (dune/fem/function/common/discretefunction.hh:825-835 ->
// space/common/communicationmanager.hh:130-150): per shared entity the owner's dof block is sent and
//   DG spaces       Copy  into the ghost copy          (space/discontinuousgalerkin/space.hh:80)
//   Lagrange spaces Add   on dofs of shared entities   (space/lagrange/space.hh:92)
// (operations: space/common/commoperations.hh:126-205).  Like the reference's cached communicator
// (space/common/cachedcommmanager.hh:943-975) the send/receive index lists are built once; each exchange is then
//   pack kernel -> ncclGroupStart; ncclSend/ncclRecv to the two neighbours of an axis; ncclGroupEnd -> unpack kernel
// axis by axis (x, y, z), later axes forwarding what earlier ones received, so edge and corner copies become
// consistent with 6 messages instead of 26.  Primary/auxiliary dofs (space/common/auxiliarydofs.hh:215-275: the lowest
// rank owning a copy is primary) are flagged for the dot products.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "lagrange_quadrature.cuh"

namespace b200fem {

struct NcclUniqueId { char internal[128]; };

// NCCL is bound at run time (dlopen) so that the library has no link-time dependency on a particular NCCL build;
// inside a torch process this resolves to the NCCL torch already loaded.
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  bool ok() const { return handle != nullptr; }
  bool load() {
    if (handle) return true;
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) return false;
    auto sym = [&](const char* n) { return dlsym(handle, n); };
    GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    CommInitRank = (int (*)(void**, int, NcclUniqueId, int))sym("ncclCommInitRank");
    CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
    GroupStart = (int (*)())sym("ncclGroupStart");
    GroupEnd = (int (*)())sym("ncclGroupEnd");
    if (!(GetUniqueId && CommInitRank && CommDestroy && Send && Recv && AllReduce && GroupStart && GroupEnd)) { handle = nullptr; return false; }
    return true;
  }
};

// one direction of one axis: `count` blocks of `block` doubles
struct HaloSide {
  int peer = -1; long long count = 0;
  long long* d_send_idx = nullptr; long long* d_recv_idx = nullptr;   // block start offsets in the dof vector
  double* d_send = nullptr; double* d_recv = nullptr;
};
struct HaloPlan { int block = 1; HaloSide side[3][2]; bool built = false; };

__global__ void halo_pack_kernel(const double* __restrict__ v, const long long* __restrict__ idx, long long count, int block, double* __restrict__ buf) {
  const long long total = count * block;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    buf[i] = v[idx[i / block] + (i % block)];
}
__global__ void halo_unpack_kernel(double* __restrict__ v, const long long* __restrict__ idx, long long count, int block, const double* __restrict__ buf, int add) {
  const long long total = count * block;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = idx[i / block] + (i % block);
    v[g] = add ? v[g] + buf[i] : buf[i];
  }
}

inline void halo_plan_free(HaloPlan& p) {
  for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) { HaloSide& h = p.side[d][s];
    for (void* q : {(void*)h.d_send_idx, (void*)h.d_recv_idx, (void*)h.d_send, (void*)h.d_recv}) if (q) cudaFree(q);
    h = HaloSide(); }
  p.built = false;
}

// Build the index lists.  DG: blocks are elements (nb doubles); the layer of owned elements next to a rank interface
// is sent, the ghost layer received; ranges in already-exchanged axes span the full local box (ghosts included).
// Lagrange: blocks are single dofs on the interface lattice planes g_d = 0 / k n_d; both sides send and add.
inline int halo_plan_build(HaloPlan& p, const int proc[3], const int pc[3], const BoxDev& box, bool lagrange, int order, int nb,
                           const LagrangeLayoutDev& layout_dev, long long size, uint8_t** d_aux_out) {
  p.block = lagrange ? 1 : nb;
  std::vector<uint8_t> aux((size_t)size, 0);
  LagrangeLayoutDev L = layout_dev;
  std::vector<long long> host_map;
  if (lagrange && L.lattice_map) {       // need the lattice map on the host
    host_map.resize((size_t)(L.lattice[0] * L.lattice[1] * L.lattice[2]));
    if (cudaMemcpy(host_map.data(), L.lattice_map, host_map.size() * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    L.lattice_map = host_map.data();
  }
  auto rank_of = [&](int c0, int c1, int c2) { return c0 + proc[0] * (c1 + proc[1] * c2); };
  for (int d = 0; d < box.dim; ++d) for (int s = 0; s < 2; ++s) {
    HaloSide& h = p.side[d][s];
    int nc[3] = {pc[0], pc[1], pc[2]}; nc[d] += s ? 1 : -1;
    if (nc[d] < 0 || nc[d] >= proc[d]) continue;
    h.peer = rank_of(nc[0], nc[1], nc[2]);
    std::vector<long long> send, recv;
    if (!lagrange) {
      int lo[3], hi[3];
      for (int a = 0; a < 3; ++a) { if (a < d) { lo[a] = 0; hi[a] = box.n[a]; } else { lo[a] = box.own_lo[a]; hi[a] = box.own_hi[a]; } }
      const int send_layer = s ? box.own_hi[d] - 1 : box.own_lo[d], recv_layer = s ? box.own_hi[d] : box.own_lo[d] - 1;
      int c[3];
      for (c[2] = lo[2]; c[2] < hi[2]; ++c[2]) for (c[1] = lo[1]; c[1] < hi[1]; ++c[1]) for (c[0] = lo[0]; c[0] < hi[0]; ++c[0]) {
        if (c[d] != lo[d]) continue;       // iterate the plane once
        int cs[3] = {c[0], c[1], c[2]}, cr[3] = {c[0], c[1], c[2]}; cs[d] = send_layer; cr[d] = recv_layer;
        send.push_back((cs[0] + (long long)box.n[0] * (cs[1] + (long long)box.n[1] * cs[2])) * nb);
        recv.push_back((cr[0] + (long long)box.n[0] * (cr[1] + (long long)box.n[1] * cr[2])) * nb);
      }
    } else {
      long long g[3];
      const long long plane = s ? L.lattice[d] - 1 : 0;
      for (g[2] = 0; g[2] < L.lattice[2]; ++g[2]) for (g[1] = 0; g[1] < L.lattice[1]; ++g[1]) for (g[0] = 0; g[0] < L.lattice[0]; ++g[0]) {
        if (g[d] != plane) continue;
        const long long dof = lagrange_dof(L, g[0], g[1], g[2]);
        send.push_back(dof); recv.push_back(dof);
        if (s == 0) aux[(size_t)dof] = 1;       // a lower rank shares this dof: auxiliary here
      }
    }
    h.count = (long long)send.size();
    if (h.count == 0) { h.peer = -1; continue; }
    const size_t ib = sizeof(long long) * send.size(), db = sizeof(double) * send.size() * p.block;
    if (cudaMalloc(&h.d_send_idx, ib) != cudaSuccess || cudaMalloc(&h.d_recv_idx, ib) != cudaSuccess || cudaMalloc(&h.d_send, db) != cudaSuccess || cudaMalloc(&h.d_recv, db) != cudaSuccess) return -1;
    cudaMemcpy(h.d_send_idx, send.data(), ib, cudaMemcpyHostToDevice); cudaMemcpy(h.d_recv_idx, recv.data(), ib, cudaMemcpyHostToDevice);
  }
  if (!lagrange) {   // ghost elements are auxiliary
    for (int c2 = 0; c2 < box.n[2]; ++c2) for (int c1 = 0; c1 < box.n[1]; ++c1) for (int c0 = 0; c0 < box.n[0]; ++c0) {
      const bool owned = c0 >= box.own_lo[0] && c0 < box.own_hi[0] && c1 >= box.own_lo[1] && c1 < box.own_hi[1] && c2 >= box.own_lo[2] && c2 < box.own_hi[2];
      if (!owned) { const long long e = c0 + (long long)box.n[0] * (c1 + (long long)box.n[1] * c2); std::memset(&aux[(size_t)(e * nb)], 1, (size_t)nb); }
    }
  }
  if (cudaMalloc(d_aux_out, (size_t)size) != cudaSuccess) return -1;
  cudaMemcpy(*d_aux_out, aux.data(), (size_t)size, cudaMemcpyHostToDevice);
  p.built = true; return 0;
}

// ---- single-phase exchange for DG spaces: every existing neighbour among the 26 (faces, edges, corners) gets its own
// message inside ONE ncclGroup, so ghost copies are fully consistent after one round trip (Copy has no ordering issue).
struct HaloNeighbour { int peer = -1; long long count = 0; long long *d_send_idx = nullptr, *d_recv_idx = nullptr; double *d_send = nullptr, *d_recv = nullptr; };
struct HaloPlanDG { int block = 1; std::vector<HaloNeighbour> nb; bool built = false; };

inline void halo_plan_dg_free(HaloPlanDG& p) {
  for (auto& h : p.nb) for (void* q : {(void*)h.d_send_idx, (void*)h.d_recv_idx, (void*)h.d_send, (void*)h.d_recv}) if (q) cudaFree(q);
  p.nb.clear(); p.built = false;
}
inline int halo_plan_dg_build(HaloPlanDG& p, const int proc[3], const int pc[3], const BoxDev& box, int nb) {
  p.block = nb;
  for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
    if (!dx && !dy && !dz) continue;
    const int dir[3] = {dx, dy, dz}; int nc[3]; bool ok = true;
    for (int a = 0; a < 3; ++a) { nc[a] = pc[a] + dir[a]; if (nc[a] < 0 || nc[a] >= proc[a]) ok = false; }
    if (!ok) continue;
    HaloNeighbour h; h.peer = nc[0] + proc[0] * (nc[1] + proc[1] * nc[2]);
    int slo[3], shi[3], rlo[3], rhi[3];
    for (int a = 0; a < 3; ++a) {
      if (dir[a] == 0) { slo[a] = rlo[a] = box.own_lo[a]; shi[a] = rhi[a] = box.own_hi[a]; }
      else if (dir[a] < 0) { slo[a] = box.own_lo[a]; shi[a] = slo[a] + 1; rlo[a] = box.own_lo[a] - 1; rhi[a] = box.own_lo[a]; }
      else { shi[a] = box.own_hi[a]; slo[a] = shi[a] - 1; rlo[a] = box.own_hi[a]; rhi[a] = rlo[a] + 1; }
    }
    std::vector<long long> send, recv;
    for (int z = slo[2]; z < shi[2]; ++z) for (int y = slo[1]; y < shi[1]; ++y) for (int x = slo[0]; x < shi[0]; ++x) send.push_back((x + (long long)box.n[0] * (y + (long long)box.n[1] * z)) * nb);
    for (int z = rlo[2]; z < rhi[2]; ++z) for (int y = rlo[1]; y < rhi[1]; ++y) for (int x = rlo[0]; x < rhi[0]; ++x) recv.push_back((x + (long long)box.n[0] * (y + (long long)box.n[1] * z)) * nb);
    h.count = (long long)send.size();
    if (h.count == 0 || send.size() != recv.size()) continue;
    const size_t ib = sizeof(long long) * send.size(), db = sizeof(double) * send.size() * nb;
    if (cudaMalloc(&h.d_send_idx, ib) != cudaSuccess || cudaMalloc(&h.d_recv_idx, ib) != cudaSuccess || cudaMalloc(&h.d_send, db) != cudaSuccess || cudaMalloc(&h.d_recv, db) != cudaSuccess) return -1;
    cudaMemcpy(h.d_send_idx, send.data(), ib, cudaMemcpyHostToDevice); cudaMemcpy(h.d_recv_idx, recv.data(), ib, cudaMemcpyHostToDevice);
    p.nb.push_back(h);
  }
  p.built = true; return 0;
}
inline int halo_exchange_dg(HaloPlanDG& p, NcclApi& nccl, void* comm, double* v, cudaStream_t st) {
  if (!p.built) return -1;
  if (p.nb.empty()) return 0;
  for (auto& h : p.nb) { const long long total = h.count * p.block; const int grid = (int)std::min<long long>(592, (total + 255) / 256);
    halo_pack_kernel<<<grid, 256, 0, st>>>(v, h.d_send_idx, h.count, p.block, h.d_send); }
  if (nccl.GroupStart() != 0) return -1;
  for (auto& h : p.nb) {
    if (nccl.Send(h.d_send, (size_t)(h.count * p.block), 8, h.peer, comm, st) != 0) return -1;
    if (nccl.Recv(h.d_recv, (size_t)(h.count * p.block), 8, h.peer, comm, st) != 0) return -1;
  }
  if (nccl.GroupEnd() != 0) return -1;
  for (auto& h : p.nb) { const long long total = h.count * p.block; const int grid = (int)std::min<long long>(592, (total + 255) / 256);
    halo_unpack_kernel<<<grid, 256, 0, st>>>(v, h.d_recv_idx, h.count, p.block, h.d_recv, 0); }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---- peer-memory exchange (all ranks on one NVSwitch box, one process per GPU) ----------------------------------------
// Every rank owns a mailbox area (plain cudaMalloc, exported with cudaIpcGetMemHandle, handles all-gathered once through
// NCCL).  Per neighbour the mailbox holds two data buffers, two `ready` sequence flags (written by the neighbour when
// its message has landed) and one `ack` flag (written by the neighbour when it has consumed my message).  An exchange is
// ONE launch and no library call (send part, then receive part):
//   send kernel:   wait ack >= seq-2  ->  gather the owned layer and store it DIRECTLY into the neighbour's mailbox over
//                  NVLink  ->  __threadfence_system  ->  last block publishes ready[seq&1] = seq in the neighbour's memory
//   recv kernel:   spin on my ready[seq&1] >= seq  ->  scatter the mailbox into the ghost layer  ->  last block
//                  publishes ack = seq in the neighbour's memory
struct P2PNeighbourDev {
  long long total;                                   // doubles per message
  const unsigned int* send_flat; const unsigned int* recv_flat;   // per double: offset in the dof vector
  int block_begin, nblocks;                          // this neighbour's slice of the grid
  int fused;                                         // 1: the compute kernel sends this message and publishes `ready` itself
  double* remote_data[2]; unsigned long long* remote_ready; unsigned long long* remote_ack;     // in the peer's mailbox
  double* local_data[2]; unsigned long long* local_ready; unsigned long long* local_ack;        // in my mailbox
  unsigned int* counters;                                                                        // [0] send, [1] recv block counters (local)
};
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
struct HaloPlanP2P {
  bool built = false; int block = 1; int nnb = 0; int grid = 0; unsigned long long seq = 0;
  P2PNeighbourDev* d_nb = nullptr; void* mailbox = nullptr; unsigned int* d_counters = nullptr; int* d_error = nullptr;
  std::vector<void*> opened;          // peer mappings to close
  std::vector<void*> owned;           // flat index arrays
  std::vector<P2PNeighbourDev> host_nb; std::vector<int> dir_code;   // host copy (+ direction (dx+1)+3(dy+1)+9(dz+1)) for the fused path
  P2PNeighbourDev* d_nb_fused = nullptr; unsigned int* d_cta_counter = nullptr;
  unsigned long long* d_ts = nullptr;  // optional timestamps (B200FEM_DEBUG_EVENTS)
};
constexpr int kP2PThreads = 256;
constexpr long long kP2PSpinLimit = 4000000000ll;   // ~2 s of SM clocks: a lost peer must not hang the box

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ int p2p_find(const P2PNeighbourDev* nbs, int nnb) {
  int k = 0; while (k + 1 < nnb && (int)blockIdx.x >= nbs[k + 1].block_begin) ++k; return k;
}

// One fused kernel per exchange: every block first sends its slice (kP2PItems independent doubles per thread: index,
// value and remote store chains overlap), then waits for the neighbour's flag and scatters the same slice of the incoming
// message.  The send part never waits on this exchange's remote state, so blocks need not be co-resident.
constexpr int kP2PItems = 8;
__global__ void __launch_bounds__(kP2PThreads) p2p_exchange_kernel(double* __restrict__ v, const P2PNeighbourDev* __restrict__ nbs, int nnb, unsigned long long seq, int* error, unsigned long long* ts) {
  const P2PNeighbourDev nb = nbs[p2p_find(nbs, nnb)];
  if (ts && blockIdx.x == 0 && threadIdx.x == 0) ts[0] = gtimer();
  if (threadIdx.x == 0 && seq > 2 && !nb.fused) {          // the buffer was last used by message seq-2: has it been consumed?
    const long long t0 = clock64();
    while (ld_acquire_sys(nb.local_ack) < seq - 2) if (clock64() - t0 > kP2PSpinLimit) { *error = 1; break; }
  }
  __syncthreads();
  const long long base = (long long)(blockIdx.x - nb.block_begin) * (kP2PThreads * kP2PItems) + threadIdx.x;
  if (!nb.fused) {
    double* dst = nb.remote_data[seq & 1];
    unsigned int idx[kP2PItems]; double val[kP2PItems];
#pragma unroll
    for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; idx[k] = i < nb.total ? nb.send_flat[i] : 0u; }
#pragma unroll
    for (int k = 0; k < kP2PItems; ++k) val[k] = v[idx[k]];
#pragma unroll
    for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; if (i < nb.total) dst[i] = val[k]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!nb.fused) {
      __threadfence_system();                              // cumulative over the block's stores (ordered by the barrier)
      if (atomicAdd(&nb.counters[0], 1u) == (unsigned)nb.nblocks - 1) { nb.counters[0] = 0; __threadfence_system(); st_release_sys(&nb.remote_ready[seq & 1], seq); if (ts) ts[1] = gtimer(); }
    }
    if (ts && blockIdx.x == 0) ts[2] = gtimer();
    const long long t0 = clock64();
    while (ld_acquire_sys(&nb.local_ready[seq & 1]) < seq) if (clock64() - t0 > kP2PSpinLimit) { *error = 2; break; }
    if (ts && blockIdx.x == 0) ts[3] = gtimer();
  }
  __syncthreads();
  {
    const double* src = nb.local_data[seq & 1];
    unsigned int idx[kP2PItems]; double val[kP2PItems];
#pragma unroll
    for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; idx[k] = i < nb.total ? nb.recv_flat[i] : 0u; val[k] = i < nb.total ? src[i] : 0.0; }
#pragma unroll
    for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; if (i < nb.total) v[idx[k]] = val[k]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&nb.counters[1], 1u) == (unsigned)nb.nblocks - 1) { nb.counters[1] = 0; __threadfence_system(); st_release_sys(nb.remote_ack, seq); if (ts) ts[4] = gtimer(); }
  }
}

inline void halo_plan_p2p_free(HaloPlanP2P& p) {
  if (p.d_ts) {
    unsigned long long h[8]; cudaDeviceSynchronize(); cudaMemcpy(h, p.d_ts, 64, cudaMemcpyDeviceToHost); cudaFree(p.d_ts);
    std::fprintf(stderr, "[b200fem p2p ns, last exchange] flag published +%lld | recv kernel start +%lld | flag seen +%lld | ack sent +%lld (from send kernel start)\n",
                 (long long)(h[1] - h[0]), (long long)(h[2] - h[0]), (long long)(h[3] - h[0]), (long long)(h[4] - h[0]));
  }
  for (void* q : p.opened) cudaIpcCloseMemHandle(q);
  for (void* q : p.owned) cudaFree(q);
  for (void* q : {(void*)p.d_nb, (void*)p.d_nb_fused, (void*)p.d_cta_counter, p.mailbox, (void*)p.d_counters, (void*)p.d_error}) if (q) cudaFree(q);
  p = HaloPlanP2P();
}

// mailbox layout of a rank: for each of its neighbours in the fixed 26-direction order: data[2][count*block] doubles,
// then ready[2] + ack (3 x u64, padded to 32 B).  `counts` = messages sizes (elements) in that order.
struct MailboxLayout { std::vector<int> dir; std::vector<long long> count; std::vector<size_t> offset; size_t bytes = 0; };
inline MailboxLayout mailbox_layout(const int proc[3], const int pc[3], const int own_n[3], int nb) {
  MailboxLayout L; size_t off = 0;
  for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
    if (!dx && !dy && !dz) continue;
    const int dir[3] = {dx, dy, dz}; bool ok = true; long long cnt = 1;
    for (int a = 0; a < 3; ++a) { const int c = pc[a] + dir[a]; if (c < 0 || c >= proc[a]) ok = false; cnt *= dir[a] == 0 ? own_n[a] : 1; }
    if (!ok || cnt == 0) continue;
    L.dir.push_back((dx + 1) + 3 * ((dy + 1) + 3 * (dz + 1))); L.count.push_back(cnt); L.offset.push_back(off);
    off += 2 * (size_t)cnt * nb * sizeof(double) + 32; off = (off + 255) / 256 * 256;
  }
  L.bytes = off; return L;
}

// own extents of rank coordinates c under the block distribution of b200fem_partition_box
inline void block_extents(const int gn[3], const int proc[3], const int c[3], int out[3]) {
  for (int a = 0; a < 3; ++a) { const int q = gn[a] / proc[a], r = gn[a] % proc[a]; out[a] = q + (c[a] < r ? 1 : 0); }
}

inline int halo_plan_p2p_build(HaloPlanP2P& p, HaloPlanDG& dg, NcclApi& nccl, void* comm, int rank, int world, const int proc[3], const int pc[3],
                               const int gn[3], int nb, cudaStream_t st) {
  if (!dg.built || !nccl.AllGather) return -1;
  int own_n[3]; block_extents(gn, proc, pc, own_n);
  MailboxLayout mine = mailbox_layout(proc, pc, own_n, nb);
  if (mine.dir.size() != dg.nb.size()) return -1;
  p.block = nb; p.nnb = (int)dg.nb.size();
  if (p.nnb == 0) { p.built = true; return 0; }
  if (cudaMalloc(&p.mailbox, mine.bytes) != cudaSuccess) return -1;
  cudaMemset(p.mailbox, 0, mine.bytes);
  cudaMalloc(&p.d_counters, sizeof(unsigned int) * 2 * p.nnb); cudaMemset(p.d_counters, 0, sizeof(unsigned int) * 2 * p.nnb);
  cudaMalloc(&p.d_error, sizeof(int)); cudaMemset(p.d_error, 0, sizeof(int));
  if (std::getenv("B200FEM_DEBUG_EVENTS")) { cudaMalloc(&p.d_ts, 64); cudaMemset(p.d_ts, 0, 64); }
  // all-gather the IPC handles
  cudaIpcMemHandle_t h; if (cudaIpcGetMemHandle(&h, p.mailbox) != cudaSuccess) return -1;
  char *d_h = nullptr, *d_all = nullptr; cudaMalloc(&d_h, sizeof(h)); cudaMalloc(&d_all, sizeof(h) * world);
  cudaMemcpy(d_h, &h, sizeof(h), cudaMemcpyHostToDevice);
  if (nccl.AllGather(d_h, d_all, sizeof(h), /*ncclChar*/ 0, comm, st) != 0) return -1;
  std::vector<cudaIpcMemHandle_t> all(world);
  cudaStreamSynchronize(st); cudaMemcpy(all.data(), d_all, sizeof(h) * world, cudaMemcpyDeviceToHost); cudaFree(d_h); cudaFree(d_all);
  std::vector<P2PNeighbourDev> host(p.nnb);
  std::vector<std::pair<int, void*>> open_cache;
  for (int i = 0; i < p.nnb; ++i) {
    HaloNeighbour& hn = dg.nb[i];
    const int code = mine.dir[i], dx = code % 3 - 1, dy = (code / 3) % 3 - 1, dz = code / 9 - 1;
    const int pcn[3] = {pc[0] + dx, pc[1] + dy, pc[2] + dz};
    int own_peer[3]; block_extents(gn, proc, pcn, own_peer);
    MailboxLayout theirs = mailbox_layout(proc, pcn, own_peer, nb);
    const int back = (-dx + 1) + 3 * ((-dy + 1) + 3 * (-dz + 1));
    int j = -1; for (size_t k = 0; k < theirs.dir.size(); ++k) if (theirs.dir[k] == back) j = (int)k;
    if (j < 0 || theirs.count[j] != hn.count || mine.count[i] != hn.count) return -1;
    void* peer_base = nullptr;
    for (auto& oc : open_cache) if (oc.first == hn.peer) peer_base = oc.second;
    if (!peer_base) {
      if (cudaIpcOpenMemHandle(&peer_base, all[hn.peer], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return -1; }
      open_cache.push_back({hn.peer, peer_base}); p.opened.push_back(peer_base);
    }
    P2PNeighbourDev& d = host[i];
    {   // flat per-double gather/scatter offsets
      std::vector<long long> si((size_t)hn.count), ri((size_t)hn.count);
      cudaMemcpy(si.data(), hn.d_send_idx, sizeof(long long) * hn.count, cudaMemcpyDeviceToHost);
      cudaMemcpy(ri.data(), hn.d_recv_idx, sizeof(long long) * hn.count, cudaMemcpyDeviceToHost);
      std::vector<unsigned int> sf((size_t)hn.count * nb), rf((size_t)hn.count * nb);
      for (long long e = 0; e < hn.count; ++e) for (int j = 0; j < nb; ++j) {
        if (si[e] + j > 0xffffffffll || ri[e] + j > 0xffffffffll) return -1;
        sf[(size_t)e * nb + j] = (unsigned int)(si[e] + j); rf[(size_t)e * nb + j] = (unsigned int)(ri[e] + j);
      }
      unsigned int *dsf = nullptr, *drf = nullptr;
      if (cudaMalloc(&dsf, sf.size() * 4) != cudaSuccess || cudaMalloc(&drf, rf.size() * 4) != cudaSuccess) return -1;
      cudaMemcpy(dsf, sf.data(), sf.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(drf, rf.data(), rf.size() * 4, cudaMemcpyHostToDevice);
      p.owned.push_back(dsf); p.owned.push_back(drf);
      d.total = hn.count * nb; d.send_flat = dsf; d.recv_flat = drf;
      d.fused = 0; d.block_begin = p.grid; d.nblocks = (int)((d.total + kP2PThreads * kP2PItems - 1) / (kP2PThreads * kP2PItems)); p.grid += d.nblocks;
    }
    char* rb = (char*)peer_base + theirs.offset[j]; char* lb = (char*)p.mailbox + mine.offset[i];
    const size_t one = (size_t)hn.count * nb * sizeof(double);
    d.remote_data[0] = (double*)rb; d.remote_data[1] = (double*)(rb + one);
    d.remote_ready = (unsigned long long*)(rb + 2 * one); d.remote_ack = d.remote_ready + 2;
    d.local_data[0] = (double*)lb; d.local_data[1] = (double*)(lb + one);
    d.local_ready = (unsigned long long*)(lb + 2 * one); d.local_ack = d.local_ready + 2;
    d.counters = p.d_counters + 2 * i;
  }
  cudaMalloc(&p.d_nb, sizeof(P2PNeighbourDev) * p.nnb);
  cudaMemcpy(p.d_nb, host.data(), sizeof(P2PNeighbourDev) * p.nnb, cudaMemcpyHostToDevice);
  // second table for exchanges whose y/z face messages are sent by the compute kernel itself
  p.host_nb = host; p.dir_code = mine.dir;
  std::vector<P2PNeighbourDev> fused = host;
  for (int i = 0; i < p.nnb; ++i) { const int c = mine.dir[i], dx = c % 3 - 1, dy = (c / 3) % 3 - 1, dz = c / 9 - 1; fused[i].fused = dx == 0 ? 1 : 0; }
  cudaMalloc(&p.d_nb_fused, sizeof(P2PNeighbourDev) * p.nnb);
  cudaMemcpy(p.d_nb_fused, fused.data(), sizeof(P2PNeighbourDev) * p.nnb, cudaMemcpyHostToDevice);
  cudaMalloc(&p.d_cta_counter, 9 * sizeof(unsigned int)); cudaMemset(p.d_cta_counter, 0, 9 * sizeof(unsigned int));
  (void)rank;
  p.built = cudaGetLastError() == cudaSuccess; return p.built ? 0 : -1;
}
// receive side of an exchange whose y/z face messages were sent by the compute kernel (sequence number `seq`); the
// remaining neighbours (edges, corners, x faces) are sent here
inline int halo_exchange_p2p_fused_tail(HaloPlanP2P& p, double* v, unsigned long long seq, cudaStream_t st) {
  if (!p.built || p.nnb == 0) return p.built ? 0 : -1;
  p2p_exchange_kernel<<<p.grid, kP2PThreads, 0, st>>>(v, p.d_nb_fused, p.nnb, seq, p.d_error, p.d_ts);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
inline int halo_exchange_p2p(HaloPlanP2P& p, double* v, cudaStream_t st) {
  if (!p.built) return -1;
  if (p.nnb == 0) return 0;
  const unsigned long long seq = ++p.seq;
  // (programmatic dependent launch of this kernel was tried: 58.8 instead of 53.2 us per step at 2 GPUs -- early-scheduled
  // blocks compete with the persistent compute kernel for SMs; plain stream order it is)
  p2p_exchange_kernel<<<p.grid, kP2PThreads, 0, st>>>(v, p.d_nb, p.nnb, seq, p.d_error, p.d_ts);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

inline int halo_exchange(HaloPlan& p, NcclApi& nccl, void* comm, double* v, bool add, cudaStream_t st) {
  if (!p.built) return -1;
  for (int d = 0; d < 3; ++d) {
    bool any = false;
    for (int s = 0; s < 2; ++s) { HaloSide& h = p.side[d][s]; if (h.peer < 0) continue; any = true;
      const long long total = h.count * p.block; const int grid = (int)std::min<long long>(1184, (total + 255) / 256);
      halo_pack_kernel<<<grid, 256, 0, st>>>(v, h.d_send_idx, h.count, p.block, h.d_send); }
    if (!any) continue;
    if (nccl.GroupStart() != 0) return -1;
    for (int s = 0; s < 2; ++s) { HaloSide& h = p.side[d][s]; if (h.peer < 0) continue;
      if (nccl.Send(h.d_send, (size_t)(h.count * p.block), 8, h.peer, comm, st) != 0) return -1;
      if (nccl.Recv(h.d_recv, (size_t)(h.count * p.block), 8, h.peer, comm, st) != 0) return -1; }
    if (nccl.GroupEnd() != 0) return -1;
    for (int s = 0; s < 2; ++s) { HaloSide& h = p.side[d][s]; if (h.peer < 0) continue;
      const long long total = h.count * p.block; const int grid = (int)std::min<long long>(1184, (total + 255) / 256);
      halo_unpack_kernel<<<grid, 256, 0, st>>>(v, h.d_recv_idx, h.count, p.block, h.d_recv, add ? 1 : 0); }
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace b200fem
