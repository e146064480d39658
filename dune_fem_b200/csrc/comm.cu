// comm.cu -- halo plans, peer-memory mailboxes, exchange kernels, NCCL fallback (see comm.cuh for the protocol).
#include "comm.cuh"

#include <algorithm>
#include <array>
#include <map>

namespace b200fem {

// ======================================================================================================================
// symmetric peer regions
int peer_region_create(NcclApi& nccl, void* comm, int rank, int world, size_t bytes, cudaStream_t st, PeerRegion& out) {
  out = PeerRegion(); out.rank = rank; out.world = world; out.bytes = bytes;
  if (world > kMaxPeers || !nccl.ok() || !comm) return -1;
  int ok = 1;
  if (cudaMalloc(&out.local, bytes) != cudaSuccess) { cudaGetLastError(); out.local = nullptr; ok = 0; }
  cudaIpcMemHandle_t h; std::memset(&h, 0, sizeof(h));
  if (ok && cudaMemset(out.local, 0, bytes) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&h, out.local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  // all-gather (ok flag, handle)
  struct Rec { int ok; int pad; cudaIpcMemHandle_t h; };
  Rec mine; mine.ok = ok; mine.pad = 0; mine.h = h;
  Rec* d_mine = nullptr; Rec* d_all = nullptr; std::vector<Rec> all((size_t)world);
  if (cudaMalloc(&d_mine, sizeof(Rec)) != cudaSuccess || cudaMalloc(&d_all, sizeof(Rec) * world) != cudaSuccess) return -1;
  cudaMemcpy(d_mine, &mine, sizeof(Rec), cudaMemcpyHostToDevice);
  int rc = nccl.AllGather(d_mine, d_all, sizeof(Rec), /*ncclChar*/ 0, comm, st);
  if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) rc = -1;
  if (rc == 0) cudaMemcpy(all.data(), d_all, sizeof(Rec) * world, cudaMemcpyDeviceToHost);
  cudaFree(d_mine); cudaFree(d_all);
  if (rc != 0) { if (out.local) cudaFree(out.local); out.local = nullptr; return -1; }
  for (int r = 0; r < world; ++r) if (!all[(size_t)r].ok) ok = 0;
  out.mapped.assign((size_t)world, nullptr);
  if (ok) {
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) { out.mapped[(size_t)r] = out.local; continue; }
      if (cudaIpcOpenMemHandle(&out.mapped[(size_t)r], all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); out.mapped[(size_t)r] = nullptr; ok = 0; }
    }
  }
  // agree on the outcome (one failing rank switches everybody to the NCCL transport); doubles as the barrier that
  // guarantees every region is zero-filled before anybody writes into it
  int* d_ok = nullptr; if (cudaMalloc(&d_ok, sizeof(int)) != cudaSuccess) return -1;
  cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice);
  rc = nccl.AllReduce(d_ok, d_ok, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, comm, st);
  if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) rc = -1;
  if (rc == 0) cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_ok);
  if (rc != 0 || !ok) { peer_region_free(out); return -1; }
  out.ok = true; return 0;
}
void peer_region_free(PeerRegion& r) {
  for (int i = 0; i < (int)r.mapped.size(); ++i) if (i != r.rank && r.mapped[(size_t)i]) cudaIpcCloseMemHandle(r.mapped[(size_t)i]);
  if (r.local) cudaFree(r.local);
  r = PeerRegion();
}

int peer_scalars_create(NcclApi& nccl, void* comm, int rank, int world, cudaStream_t st, int* d_err, PeerScalars& out) {
  out = PeerScalars();
  const size_t vals_bytes = sizeof(double) * 2 * (size_t)world * kArMax, flag_bytes = sizeof(unsigned long long) * 2 * (size_t)world;
  if (peer_region_create(nccl, comm, rank, world, vals_bytes + flag_bytes, st, out.region) != 0) return -1;
  if (cudaMalloc(&out.d_seq, sizeof(unsigned long long)) != cudaSuccess) { peer_region_free(out.region); return -1; }
  cudaMemset(out.d_seq, 0, sizeof(unsigned long long));
  out.dev.rank = rank; out.dev.world = world; out.dev.seq = out.d_seq; out.dev.err = d_err;
  for (int r = 0; r < world; ++r) {
    out.dev.vals[r] = reinterpret_cast<double*>(out.region.mapped[(size_t)r]);
    out.dev.flags[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(out.region.mapped[(size_t)r]) + vals_bytes);
  }
  out.ok = true; return 0;
}
void peer_scalars_free(PeerScalars& s) { if (s.d_seq) cudaFree(s.d_seq); peer_region_free(s.region); s = PeerScalars(); }

// ======================================================================================================================
// NCCL transport
static __global__ void halo_pack_kernel(const double* __restrict__ v, const long long* __restrict__ idx, long long count, int block, double* __restrict__ buf) {
  const long long total = count * block;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    buf[i] = v[idx[i / block] + (i % block)];
}
static __global__ void halo_unpack_kernel(double* __restrict__ v, const long long* __restrict__ idx, long long count, int block, const double* __restrict__ buf, int add) {
  const long long total = count * block;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = idx[i / block] + (i % block);
    v[g] = add ? v[g] + buf[i] : buf[i];
  }
}

void halo_plan_free(HaloPlan& p) {
  for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) { HaloSide& h = p.side[d][s];
    for (void* q : {(void*)h.d_send_idx, (void*)h.d_recv_idx, (void*)h.d_send, (void*)h.d_recv}) if (q) cudaFree(q);
    h = HaloSide(); }
  p.built = false;
}

// DG: blocks are elements (nb doubles); the layer of owned elements next to a rank interface is sent, the ghost layer
// received; ranges in already-exchanged axes span the full local box (ghosts included).  Lagrange: blocks are single dofs
// on the interface lattice planes g_d = 0 / k n_d; both sides send and add.
int halo_plan_build(HaloPlan& p, const int proc[3], const int pc[3], const BoxDev& box, bool lagrange, int order, int nb,
                    const LagrangeLayoutDev& layout_dev, long long size, uint8_t** d_aux_out) {
  (void)order;
  // DG: nb = doubles per element (n_b * dimRange); Lagrange: nb = dimRange -- a shared node carries a block of dimRange components
  p.block = nb;
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
        const long long dof = lagrange_dof(L, g[0], g[1], g[2]) * nb;
        send.push_back(dof); recv.push_back(dof);
        if (s == 0) for (int c = 0; c < nb; ++c) aux[(size_t)(dof + c)] = 1;       // a lower rank shares this node: auxiliary here
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
int halo_exchange(HaloPlan& p, NcclApi& nccl, void* comm, double* v, bool add, cudaStream_t st) {
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

void halo_plan_dg_free(HaloPlanDG& p) {
  for (auto& h : p.nb) for (void* q : {(void*)h.d_send_idx, (void*)h.d_recv_idx, (void*)h.d_send, (void*)h.d_recv}) if (q) cudaFree(q);
  p.nb.clear(); p.built = false;
}
int halo_plan_dg_build(HaloPlanDG& p, const int proc[3], const int pc[3], const BoxDev& box, int nb) {
  p.block = nb;
  for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
    if (!dx && !dy && !dz) continue;
    const int dir[3] = {dx, dy, dz}; int nc[3]; bool ok = true;
    for (int a = 0; a < 3; ++a) { nc[a] = pc[a] + dir[a]; if (nc[a] < 0 || nc[a] >= proc[a]) ok = false; }
    if (!ok) continue;
    HaloNeighbour h; h.peer = nc[0] + proc[0] * (nc[1] + proc[1] * nc[2]); h.dir = (dx + 1) + 3 * ((dy + 1) + 3 * (dz + 1));
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
int halo_exchange_dg(HaloPlanDG& p, NcclApi& nccl, void* comm, double* v, cudaStream_t st) {
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

// ======================================================================================================================
// peer-memory Copy exchange (DG)
constexpr int kP2PThreads = 256, kP2PItems = 8;

static __device__ __forceinline__ int p2p_find(const P2PNeighbourDev* nbs, int nnb) {
  int k = 0; while (k + 1 < nnb && (int)blockIdx.x >= nbs[k + 1].block_begin) ++k; return k;
}
// send: every block gathers its slice of the owned interface layer and stores it into the neighbour's mailbox over NVLink
// (kP2PItems independent doubles per thread: index, value and remote-store chains overlap); the block that completes a
// message publishes its sequence number.  Never waits.
static __global__ void __launch_bounds__(kP2PThreads) p2p_send_kernel(const double* __restrict__ v, const P2PNeighbourDev* __restrict__ nbs, int nnb,
                                                                      const unsigned long long* __restrict__ seq_p) {
  const P2PNeighbourDev nb = nbs[p2p_find(nbs, nnb)];
  const unsigned long long seq = *seq_p + 1;
  const long long base = (long long)(blockIdx.x - nb.block_begin) * (kP2PThreads * kP2PItems) + threadIdx.x;
  double* dst = nb.remote_data[seq & 1];
  unsigned int idx[kP2PItems]; double val[kP2PItems];
#pragma unroll
  for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; idx[k] = i < nb.total ? nb.send_flat[i] : 0u; }
#pragma unroll
  for (int k = 0; k < kP2PItems; ++k) val[k] = v[idx[k]];
#pragma unroll
  for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; if (i < nb.total) dst[i] = val[k]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                                // cumulative over the block's stores (ordered by the barrier)
    if (atomicAdd(nb.counter, 1u) == (unsigned)nb.nblocks - 1) { *nb.counter = 0; __threadfence_system(); st_release_sys(nb.remote_ready + (seq & 1), seq); }
  }
}
// receive: wait for the neighbour's flag, scatter the mailbox into the ghost layer; the last block of the grid advances
// the sequence number.  Blocks only wait for REMOTE send kernels, which never wait themselves.
static __global__ void __launch_bounds__(kP2PThreads) p2p_recv_kernel(double* __restrict__ v, const P2PNeighbourDev* __restrict__ nbs, int nnb,
                                                                      unsigned long long* seq_p, unsigned int* done, int* err) {
  const P2PNeighbourDev nb = nbs[p2p_find(nbs, nnb)];
  const unsigned long long seq = *seq_p + 1;
  __shared__ int ok_s;
  if (threadIdx.x == 0) ok_s = wait_flag_ge(nb.local_ready + (seq & 1), seq, err, kCommTimeoutHalo) ? 1 : 0;
  __syncthreads();
  if (ok_s) {
    const long long base = (long long)(blockIdx.x - nb.block_begin) * (kP2PThreads * kP2PItems) + threadIdx.x;
    const double* src = nb.local_data[seq & 1];
    unsigned int idx[kP2PItems]; double val[kP2PItems];
#pragma unroll
    for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; idx[k] = i < nb.total ? nb.recv_flat[i] : 0u; val[k] = i < nb.total ? __ldcg(src + i) : 0.0; }
#pragma unroll
    for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; if (i < nb.total) v[idx[k]] = val[k]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); if (atomicAdd(done, 1u) == gridDim.x - 1) { *done = 0; *seq_p = seq; } }
}

// mailbox layout of a rank: for each of its neighbours in the fixed 26-direction order: data[2][count*block] doubles,
// then ready[2] (padded to 32 B).  `count` = message sizes (elements) in that order.
struct MailboxLayout { std::vector<int> dir; std::vector<long long> count; std::vector<size_t> offset; size_t bytes = 0; };
static MailboxLayout mailbox_layout(const int proc[3], const int pc[3], const int own_n[3], int nb) {
  MailboxLayout L; size_t off = 0;
  for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
    if (!dx && !dy && !dz) continue;
    const int dir[3] = {dx, dy, dz}; bool ok = true; long long cnt = 1;
    for (int a = 0; a < 3; ++a) { const int c = pc[a] + dir[a]; if (c < 0 || c >= proc[a]) ok = false; cnt *= dir[a] == 0 ? own_n[a] : 1; }
    if (!ok || cnt == 0) continue;
    L.dir.push_back((dx + 1) + 3 * ((dy + 1) + 3 * (dz + 1))); L.count.push_back(cnt); L.offset.push_back(off);
    off += 2 * (size_t)cnt * nb * sizeof(double) + 32; off = (off + 255) / 256 * 256;
  }
  L.bytes = std::max<size_t>(off, 256); return L;
}
// own extents of rank coordinates c under the block distribution of b200fem_partition_box
static void block_extents(const int gn[3], const int proc[3], const int c[3], int out[3]) {
  for (int a = 0; a < 3; ++a) { const int q = gn[a] / proc[a], r = gn[a] % proc[a]; out[a] = q + (c[a] < r ? 1 : 0); }
}

void halo_plan_p2p_free(HaloPlanP2P& p) {
  if (p.d_march_ts) {      // diagnostics (B200FEM_MARCH_TS): tail phases of the last fused exchange, ns after the end of CTA 0's main loop
    std::vector<unsigned long long> h(16 + 4 * 160); cudaDeviceSynchronize(); cudaMemcpy(h.data(), p.d_march_ts, h.size() * 8, cudaMemcpyDeviceToHost); cudaFree(p.d_march_ts);
    std::fprintf(stderr, "[b200fem march tail, ns] CTA 0: stores complete +%lld | sends accounted +%lld | first flag seen +%lld | receive done +%lld | start -> loop end %lld | previous tail end -> this start %lld\n",
                 (long long)(h[1] - h[0]), (long long)(h[2] - h[0]), (long long)(h[5] - h[0]), (long long)(h[6] - h[0]), (long long)(h[0] - h[8]), (long long)(h[8] - h[9]));
    unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0, p1 = 0, t1 = 0; int n = 0;
    for (int b = 0; b < 160; ++b) { const unsigned long long st = h[16 + 4 * b], en = h[16 + 4 * b + 1], pe = h[16 + 4 * b + 2], te = h[16 + 4 * b + 3]; if (!st) continue; ++n; s0 = std::min(s0, st); s1 = std::max(s1, st); e0 = std::min(e0, en); e1 = std::max(e1, en); p1 = std::max(p1, pe); t1 = std::max(t1, te); }
    if (std::getenv("B200FEM_MARCH_TS_DUMP")) for (int b = 0; b < 160; ++b) { const unsigned long long st = h[16 + 4 * b], en = h[16 + 4 * b + 1], te = h[16 + 4 * b + 3]; if (st) std::fprintf(stderr, "rank %d cta %d start %lld loop %lld tail %lld\n", p.region.rank, b, (long long)(st - s0), (long long)(en - st), (long long)(te - en)); }
    std::fprintf(stderr, "[b200fem march tail, ns] %d CTAs: starts spread %lld | loop ends spread %lld | first start -> last loop end %lld | last loop end -> last tail end %lld | previous launch's last TAIL end -> first start %lld\n",
                 n, (long long)(s1 - s0), (long long)(e1 - e0), (long long)(e1 - s0), (long long)(t1 - e1), (long long)(s0 - p1));
  }
  for (void* q : p.owned) cudaFree(q);
  for (void* q : {(void*)p.d_nb, (void*)p.d_counters, (void*)p.d_seq, (void*)p.d_done, (void*)p.d_march_counters}) if (q) cudaFree(q);
  peer_region_free(p.region);
  p = HaloPlanP2P();
}

int halo_plan_p2p_build(HaloPlanP2P& p, HaloPlanDG& dg, NcclApi& nccl, void* comm, int rank, int world, const int proc[3], const int pc[3],
                        const int gn[3], const BoxDev& box, int nb, int* d_err, cudaStream_t st) {
  if (!dg.built) return -1;
  int own_n[3]; block_extents(gn, proc, pc, own_n);
  MailboxLayout mine = mailbox_layout(proc, pc, own_n, nb);
  // (collective: every rank must get here, also those whose layout turns out inconsistent below)
  if (peer_region_create(nccl, comm, rank, world, mine.bytes, st, p.region) != 0) return -1;
  if (mine.dir.size() != dg.nb.size()) return -1;
  p.block = nb; p.nnb = (int)dg.nb.size();
  if (cudaMalloc(&p.d_seq, sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&p.d_done, sizeof(unsigned int)) != cudaSuccess) return -1;
  cudaMemset(p.d_seq, 0, sizeof(unsigned long long)); cudaMemset(p.d_done, 0, sizeof(unsigned int));
  if (p.nnb == 0) { p.built = true; return 0; }
  cudaMalloc(&p.d_counters, sizeof(unsigned int) * p.nnb); cudaMemset(p.d_counters, 0, sizeof(unsigned int) * p.nnb);
  std::vector<P2PNeighbourDev> host((size_t)p.nnb);
  for (int i = 0; i < p.nnb; ++i) {
    HaloNeighbour& hn = dg.nb[(size_t)i];
    const int code = mine.dir[(size_t)i], dx = code % 3 - 1, dy = (code / 3) % 3 - 1, dz = code / 9 - 1;
    if (code != hn.dir) return -1;
    const int pcn[3] = {pc[0] + dx, pc[1] + dy, pc[2] + dz};
    int own_peer[3]; block_extents(gn, proc, pcn, own_peer);
    MailboxLayout theirs = mailbox_layout(proc, pcn, own_peer, nb);
    const int back = (-dx + 1) + 3 * ((-dy + 1) + 3 * (-dz + 1));
    int j = -1; for (size_t k = 0; k < theirs.dir.size(); ++k) if (theirs.dir[k] == back) j = (int)k;
    if (j < 0 || theirs.count[(size_t)j] != hn.count || mine.count[(size_t)i] != hn.count) return -1;
    P2PNeighbourDev& d = host[(size_t)i];
    {   // flat per-double gather/scatter offsets
      std::vector<long long> si((size_t)hn.count), ri((size_t)hn.count);
      cudaMemcpy(si.data(), hn.d_send_idx, sizeof(long long) * hn.count, cudaMemcpyDeviceToHost);
      cudaMemcpy(ri.data(), hn.d_recv_idx, sizeof(long long) * hn.count, cudaMemcpyDeviceToHost);
      std::vector<unsigned int> sf((size_t)hn.count * nb), rf((size_t)hn.count * nb);
      for (long long e = 0; e < hn.count; ++e) for (int q = 0; q < nb; ++q) {
        if (si[(size_t)e] + q > 0xffffffffll || ri[(size_t)e] + q > 0xffffffffll) return -1;
        sf[(size_t)e * nb + q] = (unsigned int)(si[(size_t)e] + q); rf[(size_t)e * nb + q] = (unsigned int)(ri[(size_t)e] + q);
      }
      unsigned int *dsf = nullptr, *drf = nullptr;
      if (cudaMalloc(&dsf, sf.size() * 4) != cudaSuccess || cudaMalloc(&drf, rf.size() * 4) != cudaSuccess) return -1;
      cudaMemcpy(dsf, sf.data(), sf.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(drf, rf.data(), rf.size() * 4, cudaMemcpyHostToDevice);
      p.owned.push_back(dsf); p.owned.push_back(drf);
      d.total = hn.count * nb; d.send_flat = dsf; d.recv_flat = drf;
      d.block_begin = p.grid; d.nblocks = (int)((d.total + kP2PThreads * kP2PItems - 1) / (kP2PThreads * kP2PItems)); p.grid += d.nblocks;
    }
    char* rb = (char*)p.region.mapped[(size_t)hn.peer] + theirs.offset[(size_t)j]; char* lb = (char*)p.region.local + mine.offset[(size_t)i];
    const size_t one = (size_t)hn.count * nb * sizeof(double);
    d.remote_data[0] = (double*)rb; d.remote_data[1] = (double*)(rb + one); d.remote_ready = (unsigned long long*)(rb + 2 * one);
    d.local_data[0] = (const double*)lb; d.local_data[1] = (const double*)(lb + one); d.local_ready = (const unsigned long long*)(lb + 2 * one);
    d.counter = p.d_counters + i;
  }
  cudaMalloc(&p.d_nb, sizeof(P2PNeighbourDev) * p.nnb);
  cudaMemcpy(p.d_nb, host.data(), sizeof(P2PNeighbourDev) * p.nnb, cudaMemcpyHostToDevice);
  p.host_nb = host; p.dir_code = mine.dir;
  // description of the exchange for the marching kernel (all neighbours must lie in the y-z plane of the process grid)
  {
    MarchCommDev& m = p.march; std::memset(&m, 0, sizeof(m));
    bool plane_only = proc[0] == 1;
    cudaMalloc(&p.d_march_counters, sizeof(unsigned int) * 20); cudaMemset(p.d_march_counters, 0, sizeof(unsigned int) * 20);
    const int on[3] = {box.own_hi[0] - box.own_lo[0], box.own_hi[1] - box.own_lo[1], box.own_hi[2] - box.own_lo[2]};
    const unsigned tiles_x = (unsigned)((on[0] + 15) / 16);
    for (int i = 0; i < p.nnb; ++i) {
      const int c = mine.dir[(size_t)i], dx = c % 3 - 1, dy = (c / 3) % 3 - 1, dz = c / 9 - 1;
      if (dx != 0) { plane_only = false; continue; }
      const int d9 = (dy + 1) + 3 * (dz + 1);
      m.enabled[d9] = 1; m.any = 1;
      m.remote[d9][0] = host[(size_t)i].remote_data[0]; m.remote[d9][1] = host[(size_t)i].remote_data[1]; m.remote_ready[d9] = host[(size_t)i].remote_ready;
      m.local[d9][0] = host[(size_t)i].local_data[0]; m.local[d9][1] = host[(size_t)i].local_data[1]; m.local_ready[d9] = host[(size_t)i].local_ready;
      m.expected[d9] = tiles_x * (unsigned)(dy == 0 ? on[1] : 1) * (unsigned)(dz == 0 ? on[2] : 1);
    }
    m.counters = p.d_march_counters; m.seq = p.d_seq; m.err = d_err; m.w = nullptr; m.ts = nullptr;
    if (std::getenv("B200FEM_MARCH_TS")) { cudaMalloc(&p.d_march_ts, (16 + 4 * 160) * sizeof(unsigned long long)); cudaMemset(p.d_march_ts, 0, (16 + 4 * 160) * sizeof(unsigned long long)); m.ts = p.d_march_ts; }
    p.march_ok = plane_only && m.any;
  }
  p.built = cudaGetLastError() == cudaSuccess; return p.built ? 0 : -1;
}
int halo_exchange_p2p(HaloPlanP2P& p, double* v, int* d_err, cudaStream_t st) {
  if (!p.built) return -1;
  if (p.nnb == 0) return 0;
  p2p_send_kernel<<<p.grid, kP2PThreads, 0, st>>>(v, p.d_nb, p.nnb, p.d_seq);
  p2p_recv_kernel<<<p.grid, kP2PThreads, 0, st>>>(v, p.d_nb, p.nnb, p.d_seq, p.d_done, d_err);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ======================================================================================================================
// peer-memory Add exchange (Lagrange)
static __device__ __forceinline__ int add_find(const AddNeighbourDev* nbs, int nnb) {
  int k = 0; while (k + 1 < nnb && (int)blockIdx.x >= nbs[k + 1].block_begin) ++k; return k;
}
static __global__ void __launch_bounds__(kP2PThreads) add_send_kernel(const double* __restrict__ v, const AddNeighbourDev* __restrict__ nbs, int nnb,
                                                                      const unsigned long long* __restrict__ seq_p) {
  const AddNeighbourDev nb = nbs[add_find(nbs, nnb)];
  const unsigned long long seq = *seq_p + 1;
  const long long base = (long long)(blockIdx.x - nb.block_begin) * (kP2PThreads * kP2PItems) + threadIdx.x;
  double* dst = nb.remote_data[seq & 1];
  unsigned int idx[kP2PItems]; double val[kP2PItems];
#pragma unroll
  for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; idx[k] = i < nb.total ? nb.send_idx[i] : 0u; }
#pragma unroll
  for (int k = 0; k < kP2PItems; ++k) val[k] = v[idx[k]];
#pragma unroll
  for (int k = 0; k < kP2PItems; ++k) { const long long i = base + (long long)k * kP2PThreads; if (i < nb.total) dst[i] = val[k]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(nb.counter, 1u) == (unsigned)nb.nblocks - 1) { *nb.counter = 0; __threadfence_system(); st_release_sys(nb.remote_ready + (seq & 1), seq); }
  }
}
// every shared dof becomes the sum over all ranks that hold a copy, added in rank order (own value at its position)
static __global__ void __launch_bounds__(kP2PThreads) add_recv_kernel(double* __restrict__ v, const AddNeighbourDev* __restrict__ nbs, int nnb,
                                                                      const double* __restrict__ mbox0, const double* __restrict__ mbox1,
                                                                      const unsigned int* __restrict__ dof, const int* __restrict__ ptr, const int* __restrict__ src,
                                                                      long long nshared, unsigned long long* seq_p, unsigned int* done, int* err) {
  const unsigned long long seq = *seq_p + 1;
  __shared__ int ok_s;
  if (threadIdx.x == 0) ok_s = 1;
  __syncthreads();
  if ((int)threadIdx.x < nnb) { if (!wait_flag_ge(nbs[threadIdx.x].local_ready + (seq & 1), seq, err, kCommTimeoutHalo)) ok_s = 0; }
  __syncthreads();
  if (ok_s) {
    const double* mbox = (seq & 1) ? mbox1 : mbox0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nshared; i += (long long)gridDim.x * blockDim.x) {
      const unsigned int g = dof[i]; double acc = 0.0;
      for (int k = ptr[i]; k < ptr[i + 1]; ++k) { const int s = src[k]; acc += s < 0 ? v[g] : __ldcg(mbox + s); }
      v[g] = acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); if (atomicAdd(done, 1u) == gridDim.x - 1) { *done = 0; *seq_p = seq; } }
}

void halo_plan_add_free(HaloPlanAddP2P& p) {
  for (void* q : p.owned) cudaFree(q);
  for (void* q : {(void*)p.d_nb, (void*)p.d_counters, (void*)p.d_seq, (void*)p.d_done, (void*)p.d_dof, (void*)p.d_ptr, (void*)p.d_src}) if (q) cudaFree(q);
  peer_region_free(p.region);
  p = HaloPlanAddP2P();
}

// shared lattice nodes with the neighbour in direction dir (local lattice coordinates, lexicographic z, y, x)
static void shared_nodes(const long long Lat[3], const int dir[3], int dim, std::vector<std::array<long long, 3>>& out) {
  long long lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    if (a >= dim || dir[a] == 0) { lo[a] = 0; hi[a] = Lat[a]; }
    else if (dir[a] < 0) { lo[a] = 0; hi[a] = 1; }
    else { lo[a] = Lat[a] - 1; hi[a] = Lat[a]; }
  }
  out.clear();
  for (long long z = lo[2]; z < hi[2]; ++z) for (long long y = lo[1]; y < hi[1]; ++y) for (long long x = lo[0]; x < hi[0]; ++x) out.push_back({x, y, z});
}

int halo_plan_add_build(HaloPlanAddP2P& p, NcclApi& nccl, void* comm, int rank, int world, const int proc[3], const int pc[3],
                        const LagrangeLayoutDev& layout_dev, const std::vector<long long>& host_lattice_map, int dim, int dim_range, int* d_err, cudaStream_t st) {
  (void)d_err;
  const int R = dim_range;            // vector-valued spaces: every (node, component) is a shared dof of its own, node * R + c
  LagrangeLayoutDev L = layout_dev; L.lattice_map = host_lattice_map.empty() ? nullptr : host_lattice_map.data();
  const long long Lat[3] = {L.lattice[0], L.lattice[1], L.lattice[2]};
  // my neighbours in the fixed 26-direction order and the size of the message exchanged with each (symmetric)
  struct Nb { int dir[3]; int peer; long long count; long long offset; };
  auto neighbours_of = [&](const int c[3], const long long lat[3]) {
    std::vector<Nb> v; long long off = 0;
    for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
      if (!dx && !dy && !dz) continue;
      const int dir[3] = {dx, dy, dz}; bool ok = true; long long cnt = 1;
      for (int a = 0; a < 3; ++a) {
        if (a >= dim) { if (dir[a] != 0) ok = false; continue; }
        const int cc = c[a] + dir[a]; if (cc < 0 || cc >= proc[a]) ok = false;
        cnt *= dir[a] == 0 ? lat[a] : 1;
      }
      if (!ok) continue;
      Nb nb; nb.dir[0] = dx; nb.dir[1] = dy; nb.dir[2] = dz; nb.peer = (c[0] + dx) + proc[0] * ((c[1] + dy) + proc[1] * (c[2] + dz)); nb.count = cnt * R; nb.offset = off;
      off += cnt * R; v.push_back(nb);
    }
    return std::make_pair(v, off);
  };
  auto mine = neighbours_of(pc, Lat);
  const long long total = mine.second;                       // doubles per parity in my mailbox
  const size_t data_bytes = (sizeof(double) * 2 * (size_t)std::max<long long>(total, 1) + 255) / 256 * 256, flag_bytes = sizeof(unsigned long long) * 2 * 32;
  // the lattice extents of a neighbour follow from ITS element extents; this needs the global element counts, which are
  // recovered from the lattice: all ranks along an axis share the block distribution of b200fem_partition_box.  The
  // message sizes are symmetric by construction (closed intersections), so only offsets inside the peer's mailbox are needed:
  // they are all-gathered instead of recomputed.
  if (peer_region_create(nccl, comm, rank, world, data_bytes + flag_bytes, st, p.region) != 0) return -1;
  // all-gather per-rank offset tables: offset of the message from direction code c (27 entries, -1 = none)
  std::vector<long long> my_tab(27, -1), all_tab((size_t)27 * world, -1);
  for (const Nb& nb : mine.first) my_tab[(size_t)((nb.dir[0] + 1) + 3 * ((nb.dir[1] + 1) + 3 * (nb.dir[2] + 1)))] = nb.offset;
  {
    long long *d_a = nullptr, *d_b = nullptr;
    if (cudaMalloc(&d_a, 27 * 8) != cudaSuccess || cudaMalloc(&d_b, (size_t)27 * 8 * world) != cudaSuccess) return -1;
    cudaMemcpy(d_a, my_tab.data(), 27 * 8, cudaMemcpyHostToDevice);
    int rc = nccl.AllGather(d_a, d_b, 27 * 8, /*ncclChar*/ 0, comm, st);
    if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) rc = -1;
    if (rc == 0) cudaMemcpy(all_tab.data(), d_b, (size_t)27 * 8 * world, cudaMemcpyDeviceToHost);
    cudaFree(d_a); cudaFree(d_b);
    if (rc != 0) return -1;
  }
  // the peer's total (for the parity stride of ITS mailbox) is not needed: parity blocks are addressed from separate bases,
  // data[par] = base + par * peer_total -- so gather the totals as well
  std::vector<long long> totals((size_t)world, 0);
  {
    long long *d_a = nullptr, *d_b = nullptr;
    if (cudaMalloc(&d_a, 8) != cudaSuccess || cudaMalloc(&d_b, (size_t)8 * world) != cudaSuccess) return -1;
    cudaMemcpy(d_a, &total, 8, cudaMemcpyHostToDevice);
    int rc = nccl.AllGather(d_a, d_b, 8, /*ncclChar*/ 0, comm, st);
    if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) rc = -1;
    if (rc == 0) cudaMemcpy(totals.data(), d_b, (size_t)8 * world, cudaMemcpyDeviceToHost);
    cudaFree(d_a); cudaFree(d_b);
    if (rc != 0) return -1;
  }
  auto data_base = [&](int r, int par, long long tot) { return reinterpret_cast<double*>(p.region.mapped[(size_t)r]) + (size_t)par * (size_t)tot; };
  auto flag_base = [&](int r, long long tot) {
    const size_t db = (sizeof(double) * 2 * (size_t)std::max<long long>(tot, 1) + 255) / 256 * 256;
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(p.region.mapped[(size_t)r]) + db);
  };
  p.nnb = (int)mine.first.size();
  if (cudaMalloc(&p.d_seq, sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&p.d_done, sizeof(unsigned int)) != cudaSuccess) return -1;
  cudaMemset(p.d_seq, 0, sizeof(unsigned long long)); cudaMemset(p.d_done, 0, sizeof(unsigned int));
  p.local_data[0] = data_base(rank, 0, total); p.local_data[1] = data_base(rank, 1, total);
  if (p.nnb == 0) { p.built = true; return 0; }
  if (p.nnb > 26) return -1;
  cudaMalloc(&p.d_counters, sizeof(unsigned int) * p.nnb); cudaMemset(p.d_counters, 0, sizeof(unsigned int) * p.nnb);
  std::vector<AddNeighbourDev> host((size_t)p.nnb);
  std::map<unsigned int, std::vector<std::pair<int, int>>> contrib;      // dof -> (source rank, mailbox position)
  std::vector<std::array<long long, 3>> nodes;
  for (int i = 0; i < p.nnb; ++i) {
    const Nb& nb = mine.first[(size_t)i];
    shared_nodes(Lat, nb.dir, dim, nodes);
    if ((long long)nodes.size() * R != nb.count) return -1;
    std::vector<unsigned int> idx(nodes.size() * (size_t)R);
    for (size_t q = 0; q < nodes.size(); ++q) for (int c = 0; c < R; ++c) {
      const long long dof = lagrange_dof(L, nodes[q][0], nodes[q][1], nodes[q][2]) * R + c; const size_t pos = q * (size_t)R + c;
      if (dof > 0xffffffffll || nb.offset + (long long)pos > 0x7fffffffll) return -1;
      idx[pos] = (unsigned int)dof;
      contrib[(unsigned int)dof].push_back({nb.peer, (int)(nb.offset + (long long)pos)});
    }
    unsigned int* d_idx = nullptr; if (cudaMalloc(&d_idx, idx.size() * 4) != cudaSuccess) return -1;
    cudaMemcpy(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice); p.owned.push_back(d_idx);
    const int back = (-nb.dir[0] + 1) + 3 * ((-nb.dir[1] + 1) + 3 * (-nb.dir[2] + 1));
    const long long roff = all_tab[(size_t)27 * nb.peer + back];
    if (roff < 0) return -1;
    const int mycode = (nb.dir[0] + 1) + 3 * ((nb.dir[1] + 1) + 3 * (nb.dir[2] + 1));
    AddNeighbourDev& d = host[(size_t)i];
    d.total = nb.count; d.send_idx = d_idx; d.block_begin = p.send_grid;
    d.nblocks = (int)((d.total + kP2PThreads * kP2PItems - 1) / (kP2PThreads * kP2PItems)); p.send_grid += d.nblocks;
    d.remote_data[0] = data_base(nb.peer, 0, totals[(size_t)nb.peer]) + roff; d.remote_data[1] = data_base(nb.peer, 1, totals[(size_t)nb.peer]) + roff;
    d.remote_ready = flag_base(nb.peer, totals[(size_t)nb.peer]) + 2 * back;          // flags of the peer: [direction the message comes from][parity]
    d.local_ready = flag_base(rank, total) + 2 * mycode;
    d.counter = p.d_counters + i;
  }
  cudaMalloc(&p.d_nb, sizeof(AddNeighbourDev) * p.nnb);
  cudaMemcpy(p.d_nb, host.data(), sizeof(AddNeighbourDev) * p.nnb, cudaMemcpyHostToDevice);
  // CSR over shared dofs, sources sorted by rank with the own value (-1) at the position of this rank
  std::vector<unsigned int> dofs; std::vector<int> ptr(1, 0), src;
  for (auto& kv : contrib) {
    auto lst = kv.second; std::sort(lst.begin(), lst.end());
    bool own_done = false;
    for (auto& e : lst) { if (!own_done && e.first > rank) { src.push_back(-1); own_done = true; } src.push_back(e.second); }
    if (!own_done) src.push_back(-1);
    dofs.push_back(kv.first); ptr.push_back((int)src.size());
  }
  p.nshared = (long long)dofs.size();
  if (cudaMalloc(&p.d_dof, std::max<size_t>(dofs.size(), 1) * 4) != cudaSuccess || cudaMalloc(&p.d_ptr, ptr.size() * 4) != cudaSuccess || cudaMalloc(&p.d_src, std::max<size_t>(src.size(), 1) * 4) != cudaSuccess) return -1;
  cudaMemcpy(p.d_dof, dofs.data(), dofs.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(p.d_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(p.d_src, src.data(), src.size() * 4, cudaMemcpyHostToDevice);
  p.built = cudaGetLastError() == cudaSuccess; return p.built ? 0 : -1;
}
int halo_exchange_add_p2p(HaloPlanAddP2P& p, double* v, int* d_err, cudaStream_t st) {
  if (!p.built) return -1;
  if (p.nnb == 0) return 0;
  add_send_kernel<<<p.send_grid, kP2PThreads, 0, st>>>(v, p.d_nb, p.nnb, p.d_seq);
  const int grid = (int)std::min<long long>(592, (p.nshared + kP2PThreads - 1) / kP2PThreads);
  add_recv_kernel<<<std::max(grid, 1), kP2PThreads, 0, st>>>(v, p.d_nb, p.nnb, p.local_data[0], p.local_data[1], p.d_dof, p.d_ptr, p.d_src, p.nshared, p.d_seq, p.d_done, d_err);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace b200fem
