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
