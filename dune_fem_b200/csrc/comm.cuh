// comm.cuh -- communication layer: halo exchange of dof vectors between the boxes of a Cartesian domain decomposition and
// the global sum of scalar products, for one process per GPU on one NVSwitch box.
//
// Replaces DiscreteFunction::communicate() (dune/fem/function/common/discretefunction.hh:825-835 ->
// space/common/communicationmanager.hh:130-150): per shared entity the owner's dof block is sent and
//   DG spaces       Copy  into the ghost copy          (space/discontinuousgalerkin/space.hh:80)
//   Lagrange spaces Add   on dofs of shared entities   (space/lagrange/space.hh:92)
// (operations: space/common/commoperations.hh:126-205), and ParallelScalarProduct's comm().sum
// (function/common/scalarproducts.hh:115-127).  Like the reference's cached communicator
// (space/common/cachedcommmanager.hh:943-975) all index lists are built once per operator.
//
// Two transports:
//  * PEER MEMORY (default): every rank owns a symmetric region (cudaMalloc, exported with cudaIpcGetMemHandle, handles
//    all-gathered once through NCCL).  A message is written straight into the receiver's mailbox over NVLink by the
//    sender's kernel, followed by a release-store of a sequence number; the receiver's kernel acquires the flag and
//    unpacks.  Sequence numbers live in device memory so that whole Krylov iterations can be captured into CUDA graphs.
//    Mailboxes are double-buffered by sequence parity and need no acknowledgements: rank A can only start message s+2
//    after its exchange s+1 has completed, i.e. after it has seen B's flag s+1, which B's kernel published after B's
//    exchange s had retired (stream order on B) -- so B has consumed message s before A overwrites its buffer.
//    Kernels never wait for a message before they have sent their own, and no block waits for another block of its own
//    grid to be scheduled: sends and receives are separate launches (or, in the marching DG kernel, the receive part
//    runs in the tail of a persistent grid whose CTAs are all resident).
//  * NCCL send/recv + ncclAllReduce: fallback when peer mappings cannot be established.
//
// Time-outs: a rank that waits longer than kCommTimeoutNs for a flag writes an error code into a host-mapped word that
// every API entry point checks (B200FEM_ERR_COMM); the kernel finishes without hanging the device.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "dg_quadrature.cuh"
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
    if (!(GetUniqueId && CommInitRank && CommDestroy && Send && Recv && AllReduce && AllGather && GroupStart && GroupEnd)) { handle = nullptr; return false; }
    return true;
  }
};

// ----------------------------------------------------------------------------------------------------------------------
// device-side primitives
constexpr unsigned long long kCommTimeoutNs = 20ull * 1000 * 1000 * 1000;   // a lost peer must not hang the box
enum CommError { kCommOk = 0, kCommTimeoutHalo = 1, kCommTimeoutScalars = 2, kCommTimeoutFused = 3 };

__device__ __forceinline__ unsigned long long gtimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
// spins until *flag >= want; on time-out records `code` in the host-mapped error word and returns false
__device__ __forceinline__ bool wait_flag_ge(const unsigned long long* flag, unsigned long long want, int* err, int code) {
  if (ld_acquire_sys(flag) >= want) return true;
  const unsigned long long t0 = gtimer_ns();
  while (ld_acquire_sys(flag) < want) {
    if (gtimer_ns() - t0 > kCommTimeoutNs) { if (err) *reinterpret_cast<volatile int*>(err) = code; return false; }
    __nanosleep(64);
  }
  return true;
}

// ----------------------------------------------------------------------------------------------------------------------
// symmetric peer-mapped memory: one allocation per rank, every rank maps all of them
constexpr int kMaxPeers = 16;
struct PeerRegion {
  void* local = nullptr; size_t bytes = 0; int rank = 0, world = 1; bool ok = false;
  std::vector<void*> mapped;           // mapped[r]: rank r's region in this process (mapped[rank] == local)
};
// collective over the communicator; zero-filled; returns 0 on success on ALL ranks, -1 on all ranks otherwise
int peer_region_create(NcclApi& nccl, void* comm, int rank, int world, size_t bytes, cudaStream_t st, PeerRegion& out);
void peer_region_free(PeerRegion& r);

// ---- global sum of up to kArMax scalars: every rank stores its values into every rank's table, then sums the table in
// rank order (bit-identical result on all ranks, independent of timing) ----
constexpr int kArMax = 8;
struct PeerScalarsDev {
  int rank, world;
  double* vals[kMaxPeers];                  // vals[r]: table of rank r: [parity][source rank][kArMax]
  unsigned long long* flags[kMaxPeers];     // flags[r]: [parity][source rank]
  unsigned long long* seq;                  // device-resident sequence number (local)
  int* err;
};
struct PeerScalars { PeerRegion region; PeerScalarsDev dev{}; unsigned long long* d_seq = nullptr; bool ok = false; };
int peer_scalars_create(NcclApi& nccl, void* comm, int rank, int world, cudaStream_t st, int* d_err, PeerScalars& out);
void peer_scalars_free(PeerScalars& s);

// To be called by ALL threads of ONE block per rank (the block that finishes a two-stage reduction last): sh[0..cnt) in
// shared memory holds this rank's values on entry and the global sums on return.
__device__ __forceinline__ void peer_allreduce_block(const PeerScalarsDev& A, double* sh, const int cnt) {
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = *A.seq + 1;
  __syncthreads();
  const unsigned long long s = s_seq; const int par = (int)(s & 1);
  if ((int)threadIdx.x < A.world) {           // thread r delivers to rank r (all NVLink round trips overlap)
    const int r = threadIdx.x;
    double* dst = A.vals[r] + ((size_t)par * A.world + A.rank) * kArMax;
    for (int i = 0; i < cnt; ++i) dst[i] = sh[i];
    __threadfence_system();
    st_release_sys(A.flags[r] + (size_t)par * A.world + A.rank, s);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc[kArMax];
#pragma unroll
    for (int i = 0; i < kArMax; ++i) acc[i] = 0.0;
    const double* tab = A.vals[A.rank] + (size_t)par * A.world * kArMax;
    for (int r = 0; r < A.world; ++r) {
      wait_flag_ge(A.flags[A.rank] + (size_t)par * A.world + r, s, A.err, kCommTimeoutScalars);
      for (int i = 0; i < cnt; ++i) acc[i] += ld_relaxed_sys_f64(tab + (size_t)r * kArMax + i);
    }
    for (int i = 0; i < cnt; ++i) sh[i] = acc[i];
    *A.seq = s;
  }
  __syncthreads();
}

// ----------------------------------------------------------------------------------------------------------------------
// NCCL transport (fallback): index lists per axis side, pack -> ncclSend/Recv -> unpack, axis by axis (later axes forward
// what earlier ones received, so edges and corners become consistent with 6 messages)
struct HaloSide {
  int peer = -1; long long count = 0;
  long long* d_send_idx = nullptr; long long* d_recv_idx = nullptr;   // block start offsets in the dof vector
  double* d_send = nullptr; double* d_recv = nullptr;
};
struct HaloPlan { int block = 1; HaloSide side[3][2]; bool built = false; };
// builds the per-axis lists AND the auxiliary-dof mask (space/common/auxiliarydofs.hh:215-275: the lowest rank owning a
// copy is primary; DG: ghost elements are auxiliary)
int halo_plan_build(HaloPlan& p, const int proc[3], const int pc[3], const BoxDev& box, bool lagrange, int order, int nb,
                    const LagrangeLayoutDev& layout_dev, long long size, uint8_t** d_aux_out);
void halo_plan_free(HaloPlan& p);
int halo_exchange(HaloPlan& p, NcclApi& nccl, void* comm, double* v, bool add, cudaStream_t st);

// single-phase exchange for DG spaces: every existing neighbour among the 26 gets its own message
struct HaloNeighbour { int peer = -1; int dir = 0; long long count = 0; long long *d_send_idx = nullptr, *d_recv_idx = nullptr; double *d_send = nullptr, *d_recv = nullptr; };
struct HaloPlanDG { int block = 1; std::vector<HaloNeighbour> nb; bool built = false; };
int halo_plan_dg_build(HaloPlanDG& p, const int proc[3], const int pc[3], const BoxDev& box, int nb);
void halo_plan_dg_free(HaloPlanDG& p);
int halo_exchange_dg(HaloPlanDG& p, NcclApi& nccl, void* comm, double* v, cudaStream_t st);

// ----------------------------------------------------------------------------------------------------------------------
// peer-memory Copy exchange of DG spaces
struct P2PNeighbourDev {
  long long total;                                   // doubles per message
  const unsigned int* send_flat; const unsigned int* recv_flat;   // per double: offset in the dof vector
  int block_begin, nblocks;                          // this neighbour's slice of the grid
  double* remote_data[2]; unsigned long long* remote_ready;       // in the peer's mailbox (ready[2], by parity)
  const double* local_data[2]; const unsigned long long* local_ready;   // in my mailbox
  unsigned int* counter;                             // blocks that have stored their slice (local)
};
// Exchange done by the marching DG kernel itself (dg_kronecker_march.cuh): directions (dy, dz) in the y-z plane of the
// process grid, code d = (dy+1) + 3 (dz+1).  A message is the dense array [z-part][y-part][on0 * n^3] of the owned
// interface layer (a part is the whole owned range for a zero direction component, one layer otherwise).
struct MarchCommDev {
  int any;                                           // 0: this launch exchanges nothing
  int enabled[9];
  double* remote[9][2]; unsigned long long* remote_ready[9];
  const double* local[9][2]; const unsigned long long* local_ready[9];
  unsigned int expected[9];                          // row segments (one per tile row and plane) that make up message d
  unsigned int* counters;                            // [0..8] segments stored so far, [9] CTAs that finished receiving, [10+d] receive items of message d handed out
  unsigned long long* seq;                           // device-resident sequence number, shared with the stand-alone kernels
  int* err;
  double* w;                                         // output vector (the receive part fills its ghost layers)
  unsigned long long* ts;                            // optional: globaltimer stamps of the tail phases of CTA 0 (diagnostics), else null
};
struct HaloPlanP2P {
  bool built = false; int block = 1; int nnb = 0; int grid = 0;
  PeerRegion region;                                 // my mailbox: per neighbour data[2][count * block], ready[2]
  P2PNeighbourDev* d_nb = nullptr; unsigned int* d_counters = nullptr; unsigned long long* d_seq = nullptr; unsigned int* d_done = nullptr;
  std::vector<void*> owned;                          // flat index arrays
  std::vector<P2PNeighbourDev> host_nb; std::vector<int> dir_code;   // host copy, direction code (dx+1) + 3 (dy+1) + 9 (dz+1)
  bool march_ok = false; MarchCommDev march{}; unsigned int* d_march_counters = nullptr; unsigned long long* d_march_ts = nullptr;
};
int halo_plan_p2p_build(HaloPlanP2P& p, HaloPlanDG& dg, NcclApi& nccl, void* comm, int rank, int world, const int proc[3], const int pc[3],
                        const int gn[3], const BoxDev& box, int nb, int* d_err, cudaStream_t st);
void halo_plan_p2p_free(HaloPlanP2P& p);
int halo_exchange_p2p(HaloPlanP2P& p, double* v, int* d_err, cudaStream_t st);      // two launches: send, receive

// ----------------------------------------------------------------------------------------------------------------------
// peer-memory Add exchange of Lagrange spaces: every rank sends its partial sums on the lattice nodes it shares with each
// of its (up to 26) neighbours; the receiver adds the contributions of ALL ranks sharing a node in rank order, its own
// included at its position -- every copy of a shared dof ends up bit-identical.
struct AddNeighbourDev {
  long long total; const unsigned int* send_idx; int block_begin, nblocks;
  double* remote_data[2]; unsigned long long* remote_ready; const unsigned long long* local_ready; unsigned int* counter;
};
struct HaloPlanAddP2P {
  bool built = false; int nnb = 0; int send_grid = 0; long long nshared = 0;
  PeerRegion region;
  AddNeighbourDev* d_nb = nullptr; unsigned int* d_counters = nullptr; unsigned long long* d_seq = nullptr; unsigned int* d_done = nullptr;
  const double* local_data[2] = {nullptr, nullptr};  // my mailbox, per parity (all neighbours concatenated)
  unsigned int* d_dof = nullptr; int* d_ptr = nullptr; int* d_src = nullptr;   // CSR over shared dofs: sources in rank order, -1 = own value
  std::vector<void*> owned;
};
int halo_plan_add_build(HaloPlanAddP2P& p, NcclApi& nccl, void* comm, int rank, int world, const int proc[3], const int pc[3],
                        const LagrangeLayoutDev& layout_dev, const std::vector<long long>& host_lattice_map, int dim, int dim_range, int* d_err, cudaStream_t st);
void halo_plan_add_free(HaloPlanAddP2P& p);
int halo_exchange_add_p2p(HaloPlanAddP2P& p, double* v, int* d_err, cudaStream_t st);

}  // namespace b200fem
