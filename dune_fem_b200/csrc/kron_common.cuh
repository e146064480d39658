// kron_common.cuh -- pieces shared by the Kronecker-form DG kernels: the compile-time dof permutation of the
// hierarchical Legendre ordering, PTX wrappers for mbarriers / the bulk-copy (TMA) engine, and the streamed 1-D operator
// application out of shared memory.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "dg_kronecker.cuh"

namespace b200fem {

// perm[tensor index] = stored local index, computed at compile time (shapefunctionset/legendre.hh:169-194, 236-250)
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copies (SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// tensor-map TMA (SASS UTMALDG / UTMASTG)
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
}  // namespace ptx

// acc[.. i ..] += sum_j M[i*N+j] src[perm(.. j ..)] along tensor axis AX, reading the source element line by line from
// shared memory (keeps only one line of N values live besides the accumulators)
template <int N, int AX, bool HIER>
__device__ __forceinline__ void apply_axis_smem(const double* __restrict__ M, const double* __restrict__ src, double (&acc)[N * N * N]) {
  constexpr PermTable<N, HIER> P{};
  constexpr int st = AX == 0 ? N * N : AX == 1 ? N : 1;
#pragma unroll
  for (int l = 0; l < N * N; ++l) {
    const int base = AX == 0 ? l : AX == 1 ? (l / N) * N * N + (l % N) : l * N;     // base tensor index of line l
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

}  // namespace b200fem
