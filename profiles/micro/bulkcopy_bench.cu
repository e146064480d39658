// bulkcopy_bench.cu -- microbenchmark: what does a persistent global->shared->global stream sustain on B200 when it is
// driven by cp.async.bulk row copies of the sizes the DG kernels use?  (Design input for dg_kronecker_pipe.cuh.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulkcopy_bench bulkcopy_bench.cu && ./bulkcopy_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
}
__device__ __forceinline__ void g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void s2g(void* dst, uint32_t src, uint32_t bytes) { asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory"); }

// one warp per CTA: stage s holds `rows` rows of `row_bytes`; chunk c = rows*row_bytes contiguous bytes of the input.
// mode 0: load only (no store), mode 1: load + store (copy), STAGES-deep ring; store of a stage must have been read
// before the stage is reloaded.
template <int STAGES>
__global__ void __launch_bounds__(32, 1) stream_kernel(const char* __restrict__ in, char* __restrict__ out, int rows, int row_bytes, long long nchunks, int mode, int halo_rows) {
  extern __shared__ __align__(128) char sm[];
  const int lane = threadIdx.x;
  const long long chunk_bytes = (long long)rows * row_bytes;
  const int stage_rows = rows + halo_rows;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)STAGES * stage_rows * row_bytes);
  if (lane == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(saddr(bars + s), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  long long issued = 0, consumed = 0;
  const long long first = blockIdx.x, step = gridDim.x;
  long long mine = (nchunks - first + step - 1) / step; if (mine < 0) mine = 0;
  while (consumed < mine) {
    // keep STAGES loads in flight
    while (issued < mine && issued < consumed + STAGES) {
      const int s = (int)(issued % STAGES);
      const long long c = first + issued * step;
      if (mode == 1) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");   // the store that last used this stage has been read
      __syncwarp();
      if (lane == 0) mbar_expect(saddr(bars + s), (uint32_t)stage_rows * row_bytes);
      __syncwarp();
      for (int r = lane; r < stage_rows; r += 32) {
        // halo rows re-read rows of a neighbouring chunk (L2 hits)
        long long src_chunk = c, rr = r;
        if (r >= rows) { src_chunk = (c + 1 + (r - rows)) % nchunks; rr = (r - rows) % rows; }
        g2s(saddr(sm + ((size_t)s * stage_rows + r) * row_bytes), in + src_chunk * chunk_bytes + rr * row_bytes, row_bytes, saddr(bars + s));
      }
      ++issued;
    }
    const int s = (int)(consumed % STAGES);
    mbar_wait(saddr(bars + s), (uint32_t)((consumed / STAGES) & 1));
    const long long c = first + consumed * step;
    if (mode == 1) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int r = lane; r < rows; r += 32) s2g(out + c * chunk_bytes + (long long)r * row_bytes, saddr(sm + ((size_t)s * stage_rows + r) * row_bytes), row_bytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    ++consumed;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// DG tile pattern: 64^3 elements of 216 B, tiles of 8x4x4 elements; per tile 16 interior rows of 12 elements (x halo),
// 16 y/z halo rows of 8 elements, 16 output rows of 8 elements.  One warp per CTA, 2-stage ring, tile = bid + it*grid.
__global__ void __launch_bounds__(32, 1) tile_kernel(const char* __restrict__ in, char* __restrict__ out, int ncell, int mode) {
  extern __shared__ __align__(128) char sm[];
  constexpr int EB = 216, RI = 12 * EB, RH = 8 * EB, STAGE = 16 * RI + 16 * RH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * STAGE);
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(saddr(bars), 1); mbar_init(saddr(bars + 1), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const int tx = ncell / 8, ty = ncell / 4, tz = ncell / 4, ntiles = tx * ty * tz;
  auto issue = [&](int tile, int s) {
    const int bx = tile % tx, by = (tile / tx) % ty, bz = tile / (tx * ty);
    const int x0 = bx * 8, y0 = by * 4, z0 = bz * 4;
    uint32_t bytes = 0;
    int ly, lz; bool interior = lane < 16;
    if (interior) { ly = y0 + lane % 4; lz = z0 + lane / 4; }
    else { const int h = lane - 16; if (h < 4) { ly = y0 - 1; lz = z0 + h; } else if (h < 8) { ly = y0 + 4; lz = z0 + h - 4; } else if (h < 12) { ly = y0 + h - 8; lz = z0 - 1; } else { ly = y0 + h - 12; lz = z0 + 4; } }
    char* dst = sm + (size_t)s * STAGE + (interior ? lane * RI : 16 * RI + (lane - 16) * RH);
    if (ly >= 0 && ly < ncell && lz >= 0 && lz < ncell) {
      const long long row_e = (long long)ncell * (ly + (long long)ncell * lz);
      int xs = interior ? (x0 >= 2 ? x0 - 2 : 0) : x0, xe = interior ? (x0 + 10 <= ncell ? x0 + 10 : ncell) : x0 + 8;
      bytes = (uint32_t)(xe - xs) * EB;
      g2s(saddr(dst), in + (row_e + xs) * EB, bytes, saddr(bars + s));
    }
    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
    if (lane == 0) mbar_expect(saddr(bars + s), bytes);
    __syncwarp();
  };
  int mine = 0; for (int t = blockIdx.x; t < ntiles; t += gridDim.x) ++mine;
  for (int it = 0; it < 2 && it < mine; ++it) issue(blockIdx.x + it * gridDim.x, it);
  for (int it = 0; it < mine; ++it) {
    const int s = it & 1, tile = blockIdx.x + it * gridDim.x;
    mbar_wait(saddr(bars + s), (uint32_t)((it >> 1) & 1));
    if (mode == 1) {
      const int bx = tile % tx, by = (tile / tx) % ty, bz = tile / (tx * ty);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane < 16) { const int ly = by * 4 + lane % 4, lz = bz * 4 + lane / 4; const long long row_e = (long long)ncell * (ly + (long long)ncell * lz);
        s2g(out + (row_e + bx * 8) * EB, saddr(sm + (size_t)s * STAGE + lane * RI + 2 * EB), RH); }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
    if (it + 2 < mine) issue(blockIdx.x + (it + 2) * gridDim.x, s);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int STAGES> float run(const char* in, char* out, int rows, int row_bytes, long long nchunks, int mode, int halo_rows, int ctas_per_sm) {
  size_t smem = (size_t)STAGES * (rows + halo_rows) * row_bytes + 64;
  cudaFuncSetAttribute(stream_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    stream_kernel<STAGES><<<148 * ctas_per_sm, 32, smem>>>(in, out, rows, row_bytes, nchunks, mode, halo_rows);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError(); if (err != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(err));
  return best;
}

int main() {
  const long long total = 432ll << 20;   // 453 MB in, 453 MB out (> L2)
  char *in, *out; cudaMalloc(&in, total); cudaMalloc(&out, total); cudaMemset(in, 1, total); cudaMemset(out, 0, total);
  printf("%-44s %10s %10s\n", "config", "ms", "GB/s(in+out algorithmic)");
  struct Cfg { int rows, row_bytes, halo, ctas; };
  std::vector<Cfg> cfgs = {{16, 1728, 0, 1}, {16, 1728, 16, 1}, {16, 1728, 0, 2}, {32, 1728, 0, 1}, {4, 13824, 0, 1}, {8, 6912, 0, 1}, {16, 1728, 16, 2}};
  for (auto c : cfgs) {
    const long long chunk = (long long)c.rows * c.row_bytes, nchunks = total / chunk;
    for (int mode = 0; mode < 2; ++mode) {
      float t2 = run<2>(in, out, c.rows, c.row_bytes, nchunks, mode, c.halo, c.ctas);
      float t3 = run<3>(in, out, c.rows, c.row_bytes, nchunks, mode, c.halo, c.ctas);
      float t4 = (size_t)4 * (c.rows + c.halo) * c.row_bytes * c.ctas < 220000 ? run<4>(in, out, c.rows, c.row_bytes, nchunks, mode, c.halo, c.ctas) : -1.f;
      const double bytes = (double)nchunks * chunk * (mode ? 2 : 1);
      printf("rows=%2d row=%5dB halo=%2d ctas/SM=%d %s  S2 %.3f ms %6.0f GB/s | S3 %.3f ms %6.0f GB/s | S4 %.3f ms %6.0f GB/s\n", c.rows, c.row_bytes, c.halo, c.ctas,
             mode ? "copy" : "load", t2, bytes / t2 / 1e6, t3, bytes / t3 / 1e6, t4, t4 > 0 ? bytes / t4 / 1e6 : 0.0);
    }
  }
  {
    const int ncell = 64; const size_t smem = 2 * (16 * 12 * 216 + 16 * 8 * 216) + 64;
    cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t vec = (size_t)ncell * ncell * ncell * 216;
    for (int mode = 0; mode < 2; ++mode) for (int grid : {148, 296}) {
      float best = 1e30f;
      for (int rep = 0; rep < 8; ++rep) {
        // rotate through the 453 MB buffers so that inputs are not L2 resident
        const size_t off = (size_t)(rep % 7) * vec;
        cudaEventRecord(e0); tile_kernel<<<grid, 32, smem>>>(in + off, out + off, ncell, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
      }
      cudaError_t err = cudaGetLastError(); if (err != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(err));
      printf("DG tile pattern 64^3 x 216 B, 8x4x4 tiles, grid=%d, %s: %.1f us  (%.0f GB/s at 16 B/dof-equivalent %s)\n", grid, mode ? "load+store" : "load only", best * 1e3,
             (mode ? 2.0 : 1.0) * vec / best / 1e6, mode ? "in+out" : "in");
    }
  }
  return 0;
}
