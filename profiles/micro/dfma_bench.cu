// dfma_bench.cu -- FP64 FMA throughput of one B200 as a function of resident warps per SM and independent chains per
// thread.  Answers: what FP64 rate can 4 / 8 / 12 / 16 warps per SM sustain (the Kronecker DG kernels run 8 consumer warps
// with 240 registers each)?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_bench dfma_bench.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int ILP>
__global__ void dfma_kernel(double* out, const double a, const double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// MIX: per DFMA, MIX/2 extra independent FP32 FMAs / integer ops in the same loop -- do they issue in the shadow of the
// half-rate FP64 pipe or do they cost issue cycles of their own?
template <int ILP, int MIX>
__global__ void dfma_mix_kernel(double* out, const double a, const double b, const float fa, int iters) {
  double acc[ILP]; float f[ILP]; int n[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x * 1e-3 + i; f[i] = i + 0.5f; n[i] = i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        acc[i] = fma(acc[i], a, b);
        if (MIX >= 1) f[i] = fmaf(f[i], fa, 0.25f);
        if (MIX >= 2) n[i] = n[i] * 3 + it;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i] + f[i] + n[i];
  if (s == 123.456) out[0] = s;
}
template <int ILP, int MIX> void run_mix(int warps, int sms) {
  double* out; cudaMalloc(&out, 8);
  const int iters = 4096 / ILP * 4;
  auto k = dfma_mix_kernel<ILP, MIX>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<sms, warps * 32, 200 * 1024>>>(out, 1.0000001, 1e-9, 1.0001f, iters);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<<<sms, warps * 32, 200 * 1024>>>(out, 1.0000001, 1e-9, 1.0001f, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double flop = 2.0 * sms * warps * 32 * (double)iters * 8 * ILP;
  std::printf("warps/SM %2d  chains %2d  + %d other op(s) per DFMA : %7.2f FP64 TFLOP/s\n", warps, ILP, MIX, flop / ms * 1e-9);
  cudaFree(out);
}

template <int ILP> void run(int warps, int sms) {
  double* out; cudaMalloc(&out, 8);
  const int iters = 4096 / ILP * 4;
  auto k = dfma_kernel<ILP>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   // one CTA per SM
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<sms, warps * 32, 200 * 1024>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<<<sms, warps * 32, 200 * 1024>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double flop = 2.0 * sms * warps * 32 * (double)iters * 8 * ILP;
  std::printf("warps/SM %2d  chains/thread %2d : %7.2f TFLOP/s  (%.1f DFMA/clk/SM at 1.965 GHz)\n", warps, ILP, flop / ms * 1e-9, flop / 2 / ms * 1e-3 / sms / 1.965e9 * 1e3 / 1e3 * 1e3);
  cudaFree(out);
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int w : {4, 8, 12, 16, 32}) { run<1>(w, sms); run<3>(w, sms); run<9>(w, sms); run<27>(w, sms); }
  for (int w : {8, 16}) { run_mix<9, 0>(w, sms); run_mix<9, 1>(w, sms); run_mix<9, 2>(w, sms); }
  return 0;
}
