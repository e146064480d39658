// dmma_bench.cu -- FP64 tensor-core (mma.sync.m8n8k4.f64, "DMMA") throughput of one B200 against the plain DFMA rate.
// Question it answers (BASELINE.json north_star: "FP64 DMMA tensor cores only where the contraction is dense enough, and
// the choice is justified by counters"): is the DMMA peak on sm_100a above the DFMA peak, i.e. can a padded 6->8 / 18->20
// sum-factorisation GEMM win anything over FMA code?       nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_kernel(double* out, const double a, const double b, int iters) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = 0.5 * i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

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

// DFMA fed from shared memory: FPL FMAs per 8-byte shared load (the balance point of the Kronecker kernels)
template <int FPL>
__global__ void dfma_lds_kernel(double* out, const double a, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-6;
  __syncthreads();
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = i;
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const double v = sm[(idx + 32 * r) & 4095];
#pragma unroll
      for (int f = 0; f < FPL; ++f) acc[(r + f) & 7] = fma(acc[(r + f) & 7], a, v);
    }
    idx += 256;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

template <class K> float time_kernel(K launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, 8);
  const int iters = 4096;
  for (int w : {4, 8, 16, 32}) {
    float ms = time_kernel([&] { dfma_kernel<8><<<sms, w * 32>>>(out, 1.0000001, 1e-9, iters); });
    std::printf("DFMA  warps/SM %2d chains 8 : %7.2f TFLOP/s\n", w, 2.0 * sms * w * 32 * (double)iters * 8 * 8 / ms * 1e-9);
  }
  for (int w : {4, 8, 16, 32}) {
    float m1 = time_kernel([&] { dmma_kernel<1><<<sms, w * 32>>>(out, 1.0000001, 1e-9, iters); });
    float m4 = time_kernel([&] { dmma_kernel<4><<<sms, w * 32>>>(out, 1.0000001, 1e-9, iters); });
    float m8 = time_kernel([&] { dmma_kernel<8><<<sms, w * 32>>>(out, 1.0000001, 1e-9, iters); });
    const double fl = 2.0 * sms * w * (double)iters * 8 * 256;      // 8*8*4 FMA per warp instruction
    std::printf("DMMA m8n8k4 warps/SM %2d : chains 1 %7.2f | 4 %7.2f | 8 %7.2f TFLOP/s\n", w, fl * 1 / m1 * 1e-9, fl * 4 / m4 * 1e-9, fl * 8 / m8 * 1e-9);
  }
  for (int w : {8, 16, 32}) {
    float a1 = time_kernel([&] { dfma_lds_kernel<1><<<sms, w * 32, 32768>>>(out, 1.0000001, iters); });
    float a2 = time_kernel([&] { dfma_lds_kernel<2><<<sms, w * 32, 32768>>>(out, 1.0000001, iters); });
    float a4 = time_kernel([&] { dfma_lds_kernel<4><<<sms, w * 32, 32768>>>(out, 1.0000001, iters); });
    float a6 = time_kernel([&] { dfma_lds_kernel<6><<<sms, w * 32, 32768>>>(out, 1.0000001, iters); });
    const double fl = 2.0 * sms * w * 32 * (double)iters * 8;
    std::printf("DFMA + 1 LDS.64 per k FMA, warps/SM %2d : k=1 %6.2f | k=2 %6.2f | k=4 %6.2f | k=6 %6.2f TFLOP/s\n", w, fl * 1 / a1 * 1e-9, fl * 2 / a2 * 1e-9, fl * 4 / a4 * 1e-9, fl * 6 / a6 * 1e-9);
  }
  return 0;
}
