// vec_types.hpp -- device-resident solver state, reduction geometry and the block reduction shared by the vector kernels
// (vec_kernels.cuh), the Lagrange lattice kernel and the cooperative CG kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace b200fem {

constexpr int kRedBlocks = 592;     // 4 per SM on 148 SMs
constexpr int kRedThreads = 256;

// device-resident CG state
struct CgState {
  double residual, prev_residual, qdoth, alpha, beta, tolerance, bnorm2;
  int iterations, done, max_iterations, tol_criteria;
  double epsilon;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double part[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) part[wid] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_sum(v);
  return v;    // valid in thread 0
}

// BiCGStab scalars (solver/linear/bicgstab.hh:64-214)
struct BicgState {
  double nu, alpha, omega, beta, res, tolerance, bnorm2;
  int iterations, done, max_iterations, tol_criteria, x_applied;
  double epsilon;
};
// GMRES: up to kGemvChunk basis vectors per sweep (solver/linear/gmres.hh:64-92)
constexpr int kGemvChunk = 8;
struct GmresVecs { const double* v[kGemvChunk]; };
// AutomaticDifferenceLinearOperator (operator/common/automaticdifferenceoperator.hh:124-166)
struct FdState { double eps_given, norm_u, eps; };

// 1-D assembled row tables of the Lagrange Kronecker form, device pointers ([L_d][2k+1]: row g = columns g-k .. g+k)
struct LagKronRows {
  const double* M[3];
  const double* T[3];
};

}  // namespace b200fem
