// cg_coop2d.cuh -- launch-bound CG (BASELINE config 1: P1 Lagrange, 2-D 256^2, 66 k dofs): a chunk of CG iterations as ONE
// cooperative kernel launch.
//
// At this size every vector fits in L2 many times over and an iteration is pure launch latency: the CUDA-graph version still
// pays four dependent kernel boundaries per iteration (17.8 us).  Here the grid stays resident and the three global
// dependencies of an iteration (p must be complete before the stencil reads its neighbours; <p,h> before alpha; <r,r> before
// beta) become grid-wide barriers.  Same recurrence and sign conventions as solver/linear/cg.hh:18-117 (unpreconditioned
// branch); same operator as lagrange_kronecker.cuh, applied as the (2k+1)^2-point lattice stencil
//     (A u)(g0,g1) = sum_{a,b} [ T0[g0][a] M1[g1][b] + M0[g0][a] T1[g1][b] ] u(g0+a-k, g1+b-k)
// with the same assembled 1-D rows.  Scalars are carried redundantly by every thread: each block sums the per-block partials
// in the same order, so all blocks hold bit-identical alpha / beta / residual and no extra barrier is needed to publish them.
// Deterministic (fixed partial order).  Single rank only (every dof is primary).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include "lagrange_kronecker.cuh"
#include "vec_types.hpp"

namespace b200fem {

constexpr int kCoopThreads = 256;

template <int K>
__global__ void __launch_bounds__(kCoopThreads) cg_coop2d_kernel(const __grid_constant__ LagrangeLayoutDev L, const __grid_constant__ LagKronRows R,
                                                                 double* __restrict__ x, double* __restrict__ r, double* __restrict__ p, double* __restrict__ h,
                                                                 const unsigned char* __restrict__ dmask, double* partial, CgState* st, double* __restrict__ history,
                                                                 const int max_iters) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  constexpr int W = 2 * K + 1;
  const int L0 = (int)L.lattice[0], L1 = (int)L.lattice[1];
  const long long nodes = (long long)L0 * L1;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  double* part_a = partial; double* part_b = partial + gridDim.x;
  // thread-local copy of the CG state (identical in all threads by construction)
  double residual = st->residual, prev_residual = st->prev_residual; const double tolerance = st->tolerance;
  int iterations = st->iterations; bool done = st->done != 0; const int max_total = st->max_iterations;

  for (int it = 0; it < max_iters && !done; ++it) {
    // ---- p <- beta p - r (cg.hh:72-85; the first direction is p = b - A x from cg_init_kernel)
    if (iterations > 0) {
      const double beta = residual / prev_residual;
      for (long long i = tid; i < nodes; i += nth) { const long long d = lagrange_dof(L, i % L0, i / L0, 0); p[d] = p[d] * beta - r[d]; }
    }
    grid.sync();
    // ---- h = A p, partial <p, h>
    double acc = 0;
    for (long long i = tid; i < nodes; i += nth) {
      const int g0 = (int)(i % L0), g1 = (int)(i / L0);
      const long long d = lagrange_dof(L, g0, g1, 0);
      const double pc = p[d];
      double w = 0;
      if (dmask && dmask[d]) w = pc;                       // DirichletWrapperOperator row of the homogeneous part
      else {
        const double* m0 = R.M[0] + (size_t)g0 * W; const double* t0 = R.T[0] + (size_t)g0 * W;
        const double* m1 = R.M[1] + (size_t)g1 * W; const double* t1 = R.T[1] + (size_t)g1 * W;
#pragma unroll
        for (int b = 0; b < W; ++b) {
          const int y = g1 + b - K; if (y < 0 || y >= L1) continue;
          double sm = 0, stt = 0;                          // sum_a M0[a] u, sum_a T0[a] u along x
#pragma unroll
          for (int a = 0; a < W; ++a) {
            const int xx = g0 + a - K; if (xx < 0 || xx >= L0) continue;
            const double uv = p[lagrange_dof(L, xx, y, 0)];
            sm = fma(m0[a], uv, sm); stt = fma(t0[a], uv, stt);
          }
          w = fma(m1[b], stt, w); w = fma(t1[b], sm, w);
        }
      }
      h[d] = w; acc = fma(pc, w, acc);
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part_a[blockIdx.x] = acc;
    grid.sync();
    double qdoth = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) qdoth += __ldcg(part_a + i);
    qdoth = block_sum(qdoth);
    __shared__ double bcast;
    if (threadIdx.x == 0) bcast = qdoth;
    __syncthreads();
    const double alpha = residual / bcast;                 // cg.hh:89-90
    // ---- x += alpha p ; r += alpha h ; partial <r, r>     (cg.hh:92, 103-107)
    double rr = 0;
    for (long long i = tid; i < nodes; i += nth) {
      const long long d = lagrange_dof(L, i % L0, i / L0, 0);
      x[d] = fma(alpha, p[d], x[d]);
      const double rv = fma(alpha, h[d], r[d]); r[d] = rv; rr = fma(rv, rv, rr);
    }
    rr = block_sum(rr);
    if (threadIdx.x == 0) part_b[blockIdx.x] = rr;
    grid.sync();
    double rsum = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) rsum += __ldcg(part_b + i);
    rsum = block_sum(rsum);
    __syncthreads();                                       // (bcast was read by everyone before it is rewritten)
    if (threadIdx.x == 0) bcast = rsum;
    __syncthreads();
    prev_residual = residual; residual = bcast;
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->qdoth = qdoth; st->alpha = alpha; if (history) history[iterations] = sqrt(residual); }
    iterations += 1;
    if (!(residual > tolerance) || iterations >= max_total) done = true;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->residual = residual; st->prev_residual = prev_residual; st->iterations = iterations; st->done = done ? 1 : 0; }
}

}  // namespace b200fem
