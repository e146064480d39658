// vec_kernels.cuh -- BLAS-1 and the CG recurrence on device dof vectors.
//
// Restates solver/linear/cg.hh:18-117 (sign conventions r = Ax - b, p = b - Ax, p <- beta p - r) with the vector
// sweeps of function/blockvectors/defaultblockvectors.hh:39-150 fused: per iteration the reference makes ~9 sweeps
// (128 B/dof); here it is  [p <- beta p - r] , [h = A p] , [<p,h>] , [x += a p ; r += a h ; <r,r>]  with all scalars
// (alpha, beta, residual, convergence flag, iteration count) resident on the device so that the loop needs no
// host round trip.  Dot products are two-stage, fixed-grid reductions with warp shuffles: deterministic.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "comm.cuh"
#include "vec_types.hpp"

namespace b200fem {

// partial[b] = sum over this block's grid-stride range of x*y restricted to primary dofs (mask may be null)
__global__ void __launch_bounds__(kRedThreads) dot_partial_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                                  const uint8_t* __restrict__ aux, long long n, double* __restrict__ partial) {
  double s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!aux || !aux[i]) s = fma(x[i], y[i], s);
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(kRedThreads) reduce_final_kernel(const double* __restrict__ partial, int nparts, double* __restrict__ out) {
  double s = 0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
  s = block_sum(s);
  if (threadIdx.x == 0) *out = s;
}
__global__ void axpy_kernel(double alpha, const double* __restrict__ x, double* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = fma(alpha, x[i], y[i]);
}
__global__ void negate_kernel(double* __restrict__ x, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = -x[i];
}

// The sweeps below move two doubles per thread and iteration (16-byte loads/stores): at 8 bytes per access the grid of
// 592 x 256 threads keeps too few bytes in flight for HBM3e (update_xr ran at 4.9 TB/s).  Vectors come from cudaMalloc
// (256-byte aligned); an odd tail element is handled by the last thread.
__device__ __forceinline__ bool primary2(const uint8_t* __restrict__ aux, long long i2, bool& p0, bool& p1) {
  if (!aux) { p0 = p1 = true; return true; }
  const uchar2 m = reinterpret_cast<const uchar2*>(aux)[i2]; p0 = !m.x; p1 = !m.y; return true;
}
__device__ __forceinline__ bool vec2_ok(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

// ---- CG pieces ----
// r = h - b ; p = b - h ; partial <p,p>          (cg.hh:39-60)
__global__ void __launch_bounds__(kRedThreads) cg_init_kernel(const double* __restrict__ h, const double* __restrict__ b, double* __restrict__ r,
                                                              double* __restrict__ p, const uint8_t* __restrict__ aux, long long n,
                                                              double* __restrict__ partial, double* __restrict__ partial_b) {
  double s = 0, sb = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double hv = h[i], bv = b[i];
    r[i] = hv - bv; const double pv = bv - hv; p[i] = pv;
    if (!aux || !aux[i]) { s = fma(pv, pv, s); sb = fma(bv, bv, sb); }
  }
  s = block_sum(s); if (threadIdx.x == 0) partial[blockIdx.x] = s;
  sb = block_sum(sb); if (threadIdx.x == 0) partial_b[blockIdx.x] = sb;
}
// residual = sum partial ; tolerance = eps^2 * {1 | <b,b> | residual}     (cg.hh:60-67). sums[0..1] hold the
// (already globally reduced) values.
__global__ void cg_init_final_kernel(const double* __restrict__ sums, CgState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->residual = sums[0]; st->bnorm2 = sums[1]; st->prev_residual = 0;
    const double scale = st->tol_criteria == 1 ? sums[1] : st->tol_criteria == 2 ? sums[0] : 1.0;
    st->tolerance = st->epsilon * st->epsilon * scale;
    st->iterations = 0;
    st->done = !(st->residual > st->tolerance) || st->max_iterations <= 0;
  }
}
// p <- beta p - r, beta = residual / prevResidual       (cg.hh:72-85, unpreconditioned: q aliases p)
__global__ void cg_update_p_kernel(double* __restrict__ p, const double* __restrict__ r, long long n, const CgState* st) {
  if (st->done || st->iterations == 0) return;            // the first search direction is p = b - A x (cg_init_kernel)
  const double beta = st->residual / st->prev_residual;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (vec2_ok(p, r, p, r)) {
    double2* p2 = reinterpret_cast<double2*>(p); const double2* r2 = reinterpret_cast<const double2*>(r);
    for (long long i = tid; i < n / 2; i += nth) { double2 pv = p2[i]; const double2 rv = r2[i]; pv.x = pv.x * beta - rv.x; pv.y = pv.y * beta - rv.y; p2[i] = pv; }
    if ((n & 1) && tid == 0) p[n - 1] = p[n - 1] * beta - r[n - 1];
  } else
    for (long long i = tid; i < n; i += nth) p[i] = p[i] * beta - r[i];
}
__global__ void __launch_bounds__(kRedThreads) cg_dot_kernel(const double* __restrict__ x, const double* __restrict__ y, const uint8_t* __restrict__ aux,
                                                             long long n, double* __restrict__ partial, const CgState* st) {
  if (st->done) return;
  double s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!aux || !aux[i]) s = fma(x[i], y[i], s);
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// sums[0] = <q,h> (globally reduced): alpha = residual / <q,h>          (cg.hh:89-90)
__global__ void cg_alpha_kernel(const double* __restrict__ sums, CgState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && !st->done) { st->qdoth = sums[0]; st->alpha = st->residual / sums[0]; }
}
// x += alpha q ; r += alpha h ; partial <r,r>                              (cg.hh:92, 103-107)
__global__ void __launch_bounds__(kRedThreads) cg_update_xr_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                                                   const double* __restrict__ h, const uint8_t* __restrict__ aux, long long n,
                                                                   double* __restrict__ partial, const CgState* st) {
  if (st->done) return;
  const double alpha = st->alpha;
  double s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double rv = fma(alpha, h[i], r[i]); r[i] = rv;
    if (!aux || !aux[i]) s = fma(rv, rv, s);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// prev = residual ; residual = <r,r> ; ++iterations ; convergence test     (cg.hh:70, 106-107, 116)
__global__ void cg_residual_kernel(const double* __restrict__ sums, CgState* st, double* __restrict__ history) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && !st->done) {
    st->prev_residual = st->residual; st->residual = sums[0];
    if (history) history[st->iterations] = sqrt(sums[0]);
    st->iterations += 1;
    if (!(st->residual > st->tolerance) || st->iterations >= st->max_iterations) st->done = 1;
  }
}

// ---- single-rank variants: the block that finishes last also does the second reduction stage and the scalar update, so an
// iteration is 4 launches (p-update, apply, <p,h>+alpha, x/r-update+residual) instead of 8.  The partials are summed by
// the same code in the same order as reduce_final_kernel: results are bit-identical to the two-kernel path and
// independent of which block happens to be last.
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
  __shared__ bool last;
  if (threadIdx.x == 0) { __threadfence(); last = atomicAdd(counter, 1u) == gridDim.x - 1; }
  __syncthreads();
  if (last) __threadfence();
  return last;
}
__device__ __forceinline__ double final_sum(const double* partial) {
  double s = 0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(partial + i);
  return block_sum(s);
}
// Several ranks: the block that has finished the local reduction also performs the global sum over peer memory (comm.cuh:
// every rank stores its value into every rank's table and adds the table in rank order), so an iteration keeps its 4
// launches and contains no library call -- it can be captured into a CUDA graph on every rank.  A.world <= 1: no-op.
// To be called by all threads of the block; t (in/out) is meaningful in thread 0.
__device__ __forceinline__ double global_sum1(const PeerScalarsDev& A, double t) {
  if (A.world <= 1) return t;
  __shared__ double sh_gs[kArMax];
  if (threadIdx.x == 0) sh_gs[0] = t;
  __syncthreads();
  peer_allreduce_block(A, sh_gs, 1);
  return sh_gs[0];
}
// second stage of `count` (<= kArMax) two-stage reductions + global sum: out[k] = sum over ranks of sum(partial[k * nparts ..])
__global__ void __launch_bounds__(kRedThreads) reduce_final_allreduce_kernel(const double* __restrict__ partial, int nparts, int count, double* __restrict__ out,
                                                                             const __grid_constant__ PeerScalarsDev A) {
  __shared__ double sh[kArMax];
  for (int k = 0; k < count; ++k) {
    double s = 0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[(size_t)k * nparts + i];
    s = block_sum(s);
    if (threadIdx.x == 0) sh[k] = s;
  }
  __syncthreads();
  if (A.world > 1) peer_allreduce_block(A, sh, count);
  if ((int)threadIdx.x < count) out[threadIdx.x] = sh[threadIdx.x];
}
__global__ void __launch_bounds__(kRedThreads) cg_dot_alpha_kernel(const double* __restrict__ x, const double* __restrict__ y, const uint8_t* __restrict__ aux,
                                                                   long long n, double* partial, CgState* st, unsigned int* counter,
                                                                   const __grid_constant__ PeerScalarsDev A) {
  if (st->done) return;
  double s = 0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (vec2_ok(x, y, x, y) && (!aux || (reinterpret_cast<uintptr_t>(aux) & 1) == 0)) {
    const double2* x2 = reinterpret_cast<const double2*>(x); const double2* y2 = reinterpret_cast<const double2*>(y);
    for (long long i = tid; i < n / 2; i += nth) {
      const double2 xv = x2[i], yv = y2[i]; bool p0, p1; primary2(aux, i, p0, p1);
      if (p0) s = fma(xv.x, yv.x, s);
      if (p1) s = fma(xv.y, yv.y, s);
    }
    if ((n & 1) && tid == 0 && (!aux || !aux[n - 1])) s = fma(x[n - 1], y[n - 1], s);
  } else
    for (long long i = tid; i < n; i += nth)
      if (!aux || !aux[i]) s = fma(x[i], y[i], s);
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (!last_block_done(counter)) return;
  const double t = global_sum1(A, final_sum(partial));
  if (threadIdx.x == 0) { st->qdoth = t; st->alpha = st->residual / t; *counter = 0; }
}
// <q,h> arrives as per-CTA partials of the apply kernel (lagrange_kronecker_kernel's fused scalar product): alpha = residual / <q,h>
__global__ void __launch_bounds__(kRedThreads) cg_alpha_partials_kernel(const double* __restrict__ partial, int nparts, CgState* st) {
  if (st->done) return;
  double s = 0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
  s = block_sum(s);
  if (threadIdx.x == 0) { st->qdoth = s; st->alpha = st->residual / s; }
}
__global__ void __launch_bounds__(kRedThreads) cg_update_xr_residual_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                                                            const double* __restrict__ h, const uint8_t* __restrict__ aux, long long n,
                                                                            double* partial, CgState* st, double* __restrict__ history, unsigned int* counter,
                                                                            const __grid_constant__ PeerScalarsDev A) {
  if (st->done) return;
  const double alpha = st->alpha;
  double s = 0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (vec2_ok(x, r, p, h) && (!aux || (reinterpret_cast<uintptr_t>(aux) & 1) == 0)) {
    double2* x2 = reinterpret_cast<double2*>(x); double2* r2 = reinterpret_cast<double2*>(r);
    const double2* p2 = reinterpret_cast<const double2*>(p); const double2* h2 = reinterpret_cast<const double2*>(h);
    for (long long i = tid; i < n / 2; i += nth) {
      double2 xv = x2[i], rv = r2[i]; const double2 pv = p2[i], hv = h2[i];
      xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y); x2[i] = xv;
      rv.x = fma(alpha, hv.x, rv.x); rv.y = fma(alpha, hv.y, rv.y); r2[i] = rv;
      bool p0, p1; primary2(aux, i, p0, p1);
      if (p0) s = fma(rv.x, rv.x, s);
      if (p1) s = fma(rv.y, rv.y, s);
    }
    if ((n & 1) && tid == 0) { const long long i = n - 1; x[i] = fma(alpha, p[i], x[i]); const double rv = fma(alpha, h[i], r[i]); r[i] = rv; if (!aux || !aux[i]) s = fma(rv, rv, s); }
  } else
    for (long long i = tid; i < n; i += nth) {
      x[i] = fma(alpha, p[i], x[i]);
      const double rv = fma(alpha, h[i], r[i]); r[i] = rv;
      if (!aux || !aux[i]) s = fma(rv, rv, s);
    }
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (!last_block_done(counter)) return;
  const double t = global_sum1(A, final_sum(partial));
  if (threadIdx.x == 0) {
    st->prev_residual = st->residual; st->residual = t;
    if (history) history[st->iterations] = sqrt(t);
    st->iterations += 1;
    if (!(st->residual > st->tolerance) || st->iterations >= st->max_iterations) st->done = 1;
    *counter = 0;
  }
}

// ---- BiCGStab (solver/linear/bicgstab.hh:64-214, unpreconditioned; five-fold scalar product of :19-52) ----
// Scalars live on the device; an iteration is  [tmp = A p] [<tmp,r*> -> alpha] [s = r - alpha tmp] [r = A s]
// [5 dots -> omega, res, beta, nu, convergence] [x += alpha p + omega s ; r = s - omega r ; p = r + beta (p - omega tmp)].
// r = b - r ; p = r ; r* = r ; partial <r,r*> and <b,b>                    (bicgstab.hh:94-122)
__global__ void __launch_bounds__(kRedThreads) bicg_init_kernel(double* __restrict__ r, const double* __restrict__ b, double* __restrict__ p, double* __restrict__ rstar,
                                                                const uint8_t* __restrict__ aux, long long n, double* __restrict__ partial, double* __restrict__ partial_b) {
  double s = 0, sb = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double bv = b[i], rv = -r[i] + bv;
    r[i] = rv; p[i] = rv; rstar[i] = rv;
    if (!aux || !aux[i]) { s = fma(rv, rv, s); sb = fma(bv, bv, sb); }
  }
  s = block_sum(s); if (threadIdx.x == 0) partial[blockIdx.x] = s;
  sb = block_sum(sb); if (threadIdx.x == 0) partial_b[blockIdx.x] = sb;
}
__global__ void bicg_init_final_kernel(const double* __restrict__ sums, BicgState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->nu = sums[0]; st->bnorm2 = sums[1];
    st->tolerance = st->epsilon * (st->tol_criteria == 1 ? sqrt(sums[1]) : st->tol_criteria == 2 ? sqrt(sums[0]) : 1.0);
    st->iterations = 0; st->x_applied = 0; st->done = 0;        // the reference tests convergence only after an iteration
  }
}
__device__ __forceinline__ void bicg_scalars(const double* gd, BicgState* st, double* history) {     // bicgstab.hh:166-183
  const double omega = gd[0] / gd[1];
  const double res = sqrt(gd[2] - omega * (2.0 * gd[0] - omega * gd[1]));
  const double nu_new = gd[3] - omega * gd[4];
  st->beta = nu_new * st->alpha / (omega * st->nu); st->nu = nu_new; st->omega = omega; st->res = res;
  if (history) history[st->iterations] = res;
  st->iterations += 1;
  if (res < st->tolerance || st->iterations >= st->max_iterations || !(res == res)) st->done = 1;
}
// partial <tmp, r*>; single rank: the last block computes alpha = nu / <tmp,r*>
__global__ void __launch_bounds__(kRedThreads) bicg_dot_alpha_kernel(const double* __restrict__ tmp, const double* __restrict__ rstar, const uint8_t* __restrict__ aux,
                                                                     long long n, double* partial, BicgState* st, unsigned int* counter,
                                                                     const __grid_constant__ PeerScalarsDev A) {
  if (st->done) return;
  double s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!aux || !aux[i]) s = fma(tmp[i], rstar[i], s);
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (!counter || !last_block_done(counter)) return;
  const double t = global_sum1(A, final_sum(partial));
  if (threadIdx.x == 0) { st->alpha = st->nu / t; *counter = 0; }
}
__global__ void bicg_alpha_kernel(const double* __restrict__ sums, BicgState* st) { if (threadIdx.x == 0 && blockIdx.x == 0 && !st->done) st->alpha = st->nu / sums[0]; }
// s = r - alpha tmp                                                        (bicgstab.hh:145-146)
__global__ void bicg_s_kernel(double* __restrict__ s, const double* __restrict__ r, const double* __restrict__ tmp, long long n, const BicgState* st) {
  if (st->done) return;
  const double ma = -st->alpha;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s[i] = fma(ma, tmp[i], r[i]);
}
// the five scalar products r.s, r.r, s.s, s.r*, r.r* in one sweep (scalarProductVecs); partial[k * gridDim.x + block]
__global__ void __launch_bounds__(kRedThreads) bicg_dots5_kernel(const double* __restrict__ r, const double* __restrict__ s, const double* __restrict__ rstar,
                                                                 const uint8_t* __restrict__ aux, long long n, double* partial, BicgState* st, double* history, unsigned int* counter,
                                                                 const __grid_constant__ PeerScalarsDev A) {
  if (st->done) return;
  double d[5] = {0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!aux || !aux[i]) { const double rv = r[i], sv = s[i], qv = rstar[i]; d[0] = fma(rv, sv, d[0]); d[1] = fma(rv, rv, d[1]); d[2] = fma(sv, sv, d[2]); d[3] = fma(sv, qv, d[3]); d[4] = fma(rv, qv, d[4]); }
#pragma unroll
  for (int k = 0; k < 5; ++k) { const double t = block_sum(d[k]); if (threadIdx.x == 0) partial[(size_t)k * gridDim.x + blockIdx.x] = t; }
  if (!counter || !last_block_done(counter)) return;
  __shared__ double gd[kArMax];
#pragma unroll
  for (int k = 0; k < 5; ++k) { const double t = final_sum(partial + (size_t)k * gridDim.x); if (threadIdx.x == 0) gd[k] = t; }
  __syncthreads();
  if (A.world > 1) peer_allreduce_block(A, gd, 5);
  if (threadIdx.x == 0) { bicg_scalars(gd, st, history); *counter = 0; }
}
__global__ void bicg_scalars_kernel(const double* __restrict__ sums, BicgState* st, double* history) { if (threadIdx.x == 0 && blockIdx.x == 0 && !st->done) bicg_scalars(sums, st, history); }
// x += alpha p ; x += omega s ; unless the iteration was the last one: r = s - omega r ; p = r + beta (p - omega tmp)
// (bicgstab.hh:185-199).  Runs once per executed iteration (x_applied lags the iteration counter by one while pending).
__global__ void __launch_bounds__(kRedThreads) bicg_update_kernel(double* __restrict__ x, double* __restrict__ r, double* __restrict__ p, const double* __restrict__ s,
                                                                  const double* __restrict__ tmp, long long n, BicgState* st, unsigned int* counter) {
  if (st->x_applied == st->iterations) return;
  const double alpha = st->alpha, omega = st->omega, beta = st->beta; const bool last = st->done != 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double sv = s[i], pv = p[i];
    x[i] = fma(omega, sv, fma(alpha, pv, x[i]));
    if (!last) { const double rv = fma(-omega, r[i], sv); r[i] = rv; p[i] = fma(-omega * beta, tmp[i], beta * pv) + rv; }
  }
  if (last_block_done(counter) && threadIdx.x == 0) { st->x_applied = st->iterations; *counter = 0; }
}

// ---- GMRES (solver/linear/gmres.hh:117-301) vector work; the (m+1) x m Hessenberg matrix, the Givens rotations and the
// back substitution are O(m^2) scalars and stay on the host like in the reference ----
// y[l] = <vjp, v_l> over primary dofs for up to kGemvChunk basis vectors in ONE sweep over vjp (gemv, gmres.hh:64-92);
// partial[l * gridDim.x + block]
__global__ void __launch_bounds__(kRedThreads) gmres_gemv_kernel(const double* __restrict__ vjp, const GmresVecs V, int count, const uint8_t* __restrict__ aux,
                                                                 long long n, double* __restrict__ partial) {
  double d[kGemvChunk];
#pragma unroll
  for (int l = 0; l < kGemvChunk; ++l) d[l] = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (aux && aux[i]) continue;
    const double x = vjp[i];
#pragma unroll
    for (int l = 0; l < kGemvChunk; ++l) if (l < count) d[l] = fma(x, V.v[l][i], d[l]);
  }
#pragma unroll
  for (int l = 0; l < kGemvChunk; ++l) { const double t = block_sum(d[l]); if (threadIdx.x == 0 && l < count) partial[(size_t)l * gridDim.x + blockIdx.x] = t; }
}
// y += sign * sum_l coef[l] v_l for up to kGemvChunk vectors, coefficients read from device memory, applied in order
// (vjp.axpy(-global_dot[l], v[l]) for l = 0..j, gmres.hh:214-217; u.axpy(y[i], v[i]), :289-292)
__global__ void gmres_axpys_kernel(double* __restrict__ y, const GmresVecs V, int count, const double* __restrict__ coef, double sign, long long n) {
  double c[kGemvChunk];
#pragma unroll
  for (int l = 0; l < kGemvChunk; ++l) c[l] = l < count ? sign * coef[l] : 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = y[i];
#pragma unroll
    for (int l = 0; l < kGemvChunk; ++l) if (l < count) s = fma(c[l], V.v[l][i], s);
    y[i] = s;
  }
}
__global__ void scale_kernel(double* __restrict__ x, double a, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= a;
}

// ---- AutomaticDifferenceLinearOperator (operator/common/automaticdifferenceoperator.hh:124-149): dest = (L[u + eps arg] - L[u]) / eps ----
// eps = eps_given > 0 ? eps_given : sqrt((1 + |u|) macheps / |arg|^2)   (|arg|^2 = sums[0], globally reduced)
__global__ void fd_eps_kernel(const double* __restrict__ sums, FdState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double me = 2.220446049250313e-16;
    st->eps = st->eps_given > 0 ? st->eps_given : (sums[0] > me ? sqrt((1.0 + st->norm_u) * me / sums[0]) : sqrt(me));
  }
}
// b = u + eps arg
__global__ void fd_perturb_kernel(double* __restrict__ b, const double* __restrict__ u, const double* __restrict__ arg, long long n, const FdState* st) {
  const double eps = st->eps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) b[i] = fma(eps, arg[i], u[i]);
}
// dest = (dest - op_u) * (1 / eps)
__global__ void fd_quotient_kernel(double* __restrict__ dest, const double* __restrict__ op_u, long long n, const FdState* st) {
  const double inv = 1.0 / st->eps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dest[i] = (dest[i] - op_u[i]) * inv;
}

// ---- Jacobi-preconditioned CG: the preconditioned branch of LinearSolver::cg (cg.hh:52-56, 72-107) with B = diag(A)^-1 ----
// after h = A x:  p = b - h ; q = s = B p ; partial <p,q> and <b,b>
__global__ void __launch_bounds__(kRedThreads) pcg_init_kernel(const double* __restrict__ h, const double* __restrict__ b, const double* __restrict__ dinv,
                                                               double* __restrict__ p, double* __restrict__ q, double* __restrict__ s_, const uint8_t* __restrict__ aux,
                                                               long long n, double* __restrict__ partial, double* __restrict__ partial_b) {
  double s = 0, sb = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double bv = b[i], pv = bv - h[i], qv = dinv[i] * pv;
    p[i] = pv; q[i] = qv; s_[i] = qv;
    if (!aux || !aux[i]) { s = fma(pv, qv, s); sb = fma(bv, bv, sb); }
  }
  s = block_sum(s); if (threadIdx.x == 0) partial[blockIdx.x] = s;
  sb = block_sum(sb); if (threadIdx.x == 0) partial_b[blockIdx.x] = sb;
}
// q <- beta q + s   (cg.hh:76-80)
__global__ void pcg_update_q_kernel(double* __restrict__ q, const double* __restrict__ s_, long long n, const CgState* st) {
  if (st->done || st->iterations == 0) return;
  const double beta = st->residual / st->prev_residual;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) q[i] = q[i] * beta + s_[i];
}
// x += alpha q ; p -= alpha h ; s = B p ; partial <p,s>   (cg.hh:92-101); the last block closes the iteration
__global__ void __launch_bounds__(kRedThreads) pcg_update_kernel(double* __restrict__ x, double* __restrict__ p, double* __restrict__ s_, const double* __restrict__ q,
                                                                 const double* __restrict__ h, const double* __restrict__ dinv, const uint8_t* __restrict__ aux, long long n,
                                                                 double* partial, CgState* st, double* __restrict__ history, unsigned int* counter,
                                                                 const __grid_constant__ PeerScalarsDev A) {
  if (st->done) return;
  const double alpha = st->alpha;
  double s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, q[i], x[i]);
    const double pv = fma(-alpha, h[i], p[i]), sv = dinv[i] * pv;
    p[i] = pv; s_[i] = sv;
    if (!aux || !aux[i]) s = fma(pv, sv, s);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (!counter || !last_block_done(counter)) return;
  const double t = global_sum1(A, final_sum(partial));
  if (threadIdx.x == 0) {
    st->prev_residual = st->residual; st->residual = t;
    if (history) history[st->iterations] = sqrt(t);
    st->iterations += 1;
    if (!(st->residual > st->tolerance) || st->iterations >= st->max_iterations) st->done = 1;
    *counter = 0;
  }
}
__global__ void set_masked_kernel(double* __restrict__ d, const uint8_t* __restrict__ mask, double value, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) if (mask[i]) d[i] = value;
}
__global__ void invert_kernel(double* __restrict__ d, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) d[i] = 1.0 / d[i];
}

// strong Dirichlet rows: w_d = u_d - g_d   (schemes/dirichletwrapper.hh:101-105; Operation::sub)
__global__ void dirichlet_sub_kernel(const double* __restrict__ u, double* __restrict__ w, const uint8_t* __restrict__ mask,
                                     const double* __restrict__ g, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (mask[i]) w[i] = u[i] - (g ? g[i] : 0.0);
}

}  // namespace b200fem
