// Test integrands for the run-time compiled path: advection-diffusion-reaction with VARIABLE coefficients and a non-linear
// reaction -- a form the built-in constant-coefficient family (integrands.cuh) cannot express:
//   -div(k(x) grad u) + div(b(x) u) + c0 u + c1 u^3 = f,   k(x) = c[0] (1 + 0.5 sin(x0) cos(x1)),  b(x) = (c[1] + x1, c[2] - x0, c[3])
// SIPG with penalty c[4] * k / he, upwinding, weak Dirichlet data g = sin(x0 x1) on every boundary side, f = 1 + x0.
// The same text is compiled by NVRTC into the device kernel and by g++ (with __device__ defined away) into the callbacks the CPU
// oracle integrates -- see tests/test_gpu_jit.py.
__device__ inline double kdiff(const double* x, const double* c) { return c[0] * (1.0 + 0.5 * sin(x[0]) * cos(x[1])); }
__device__ inline double bvel(const double* x, const double* c, int d) { return d == 0 ? c[1] + x[1] : d == 1 ? c[2] - x[0] : c[3]; }

__device__ void interior(const double* x, const PointValue& u, PointRange& r, const double* c, int dim) {
  const double k = kdiff(x, c);
  r.s = c[5] * u.u + c[6] * u.u * u.u * u.u - (1.0 + x[0]);
  for (int d = 0; d < dim; ++d) r.F[d] = k * u.du[d] - bvel(x, c, d) * u.u;
}

__device__ void skeleton(const double* x, int axis, double sign, double ihe, const PointValue& in, const PointValue& out,
                         PointRange& rin, PointRange& rout, const double* c, int dim) {
  const double k = kdiff(x, c), jump = in.u - out.u;
  const double avg_dn = 0.5 * (in.du[axis] + out.du[axis]) * sign;
  const double bn = bvel(x, c, axis) * sign;
  const double flux = bn > 0 ? bn * in.u : bn * out.u;          // upwind
  const double cj = k * c[4] * ihe * jump - k * avg_dn + flux;
  rin.s = cj; rout.s = -cj;
  rin.F[axis] = rout.F[axis] = -k * jump * 0.5 * sign;
}

__device__ void boundary(const double* x, int axis, int side, double ihbnd, const PointValue& u, PointRange& r, const double* c, int dim) {
  const double sign = side ? 1.0 : -1.0, k = kdiff(x, c), g = sin(x[0] * x[1]);
  const double bn = bvel(x, c, axis) * sign;
  r.s = k * c[4] * ihbnd * (u.u - g) - k * u.du[axis] * sign + (bn > 0 ? bn * u.u : bn * g);
  r.F[axis] = -k * (u.u - g) * sign;
}
