// A coupled non-linear reaction-diffusion-advection SYSTEM (dimRange components) in SIPG form, variable coefficients:
//   -div(k(x) grad u_k + c[1] grad u_{k+1}) + div(b u_k) + c[2] u_k u_{k+1} = f_k,   k(x) = c[0] (1 + 0.5 sin(x0) cos(x1)),  b = (1, -0.5, 0.25)
// penalty c[3] * k / he on the jumps, upwinding, weak Dirichlet data g_k = sin(x0 x1 + k) on every side, f_k = 1 + k + x0.
// Components couple in the interior flux, the reaction and the skeleton consistency term.
__device__ inline double kdiff(const double* x, const double* c) { return c[0] * (1.0 + 0.5 * sin(x[0]) * cos(x[1])); }
__device__ inline double bvel(int d) { return d == 0 ? 1.0 : d == 1 ? -0.5 : 0.25; }

__device__ void interior(const double* x, const VectorValue& u, VectorRange& r, const double* c, int dim) {
  const double k = kdiff(x, c);
  for (int i = 0; i < dimRange; ++i) {
    const int j = (i + 1) % dimRange;
    r.s[i] = c[2] * u.u[i] * u.u[j] - (1.0 + i + x[0]);
    for (int d = 0; d < dim; ++d) r.F[i][d] = k * u.du[i][d] + c[1] * u.du[j][d] - bvel(d) * u.u[i];
  }
}

__device__ void skeleton(const double* x, int axis, double sign, double ihe, const VectorValue& in, const VectorValue& out,
                         VectorRange& rin, VectorRange& rout, const double* c, int dim) {
  const double k = kdiff(x, c), bn = bvel(axis) * sign;
  for (int i = 0; i < dimRange; ++i) {
    const int j = (i + 1) % dimRange;
    const double jump = in.u[i] - out.u[i];
    const double avg_dn = 0.5 * (k * (in.du[i][axis] + out.du[i][axis]) + c[1] * (in.du[j][axis] + out.du[j][axis])) * sign;
    const double flux = bn > 0 ? bn * in.u[i] : bn * out.u[i];
    const double cj = k * c[3] * ihe * jump - avg_dn + flux;
    rin.s[i] = cj; rout.s[i] = -cj;
    rin.F[i][axis] = rout.F[i][axis] = -k * jump * 0.5 * sign;
  }
}

__device__ void boundary(const double* x, int axis, int side, double ihbnd, const VectorValue& u, VectorRange& r, const double* c, int dim) {
  const double sign = side ? 1.0 : -1.0, k = kdiff(x, c), bn = bvel(axis) * sign;
  for (int i = 0; i < dimRange; ++i) {
    const int j = (i + 1) % dimRange;
    const double g = sin(x[0] * x[1] + i);
    r.s[i] = k * c[3] * ihbnd * (u.u[i] - g) - (k * u.du[i][axis] + c[1] * u.du[j][axis]) * sign + (bn > 0 ? bn * u.u[i] : bn * g);
    r.F[i][axis] = -k * (u.u[i] - g) * sign;
  }
}
