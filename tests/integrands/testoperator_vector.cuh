// The form of the reference's own matrix-free operator check on a vector-valued space (dune/fempy/test/testoperator.py:33-34):
//   a = ( inner(0.5*dot(u,u), v[0]) + inner(u[0]*grad(u), grad(v)) ) * dx
// written over VectorValue / VectorRange for any dimRange.
__device__ void interior(const double* x, const VectorValue& u, VectorRange& r, const double* c, int dim) {
  double uu = 0;
  for (int k = 0; k < dimRange; ++k) uu += u.u[k] * u.u[k];
  r.s[0] = 0.5 * uu;
  for (int k = 0; k < dimRange; ++k)
    for (int d = 0; d < dim; ++d) r.F[k][d] = u.u[0] * u.du[k][d];
}
