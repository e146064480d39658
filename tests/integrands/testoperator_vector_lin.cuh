// dune/fempy/test/testoperator.py:59-60: the same form with the first factor frozen at ubar = dot(x,x) in every component,
//   lina = ( inner(0.5*dot(ubar,u), v[0]) + inner(ubar[0]*grad(u), grad(v)) ) * dx
// (ubar is the P2 interpolant of a quadratic: the pointwise value is exact)
__device__ void interior(const double* x, const VectorValue& u, VectorRange& r, const double* c, int dim) {
  double ubar = 0, su = 0;
  for (int d = 0; d < dim; ++d) ubar += x[d] * x[d];
  for (int k = 0; k < dimRange; ++k) su += u.u[k];
  r.s[0] = 0.5 * ubar * su;
  for (int k = 0; k < dimRange; ++k)
    for (int d = 0; d < dim; ++d) r.F[k][d] = ubar * u.du[k][d];
}
