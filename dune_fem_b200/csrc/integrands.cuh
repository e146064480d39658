// integrands.cuh -- device integrands of the advection-diffusion-reaction family.
//
// Mirrors the Integrands interface of the reference (dune/fem/schemes/integrands.hh:152-375):
//   interior(x, (u, grad u))                       -> (s, F)         tested as  s*phi_i + F.grad phi_i
//   skeleton(xIn, (u,grad u)_in, xOut, (u,grad u)_out) -> ((s,F)_in, (s,F)_out)
//   boundary(x, (u, grad u))                       -> (s, F)
// The reference JIT-generates such a class from UFL (python/dune/models/integrands/model.py:72-106); the form
// implemented here is pydemo/advectiondiffusion.py:33-60 (SIPG + upwind, weak Dirichlet on masked sides, Neumann
// data elsewhere) extended by a reaction term c u + gamma u^3.  The generic quadrature kernel is templated on
// this struct, so another integrand family is another struct with the same three members.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#include "../../include/b200fem.h"

namespace b200fem {

struct PointValue { double u; double du[3]; };   // DomainValueType  = tuple<RangeType, JacobianRangeType>
struct PointRange { double s; double F[3]; };    // RangeValueType

// kData = false compiles the data terms out (g = f = 0): the apply kernels of the generic path take the affine part from
// the load vector built once per operator by the kData = true instantiation, so their code carries no sincos expansion.
template <bool kData>
struct AdrIntegrandsT {
  b200fem_model m;
  int dim;
  bool with_data;    // false: homogeneous part (g = f = 0)

  __device__ bool has_data() const { if constexpr (kData) return with_data && m.data; else return false; }

  __host__ __device__ bool linear() const { return m.gamma == 0.0; }

  // (not inlined: the analytic data are only evaluated when the load vector is built, once per operator -- the apply kernels
  // carry one call site instead of a sincos expansion per quadrature point)
  __device__ __noinline__ void data(const double* x, double& g, double dg[3], double& lap) const {
    g = 0; lap = 0; dg[0] = dg[1] = dg[2] = 0;
    if (!with_data) return;
    if (m.data == 1) {
      double sn, cs; sincos(x[0] * x[1], &sn, &cs);
      g = sn; dg[0] = x[1] * cs; dg[1] = x[0] * cs; lap = -(x[0] * x[0] + x[1] * x[1]) * sn;
    } else if (m.data == 2) {
      const double pi = 3.14159265358979323846;
      double sn[3] = {1, 1, 1}, cs[3] = {1, 1, 1};
      for (int d = 0; d < dim; ++d) sincos(pi * x[d], &sn[d], &cs[d]);
      g = sn[0] * sn[1] * sn[2];
      for (int d = 0; d < dim; ++d) { double v = pi * cs[d]; for (int k = 0; k < dim; ++k) if (k != d) v *= sn[k]; dg[d] = v; }
      lap = -dim * pi * pi * g;
    }
  }

  __device__ PointRange interior(const double* x, const PointValue& v) const {
    PointRange r; double f = 0;
    if (has_data()) {
      double g, dg[3], lap; data(x, g, dg, lap);
      f = -m.eps * lap + m.c * g + m.gamma * g * g * g;
      for (int d = 0; d < dim; ++d) f += m.b[d] * dg[d];
    }
    r.s = m.c * v.u + m.gamma * v.u * v.u * v.u - f;
    for (int d = 0; d < 3; ++d) r.F[d] = (d < dim) ? m.eps * v.du[d] - m.b[d] * v.u : 0.0;
    return r;
  }

  // unit outer normal of the inside element = sign * e_axis; ihe = 1 / he, he = avg(CellVolume)/FacetArea (the caller
  // passes the reciprocal: one division per face instead of one per quadrature point)
  // (the axis may be a run-time value: it selects components, it never indexes a register array)
  // (x: the face point -- unused by this constant-coefficient family, part of the interface for generic integrands)
  __device__ void skeleton(const double* x, int axis, double sign, double ihe, const PointValue& in, const PointValue& out,
                           PointRange& rin, PointRange& rout) const {
    const double jump = in.u - out.u;
    const double in_dn = axis == 0 ? in.du[0] : axis == 1 ? in.du[1] : in.du[2], out_dn = axis == 0 ? out.du[0] : axis == 1 ? out.du[1] : out.du[2];
    const double avg_dn = 0.5 * (in_dn + out_dn) * sign;
    const double bn = m.b[axis] * sign;
    const double hat_in = 0.5 * (bn + fabs(bn)), hat_out = 0.5 * (-bn + fabs(bn));
    const double cj = m.eps * m.beta * ihe * jump - m.eps * avg_dn + (hat_in * in.u - hat_out * out.u);
    rin.s = cj; rout.s = -cj;
    const double fn = -m.eps * jump * 0.5 * sign;
    rin.F[0] = rout.F[0] = axis == 0 ? fn : 0.0; rin.F[1] = rout.F[1] = axis == 1 ? fn : 0.0; rin.F[2] = rout.F[2] = axis == 2 ? fn : 0.0;
  }

  __device__ PointRange boundary(int axis, int side, double ihbnd, const double* x, const PointValue& v) const {
    PointRange r; r.F[0] = r.F[1] = r.F[2] = 0;
    const double sign = side ? 1.0 : -1.0;
    double g = 0, dg[3] = {0, 0, 0}, lap = 0;
    if (has_data()) data(x, g, dg, lap);
    r.s = -m.eps * (axis == 0 ? dg[0] : axis == 1 ? dg[1] : dg[2]) * sign;
    if ((m.dirichlet_mask >> (2 * axis + side)) & 1) {
      const double bn = m.b[axis] * sign, hatb = 0.5 * (bn + fabs(bn));
      r.s += m.eps * m.beta * ihbnd * (v.u - g) + hatb * v.u + (bn - hatb) * g;
    }
    return r;
  }
};
using AdrIntegrands = AdrIntegrandsT<true>;       // data terms selected at run time (with_data)
using AdrIntegrandsHom = AdrIntegrandsT<false>;   // homogeneous part only

}  // namespace b200fem
