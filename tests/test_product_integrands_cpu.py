"""The PRODUCT's built-in integrands (dune_fem_b200/csrc/integrands.cuh: interior / skeleton / boundary of the advection-diffusion-
reaction family, the code the generic quadrature kernels are instantiated with) checked on the CPU: the header is compiled for the
host (g++, __device__ defined away -- the way the oracle compiles user-supplied integrand text) and integrated by the oracle's
restatement of the reference loop, against the oracle's own integrands (fem_oracle.cpp, written from pydemo/advectiondiffusion.py:33-60
independently of the product).  Covers data terms, upwinding in both directions, weak Dirichlet / Neumann sides and the cubic reaction."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "dune_fem_b200", "csrc", "integrands.cuh")

# the oracle wraps user text in `namespace user { ... }`: leave it for the include, re-enter it for the three entry points
SOURCE = """}
// integrands.cuh sha1 %(sha)s
#include "%(abi)s"          /* first, as a host compiler sees it (the header has its own NVRTC branch) */
#define __CUDACC_RTC__      /* makes integrands.cuh skip <cuda_runtime.h>: host build */
#define __host__
#define __noinline__
#include "%(header)s"
#undef __CUDACC_RTC__
namespace user {
static b200fem::AdrIntegrandsT<true> model(const double* c, int dim) {
  b200fem::AdrIntegrandsT<true> I{};
  I.m.eps = c[0]; I.m.b[0] = c[1]; I.m.b[1] = c[2]; I.m.b[2] = c[3]; I.m.c = c[4]; I.m.gamma = c[5]; I.m.beta = c[6];
  I.m.dirichlet_mask = (int)c[7]; I.m.data = (int)c[8]; I.m.has_skeleton = 1; I.m.has_boundary = 1; I.dim = dim; I.with_data = true;
  return I;
}
static b200fem::PointValue in(const PointValue& u) { b200fem::PointValue v; v.u = u.u; for (int d = 0; d < 3; ++d) v.du[d] = u.du[d]; return v; }
static void out(const b200fem::PointRange& a, PointRange& r) { r.s = a.s; for (int d = 0; d < 3; ++d) r.F[d] = a.F[d]; }
void interior(const double* x, const PointValue& u, PointRange& r, const double* c, int dim) { out(model(c, dim).interior(x, in(u)), r); }
void skeleton(const double* x, int axis, double sign, double ihe, const PointValue& ui, const PointValue& uo, PointRange& ri, PointRange& ro, const double* c, int dim) {
  b200fem::PointRange a, b; model(c, dim).skeleton(x, axis, sign, ihe, in(ui), in(uo), a, b); out(a, ri); out(b, ro);
}
void boundary(const double* x, int axis, int side, double ihbnd, const PointValue& u, PointRange& r, const double* c, int dim) {
  out(model(c, dim).boundary(axis, side, ihbnd, x, in(u)), r);
}
""" % {"header": HEADER, "abi": os.path.join(ROOT, "include", "b200fem.h"), "sha": hashlib.sha1(open(HEADER, "rb").read()).hexdigest()}


@pytest.mark.parametrize("dim,order,n,b,mask,data,gamma", [
    (3, 2, [3, 3, 2], (1.0, -0.5, 0.25), 0b000011, 1, 0.0),        # the pydemo's form: upwind both ways, Dirichlet on two sides, Neumann elsewhere
    (3, 1, [3, 2, 3], (-0.7, 0.4, -1.0), 0b111111, 2, 0.0),        # Dirichlet everywhere, product-of-sines data
    (3, 2, [2, 3, 2], (0.3, 0.0, -0.2), 0b100100, 1, 2.5),         # cubic reaction
    (2, 2, [5, 4], (1.0, 0.6, 0.0), 0b0110, 1, 1.5),               # 2-D
    (3, 1, [3, 3, 3], (0.0, 0.0, 0.0), 0, 0, 0.0),                 # pure diffusion-reaction, no data, Neumann
])
def test_builtin_integrands_of_the_product_match_the_oracle(dim, order, n, b, mask, data, gamma):
    lo, hi = [-1.0, 0.0, 0.5][:dim], [1.0, 0.5, 2.0][:dim]
    sp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, order)
    eps, c, beta = 0.3, 0.7, 12.0 * order * order
    own = ol.Operator(sp, eps=eps, b=b, c=c, gamma=gamma, beta=beta, dirichlet_mask=mask, data=data, skeleton=True, boundary=True)
    prod = ol.UserOperator(sp, SOURCE, [eps, *b, c, gamma, beta, mask, data])
    u = np.random.default_rng(dim * 10 + order).uniform(-1, 1, sp.size)
    for v in (u, np.zeros(sp.size)):                                # L[u] and the load vector -L[0]
        w_own, w_prod = own.apply(v), prod.apply(v)
        assert np.abs(w_prod - w_own).max() <= 1e-14 * max(np.abs(w_own).max(), 1.0)
