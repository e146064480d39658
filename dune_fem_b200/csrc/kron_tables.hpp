// kron_tables.hpp -- 1-D operator matrices of the Kronecker (constant-coefficient, uniform box) form of the
// SIPG/upwind advection-diffusion-reaction operator.  Built from the same 1-D tabulations and Gauss weights the
// quadrature kernel uses, i.e. every entry is the quadrature sum the reference would compute
// (dune/fem/schemes/galerkin.hh:332-360, 475-537 with the integrands of pydemo/advectiondiffusion.py:33-60),
// factored along the axes.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "../../include/b200fem.h"
#include "tables.hpp"

namespace b200fem {

struct KronHost {
  int n = 0;
  std::vector<double> S[3], Dlo[3], Dhi[3], L[3], R[3];   // n*n each, row = test index, col = trial index
};

// `scale` multiplies every matrix (inverse mass of MOLGalerkinOperator: 1 / detJ)
inline KronHost build_kron_tables(const Tab1D& t, const b200fem_model& m, int dim, const double* h, double scale = 1.0) {
  KronHost k; const int n = t.n; k.n = n;
  double detJ = 1; for (int d = 0; d < dim; ++d) detJ *= h[d];
  // 1-D quadrature matrices
  std::vector<double> K1(n * n, 0.0), C1(n * n, 0.0);
  for (int q = 0; q < t.m; ++q) for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
    K1[i * n + j] += t.w[q] * t.G[q * n + i] * t.G[q * n + j];
    C1[i * n + j] += t.w[q] * t.G[q * n + i] * t.B[q * n + j];
  }
  for (int d = 0; d < 3; ++d) {
    k.S[d].assign(n * n, 0.0); k.Dlo[d].assign(n * n, 0.0); k.Dhi[d].assign(n * n, 0.0); k.L[d].assign(n * n, 0.0); k.R[d].assign(n * n, 0.0);
    if (d >= dim) continue;
    const double hd = h[d], area = detJ / hd;
    const double pen = m.eps * m.beta / hd, e2 = m.eps / (2 * hd);
    const double hbp = std::max(m.b[d], 0.0), hbm = std::max(-m.b[d], 0.0);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
      const double p0i = t.phi0[i], p1i = t.phi1[i], d0i = t.dphi0[i], d1i = t.dphi1[i];
      const double p0j = t.phi0[j], p1j = t.phi1[j], d0j = t.dphi0[j], d1j = t.dphi1[j];
      double vol = detJ * (m.eps / (hd * hd) * K1[i * n + j] - m.b[d] / hd * C1[i * n + j]);
      if (d == 0 && i == j) vol += m.c * detJ;
      double HH = 0, LL = 0, Lm = 0, Rm = 0, BH = 0, BL = 0;
      if (m.has_skeleton) {
        HH = area * ((pen + hbp) * p1i * p1j - e2 * p1i * d1j - e2 * d1i * p1j);
        Rm = area * (-(pen + hbm) * p1i * p0j - e2 * p1i * d0j + e2 * d1i * p0j);
        Lm = area * (-(pen + hbp) * p0i * p1j + e2 * p0i * d1j - e2 * d0i * p1j);
        LL = area * ((pen + hbm) * p0i * p0j + e2 * p0i * d0j + e2 * d0i * p0j);
      }
      if (m.has_boundary) {
        if ((m.dirichlet_mask >> (2 * d + 1)) & 1) BH = area * (pen + hbp) * p1i * p1j;
        if ((m.dirichlet_mask >> (2 * d)) & 1)     BL = area * (pen + hbm) * p0i * p0j;
      }
      k.S[d][i * n + j] = scale * (vol + HH + LL);
      k.Dhi[d][i * n + j] = scale * (BH - HH);
      k.Dlo[d][i * n + j] = scale * (BL - LL);
      k.L[d][i * n + j] = scale * Lm;
      k.R[d][i * n + j] = scale * Rm;
    }
  }
  return k;
}


// 1-D assembled row tables of the Lagrange Kronecker form (lagrange_kronecker.cuh): for axis d, rows[g*(2k+1) + j] is
// the entry (g, g-k+j) of  M_d = assembled mass  and  T_d = eps K_d - b_d C_d (+ c M_0 on axis 0) (+ the u-dependent
// boundary term (eps beta / h_d + hat b) on the end nodes of masked DOMAIN boundaries).  Element matrices are the
// 1-D quadrature sums of the reference loop (galerkin.hh:332-360) with the tabulation `t` of the space:
//   Me[i][j] = h sum_q w_q B_qi B_qj,   Ke[i][j] = 1/h sum_q w_q G_qi G_qj,   Ce[i][j] = sum_q w_q G_qi B_qj
// (i test, j trial; F = eps grad u - b u is tested with grad phi_i).  Axes >= dim get M = [1], T = [0].
struct LagRowsHost { std::vector<double> M[3], T[3]; int k = 0; long long L[3] = {1, 1, 1}; };

inline LagRowsHost build_lagrange_rows(const Tab1D& t, const b200fem_model& m, int dim, int order, const int* n, const int* origin, const int* gn, const double* h) {
  LagRowsHost r; const int k = order, W = 2 * k + 1, nb = k + 1; r.k = k;
  for (int d = 0; d < 3; ++d) {
    const long long Ld = d < dim ? (long long)k * n[d] + 1 : 1; r.L[d] = Ld;
    r.M[d].assign((size_t)Ld * W, 0.0); r.T[d].assign((size_t)Ld * W, 0.0);
    if (d >= dim) { r.M[d][k] = 1.0; continue; }
    std::vector<double> Me(nb * nb, 0.0), Te(nb * nb, 0.0);
    for (int q = 0; q < t.m; ++q) for (int i = 0; i < nb; ++i) for (int j = 0; j < nb; ++j) {
      const double mm = h[d] * t.w[q] * t.B[q * nb + i] * t.B[q * nb + j];
      Me[i * nb + j] += mm;
      Te[i * nb + j] += m.eps / h[d] * t.w[q] * t.G[q * nb + i] * t.G[q * nb + j] - m.b[d] * t.w[q] * t.G[q * nb + i] * t.B[q * nb + j];
      if (d == 0) Te[i * nb + j] += m.c * mm;
    }
    for (int e = 0; e < n[d]; ++e) for (int i = 0; i < nb; ++i) for (int j = 0; j < nb; ++j) {
      const long long g = (long long)k * e + i; const int col = j - i + k;          // column g - k + col = k e + j
      r.M[d][(size_t)g * W + col] += Me[i * nb + j]; r.T[d][(size_t)g * W + col] += Te[i * nb + j];
    }
    if (m.has_boundary) for (int side = 0; side < 2; ++side) {
      const bool domain_bnd = side == 0 ? origin[d] == 0 : origin[d] + n[d] == gn[d];
      if (!domain_bnd || !((m.dirichlet_mask >> (2 * d + side)) & 1)) continue;
      const double bn = m.b[d] * (side ? 1.0 : -1.0), hatb = 0.5 * (bn + std::fabs(bn));
      const long long g = side ? Ld - 1 : 0;
      r.T[d][(size_t)g * W + k] += m.eps * m.beta / h[d] + hatb;
    }
  }
  return r;
}


// The same operators in the two-row-type form of lagrange_lattice.cuh: an assembled row depends only on the node type (element
// vertex: contributions of the element to the left, local node k, and to the right, local node 0; element-interior node,
// k = 2: one element) -- except on the first / last lattice plane of the local box, where one element contribution is missing.
// Nodes outside the box read as zero, so those rows are the interior row plus a correction of the diagonal entry:
//   lo: -Xe[k][k] (no left element),  hi: -Xe[0][0] (no right element),  T additionally + the boundary term on masked DOMAIN sides.
// Offsets -k..k sit in slots 0..2k.  Axes >= dim: M = identity (centre 1), T = 0.
struct LagStencilHost { int k = 0; double M[3][2][5] = {}, T[3][2][5] = {}, Mlo[3] = {}, Mhi[3] = {}, Tlo[3] = {}, Thi[3] = {}; };

inline LagStencilHost build_lagrange_stencil(const Tab1D& t, const b200fem_model& m, int dim, int order, const int* n, const int* origin, const int* gn, const double* h) {
  LagStencilHost r; const int k = order, nb = k + 1; r.k = k;
  for (int d = 0; d < 3; ++d) {
    if (d >= dim) { r.M[d][0][k] = 1.0; r.M[d][1][k] = 1.0; continue; }
    std::vector<double> Me(nb * nb, 0.0), Te(nb * nb, 0.0);
    for (int q = 0; q < t.m; ++q) for (int i = 0; i < nb; ++i) for (int j = 0; j < nb; ++j) {
      const double mm = h[d] * t.w[q] * t.B[q * nb + i] * t.B[q * nb + j];
      Me[i * nb + j] += mm;
      Te[i * nb + j] += m.eps / h[d] * t.w[q] * t.G[q * nb + i] * t.G[q * nb + j] - m.b[d] * t.w[q] * t.G[q * nb + i] * t.B[q * nb + j];
      if (d == 0) Te[i * nb + j] += m.c * mm;
    }
    for (int j = 0; j < nb; ++j) {               // vertex row: left element (test node k) on offsets -k..0, right element (test node 0) on 0..k
      r.M[d][0][j] += Me[k * nb + j]; r.T[d][0][j] += Te[k * nb + j];
      r.M[d][0][k + j] += Me[0 * nb + j]; r.T[d][0][k + j] += Te[0 * nb + j];
    }
    if (k == 2) for (int j = 0; j < nb; ++j) { r.M[d][1][1 + j] = Me[1 * nb + j]; r.T[d][1][1 + j] = Te[1 * nb + j]; }   // interior node: offsets -1..1
    r.Mlo[d] = -Me[k * nb + k]; r.Mhi[d] = -Me[0]; r.Tlo[d] = -Te[k * nb + k]; r.Thi[d] = -Te[0];
    if (m.has_boundary) for (int side = 0; side < 2; ++side) {
      const bool domain_bnd = side == 0 ? origin[d] == 0 : origin[d] + n[d] == gn[d];
      if (!domain_bnd || !((m.dirichlet_mask >> (2 * d + side)) & 1)) continue;
      const double bn = m.b[d] * (side ? 1.0 : -1.0), hatb = 0.5 * (bn + std::fabs(bn));
      (side ? r.Thi[d] : r.Tlo[d]) += m.eps * m.beta / h[d] + hatb;
    }
  }
  return r;
}

}  // namespace b200fem
