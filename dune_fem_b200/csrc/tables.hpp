// tables.hpp -- host-side 1-D building blocks of the product library: Gauss rules, Legendre / Lagrange
// 1-D bases and their tabulations.  Everything the device kernels need about a space is a handful of small
// 1-D arrays (tensor-product structure of cube elements), built once per space on the host.
//
// Reference semantics restated here (paths under /root/reference/dune/fem):
//   Gauss rules on [0,1], rule choice 2m-1 >= order ... quadrature/gausspoints.hh:107-128, femquadratures_inline.hh:59-70
//   orthonormal Legendre P_n on [0,1], Horner form ... space/shapefunctionset/legendrepolynomials.hh:24-46
//   Legendre multi-index order / hierarchical sort .. space/shapefunctionset/legendre.hh:169-194, 236-250
//   Lagrange Q_k equidistant nodal basis, x fastest . space/lagrange/genericlagrangepoints.hh:862-876
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace b200fem {

struct Rule1D { std::vector<double> x, w; };

// m-point Gauss-Legendre rule on [0,1], ascending points.  Golub-Welsch would do as well; Newton on the
// three-term recurrence in extended precision reproduces the reference's 70-digit tables after rounding.
inline Rule1D gauss_rule(int m) {
  Rule1D r; r.x.assign(m, 0.0); r.w.assign(m, 0.0);
  const long double pi = acosl(-1.0L);
  for (int k = 0; k < (m + 1) / 2; ++k) {
    long double t = cosl(pi * (k + 0.75L) / (m + 0.5L)), dp = 1;
    for (int sweep = 0; sweep < 64; ++sweep) {
      long double a = 1, b = t;                       // P_0, P_1 at t
      for (int j = 2; j <= m; ++j) { long double c = ((2 * j - 1) * t * b - (j - 1) * a) / j; a = b; b = c; }
      if (m == 1) { b = t; a = 1; }
      dp = m * (a - t * b) / (1 - t * t);            // P_m'(t)
      long double step = b / dp; t -= step;
      if (fabsl(step) < 1e-20L) break;
    }
    long double wt = 2 / ((1 - t * t) * dp * dp);
    r.x[m - 1 - k] = (double)((1 + t) / 2); r.x[k] = (double)((1 - t) / 2);
    r.w[m - 1 - k] = r.w[k] = (double)(wt / 2);
  }
  if (m & 1) r.x[m / 2] = 0.5;
  return r;
}

inline int gauss_points_for_order(int order) {
  if (order <= 0) order = 1;
  for (int m = 1; m <= 10; ++m) if (2 * m - 1 >= order) return m;
  throw std::runtime_error("quadrature order not implemented (max 19)");
}

// Shifted Legendre polynomials through their integer monomial coefficients, evaluated by Horner exactly as
// the reference does (same operation order, so tabulated values agree bit for bit with a plain x86-64 build).
class Legendre1D {
 public:
  static constexpr int kMax = 11;
  Legendre1D() {
    for (int n = 0; n < kMax; ++n) {
      scale_[n] = std::sqrt((double)(2 * n + 1));
      for (int i = 0; i < kMax; ++i) coef_[n][i] = 0;
      for (int i = 0; i <= n; ++i) {
        long double bin1 = 1, bin2 = 1;
        for (int t = 1; t <= i; ++t) { bin1 = bin1 * (n - i + t) / t; bin2 = bin2 * (n + t) / t; }
        coef_[n][i] = (double)(((n + i) & 1) ? -(bin1 * bin2) : (bin1 * bin2));
      }
    }
    coef_[10][3] = -34920.0;   // as tabulated by the reference (legendrepolynomials.cc:26); mathematically -34320
  }
  double value(int n, double x) const {
    double p = coef_[n][n];
    for (int i = n - 1; i >= 0; --i) p = p * x + coef_[n][i];
    return scale_[n] * p;
  }
  double derivative(int n, double x) const {
    double p = 0;
    if (n >= 1) { p = coef_[n][n] * n; for (int i = n - 1; i >= 1; --i) p = p * x + coef_[n][i] * i; }
    return scale_[n] * p;
  }
 private:
  double coef_[kMax][kMax], scale_[kMax];
};

// equidistant Lagrange basis of degree k on [0,1]
inline double lagrange_value(int k, int a, double x) {
  double v = 1; for (int b = 0; b <= k; ++b) if (b != a) v *= (x * k - b) / (double)(a - b); return v;
}
inline double lagrange_derivative(int k, int a, double x) {
  double s = 0;
  for (int c = 0; c <= k; ++c) if (c != a) {
    double v = (double)k / (double)(a - c);
    for (int b = 0; b <= k; ++b) if (b != a && b != c) v *= (x * k - b) / (double)(a - b);
    s += v;
  }
  return s;
}

enum class Basis { Lagrange, Legendre };

// 1-D tabulation of a basis with n functions at an m-point Gauss rule, plus end-point traces.
struct Tab1D {
  int n = 0, m = 0;
  std::vector<double> x, w;          // rule
  std::vector<double> B, G;          // B[q*n+i] = phi_i(x_q), G[q*n+i] = phi_i'(x_q)
  std::vector<double> phi0, phi1, dphi0, dphi1;   // traces at 0 and 1
};

inline Tab1D tabulate_1d(Basis basis, int order, int m) {
  static const Legendre1D leg;
  Tab1D t; t.n = order + 1; t.m = m;
  Rule1D r = gauss_rule(m); t.x = r.x; t.w = r.w;
  auto val = [&](int i, double x) { return basis == Basis::Legendre ? leg.value(i, x) : lagrange_value(order, i, x); };
  auto der = [&](int i, double x) { return basis == Basis::Legendre ? leg.derivative(i, x) : lagrange_derivative(order, i, x); };
  t.B.resize((size_t)m * t.n); t.G.resize((size_t)m * t.n);
  for (int q = 0; q < m; ++q) for (int i = 0; i < t.n; ++i) { t.B[q * t.n + i] = val(i, t.x[q]); t.G[q * t.n + i] = der(i, t.x[q]); }
  for (int i = 0; i < t.n; ++i) { t.phi0.push_back(val(i, 0.0)); t.phi1.push_back(val(i, 1.0)); t.dphi0.push_back(der(i, 0.0)); t.dphi1.push_back(der(i, 1.0)); }
  return t;
}

// local numbering of a DG Legendre space: perm[tensor index (i0*n+i1)*n+i2 ...] = stored local index
inline std::vector<int> legendre_local_permutation(int dim, int order, bool hierarchical) {
  const int n = order + 1; int nb = 1; for (int d = 0; d < dim; ++d) nb *= n;
  std::vector<std::array<int, 3>> mi(nb);
  for (int t = 0; t < nb; ++t) { int z = t; std::array<int, 3> a = {0, 0, 0}; for (int d = dim - 1; d >= 0; --d) { a[d] = z % n; z /= n; } mi[t] = a; }
  std::vector<int> order_of(nb); for (int t = 0; t < nb; ++t) order_of[t] = t;
  if (hierarchical) {
    auto key = [&](int t) { return *std::max_element(mi[t].begin(), mi[t].begin() + dim); };
    std::sort(order_of.begin(), order_of.end(), [&](int a, int b) {
      if (key(a) != key(b)) return key(a) < key(b);
      return std::lexicographical_compare(mi[a].begin(), mi[a].begin() + dim, mi[b].begin(), mi[b].begin() + dim);
    });
  }
  std::vector<int> perm(nb);
  for (int stored = 0; stored < nb; ++stored) perm[order_of[stored]] = stored;
  return perm;
}

// Number of local dofs of a DG space and the map tensor index -> stored local index over the FULL 3-D tensor basis
// (t = (m0*n + m1)*n + m2, n = order+1), -1 where the space holds no such function.  The device kernels work on the full
// tensor-product Legendre basis of the cube; the spaces are sub-bases of it:
//   Q_k Legendre in 3-D: all n^3 functions (legendre.hh:169-194, hierarchical sort :236-250)
//   Q_k Legendre in 2-D: the functions that are constant in x2 (m2 = 0); the mesh is one layer of unit-height cells
//   `dgonb` P_k (orthonormal.hh:55-60, orthonormal/orthonormalbase_{2,3}d.hh): total degree <= k, graded by degree, the
//   exponent of x0 descending first -- on cubes these ARE products of the orthonormal 1-D Legendre polynomials
inline std::vector<int> dg_tensor_map(int dim, int order, int kind /* 1 lexicographic, 2 hierarchical, 3 dgonb */, int* nb_out) {
  const int n = order + 1;
  std::vector<std::array<int, 3>> stored;
  if (kind == 3) {
    for (int p = 0; p <= order; ++p)
      for (int a = p; a >= 0; --a)
        for (int b = p - a; b >= 0; --b) {
          if (dim == 2) { if (a + b == p) stored.push_back({a, b, 0}); }
          else stored.push_back({a, b, p - a - b});
        }
  } else {
    int nb = 1; for (int d = 0; d < dim; ++d) nb *= n;
    stored.resize(nb);
    for (int t = 0; t < nb; ++t) { int z = t; std::array<int, 3> a = {0, 0, 0}; for (int d = dim - 1; d >= 0; --d) { a[d] = z % n; z /= n; } stored[t] = a; }
    if (kind == 2) {
      auto key = [&](const std::array<int, 3>& a) { return *std::max_element(a.begin(), a.begin() + dim); };
      std::stable_sort(stored.begin(), stored.end(), [&](const std::array<int, 3>& a, const std::array<int, 3>& b) {
        if (key(a) != key(b)) return key(a) < key(b);
        return std::lexicographical_compare(a.begin(), a.begin() + dim, b.begin(), b.begin() + dim);
      });
    }
  }
  std::vector<int> map((size_t)n * n * n, -1);
  for (size_t l = 0; l < stored.size(); ++l) map[(size_t)(stored[l][0] * n + stored[l][1]) * n + stored[l][2]] = (int)l;
  if (nb_out) *nb_out = (int)stored.size();
  return map;
}

}  // namespace b200fem
