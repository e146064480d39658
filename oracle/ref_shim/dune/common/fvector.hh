// Stand-in for dune-common's fvector.hh (absent from this image): the slice of Dune::FieldVector the reference's generic Lagrange
// points / base functions use (dune/fem/space/lagrange/generic*.hh).  Environment glue for oracle/_ref, no algorithm lives here.
#ifndef B200FEM_REF_SHIM_FVECTOR_HH
#define B200FEM_REF_SHIM_FVECTOR_HH
#include <cstddef>
#include <dune/common/exceptions.hh>
#include <dune/common/ftraits.hh>
namespace Dune {
template <class T, int n>
class FieldVector {
  T d_[n > 0 ? n : 1];
 public:
  typedef T value_type; typedef T field_type;
  static constexpr int dimension = n;
  FieldVector() { for (int i = 0; i < n; ++i) d_[i] = T(0); }
  FieldVector(const T& v) { for (int i = 0; i < n; ++i) d_[i] = v; }
  template <class U> FieldVector(const FieldVector<U, n>& o) { for (int i = 0; i < n; ++i) d_[i] = o[i]; }
  FieldVector& operator=(const T& v) { for (int i = 0; i < n; ++i) d_[i] = v; return *this; }
  T& operator[](std::size_t i) { return d_[i]; }
  const T& operator[](std::size_t i) const { return d_[i]; }
  FieldVector& operator*=(const T& v) { for (int i = 0; i < n; ++i) d_[i] *= v; return *this; }
  FieldVector& operator+=(const FieldVector& o) { for (int i = 0; i < n; ++i) d_[i] += o.d_[i]; return *this; }
  FieldVector& operator-=(const FieldVector& o) { for (int i = 0; i < n; ++i) d_[i] -= o.d_[i]; return *this; }
  FieldVector& axpy(const T& a, const FieldVector& o) { for (int i = 0; i < n; ++i) d_[i] += a * o.d_[i]; return *this; }
  static constexpr std::size_t size() { return n; }
};
template <class T, int n> struct FieldTraits<FieldVector<T, n>> { typedef T field_type; typedef T real_type; };
}
#endif
