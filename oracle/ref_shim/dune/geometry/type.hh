// Stand-in for dune-geometry's type.hh (absent from this image): the slice of Dune::GeometryType / Dune::GeometryTypes that
// dune/fem/quadrature/{quadratureimp,femquadratures}.hh name (a tag carried around; no algorithm lives here).
#ifndef B200FEM_REF_SHIM_GEOMETRY_TYPE_HH
#define B200FEM_REF_SHIM_GEOMETRY_TYPE_HH
namespace Dune {
class GeometryType {
  unsigned int id_ = 0, dim_ = 0;
 public:
  GeometryType() = default;
  GeometryType(unsigned int id, unsigned int dim) : id_(id), dim_(dim) {}
  unsigned int dim() const { return dim_; }
  unsigned int id() const { return id_; }
  bool isCube() const { return id_ == (1u << dim_) - 1u; }
  bool isSimplex() const { return id_ == 0; }
  bool operator==(const GeometryType& o) const { return id_ == o.id_ && dim_ == o.dim_; }
};
namespace GeometryTypes {
inline GeometryType cube(unsigned int dim) { return GeometryType((1u << dim) - 1u, dim); }
inline GeometryType simplex(unsigned int dim) { return GeometryType(0, dim); }
static const GeometryType prism(5, 3), pyramid(3, 3);
}
}
#endif
