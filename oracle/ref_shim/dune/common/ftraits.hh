// Stand-in for dune-common's ftraits.hh (absent from this image).  dune/fem/solver/linear/cg.hh only asks for
// FieldTraits<double>::real_type; this is environment glue for oracle/_ref, no algorithm lives here.
#ifndef B200FEM_REF_SHIM_FTRAITS_HH
#define B200FEM_REF_SHIM_FTRAITS_HH
namespace Dune {
template <class T> struct FieldTraits { typedef T field_type; typedef T real_type; };
}
#endif
