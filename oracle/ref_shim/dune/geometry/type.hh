// Stand-in for dune-geometry's type.hh (absent from this image): dune/fem/space/lagrange/genericgeometry.hh includes it but uses
// nothing of it.  Environment glue for oracle/_ref.
#ifndef B200FEM_REF_SHIM_GEOMETRY_TYPE_HH
#define B200FEM_REF_SHIM_GEOMETRY_TYPE_HH
#endif
