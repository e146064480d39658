// Stand-in for dune-grid's common/geometry.hh (absent from this image): dune/fem/common/coordinate.hh includes it and uses nothing of it
// for plain coordinates.  Environment glue for oracle/_ref.
#ifndef B200FEM_REF_SHIM_GRID_GEOMETRY_HH
#define B200FEM_REF_SHIM_GRID_GEOMETRY_HH
#endif
