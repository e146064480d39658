// Stand-in for dune-common's deprecated.hh (absent from this image; pulled in by dune/fem/version.hh).  Environment glue for oracle/_ref.
#ifndef B200FEM_REF_SHIM_DEPRECATED_HH
#define B200FEM_REF_SHIM_DEPRECATED_HH
#endif
