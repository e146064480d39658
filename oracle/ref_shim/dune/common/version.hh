// Stand-in for dune-common's version.hh (absent from this image; pulled in by dune/fem/version.hh).  Environment glue for oracle/_ref.
#ifndef B200FEM_REF_SHIM_VERSION_HH
#define B200FEM_REF_SHIM_VERSION_HH
#endif
