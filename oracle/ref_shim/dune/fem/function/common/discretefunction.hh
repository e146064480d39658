// Stand-in for dune/fem/function/common/discretefunction.hh when oracle/_ref is built (the real header needs dune-grid): the solver
// headers compiled here (solver/cginverseoperator.hh) are templates over the discrete function and only include this header;
// ref_bind.cpp supplies the slice of the interface they call over a plain array.
#ifndef B200FEM_REF_SHIM_DISCRETEFUNCTION_HH
#define B200FEM_REF_SHIM_DISCRETEFUNCTION_HH
#endif
