// Shadows dune/fem/io/parameter.hh when oracle/_ref is built: the real header is the parameter/IO stack on top of dune-common.
// The code under test (dune/fem/operator/common/automaticdifferenceoperator.hh:99-101) only reads ONE key with a default,
// "fem.differenceoperator.eps" = 0, from Parameter::container(); this stand-in answers every query with the default handed in,
// i.e. an empty parameter file (reference: dune/fem/io/parameter/reader.hh getValue(key, defaultValue)).
#ifndef B200FEM_REF_SHIM_IO_PARAMETER_HH
#define B200FEM_REF_SHIM_IO_PARAMETER_HH
#include <string>
namespace Dune { namespace Fem {
struct ParameterReader {
  template <class T> T getValue(const std::string&, const T& defaultValue) const { return defaultValue; }
};
struct Parameter {
  static const ParameterReader& container() { static ParameterReader r; return r; }
};
}}
#endif
