// Shadows dune/fem/io/parameter.hh when oracle/_ref is built: the real header is the parameter-file / IO stack on top of dune-common.
// The code under test reads its settings through it by KEY -- dune/fem/solver/parameter.hh (SolverParameter, compiled from the reference),
// dune/fem/solver/newtoninverseoperator.hh:103-380 (NewtonParameter) and operator/common/automaticdifferenceoperator.hh:99-101 -- so this
// stand-in is a key -> string table (filled by ref_bind.cpp from the test) behind the slice of the reader's interface those headers call
// (reference: dune/fem/io/parameter/reader.hh:60-260 getValue / getEnum / exists, dune/fem/io/parameter.hh:570-605 LocalParameter).
// Environment glue for oracle/_ref; no algorithm of the path lives here.
#ifndef B200FEM_REF_SHIM_IO_PARAMETER_HH
#define B200FEM_REF_SHIM_IO_PARAMETER_HH
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <dune/common/exceptions.hh>
namespace Dune { namespace Fem {
struct ParameterNotFound : Dune::Exception { using Dune::Exception::Exception; };
struct ParameterInvalid : Dune::Exception { using Dune::Exception::Exception; };
namespace RefShim {
inline std::map<std::string, std::string>& table() { static std::map<std::string, std::string> t; return t; }
template <class T> inline T parse(const std::string& key, const std::string& s) {
  std::istringstream in(s); T v; in >> std::boolalpha >> v;
  if (in.fail()) { in.clear(); in.str(s); in >> std::noboolalpha >> v; }
  if (in.fail()) DUNE_THROW(ParameterInvalid, "Parameter '" << key << "' invalid.");
  return v;
}
template <> inline std::string parse<std::string>(const std::string&, const std::string& s) { return s; }
}
struct ParameterReader {
  bool exists(const std::string& key) const { return RefShim::table().count(key) != 0; }
  template <class T> T getValue(const std::string& key) const {
    auto it = RefShim::table().find(key);
    if (it == RefShim::table().end()) DUNE_THROW(ParameterNotFound, "Parameter '" << key << "' not found.");
    return RefShim::parse<T>(key, it->second);
  }
  template <class T> T getValue(const std::string& key, const T& defaultValue) const {
    auto it = RefShim::table().find(key);
    return it == RefShim::table().end() ? defaultValue : RefShim::parse<T>(key, it->second);
  }
  template <int n> int getEnum(const std::string& key, const std::string (&values)[n], int defaultValue) const {
    return enumeration(key, values, n, defaultValue);
  }
  int getEnum(const std::string& key, const std::vector<std::string>& values, int defaultValue) const {
    return enumeration(key, values, (int)values.size(), defaultValue);
  }
private:
  template <class Values> int enumeration(const std::string& key, const Values& values, int n, int defaultValue) const {
    auto it = RefShim::table().find(key);
    if (it == RefShim::table().end()) return defaultValue;
    for (int i = 0; i < n; ++i) if (it->second == values[i]) return i;
    int j = -1; { std::istringstream in(it->second); in >> j; if (in.fail()) j = -1; }
    if (j < 0 || j >= n) DUNE_THROW(ParameterInvalid, "Parameter '" << key << "' invalid.");
    return j;
  }
};
struct Parameter {
  static const int solverStatistics = 1, extendedStatistics = 2;
  static const ParameterReader& container() { static ParameterReader r; return r; }
  static bool verbose(int = 1) { return false; }
  static void append(const std::string& key, const std::string& value, bool = false) { RefShim::table()[key] = value; }
};
template <class ParamDefault, class ParamImpl> struct LocalParameter : public ParamDefault {
  virtual ~LocalParameter() {}
  virtual ParamDefault* clone() const { return new ParamImpl(static_cast<const ParamImpl&>(*this)); }
  template <class... Args> LocalParameter(Args... args) : ParamDefault(args...) {}
};
template <class ParamDefault> struct LocalParameter<ParamDefault, ParamDefault> {
  virtual ~LocalParameter() {}
  virtual ParamDefault* clone() const { return new ParamDefault(static_cast<const ParamDefault&>(*this)); }
};
}}
#endif
