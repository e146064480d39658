// Stand-in for dune-common's exceptions.hh (absent from this image; reached through fvector.hh there): DUNE_THROW and the exception
// types the compiled reference pieces name.  Environment glue for oracle/_ref, no algorithm lives here.
#ifndef B200FEM_REF_SHIM_EXCEPTIONS_HH
#define B200FEM_REF_SHIM_EXCEPTIONS_HH
#include <sstream>
#include <stdexcept>
namespace Dune {
struct Exception : std::runtime_error { Exception() : std::runtime_error("Dune::Exception") {} explicit Exception(const std::string& m) : std::runtime_error(m) {} };
struct RangeError : Exception { using Exception::Exception; };
struct NotImplemented : Exception { using Exception::Exception; };
struct InvalidStateException : Exception { using Exception::Exception; };
}
#ifndef DUNE_THROW
#define DUNE_THROW(E, m) do { std::ostringstream dune_throw_msg; dune_throw_msg << m; throw E(dune_throw_msg.str()); } while (0)
#endif
#endif
