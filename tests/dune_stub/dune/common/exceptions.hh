// tests/dune_stub: stand-ins for the few DUNE headers dune/fem/schemes/b200galerkin.hh includes, so that the reference-side
// binding can be compiled and run in this image (dune-common/-grid/-fem are not installable here).  Names, signatures and
// semantics follow the real headers (dune/common/exceptions.hh); test infrastructure only.
#ifndef B200FEM_DUNE_STUB_EXCEPTIONS_HH
#define B200FEM_DUNE_STUB_EXCEPTIONS_HH
#include <sstream>
#include <stdexcept>
#include <string>
namespace Dune {
struct Exception : std::runtime_error { Exception() : std::runtime_error("") {} void message(const std::string& m) { msg_ = m; } const char* what() const noexcept override { return msg_.c_str(); } std::string msg_; };
struct InvalidStateException : Exception {};
struct NotImplemented : Exception {};
}
#define DUNE_THROW(E, m) do { E th__ex; std::ostringstream th__out; th__out << #E << ": " << m; th__ex.message(th__out.str()); throw th__ex; } while (0)
#endif
