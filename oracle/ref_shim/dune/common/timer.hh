// Stand-in for dune-common's timer.hh (absent from this image): the Newton loop only takes timings with it.
#ifndef B200FEM_REF_SHIM_TIMER_HH
#define B200FEM_REF_SHIM_TIMER_HH
#include <chrono>
namespace Dune {
class Timer {
  std::chrono::steady_clock::time_point t0_ = std::chrono::steady_clock::now();
public:
  void reset() { t0_ = std::chrono::steady_clock::now(); }
  double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
};
}
#endif
