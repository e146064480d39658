// jit_integrands.cuh -- the kernel-parameter layout of run-time compiled integrands (jit.cu).  Included by the host
// translation unit and by the NVRTC program, so that both sides agree on the bytes that travel as the `Integrands`
// kernel argument.  The device class (with the three integrand methods wrapping the user's functions) derives from this
// base inside the NVRTC program and adds no members.
#pragma once
#include "../../include/b200fem.h"

namespace b200fem {

constexpr int kJitMaxConstants = 32;

struct JitIntegrandsBase {
  b200fem_model m;     // only has_skeleton / has_boundary are read (by the kernel itself)
  int dim;
  int with_data;       // kept for layout symmetry with AdrIntegrandsT; user integrands always carry their data terms
  double c[kJitMaxConstants];
};

}  // namespace b200fem
