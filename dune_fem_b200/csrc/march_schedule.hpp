// march_schedule.hpp -- host side: static work schedule of the z-marching DG kernel (dg_kronecker_march.cuh).  No CUDA in here:
// the schedule is plain host logic (checked on the CPU through b200fem_march_schedule).
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

namespace b200fem {

// one column of TX x TY elements, planes z in [za, zb); flush: account / publish the rank-interface rows right after this run
struct MarchRun { int col, za, zb, flush; };

// run lists (dg_kronecker_march.cuh: MarchRun), one per (owned extents, grid, interface pattern)
struct MarchScheduleKey { int on[3], grid, ifz_lo, ifz_hi, gx_lo, gx_hi, gy_lo, gy_hi, ify_lo, ify_hi; bool operator==(const MarchScheduleKey& o) const { return std::memcmp(this, &o, sizeof(*this)) == 0; } };
// Static schedule of the persistent grid.  Costs in units of one interior plane-tile: a plane of a column that touches the
// domain boundary in x costs ~10 % more (boundary corrections diverge in every warp), in y ~3 %; a run costs 0.8 on top of
// its planes (two edge steps).  Rank-interface planes in z become single-plane runs that are dealt out first, one per CTA,
// with the flush flag set; the remaining planes of all columns form one tape that is cut into pieces of equal cost (a CTA
// that already holds an interface run gets correspondingly less).  A piece that crosses a column boundary is two runs.
inline void march_schedule(const MarchScheduleKey& k, int tx, int ty, std::vector<MarchRun>& runs, std::vector<int>& begin) {
  const int ncols = tx * ty, nz = k.on[2], grid = k.grid;
  const int zlo = k.ifz_lo ? 1 : 0, zhi = nz - (k.ifz_hi ? 1 : 0);
  auto plane_cost = [&](int col) { const int bx = col % tx, by = col / tx; double c = 1.0;
    if ((bx == 0 && k.gx_lo) || (bx == tx - 1 && k.gx_hi)) c += 0.10;
    if ((by == 0 && k.gy_lo) || (by == ty - 1 && k.gy_hi)) c += 0.03;
    if ((by == 0 && k.ify_lo) || (by == ty - 1 && k.ify_hi)) c += 0.35;     // rows for a y-neighbour: these CTAs are to finish early
    return c; };
  constexpr double kRun = 0.8, kFlush = 1.0;
  std::vector<std::vector<MarchRun>> mine((size_t)grid);
  std::vector<double> load((size_t)grid, 0.0);
  // interface planes first
  int nif = 0;
  if (nz >= 3) for (int side = 0; side < 2; ++side) {
    if (!(side ? k.ifz_hi : k.ifz_lo)) continue;
    for (int c = 0; c < ncols; ++c, ++nif) {
      const int b = (int)(((long long)nif * grid) / (2 * ncols)) % grid, z = side ? nz - 1 : 0;
      mine[(size_t)b].push_back(MarchRun{c, z, z + 1, 1}); load[(size_t)b] += plane_cost(c) + kRun + kFlush;
    }
  }
  const int za0 = nz >= 3 ? zlo : 0, zb0 = nz >= 3 ? zhi : nz;
  // cut the tape of the remaining planes: CTA b is filled up to the quota q (its interface run counts), the last CTA takes what
  // is left; q is the smallest quota (bisection) for which the last CTA does not end up above it
  const std::vector<std::vector<MarchRun>> pre = mine;
  auto fill = [&](double q, bool keep) -> double {
    if (keep) mine = pre;
    int b = 0; double acc = load[0];
    for (int c = 0; c < ncols; ++c) {
      const double pc = plane_cost(c);
      int z = za0;
      while (z < zb0) {
        if (acc + kRun + 0.5 * pc > q && acc > load[(size_t)b] && b < grid - 1) { ++b; acc = load[(size_t)b]; }
        int z1 = z; acc += kRun;
        while (z1 < zb0 && (acc + 0.5 * pc <= q || z1 == z || b == grid - 1)) { acc += pc; ++z1; }
        if (keep) mine[(size_t)b].push_back(MarchRun{c, z, z1, 0});
        z = z1;
      }
    }
    return b == grid - 1 ? acc : 0.0;          // load of the CTA that takes the remainder (0: the tape ended before the last CTA)
  };
  double total = 0; for (double l : load) total += l;
  for (int c = 0; c < ncols; ++c) total += plane_cost(c) * std::max(0, zb0 - za0);
  double lo_q = total / grid, hi_q = total / grid + 4.0 + 2.0 * kRun;
  for (int it = 0; it < 40; ++it) { const double mid = 0.5 * (lo_q + hi_q); if (fill(mid, false) <= mid + 0.55) hi_q = mid; else lo_q = mid; }
  fill(hi_q, true);
  runs.clear(); begin.assign((size_t)grid + 1, 0);
  for (int i = 0; i < grid; ++i) { begin[(size_t)i] = (int)runs.size(); for (const MarchRun& r : mine[(size_t)i]) runs.push_back(r); }
  begin[(size_t)grid] = (int)runs.size();
}


}  // namespace b200fem
