"""Host logic of the z-marching DG kernel: the static run list (b200fem_march_schedule, csrc/march_schedule.hpp).
No device needed.  Every plane of every column must be covered exactly once, the CTAs must be balanced under the cost
model, and the single-plane runs of rank-interface planes must come first in their CTA (they are sent to the neighbour
while the rest of the box is still being computed)."""
import ctypes as C

import numpy as np
import pytest

from dune_fem_b200 import _capi


def schedule(on, grid, flags):
    on_a = (C.c_int32 * 3)(*on)
    begin = (C.c_int32 * (grid + 1))()
    n = C.c_int32()
    runs = (C.c_int32 * (4 * 16384))()
    _capi.check(_capi.lib().b200fem_march_schedule(on_a, grid, flags, runs, 16384, begin, C.byref(n)))
    return np.array(runs[:4 * n.value]).reshape(-1, 4), np.array(begin[:])


@pytest.mark.parametrize("on,grid,flags", [
    ([64, 64, 64], 148, 0b00111100), ([64, 64, 64], 148, 0b00111111), ([64, 64, 64], 148, 0b01001101), ([64, 64, 64], 148, 0b10110010),
    ([34, 20, 14], 148, 0b10), ([16, 16, 1], 1, 0), ([18, 4, 9], 148, 3), ([48, 40, 56], 148, 0), ([2, 1, 2], 148, 3),
    ([64, 64, 2], 148, 3), ([64, 64, 3], 148, 3), ([64, 64, 64], 132, 0b111100), ([128, 128, 128], 148, 0b111100), ([8, 6, 4], 148, 1)])
def test_schedule_covers_every_plane_once_and_is_balanced(on, grid, flags):
    runs, begin = schedule(on, grid, flags)
    tx, ty, nz = (on[0] + 15) // 16, (on[1] + 15) // 16, on[2]
    cover = np.zeros((tx * ty, nz), dtype=int)
    for col, za, zb, flush in runs:
        assert 0 <= col < tx * ty and 0 <= za < zb <= nz
        cover[col, za:zb] += 1
    assert (cover == 1).all()
    assert begin[0] == 0 and begin[-1] == len(runs) and (np.diff(begin) >= 0).all()
    # interface planes: single-plane runs with the flush flag, first in their CTA
    lo_if, hi_if = flags & 1, (flags >> 1) & 1
    expect_flush = (tx * ty) * (lo_if + hi_if) if nz >= 3 else 0
    assert int(runs[:, 3].sum()) == expect_flush
    for b in range(grid):
        mine = runs[begin[b]:begin[b + 1]]
        seen_plain = False
        for col, za, zb, flush in mine:
            if flush:
                assert zb - za == 1 and za in (0, nz - 1) and not seen_plain
            else:
                seen_plain = True
    # balance (the cost model gives boundary columns up to 30 % fewer planes): no CTA far above the mean once there is enough work
    planes = np.array([sum(zb - za for _, za, zb, _ in runs[begin[b]:begin[b + 1]]) for b in range(grid)])
    if planes.sum() >= 4 * grid:
        assert planes.max() <= 1.12 * planes.sum() / grid + 2.0


def test_schedule_rejects_bad_arguments():
    on_a = (C.c_int32 * 3)(0, 4, 4)
    begin = (C.c_int32 * 3)()
    n = C.c_int32()
    assert _capi.lib().b200fem_march_schedule(on_a, 2, 0, None, 0, begin, C.byref(n)) == _capi.ERR_INVALID
