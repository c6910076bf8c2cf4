"""The launch plan of the two E-step passes (host logic, no GPU): every (tile, unit) is covered exactly once -- a unit is a sub-chunk
of 64 j-records, items may begin and end inside a TMA stage --, partial slots are dense per tile, the resident CTA slots are filled
to within a few units of the optimum, and a last tile that is mostly padding gets correspondingly fewer, longer items."""
import math

import numpy as np
import pytest

from probreg_b200 import _cabi

SHAPES = [(98, 1563, 296, 0.75), (13, 1563, 296, 0.25), (25, 1563, 296, 0.5), (49, 1563, 296, 0.875), (98, 196, 296, 1.0),
          (977, 15625, 296, 0.625), (245, 3907, 296, 1.0), (1, 1, 296, 1.0), (1, 2, 296, 0.125), (3, 7, 296, 1.0), (300, 1, 296, 1.0),
          (1, 32000, 296, 1.0), (98, 1563, 444, 0.75), (13, 1563, 296, 1.0), (400, 8, 296, 0.5)]


@pytest.mark.parametrize("ntiles,nunits,slots,last", SHAPES)
def test_plan_covers_everything_once(ntiles, nunits, slots, last):
    items, max_slots = _cabi.plan_work(ntiles, nunits, slots, last)
    cover = np.zeros((ntiles, nunits), dtype=np.int32)
    per_tile = {}
    for t, a, b, s in items:
        assert 0 <= a < b <= nunits
        cover[t, a:b] += 1
        per_tile.setdefault(int(t), []).append(int(s))
    assert (cover == 1).all()
    for t, sl in per_tile.items():
        assert sorted(sl) == list(range(len(sl))) and len(sl) <= max_slots
    cost = np.where(items[:, 0] == ntiles - 1, last, 1.0) * (items[:, 2] - items[:, 1])
    assert (np.diff(cost) <= 1e-12).all()                   # most expensive first
    # makespan over `slots` CTAs (list scheduling in launch order, 4 units of overhead per item) vs the ideal total / slots
    loads = np.zeros(min(slots, len(items)))
    for c in cost:
        loads[loads.argmin()] += c + 4.0
    total = (ntiles - 1 + last) * nunits
    ideal = total / min(slots, ntiles * nunits)
    assert loads.max() <= 1.25 * ideal + 12.0, (loads.max(), ideal)


def test_bench_shapes_are_tight():
    """N = M = 100k on 1, 2, 4 and 8 GPUs (98 / 49 / 25 / 13 target tiles of 1024, 1563 sub-chunks of sources, 296 resident CTAs): one
    wave, and the longest item within 2 % (+ 1 unit) of the cost-weighted mean.  Before the cuts followed the cost and the sub-chunk
    grain the 8-GPU shape ran 72 units per CTA against a mean of 64 (a 13th tile with 2 of 8 warps alive cut 22 ways like the rest,
    items of whole 512-record stages), the 4-GPU one 144 against 129."""
    for ntiles, last in [(98, 0.75), (49, 0.875), (25, 0.5), (13, 0.25)]:
        items, _ = _cabi.plan_work(ntiles, 1563, 296, last)
        assert len(items) <= 296
        cost = np.where(items[:, 0] == ntiles - 1, last, 1.0) * (items[:, 2] - items[:, 1])
        mean = (ntiles - 1 + last) * 1563 / 296
        assert cost.max() <= 1.02 * mean + 1.0, (ntiles, cost.max(), mean)


def test_argument_errors():
    for bad in [(0, 5, 10, 1.0), (3, 0, 10, 1.0), (3, 5, 0, 1.0), (3, 5, 10, 0.0), (3, 5, 10, 1.5)]:
        with pytest.raises(_cabi.CpdError):
            _cabi.plan_work(*bad)
