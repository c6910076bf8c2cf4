"""The launch plan of the two E-step passes (host logic, no GPU): every (tile, stage) is covered exactly once,
partial slots are dense per tile, and the resident CTA slots are filled to within a stage of the optimum."""
import math

import numpy as np
import pytest

from probreg_b200 import _cabi

SHAPES = [(98, 196, 296), (13, 196, 296), (25, 196, 296), (49, 196, 296), (98, 25, 296), (977, 1954, 296), (245, 489, 296),
          (1, 1, 296), (1, 2, 296), (3, 7, 296), (300, 1, 296), (1, 4000, 296), (98, 196, 444)]


@pytest.mark.parametrize("ntiles,nstages,slots", SHAPES)
def test_plan_covers_everything_once(ntiles, nstages, slots):
    items, max_slots = _cabi.plan_work(ntiles, nstages, slots)
    cover = np.zeros((ntiles, nstages), dtype=np.int32)
    per_tile = {}
    for t, a, b, s in items:
        assert 0 <= a < b <= nstages
        cover[t, a:b] += 1
        per_tile.setdefault(int(t), []).append(int(s))
    assert (cover == 1).all()
    for t, sl in per_tile.items():
        assert sorted(sl) == list(range(len(sl))) and len(sl) <= max_slots
    lens = items[:, 2] - items[:, 1]
    assert (np.diff(lens) <= 0).all()                       # longest first
    # makespan over `slots` CTAs (longest-first list scheduling) vs the ideal total / slots
    loads = np.zeros(min(slots, len(items)))
    for l in lens:
        loads[loads.argmin()] += l + 0.5
    ideal = ntiles * nstages / min(slots, ntiles * nstages)
    assert loads.max() <= 1.25 * ideal + 2.0, (loads.max(), ideal)


def test_bench_shapes_are_tight():
    # N = M = 100k on 1 and 8 GPUs: 98 x 196 and 13 x 196 stages on 296 slots
    for ntiles, nstages, bound in [(98, 196, 1.03), (13, 196, 1.12)]:
        items, _ = _cabi.plan_work(ntiles, nstages, 296)
        assert len(items) <= 296
        assert (items[:, 2] - items[:, 1]).max() <= math.ceil(bound * ntiles * nstages / 296)
