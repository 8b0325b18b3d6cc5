"""The batched weight-gradient kernel (csrc/wgrad_batch.cu) cuts the (job, 64-sample block) work list of one backward
pass into 148 pieces of equal weight (the measured time per block of each job shape); a CTA's piece may straddle job
boundaries.  The cut is replayed here on
the host (no GPU): every block of every job must be owned exactly once, the segments of one job must hold
consecutive workspace slots (the reduce kernel sums slot ranges), and no CTA may carry more than its share plus one
block."""
import ctypes

import pytest

from panonerf_b200 import _lib

# (Nw, Kw) of the jobs of one level of the configs/panonerf.yaml backward pass (field.py:_backward_fused)
BASE = [(128, 64), (128, 256), (256, 256), (256, 64)] + [(256, 256)] * 2 + [(256, 96)] + [(256, 256)] * 5 + [(256, 96)]
JADJ = [(256, 96)] + [(256, 256)] * 4 + [(256, 96)] + [(256, 256)] * 3 + [(256, 16)]


def _cost(nw, kw, colsum):
    """ns per 64-sample block (csrc/wgrad_batch.cu: wb_block_cost)"""
    xb = (kw + 63) // 64
    return 1075 + 13 * xb + (140 if colsum else 0) if nw == 256 else 685 + 5 * xb + (70 if colsum else 0)


def _colsum(j):
    return j % 3 != 0


def _plan(M, shapes):
    n = len(shapes)
    jobs = (ctypes.c_longlong * (8 * n))()
    for j, (nw, kw) in enumerate(shapes):
        jobs[8 * j + 4], jobs[8 * j + 5], jobs[8 * j + 7] = nw, kw, int(_colsum(j))
    cap = 148 + n + 8
    out = (ctypes.c_longlong * (5 * cap))()
    k = _lib.lib().pnb_wgrad_batch_plan(M, n, jobs, out, cap)
    assert k > 0
    return [tuple(out[5 * i + c] for c in range(5)) for i in range(k)]


@pytest.mark.parametrize("M", [1, 63, 64, 100, 5000, 8192 * 64, 8192 * 100, 8192 * 64 + 17])
@pytest.mark.parametrize("shapes", [BASE, JADJ + BASE, [(256, 256)], [(128, 16), (256, 256)]])
def test_wgrad_batch_plan_covers_every_block_once_and_is_balanced(M, shapes):
    segs = _plan(M, shapes)
    blocks = (M + 63) // 64
    owned = {j: [] for j in range(len(shapes))}
    load = [0] * 148
    slots_of = {}
    for i, (cta, job, b0, b1, slot) in enumerate(segs):
        assert 0 <= cta < 148 and 0 <= b0 < b1 <= blocks and slot == i
        owned[job].append((b0, b1))
        load[cta] += (b1 - b0) * _cost(*shapes[job], _colsum(job))
        slots_of.setdefault(job, []).append(slot)
    for j, rng in owned.items():                       # contiguous cover of [0, blocks), in order, without overlap
        assert rng and rng[0][0] == 0 and rng[-1][1] == blocks, (j, rng[:3])
        assert all(a[1] == b[0] for a, b in zip(rng, rng[1:]))
        s = slots_of[j]
        assert s == list(range(s[0], s[0] + len(s)))   # consecutive slots
    assert len(segs) <= 148 + len(shapes)
    total = blocks * sum(_cost(nw, kw, _colsum(j)) for j, (nw, kw) in enumerate(shapes))
    biggest = max(_cost(nw, kw, True) for nw, kw in shapes)
    assert max(load) <= total / 148 + biggest + 1      # nobody carries more than its share + one block
