"""CPU model of the slot insertion of `k_pillar_build` (csrc/pillars.cu, step d) under arbitrary interleavings.

The GPU tests compare the kernel with the oracle, but a run only samples the schedules the hardware happens to produce.  This
model restates the per-point protocol as a sequence of atomic memory steps -- two snapshot reads {first quad, last slot}, seven
more quad reads when the first quad is all below, then the conserving atomicMin chain with its early-exit read of the last slot
every four steps -- and interleaves the points of a cell in random order, step by step.  Whatever the schedule, the row must end
as the S smallest point indices in ascending order (spconv's first-come result), i.e. what
oracle/pillar_ref.c computes.  The invariants the kernel relies on are asserted after every step: the row is ascending and its
values never increase.
"""
import random

import pytest

EMPTY = 0xFFFFFFFF


class Point:
    """One non-first point of a cell: a generator of atomic steps over the shared row."""

    def __init__(self, v, row, S):
        self.v, self.row, self.S = v, row, S
        self.steps = self._run()

    def _run(self):
        row, S, v = self.row, self.S, self.v
        own = v
        # snapshot: first quad and last slot, in either order (two independent loads)
        order = [0, 1]
        random.shuffle(order)
        q0 = last = None
        for what in order:
            if what == 0:
                q0 = list(row[0:4])
            else:
                last = row[S - 1]
            yield
        if last < own:
            return                      # S smaller indices in place
        k = sum(1 for x in q0 if x < own)
        if k == 4:
            quads = list(range(1, S // 4))
            random.shuffle(quads)       # the seven loads are in flight together: any completion order
            for q in quads:
                k += sum(1 for x in row[4 * q:4 * q + 4] if x < own)
                yield
            if k >= S:
                return
        while True:
            old = row[k]                # atomicMin(row + k, v): one indivisible step
            row[k] = min(old, v)
            yield
            if old == EMPTY:
                return
            v = max(old, v)
            k += 1
            if k == S:
                return
            if k % 4 == 0:
                seen = row[S - 1]
                yield
                if seen < v:
                    return


def run_cell(indices, S, rng):
    """indices: point indices of one cell in any order; returns the final row."""
    first = min(indices)
    row = [first] + [EMPTY] * (S - 1)           # opened by the first point before anyone else starts (release / acquire)
    live = [Point(v, row, S) for v in indices if v != first]
    prev = list(row)
    while live:
        p = rng.choice(live)
        try:
            next(p.steps)
        except StopIteration:
            live.remove(p)
        assert all(a <= b for a, b in zip(row, row[1:])), "row not ascending"
        assert all(a <= b for a, b in zip(row, prev)), "a slot increased"
        prev = list(row)
    return row


@pytest.mark.parametrize("S", [8, 32])
def test_slot_insertion_any_interleaving(S):
    rng = random.Random(1234 + S)
    random.seed(99 + S)
    for trial in range(150):
        n = rng.choice([1, 2, 3, 5, S - 1, S, S + 1, 2 * S, 3 * S + 1])
        indices = rng.sample(range(100000), n)
        # arrival bias: mostly in index order (the kernel's sweep), sometimes reversed or random
        mode = trial % 3
        live_order = sorted(indices) if mode == 0 else sorted(indices, reverse=True) if mode == 1 else indices
        row = run_cell(live_order, S, rng)
        want = sorted(indices)[:S]
        assert row == want + [EMPTY] * (S - len(want))


def test_late_points_leave_without_an_atomic():
    """A point that finds S smaller indices in place must not touch the row (the property the in-order sweep exploits)."""
    S = 8
    row = list(range(10, 10 + S))
    p = Point(1000, row, S)
    n_steps = sum(1 for _ in p.steps)
    assert n_steps == 2 and row == list(range(10, 10 + S))
