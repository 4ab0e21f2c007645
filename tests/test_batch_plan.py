"""Host logic of the batch scheduler: the sub-batch plan of the end-to-end path (batch.plan_subbatches)."""
import pytest

from nightcore_analyzer import batch


@pytest.mark.parametrize("n,sub", [(1000, 125), (250, 63), (125, 32), (40, 16), (3, 16), (1, 1), (0, 16), (999, 125)])
def test_plan_covers_every_pair_once(n, sub):
    sizes = batch.plan_subbatches(n, sub)
    assert sum(sizes) == n
    assert all(0 < s <= sub for s in sizes)


def test_plan_grows_geometrically_then_splits_evenly():
    sizes = batch.plan_subbatches(1000, 125)
    head = [s for s in sizes if s < 91]
    assert head == sorted(head) and head[0] == 8
    assert all(b <= 2 * a for a, b in zip(sizes, sizes[1:]))       # an upload never outgrows the compute before it
    tail = sizes[len(head) + 1:]
    assert max(tail) - min(tail) <= 1                              # no short tail job
