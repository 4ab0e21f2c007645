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


@pytest.mark.parametrize("n,sub,workers", [(1000, 125, 2), (125, 32, 2), (500, 125, 3), (90, 125, 2), (1, 1, 2), (0, 16, 2), (1000, 125, 1)])
def test_staggered_resident_plan(n, sub, workers):
    sizes = batch.stagger_sizes(n, sub, workers)
    assert sum(sizes) == n
    assert all(0 < s <= sub for s in sizes)
    if workers >= 2 and n > sub:
        assert sizes[0] == max(1, sub // workers)                  # the first job is short: workers run out of phase
        assert max(sizes[1:]) - min(sizes[1:]) <= 1


def test_gil_handoff_is_restored():
    import sys
    before = sys.getswitchinterval()
    with batch._fast_gil_handoff():
        assert sys.getswitchinterval() <= 2e-4
    assert sys.getswitchinterval() == before


@pytest.mark.parametrize("n", [40, 60, 125, 130, 200, 250])
def test_small_batches_stream_up_then_down(n):
    """A batch of one or two jobs' worth (a rank's share of 1000 pairs on 8 GPUs) is streamed as five jobs that grow and
    then shrink: the first upload is the only exposed one, and the call ends a short job after the last byte arrives."""
    sizes = batch.plan_subbatches(n, 125)
    assert sum(sizes) == n and len(sizes) == 5 and all(s > 0 for s in sizes)
    assert sizes[0] < sizes[1] < sizes[2] > sizes[3] > sizes[4]
    assert sizes[0] <= max(1, round(0.12 * n)) and sizes[4] <= round(0.2 * n)


def test_large_batches_keep_the_ramp():
    assert batch.plan_subbatches(1000, 125)[:7] == [8, 12, 18, 27, 40, 60, 91]
    assert batch.plan_subbatches(251, 125)[0] == 8            # just above two jobs' worth: the plain ramp
    assert batch.plan_subbatches(39, 125) == [8, 12, 19]      # below five first-jobs: ramp with the remainder folded in
