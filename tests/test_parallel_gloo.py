"""CPU: the N>1 path of SURVEY.md §8(e) — pairs sharded round-robin over ranks, one all_gather of the
fixed-size per-pair records — on the gloo backend with world_size 2 (NCCL over NVLink on the GPU box)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_result(i):
    from nightcore_analyzer import consensus as c
    if i % 5 == 3:
        return RuntimeError("All windows were discarded by the energy gate.")
    t = 1.2 + 0.001 * i
    return c._assemble([440.0] * 3, [440.0 * 2 ** (1 / 36)] * 3, [100.0] * 4, [100.0 * t] * 4, np.full(4, 100.0),
                       np.full(4, 100.0 * t), (1.02, (1.01, 1.03)), (t, (t - 0.01, t + 0.01)), (3, 3), 100.0, 100.0 * t)


def _worker(rank, world, port, n_pairs, q):
    for p in (ROOT, os.path.join(ROOT, "nightcore-to-flac-analyzer_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from nightcore_analyzer import parallel as npar
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    my_ids = npar.shard_indices(n_pairs, rank, world)
    results = [_fake_result(i) for i in my_ids]
    table = npar.gather_result_records(results, my_ids, n_pairs)
    q.put((rank, my_ids, table))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    n_pairs, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from nightcore_analyzer import parallel as npar
    ids = sorted(i for _, my, _ in out for i in my)
    assert ids == list(range(n_pairs))                         # every pair analysed exactly once
    want = npar.records_of([_fake_result(i) for i in range(n_pairs)])
    for _, _, table in out:                                    # every rank holds the full table, in pair order
        assert table.shape == (n_pairs, npar.RECORD_F64)
        assert np.array_equal(np.isnan(table), np.isnan(want))
        assert np.array_equal(np.nan_to_num(table), np.nan_to_num(want))
    assert want[3, 0] == 1.0 and want[0, 0] == 0.0 and want[0, 1] == 1.2


def test_shard_indices_cover_and_balance():
    from nightcore_analyzer import parallel as npar
    for n in (0, 1, 7, 1000):
        for w in (1, 2, 4, 8):
            parts = [npar.shard_indices(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_numa_binding_is_a_no_op_without_nvml():
    """parallel.bind_to_gpu_cpus is an optimisation: without a driver it must change nothing and raise nothing."""
    import os
    from nightcore_analyzer import parallel
    before = os.sched_getaffinity(0)
    cpus = parallel.bind_to_gpu_cpus(0)
    assert isinstance(cpus, list)
    assert os.sched_getaffinity(0) == (set(cpus) if cpus else before)
