"""CPU: the task-parallel runner of the CPU port (oracle/pipeline_tasks.py — what `bench.py --impl reference` times)
returns exactly what the serial port returns."""
import numpy as np

from oracle import pipeline_port as port
from oracle import pipeline_tasks, synth

SR = 22050


def test_task_runner_equals_serial_port():
    src, nc = synth.make_pair(7, 25.0, SR)
    quiet = (np.zeros(SR * 12, np.float32), np.zeros(SR * 15, np.float32))
    runner = pipeline_tasks.TaskRunner([(nc, src), quiet], SR, procs=2, faithful_cost=False)
    try:
        (got, failed), windows = runner.run([0, 1])
    finally:
        runner.close()
    want, n_win = port.run_arrays(nc, src, SR, faithful_cost=False, return_window_count=True)
    assert isinstance(failed, (RuntimeError, ValueError))
    assert windows >= n_win
    for key in ("src_tempos", "nc_tempos", "nc_start_bpm", "nc_duration", "src_duration", "chunk_lags", "src_hz", "nc_hz",
                "pitch", "tempo", "ibi", "pitch_point_st"):
        assert got[key] == want[key], key
    assert tuple(got["pitch_ci_st"]) == tuple(want["pitch_ci_st"])
