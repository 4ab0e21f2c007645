"""GPU: the whole pipeline (pipeline.run / run_arrays / run_batch) vs the CPU port of the reference's
control flow (oracle/pipeline_port.py) on a synthetic pair — BASELINE config 1 in miniature.

Integer outputs (kept windows, per-window tempo lags → quantised BPM values, chunk lags, beat frames →
IBIs, PCG64 resample streams) are compared exactly; they all feed float64 medians/percentiles that are
then bit-identical."""
import os

import numpy as np
import pytest

from oracle import pipeline_port as port
from oracle import synth

pytestmark = pytest.mark.gpu
SR = 22050


@pytest.fixture(scope="module")
def pair():
    src, nc = synth.make_pair(1000, 45.0, SR)
    return nc, src


@pytest.fixture(scope="module")
def want(pair):
    return port.run_arrays(pair[0], pair[1], SR, faithful_cost=False)


def check(res, want):
    assert res.src_tempos_raw == want["src_tempos"]
    assert res.nc_tempos_raw == want["nc_tempos"]
    assert res.nc_pitches_raw == want["nc_hz"] and res.src_pitches_raw == want["src_hz"]
    assert (res.tempo_ratio, res.tempo_ci) == want["tempo"]
    if want["pitch"] is not None:
        assert (res.pitch_ratio, res.pitch_ci) == want["pitch"]
    else:
        assert res.pitch_ratio == 1.0 and res.n_source_pitch_windows == 0
    if want["ibi"] is not None:
        assert (res.ibi_ratio, res.ibi_ci) == want["ibi"]
    else:
        assert res.ibi_ratio is None
    assert res.nc_duration == want["nc_duration"] and res.src_duration == want["src_duration"]


def test_run_arrays_matches_port(engine, pair, want):
    import nightcore_analyzer as na
    logs = []
    res = na.run_arrays(pair[0], pair[1], SR, log=logs.append)
    check(res, want)
    assert res.pitch_method == "chroma_xcorr" and res.intro_offset_sec is None
    assert logs[0].startswith("Stripping silence") and logs[-1] == "Done."
    assert any(m.startswith("  NC tempo prior:") for m in logs)
    assert "Classification" in str(res)


def test_run_from_files(engine, pair, want, tmp_path):
    import nightcore_analyzer as na
    np.save(tmp_path / "nc.npy", pair[0])
    np.save(tmp_path / "src.npy", pair[1])
    res = na.run(str(tmp_path / "nc.npy"), str(tmp_path / "src.npy"), log=None)
    check(res, want)


def test_run_batch_equals_single_runs(engine, pair, want):
    import nightcore_analyzer as na
    src2, nc2 = synth.make_pair(1001, 40.0, SR)
    quiet = (np.zeros(SR * 12, np.float32), np.zeros(SR * 15, np.float32))
    out = na.run_batch([pair, (nc2, src2), quiet], SR)
    check(out[0], want)
    check(out[1], port.run_arrays(nc2, src2, SR, faithful_cost=False))
    assert isinstance(out[2], (RuntimeError, ValueError))     # a failing pair does not poison the batch
    single = na.run_arrays(nc2, src2, SR, log=None)
    assert single.tempo_ratio == out[1].tempo_ratio and single.ibi_ratio == out[1].ibi_ratio
    assert single.nc_pitches_raw == out[1].nc_pitches_raw


def test_gate_failure_raises(engine):
    import nightcore_analyzer as na
    with pytest.raises(RuntimeError, match="energy gate"):
        na.run_arrays(np.zeros(1000, np.float32), np.zeros(1000, np.float32), SR, log=None)


def test_manual_trim_and_no_pitch(engine, pair):
    import nightcore_analyzer as na
    res = na.run_arrays(pair[0], pair[1], SR, src_trim_sec=2.5, compute_pitch=False, log=None)
    assert res.intro_offset_sec == 2.5 and res.pitch_method is None and res.pitch_ratio == 1.0
    assert res.src_duration < len(pair[1]) / SR - 2.4


def test_verification_library_call(engine):
    """workflow.py:778-833 as a library call — a faithful HQNC passes, a 1.5 % slow one fails on IBI."""
    import scipy.signal
    from nightcore_analyzer import verify as nverify
    src = synth.synth(2100, 60.0, SR, bpm=128.0)
    ncog = scipy.signal.resample_poly(src, 4, 5).astype(np.float32)
    ncog = (ncog + np.random.default_rng(3).standard_normal(len(ncog)).astype(np.float32) * 0.003).astype(np.float32)
    good = scipy.signal.resample_poly(src, 4, 5).astype(np.float32)
    v = nverify.verify_arrays(good, ncog, SR)
    w = port.run_arrays(ncog, good, SR, compute_pitch=False, faithful_cost=False)
    wx = port.speed_xcorr_arrays(good, ncog, SR)
    assert (v.result.tempo_ratio, v.result.tempo_ci) == w["tempo"]
    assert (v.result.ibi_ratio, v.result.ibi_ci) == w["ibi"]
    assert v.result.xcorr_ratio == wx[0] and abs(v.result.xcorr_quality - wx[1]) <= 1e-5
    assert v.estimator == "IBI" and v.tempo_ok and v.pitch_ok and not v.length_warn
    slow = scipy.signal.resample_poly(src, 203, 250).astype(np.float32)     # speed 1.2315 instead of 1.25
    v2 = nverify.verify_arrays(slow, ncog, SR)
    assert not v2.tempo_ok and v2.length_warn
    assert abs(v2.corrected_speed_factor - 1.25 / (250 / 203)) < 5e-3
