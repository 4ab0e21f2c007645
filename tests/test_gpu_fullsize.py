"""GPU: the BASELINE.json configurations at FULL size, checked through size-independent properties (the CPU oracle
needs minutes to hours at these sizes) and, where the oracle is cheap, exactly.

  config 2  60-min track, hop-64 onset + beat tracking      → beats strictly increasing, IBI median = beat period
  config 4  10-min pair at 44.1 kHz, waveform xcorr          → identical picks / slope to the CPU port (fast on CPU)
  config 5  pair batch                                       → tiling idempotence: equal pairs give bit-equal results,
                                                               batch order does not matter"""
import numpy as np
import pytest
import scipy.signal

from oracle import pipeline_port as port
from oracle import synth

pytestmark = pytest.mark.gpu
SR = 22050


def click_track(dur_s, bpm, sr=SR, seed=0):
    """Cheap long input: decaying noise bursts on every beat over a quiet noise floor (vectorised)."""
    rng = np.random.default_rng(seed)
    n = int(dur_s * sr)
    y = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    period = 60.0 / bpm
    starts = (np.arange(int(dur_s / period)) * period * sr).astype(np.int64)
    burst = (rng.standard_normal(400) * np.exp(-np.arange(400) / 80.0)).astype(np.float32)
    for k in range(400):
        idx = starts + k
        idx = idx[idx < n]
        y[idx] += 0.5 * burst[k]
    return y


def test_config2_sixty_minute_hop64_pass(engine):
    from nightcore_analyzer import tempo as ntempo
    bpm = 123.0
    y = click_track(3600.0, bpm)
    assert len(y) == 79_380_000
    ibis = ntempo.estimate_ibis_global(y, SR)
    assert ibis is not None and len(ibis) > 0.9 * 3600 * bpm / 60
    assert np.all(ibis > 0.05)                                   # tempo.py:170 filter
    period = 60.0 / bpm
    assert abs(np.median(ibis) - period) < 64 / SR               # within one hop-64 frame of the true beat period
    frames = np.round(np.cumsum(ibis) * SR / 64)
    assert np.all(np.diff(frames) > 0)                           # strictly increasing beat frames
    # the first three minutes alone give the same beat period (the pass does not degrade with length)
    ibis3 = ntempo.estimate_ibis_global(y[: 180 * SR], SR)
    assert abs(np.median(ibis3) - np.median(ibis)) < 64 / SR


def test_config4_ten_minute_44k_xcorr_matches_port(engine):
    from nightcore_analyzer import xcorr as nx
    sr = 44100
    a = np.tile(synth.synth(4000, 60.0, sr, bpm=124.0), 10)       # 10 minutes = 26 460 000 samples
    assert len(a) == 26_460_000
    b = scipy.signal.resample_poly(a, 1000, 1003).astype(np.float32)
    b = (b + np.random.default_rng(1).standard_normal(len(b)).astype(np.float32) * 0.01).astype(np.float32)
    (slope, quality), (pos, pb) = nx.estimate_speed_xcorr_batch([(a, b)], sr, return_indices=True)[0]
    (w_slope, w_quality), (w_pos, w_pb) = port.speed_xcorr_arrays(a, b, sr, return_indices=True)
    assert pos.tolist() == w_pos.tolist() and pb.tolist() == w_pb.tolist()
    assert slope == w_slope and abs(quality - w_quality) <= 1e-5
    # xcorr.py:57,92: the path-based reference call would load at 22 050 Hz — same check on 2:1 decimated copies
    a2, b2 = a[::2].copy(), b[::2].copy()
    got = nx.estimate_speed_xcorr_batch([(a2, b2)], 22050, return_indices=True)[0]
    want = port.speed_xcorr_arrays(a2, b2, 22050, return_indices=True)
    assert got[1][1].tolist() == want[1][1].tolist() and got[0][0] == want[0][0]


def test_config5_batch_tiling_idempotence(engine):
    import nightcore_analyzer as na
    distinct = []
    for i in range(3):
        src, nc = synth.make_pair(5000 + i, 60.0, SR)
        distinct.append((nc, src))
    order = [0, 1, 2, 1, 0, 2, 2, 0]
    out = na.run_batch([distinct[i] for i in order], SR)
    first = {}
    for i, r in zip(order, out):
        assert not isinstance(r, Exception), r
        key = (r.tempo_ratio, r.tempo_ci, r.pitch_ratio, r.pitch_ci, r.ibi_ratio, r.ibi_ci, tuple(r.nc_tempos_raw),
               tuple(r.src_tempos_raw), tuple(r.nc_pitches_raw), r.classification)
        if i in first:
            assert key == first[i]                               # bit-equal regardless of position in the batch
        first[i] = key
    # and equal to analysing the pair alone
    alone = na.run_batch([distinct[1]], SR)[0]
    assert (alone.tempo_ratio, alone.ibi_ratio, tuple(alone.nc_pitches_raw)) == (first[1][0], first[1][4], first[1][8])
