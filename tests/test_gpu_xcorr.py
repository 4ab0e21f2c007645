"""GPU: waveform cross-correlation speed search and intro alignment vs the CPU restatement
(oracle/pipeline_port.py: speed_xcorr_arrays / content_offset).

best_pb per window and the alignment peak index are integers and must be identical; slope is a
polyfit of identical integer tables (exact); quality is a float32-rounded cosine: abs 1e-5."""
import numpy as np
import pytest
import scipy.signal

from oracle import pipeline_port as port
from oracle import synth

pytestmark = pytest.mark.gpu
SR = 22050


def make_pair(seed, dur, sr=SR, up=1000, down=1003):
    a = synth.synth(seed, dur, sr, bpm=124.0)
    b = scipy.signal.resample_poly(a, up, down).astype(np.float32)
    b = (b + np.random.default_rng(seed).standard_normal(len(b)).astype(np.float32) * 0.01).astype(np.float32)
    return a, b


def test_speed_xcorr_matches_port(engine):
    from nightcore_analyzer import xcorr as nx
    a, b = make_pair(4000, 120.0)
    (slope, quality), (a_pos, best_pb) = nx.estimate_speed_xcorr_batch([(a, b)], SR, return_indices=True)[0]
    (w_slope, w_quality), (w_pos, w_pb) = port.speed_xcorr_arrays(a, b, SR, return_indices=True)
    assert a_pos.tolist() == w_pos.tolist()
    assert best_pb.tolist() == w_pb.tolist()
    assert slope == w_slope
    assert abs(quality - w_quality) <= 1e-5
    assert nx.estimate_speed_xcorr_arrays(a, b, SR) == (slope, quality)


def test_speed_xcorr_44k_and_batch(engine):
    from nightcore_analyzer import xcorr as nx
    a, b = make_pair(4001, 40.0, sr=44100)
    c, d = make_pair(4002, 60.0)
    res = nx.estimate_speed_xcorr_batch([(a, b)], 44100, return_indices=True)[0]
    want = port.speed_xcorr_arrays(a, b, 44100, return_indices=True)
    assert res[1][1].tolist() == want[1][1].tolist() and res[0][0] == want[0][0]
    both = nx.estimate_speed_xcorr_batch([(c, d), (c, c)], SR)
    assert both[0] == nx.estimate_speed_xcorr_arrays(c, d, SR)
    w_same = port.speed_xcorr_arrays(c, c, SR)     # identical files: the strided grid need not contain the true offset
    assert both[1][0] == w_same[0] and abs(both[1][1] - w_same[1]) <= 1e-5


def test_speed_xcorr_sentinels(engine):
    from nightcore_analyzer import xcorr as nx
    short = synth.synth(1, 3.0, SR, bpm=120.0)
    assert nx.estimate_speed_xcorr_arrays(short, short, SR) == (1.0, 0.0)            # shorter than one window
    z = np.zeros(SR * 30, np.float32)
    assert nx.estimate_speed_xcorr_arrays(z, z, SR) == (1.0, 0.0)                    # RMS gate drops every window
    assert nx.quality_label(0.8) == "good match" and nx.quality_label(0.5) == "moderate match"
    assert nx.quality_label(0.1).startswith("poor match")


def test_find_content_offset_matches_port(engine):
    from nightcore_analyzer import xcorr as nx
    body = synth.synth(77, 60.0, SR, bpm=110.0)
    intro = synth.synth(78, 12.0, SR, bpm=90.0) * 0.3
    src = np.concatenate([intro, body]).astype(np.float32)
    nc = scipy.signal.resample_poly(body, 4, 5).astype(np.float32)
    got = nx.find_content_offset(src, nc, SR)
    want, dbg = port.content_offset(src, nc, SR, return_debug=True)
    assert got == want
    assert got[0] > 1.0
