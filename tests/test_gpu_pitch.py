"""GPU: tuning estimate, CQT chroma, cyclic xcorr and estimate_pitch_chroma vs the CPU restatement
(oracle/librosa_restated.py, oracle/pipeline_port.py).

Tolerances: tuning histogram bin and chroma lags are integers and must be identical; the mean chroma
vector is float (float32 contraction on the device, complex64/float64 mix on the oracle): max abs
error <= 1e-4 of the vector's maximum."""
import numpy as np
import pytest

from oracle import librosa_restated as lr
from oracle import pipeline_port as port
from oracle import synth
import golden_inputs as gi

pytestmark = pytest.mark.gpu
SR = 22050


def tuning_index(t):
    return int(round((t + 0.5) * 100))


@pytest.fixture(scope="module")
def chunks():
    out = []
    for seed, mult in ((31, 1.0), (32, 2.0 ** (0.3 / 12)), (33, 1.25 * 2.0 ** (1.5 / 12))):
        out.append(synth.synth(seed, 20.0, SR, bpm=118.0, pitch_mult=mult))
    out.append(synth.synth(34, 7.3, SR, bpm=100.0)[:150001])      # ragged, odd length
    return out


def test_tuning_index_matches_oracle(engine, chunks):
    audio, off, ln = engine.pack(chunks)
    _, tun = engine.chroma_mean_dev(audio, off, ln, SR)
    got = tun.cpu().numpy()
    want = [tuning_index(lr.estimate_tuning(y, SR, bins_per_octave=36)) for y in chunks]
    assert got.tolist() == want


def test_tuning_of_silence_is_zero(engine):
    audio, off, ln = engine.pack([np.zeros(30000, np.float32)])
    chroma, tun = engine.chroma_mean_dev(audio, off, ln, SR)
    assert tun.cpu().numpy().tolist() == [50]
    assert np.all(chroma.cpu().numpy() == 0.0)


def test_mean_chroma_matches_oracle_given_tuning(engine, chunks):
    import torch
    want_t = [lr.estimate_tuning(y, SR, bins_per_octave=36) for y in chunks]
    audio, off, ln = engine.pack(chunks)
    tun = torch.tensor([tuning_index(t) for t in want_t], dtype=torch.int32)
    chroma, _ = engine.chroma_mean_dev(audio, off, ln, SR, tuning_idx=tun)
    got = chroma.cpu().numpy()
    for y, t, g in zip(chunks, want_t, got):
        want = lr.chroma_cqt(y, SR, 512, 36, tuning=t).mean(axis=1)
        assert g.shape == want.shape == (12,)
        assert np.max(np.abs(g - want)) <= 1e-4 * np.max(want)


def test_mean_chroma_api(engine, chunks):
    from nightcore_analyzer import pitch as npitch
    g = npitch._mean_chroma(chunks[0], SR)
    want = port.mean_chroma(chunks[0], SR)
    assert np.max(np.abs(g - want)) <= 1e-4 * np.max(want)


def test_cyclic_xcorr_golden(engine, golden):
    from nightcore_analyzer import pitch as npitch
    for k, want in golden["GV7"].items():
        assert npitch._cyclic_xcorr_peak(gi.GV7_A, np.roll(gi.GV7_A, int(k))) == want
    r = np.random.default_rng(5)
    for _ in range(20):
        a, b = r.random(12), r.random(12)
        assert npitch._cyclic_xcorr_peak(a, b) == port.cyclic_xcorr_peak(a, b)


def test_estimate_pitch_chroma_matches_port(engine):
    """BASELINE config 3 in miniature: independent speed (1.25x) and +1.5 st pitch shift, 3 chunk pairs."""
    from nightcore_analyzer import pitch as npitch
    src = synth.synth(3000, 75.0, SR, bpm=112.0)
    nc = synth.synth(3000, 60.0, SR, bpm=112.0, speed=1.25, pitch_mult=1.25 * 2.0 ** (1.5 / 12))
    logs = []
    src_hz, nc_hz, point, ci, n = npitch.estimate_pitch_chroma(src, nc, SR, log=logs.append)
    w_src, w_nc, w_point, w_ci, w_n, w_lags = port.estimate_pitch_chroma(src, nc, SR)
    assert n == w_n == 3
    assert src_hz == w_src and nc_hz == w_nc            # integer lags / 3 → identical floats
    assert point == w_point and ci == w_ci              # PCG64-exact bootstrap of the chunk shifts
    assert logs and logs[-1].startswith("    Chroma xcorr:")
    hz_s, hz_n, method = npitch.estimate_pitch_combined(src, nc, SR)
    assert method == "chroma_xcorr" and hz_n == w_nc


def test_short_files_fall_back_to_one_chunk(engine):
    from nightcore_analyzer import pitch as npitch
    src = synth.synth(41, 9.0, SR, bpm=120.0)
    nc = synth.synth(41, 7.2, SR, bpm=120.0, speed=1.25, pitch_mult=1.25)
    logs = []
    src_hz, nc_hz, point, ci, n = npitch.estimate_pitch_chroma(src, nc, SR, log=logs.append)
    w = port.estimate_pitch_chroma(src, nc, SR)
    assert n == 1 and nc_hz == w[1] and ci == (point, point)
    assert any("Only 1 chunk(s)" in m for m in logs)
