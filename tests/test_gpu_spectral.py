"""GPU: whole-file spectral statistics vs the CPU restatement of spectral.analyze
(oracle/pipeline_port.spectral_stats → librosa_restated.spectral_centroid / spectral_rolloff / amplitude_to_db).

Tolerances: centroid, roll-off, band means, RMS statistics 1e-5 relative (float32 FFT on the device, float64 FFT in the
oracle; a roll-off bin may flip in a frame whose cumulative sum crosses the threshold within rounding — the mean moves
by < 1e-6 relative); per-bin mean dB 2e-3 dB absolute; the effective-bandwidth bin must be identical."""
import numpy as np
import pytest

from oracle import pipeline_port as port
from oracle import synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


@pytest.mark.parametrize("sr,dur,seed", [(22050, 30.0, 9), (44100, 12.0, 10)])
def test_spectral_stats_match_port(engine, sr, dur, seed):
    from nightcore_analyzer import spectral as nsp
    y = synth.synth(seed, dur, sr, bpm=120.0)
    if sr == 44100:                      # band-limit like a lossy transcode: the effective bandwidth must see it
        import scipy.signal
        sos = scipy.signal.butter(12, 15000, btype="low", fs=sr, output="sos")
        y = scipy.signal.sosfilt(sos, y).astype(np.float32)
    got = nsp.analyze_arrays([y], sr)[0]
    want = port.spectral_stats(y, sr)
    for k in ("centroid", "rolloff", "rms_mean", "rms_variance", "sub_bass", "bass", "midrange", "presence", "brilliance"):
        assert rel(getattr(got, k), want[k]) < 1e-5, (k, getattr(got, k), want[k])
    assert abs(got.decay_rate - want["decay_rate"]) < 1e-6 * max(1.0, abs(want["decay_rate"])) + 1e-9
    assert got.duration == want["duration"]
    assert got.effective_bandwidth_hz == want["effective_bandwidth_hz"]
    if sr == 44100:
        assert got.effective_bandwidth_hz < 18000


def test_per_bin_db_and_batch(engine):
    sr = 22050
    ys = [synth.synth(11, 8.0, sr, bpm=100.0), synth.synth(12, 5.5, sr, bpm=140.0)[:100001], np.zeros(4096, np.float32)]
    audio, off, ln = engine.pack(ys)
    stats, bins = engine.spectral_stats_dev(audio, off, ln, sr)
    stats, bins = stats.cpu().numpy(), bins.cpu().numpy()
    for i, y in enumerate(ys[:2]):
        want = port.spectral_stats(y, sr)
        assert np.max(np.abs(bins[i] - want["freq_avg_db"])) < 2e-3
        assert stats[i, 7] == 1 + len(y) // 512
    # silence: every frame sums to 0 → centroid 0, roll-off bin 0, all dB at the floor relation 0 − 0
    assert stats[2, 0] == 0.0 and stats[2, 1] == 0.0 and stats[2, 8] == 0.0
    assert np.all(bins[2] == 0.0)


def test_analyze_from_file(engine, tmp_path):
    from nightcore_analyzer import spectral as nsp
    y = synth.synth(13, 6.0, 22050, bpm=120.0)
    np.save(tmp_path / "a.npy", y)
    st = nsp.analyze(str(tmp_path / "a.npy"))
    want = port.spectral_stats(y, 22050)
    assert rel(st.centroid, want["centroid"]) < 1e-5 and st.duration == 6.0
