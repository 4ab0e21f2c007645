"""Optional tier (test-plan item 5): the CPU restatement of librosa (oracle/librosa_restated.py) against REAL
librosa, function by function, on seeded synthetic audio.

librosa is not installed in the build container or on the GPU box and cannot be installed there (no network, not in the
wheelhouse), so these tests are normally SKIPPED and parity with librosa stays "unpinned" (DESIGN.md).  The moment an
operator provides librosa — site-wide, or unpacked under ``baseline/_ref`` (tests/conftest.py puts that directory on
``sys.path``) — the same restatement the CUDA kernels are held to is held to librosa itself, and the pytest report says
so.  Tolerances: integers (trim bounds, tempo lag / BPM, beat frames, tuning) exact; spectra and envelopes to float32
round-off; chroma to 5e-3 of its maximum because the restatement's 2:1 resampler is a Kaiser half-band, not soxr_hq (the one
documented deviation)."""
import numpy as np
import pytest

librosa = pytest.importorskip("librosa", reason="librosa not available: parity with real librosa stays unpinned")

from oracle import librosa_restated as lr  # noqa: E402
from oracle import synth  # noqa: E402

SR = 22050


@pytest.fixture(scope="module")
def y():
    return synth.synth(4242, 20.0, SR, bpm=117.0)


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))) /
                 max(1e-30, float(np.max(np.abs(np.asarray(b, dtype=np.float64))))))


def test_reports_the_librosa_version():
    print(f"real-librosa tier active: librosa {librosa.__version__}")


def test_stft_mel_db(y):
    S_ref = librosa.stft(y, n_fft=2048, hop_length=512, window="hann", center=True, pad_mode="constant")
    assert rel(lr.stft(y, 2048, 512), S_ref) < 1e-5
    assert rel(lr.mel_filter(SR, 2048, 128), librosa.filters.mel(sr=SR, n_fft=2048, n_mels=128)) < 1e-6
    P = np.abs(S_ref) ** 2
    assert np.max(np.abs(lr.power_to_db(P) - librosa.power_to_db(P, top_db=80.0))) < 1e-4


@pytest.mark.parametrize("hop", [512, 64])
def test_onset_strength(y, hop):
    ref = librosa.onset.onset_strength(y=y, sr=SR, hop_length=hop)
    got = lr.onset_strength(y, SR, hop)
    assert got.shape == ref.shape and rel(got, ref) < 1e-4


@pytest.mark.parametrize("hop,bpm", [(512, 120.0), (512, 93.7), (64, 120.0)])
def test_tempo_and_beats(y, hop, bpm):
    env = librosa.onset.onset_strength(y=y, sr=SR, hop_length=hop)
    t_ref = float(np.atleast_1d(librosa.feature.tempo(onset_envelope=env, sr=SR, hop_length=hop, start_bpm=bpm))[0])
    assert float(np.atleast_1d(lr.tempo(env, SR, hop, bpm))[0]) == t_ref
    tempo_ref, beats_ref = librosa.beat.beat_track(onset_envelope=env, sr=SR, hop_length=hop, start_bpm=bpm)
    tempo_got, beats_got = lr.beat_track(env, SR, hop, bpm)
    assert float(np.atleast_1d(tempo_got)[0]) == float(np.atleast_1d(tempo_ref)[0])
    assert np.asarray(beats_got).tolist() == np.asarray(beats_ref).tolist()
    assert np.array_equal(lr.frames_to_time(beats_ref, SR, hop), librosa.frames_to_time(beats_ref, sr=SR, hop_length=hop))


def test_trim_and_rms(y):
    padded = np.concatenate([np.zeros(7000, np.float32), y, np.zeros(12345, np.float32)])
    _, idx = librosa.effects.trim(padded, top_db=60.0)
    _, (s, e) = lr.trim(padded, top_db=60.0)
    assert (int(s), int(e)) == (int(idx[0]), int(idx[1]))
    assert rel(lr.rms(y, 2048, 512), librosa.feature.rms(y=y, frame_length=2048, hop_length=512)[0]) < 1e-5


def test_tuning_and_chroma(y):
    assert float(lr.estimate_tuning(y, SR, 12)) == float(librosa.estimate_tuning(y=y, sr=SR, bins_per_octave=12))
    ref = librosa.feature.chroma_cqt(y=y, sr=SR, hop_length=512, bins_per_octave=36)
    got = lr.chroma_cqt(y, SR, 512, 36)
    assert got.shape == ref.shape
    assert rel(got.mean(axis=1), ref.mean(axis=1)) < 5e-3      # resampler deviation (Kaiser half-band vs soxr_hq)
