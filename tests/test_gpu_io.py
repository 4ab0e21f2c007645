"""GPU: io.py drop-in (slice_windows energies in float64, energy gate mask, strip_silence indices)."""
import numpy as np
import pytest

from oracle import librosa_restated as lr
from oracle import synth

pytestmark = pytest.mark.gpu
SR = 22050


def ref_rms_db(audio):
    rms = float(np.sqrt(np.mean(audio.astype(np.float64) ** 2)))   # io.py:38-40
    return 20.0 * np.log10(max(rms, 1e-10))


def test_slice_windows_and_gate(engine):
    from nightcore_analyzer import io as nio
    y = synth.synth(21, 47.0, SR, bpm=100.0).copy()
    y[10 * SR : 22 * SR] *= 1e-4        # a quiet stretch the gate must remove
    wins = nio.slice_windows(y, SR)
    assert len(wins) == 1 + (len(y) - 220500) // 110250
    for i, w in enumerate(wins):
        assert w.audio.base is y or w.audio.base is y.base
        assert len(w.audio) == 220500 and w.start_sec == i * 110250 / SR
        assert abs(w.energy_db - ref_rms_db(w.audio)) < 1e-9
    kept = nio.energy_gate(wins)
    peak = max(ref_rms_db(w.audio) for w in wins)
    want_mask = [ref_rms_db(w.audio) >= peak - 40.0 for w in wins]
    assert [any(w is k for k in kept) for w in wins] == want_mask
    assert 0 < len(kept) < len(wins)
    assert nio.slice_windows(y[:1000], SR) == []
    assert nio.energy_gate([]) == []


def test_strip_silence(engine):
    from nightcore_analyzer import io as nio
    y = np.concatenate([np.zeros(30000, np.float32), synth.synth(22, 8.0, SR, bpm=120.0), np.zeros(41234, np.float32)])
    trimmed, lead, trail = nio.strip_silence(y, SR, 60.0)
    wt, (s, e) = lr.trim(y, 60.0)
    assert len(trimmed) == len(wt) and lead == s / SR and trail == (len(y) - e) / SR
    z = np.zeros(5000, np.float32)
    t, a, b = nio.strip_silence(z, SR)
    wt, (s, e) = lr.trim(z, 60.0)
    # librosa quirk: an all-zero signal is 0 dB below its own (zero) peak, i.e. "non-silent" everywhere
    assert len(t) == len(wt) == 5000 and a == s / SR


def test_rms_frames(engine):
    y = synth.synth(23, 5.0, SR, bpm=120.0)
    got = engine.rms_frames_dev(engine.to_dev(y), len(y), 2048, 512).cpu().numpy()
    want = lr.rms(y, 2048, 512)
    assert got.shape == want.shape
    # oracle: float32 mean (numpy pairwise); device: float64 accumulation rounded once — a few float32 ulps
    assert np.max(np.abs(got - want)) <= 2e-6 * float(np.max(want)) + 1e-12


def test_engines_of_ended_threads_are_released(engine):
    """An engine (pinned parameter ring + device workspaces) belongs to the host thread that created it and goes away
    with it; the launch counter stays monotonic."""
    import threading
    from nightcore_analyzer import _engine
    before = _engine.total_launches()
    seen = {}

    def work():
        eng = _engine.get_engine()
        seen["id"] = id(eng)
        eng.window_energy_dev(eng.pack([np.ones(4096, np.float32)])[0], eng.to_dev(np.array([0], np.int64)),
                              eng.to_dev(np.array([4096], np.int32)))
        seen["launches"] = eng.launches

    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert seen["launches"] > 0
    mine = _engine.get_engine()                 # same thread as the fixture: the cached engine, not a new one
    assert mine is engine
    with _engine._ENG_LOCK:
        _engine._drop_dead_engines()
        assert all(id(e) != seen["id"] for e, _ in _engine._ENGINES.values())
    assert _engine.total_launches() >= before + seen["launches"]
