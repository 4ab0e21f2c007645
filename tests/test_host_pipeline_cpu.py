"""CPU: the product's HOST logic end to end on an oracle-backed stand-in engine (tests/fake_engine.py).

The drop-in modules (io / tempo / pitch / consensus / pipeline) and the batch path (batch.analyse_staged) build segment
tables, gate windows, derive priors, format the reference's log lines and assemble AnalysisResult objects on the host;
only the arithmetic is on the device.  With the engine replaced by one that computes the same quantities with the
oracle, that host logic is held — without a GPU — to the golden vectors produced by the reference's own modules
(tests/golden/pipeline_golden.json): every value via float.hex, every log line, str(result)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import fromhex
from fake_engine import FakeEngine
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SR = 22050


def unhex(v):
    return None if v is None else fromhex(v)


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "pipeline_golden.json")) as f:
        return json.load(f)["pipeline_A"]


@pytest.fixture()
def fake(monkeypatch):
    from nightcore_analyzer import _engine
    eng = FakeEngine()
    monkeypatch.setattr(_engine, "get_engine", lambda device=None: eng)
    return eng


@pytest.fixture(scope="module")
def pair_a():
    src, nc = synth.make_pair(1000, 45.0, SR)
    return nc, src


def check_against_golden(res, A):
    assert res.src_tempos_raw == [unhex(v) for v in A["src_tempos"]]
    assert res.nc_tempos_raw == [unhex(v) for v in A["nc_tempos"]]
    assert res.nc_pitches_raw == [unhex(v) for v in A["nc_hz"]] and res.src_pitches_raw == [unhex(v) for v in A["src_hz"]]
    assert res.tempo_ratio == unhex(A["tempo_ratio"]) and list(res.tempo_ci) == [unhex(v) for v in A["tempo_ci"]]
    assert res.pitch_ratio == unhex(A["pitch_ratio"]) and list(res.pitch_ci) == [unhex(v) for v in A["pitch_ci"]]
    assert res.ibi_ratio == unhex(A["ibi_ratio"]) and list(res.ibi_ci) == [unhex(v) for v in A["ibi_ci"]]
    assert res.classification == A["classification"] and res.pitch_method == A["pitch_method"]
    assert res.warnings == A["warnings"] and res.rubberband == A["rubberband"]
    assert str(res) == A["str"]
    assert [res.n_source_pitch_windows, res.n_nc_pitch_windows, res.n_source_tempo_windows,
            res.n_nc_tempo_windows] == A["n"]


def test_run_arrays_host_logic_matches_reference_flow(fake, pair_a, golden):
    """pipeline.run_arrays → io / pitch / tempo / consensus drop-ins, stage order and log lines of pipeline.py:91-216."""
    import nightcore_analyzer as na
    nc, src = pair_a
    logs = []
    res = na.run_arrays(nc, src, SR, log=logs.append)
    check_against_golden(res, golden)
    ref_logs = golden["logs"][golden["logs"].index(next(m for m in golden["logs"] if m.startswith("Stripping silence"))):]
    assert logs == ref_logs


def test_batch_path_host_logic_equals_the_single_pair_path(fake, pair_a, golden):
    """batch.analyse_staged is pipeline.run turned inside out (every stage once over all pairs): on the same engine it
    must give the single-pair results — here for a batch that also holds a pair the energy gate empties and a pair too
    short for any window, which must fail alone without disturbing the others."""
    from nightcore_analyzer import batch
    nc, src = pair_a
    silent = np.zeros(30 * SR, np.float32)
    short = synth.synth(5, 4.0, SR, bpm=100.0)
    tracks = [nc, src, silent, silent, nc, src, short, short]
    audio, off, ln = fake.pack(tracks)
    st = batch.StagedBatch(audio=audio, off=off, length=ln.astype(np.int64), sr=SR, h2d_bytes=0)
    stats = {}
    out = batch.analyse_staged(st, stats=stats)
    assert len(out) == 4
    check_against_golden(out[0], golden)
    check_against_golden(out[2], golden)
    assert out[0].intro_offset_sec is None
    # the failing pairs fail exactly as they do alone through run_arrays (same exception type and message)
    import nightcore_analyzer as na
    for k, (a, b) in ((1, (silent, silent)), (3, (short, short))):
        with pytest.raises(Exception) as alone:
            na.run_arrays(a, b, SR, log=None)
        assert type(out[k]) is type(alone.value) and str(out[k]) == str(alone.value)
    assert isinstance(out[3], RuntimeError) and "All windows were discarded by the energy gate" in str(out[3])
    assert isinstance(out[1], ValueError) and "Insufficient valid tempo windows" in str(out[1])
    # analysed windows: the two good pairs + the 5 + 5 windows of the all-zero pair (equal energies: the gate keeps them all)
    assert stats["tracks"] == 8 and stats["windows"] == 2 * (len(golden["src_tempos"]) + len(golden["nc_tempos"])) + 10


def test_no_pitch_and_manual_trim_flow(fake, pair_a):
    """compute_pitch=False and src_trim_sec (pipeline.py:106-112, 148-159): log lines and fields of the reference."""
    import nightcore_analyzer as na
    nc, src = pair_a
    logs = []
    res = na.run_arrays(nc, np.concatenate([np.zeros(2 * SR, np.float32) + 1e-3, src]), SR, src_trim_sec=2.0,
                        silence_strip_db=None, compute_pitch=False, log=logs.append)
    assert res.intro_offset_sec == 2.0 and res.pitch_method is None
    assert res.src_pitches_raw == [] and res.nc_pitches_raw == [] and res.pitch_ratio == 1.0
    assert "Manual source trim: skipping 2.00s from source start" in logs and "Skipping pitch estimation." in logs
    assert not any(m.startswith("Stripping silence") for m in logs) and logs[-1] == "Done."


def test_xcorr_host_logic_matches_reference_flow(fake):
    """xcorr.estimate_speed_xcorr_arrays: trimming, window / candidate tables, polyfit slope and median quality
    (xcorr.py:95-162) on the stand-in engine vs the reference's own result for the same pair."""
    import scipy.signal
    from nightcore_analyzer import xcorr as nx
    with open(os.path.join(ROOT, "tests", "golden", "pipeline_golden.json")) as f:
        C = json.load(f)["xcorr_C"]
    a = synth.synth(4000, 120.0, SR, bpm=124.0)
    b = scipy.signal.resample_poly(a, 1000, 1003).astype(np.float32)
    b = (b + np.random.default_rng(4000).standard_normal(len(b)).astype(np.float32) * 0.01).astype(np.float32)
    ratio, quality = nx.estimate_speed_xcorr_arrays(a, b, SR)
    assert ratio == unhex(C["ratio"])
    assert abs(quality - unhex(C["quality"])) <= 1e-6
    assert nx.quality_label(quality) == C["label"]
    # too short for a single 3 s window after edge trimming: the reference's (1.0, 0.0) sentinel
    assert nx.estimate_speed_xcorr_arrays(a[: 2 * SR], b[: 2 * SR], SR) == (1.0, 0.0)


def test_run_from_files_emits_the_reference_log_from_its_first_line(fake, pair_a, golden, tmp_path):
    """pipeline.run (pipeline.py:23-216) on .npy files: the complete log of the reference's own run, loading stanzas
    included, and the same result object."""
    import nightcore_analyzer as na
    nc, src = pair_a
    np.save(tmp_path / "nc.npy", nc)
    np.save(tmp_path / "src.npy", src)
    logs = []
    res = na.run(str(tmp_path / "nc.npy"), str(tmp_path / "src.npy"), log=logs.append)
    assert logs == golden["logs"]
    check_against_golden(res, golden)
    # log=None is silent and changes nothing
    assert str(na.run(str(tmp_path / "nc.npy"), str(tmp_path / "src.npy"), log=None)) == golden["str"]


def test_auto_align_branches(fake, pair_a, monkeypatch):
    """pipeline.py:113-125: an intro offset at or above ALIGN_MIN_OFFSET trims the source and is reported; a smaller one is
    only logged.  (find_content_offset itself is held to the reference on the GPU and in the oracle tests.)"""
    import nightcore_analyzer as na
    from nightcore_analyzer import pipeline, xcorr
    nc, src = pair_a
    padded = np.concatenate([synth.synth(9, 3.0, SR, bpm=90.0) * 0.2, src]).astype(np.float32)

    monkeypatch.setattr(pipeline, "find_content_offset", lambda s, n, sr: (3.0, 1.2512))
    logs = []
    res = na.run_arrays(nc, padded, SR, auto_align=True, compute_pitch=False, silence_strip_db=None, log=logs.append)
    assert res.intro_offset_sec == 3.0
    assert logs[0] == "Detecting intro offset (RMS envelope alignment)…"
    assert logs[1] == "  Intro detected — trimming 3.00s from source start  (speed hint: 1.2512×)"
    assert res.src_duration == len(src) / SR

    small = xcorr.ALIGN_MIN_OFFSET / 2
    monkeypatch.setattr(pipeline, "find_content_offset", lambda s, n, sr: (small, 1.25))
    logs = []
    res = na.run_arrays(nc, src, SR, auto_align=True, compute_pitch=False, silence_strip_db=None, log=logs.append)
    assert res.intro_offset_sec is None
    assert logs[1] == (f"  No significant intro offset detected  (raw: {small:.2f}s < "
                       f"{xcorr.ALIGN_MIN_OFFSET:.1f}s threshold)")
    # a manual trim wins over auto-align and never calls the detector
    monkeypatch.setattr(pipeline, "find_content_offset", lambda *a: (_ for _ in ()).throw(AssertionError("called")))
    res = na.run_arrays(nc, padded, SR, src_trim_sec=3.0, auto_align=True, compute_pitch=False, silence_strip_db=None,
                        log=None)
    assert res.intro_offset_sec == 3.0


def test_cli_single_pair_on_the_stand_in_engine(fake, pair_a, golden, tmp_path, capsys):
    """cli.main → pipeline.run → export.cli_dict: exit code 0, the JSON of the reference's cli for this result; --quiet
    prints nothing; an emptied gate is exit code 1 with the reference's message."""
    from nightcore_analyzer import cli, export
    import nightcore_analyzer as na
    nc, src = pair_a
    np.save(tmp_path / "nc.npy", nc)
    np.save(tmp_path / "src.npy", src)
    out = tmp_path / "res.json"
    assert cli.main(["-n", str(tmp_path / "nc.npy"), "-s", str(tmp_path / "src.npy"), "-o", str(out), "--quiet"]) == 0
    assert capsys.readouterr().out == ""
    want = na.run_arrays(nc, src, SR, log=None)
    assert out.read_text(encoding="utf-8") == json.dumps(export.cli_dict(want), indent=2)
    assert str(want) == golden["str"]
    np.save(tmp_path / "short.npy", synth.synth(3, 4.0, SR, bpm=100.0))
    assert cli.main(["-n", str(tmp_path / "short.npy"), "-s", str(tmp_path / "short.npy"), "-q"]) == 1
    assert "ERROR: All windows were discarded by the energy gate." in capsys.readouterr().err


def test_find_content_offset_host_logic_matches_reference_flow(fake):
    """xcorr.find_content_offset: per-speed geometry (which speeds are skipped, search lengths), best-score selection and
    the frames → seconds conversion (xcorr.py:165-259) against the reference's own answer for the same pair."""
    import scipy.signal
    from nightcore_analyzer import xcorr as nx
    with open(os.path.join(ROOT, "tests", "golden", "pipeline_golden.json")) as f:
        D = json.load(f)["align_D"]
    body = synth.synth(77, 60.0, SR, bpm=110.0)
    intro = synth.synth(78, 12.0, SR, bpm=90.0) * 0.3
    src = np.concatenate([intro, body]).astype(np.float32)
    nc = scipy.signal.resample_poly(body, 4, 5).astype(np.float32)
    off, speed = nx.find_content_offset(src, nc, SR)
    assert off == unhex(D["offset_sec"]) and float(speed) == unhex(D["speed"])
    # a nightcore longer than the source at every candidate speed: nothing to search, the reference's neutral answer
    assert nx.find_content_offset(nc[: 3 * SR], np.tile(nc, 3), SR) == (0.0, (nx.ALIGN_SPEED_LO + nx.ALIGN_SPEED_HI) / 2.0)


def test_verification_library_call_host_logic(fake):
    """verify.verify_arrays (workflow.py:160-163, 299-400, 778-833 as a library call): a re-render of the same nightcore
    (tiny added noise) must verify — IBI estimator, tempo and pitch within tolerance, no length warning, and the xcorr
    verdict follows the 0.30 quality gate; a copy played 3 % faster must fail the tempo check, warn about the length and
    report the corrective factor."""
    import scipy.signal
    from nightcore_analyzer import verify, xcorr as nx
    ncog = synth.synth(31, 32.0, SR, bpm=150.0)
    hqnc = (ncog + np.random.default_rng(5).standard_normal(len(ncog)).astype(np.float32) * 1e-3).astype(np.float32)
    v = verify.verify_arrays(hqnc, ncog, SR)
    assert v.estimator == "IBI" and v.tempo_ok and v.pitch_ok and not v.length_warn
    assert v.result.pitch_method is None and v.corrected_speed_factor == v.best_ratio
    assert (v.result.xcorr_ratio, v.result.xcorr_quality) == nx.estimate_speed_xcorr_arrays(hqnc, ncog, SR)
    assert v.xcorr_discarded == ((v.result.xcorr_quality or 0.0) < verify.XCORR_QUALITY_GATE)
    assert v.xcorr_label == (None if v.xcorr_discarded else nx.quality_label(v.result.xcorr_quality))
    fast = scipy.signal.resample_poly(ncog, 100, 103).astype(np.float32)       # HQNC 3 % too fast
    w = verify.verify_arrays(fast, ncog, SR)
    assert not w.tempo_ok and w.length_warn and abs(w.best_ratio - 1 / 1.03) < 5e-3
