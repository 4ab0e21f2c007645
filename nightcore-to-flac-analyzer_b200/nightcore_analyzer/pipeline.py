"""
Full windowed consensus pipeline — drop-in for the reference's ``nightcore_analyzer/pipeline.py``
(pipeline.py:23-216: same ``run`` signature, stage order, log lines, exceptions and result object).

``run`` follows the reference stage by stage through the drop-in modules (io → pitch → tempo →
consensus → IBI), each of which batches its work onto the GPU.  ``run_arrays`` is the same pipeline
from already-loaded arrays; ``run_batch`` analyses many pairs in one data-parallel pass
(``batch.analyse_staged``) and is what ``bench.py`` measures.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

from .io import (
    load_audio, strip_silence, slice_windows, energy_gate,
    WINDOW_SEC, HOP_SEC, ENERGY_GATE_DB, SILENCE_STRIP_DB, SAMPLE_RATE,
)
from .pitch import estimate_pitch_combined
from .tempo import batch_estimate_tempo, estimate_ibis_global
from .consensus import build_result, compute_ibi_ratio, AnalysisResult
from .xcorr import find_content_offset, ALIGN_MIN_OFFSET


def run_arrays(
    nc_audio: np.ndarray,
    src_audio: np.ndarray,
    sr: int = SAMPLE_RATE,
    *,
    window_sec: float = WINDOW_SEC,
    hop_sec: float = HOP_SEC,
    energy_gate_db: float = ENERGY_GATE_DB,
    silence_strip_db: Optional[float] = SILENCE_STRIP_DB,
    src_trim_sec: float = 0.0,
    auto_align: bool = False,
    compute_pitch: bool = True,
    log: Optional[Callable[[str], None]] = print,
) -> AnalysisResult:
    """pipeline.py:91-216 — everything of ``run`` after the two files are loaded."""
    def _log(msg: str) -> None:
        if log is not None:
            log(msg)

    # ── 1b. strip leading/trailing silence (pipeline.py:91-104)
    if silence_strip_db is not None:
        _log(f"Stripping silence (top_db={silence_strip_db} dB)…")
        nc_audio, nc_lead, nc_trail = strip_silence(nc_audio, sr, silence_strip_db)
        src_audio, src_lead, src_trail = strip_silence(src_audio, sr, silence_strip_db)
        _log(
            f"  nightcore: −{nc_lead:.2f}s leading, −{nc_trail:.2f}s trailing"
            f"  →  {len(nc_audio)/sr:.1f} s"
        )
        _log(
            f"  source:    −{src_lead:.2f}s leading, −{src_trail:.2f}s trailing"
            f"  →  {len(src_audio)/sr:.1f} s"
        )

    # ── 1c. source trim: manual override takes priority over auto-align (pipeline.py:106-125)
    intro_offset_sec: Optional[float] = None
    if src_trim_sec > 0.0:
        trim_samples = int(src_trim_sec * sr)
        src_audio = src_audio[trim_samples:]
        intro_offset_sec = src_trim_sec
        _log(f"Manual source trim: skipping {src_trim_sec:.2f}s from source start")
    elif auto_align:
        _log("Detecting intro offset (RMS envelope alignment)…")
        raw_offset, align_speed = find_content_offset(src_audio, nc_audio, sr)
        if raw_offset >= ALIGN_MIN_OFFSET:
            src_audio = src_audio[int(raw_offset * sr):]
            intro_offset_sec = raw_offset
            _log(
                f"  Intro detected — trimming {raw_offset:.2f}s from source start"
                f"  (speed hint: {align_speed:.4f}×)"
            )
        else:
            _log(
                f"  No significant intro offset detected"
                f"  (raw: {raw_offset:.2f}s < {ALIGN_MIN_OFFSET:.1f}s threshold)"
            )

    # ── 2. window (pipeline.py:127-131)
    _log(f"Slicing into {window_sec:.0f} s windows (hop {hop_sec:.0f} s)…")
    nc_windows = slice_windows(nc_audio, sr, window_sec, hop_sec)
    src_windows = slice_windows(src_audio, sr, window_sec, hop_sec)
    _log(f"  nightcore: {len(nc_windows)} windows  |  source: {len(src_windows)} windows")

    # ── 3. energy gate (pipeline.py:133-146)
    _log(f"Energy gating (threshold {energy_gate_db} dB below peak)…")
    nc_windows = energy_gate(nc_windows, energy_gate_db)
    src_windows = energy_gate(src_windows, energy_gate_db)
    _log(
        f"  after gating — nightcore: {len(nc_windows)} windows"
        f"  |  source: {len(src_windows)} windows"
    )
    if not nc_windows or not src_windows:
        raise RuntimeError(
            "All windows were discarded by the energy gate.  "
            "Try raising --energy-gate (e.g. --energy-gate -60)."
        )

    # ── 4. pitch (pipeline.py:148-159)
    if compute_pitch:
        _log("Estimating pitch (chromagram cross-correlation)…")
        src_pitches, nc_pitches, pitch_method = estimate_pitch_combined(src_audio, nc_audio, sr, log=_log)
        _log(f"  Pitch method: {pitch_method}")
    else:
        _log("Skipping pitch estimation.")
        src_pitches, nc_pitches, pitch_method = [], [], None

    # ── 5. tempo: source first, then the nightcore with the source-derived prior (pipeline.py:161-186)
    _log("Estimating tempo (librosa)…")
    _log("  ← source →")
    src_tempos = batch_estimate_tempo(src_windows, log=_log)

    nc_duration = len(nc_audio) / sr
    src_duration = len(src_audio) / sr

    nc_start_bpm = 120.0
    valid_src = [t for t in src_tempos if t is not None]
    if valid_src and nc_duration > 0 and src_duration > 0:
        median_src = float(np.median(valid_src))
        nc_start_bpm = median_src * (src_duration / nc_duration)
        _log(
            f"  NC tempo prior: {nc_start_bpm:.1f} BPM  "
            f"(src median {median_src:.1f} BPM × dur ratio "
            f"{src_duration / nc_duration:.4f})"
        )

    _log("  ← nightcore →")
    nc_tempos = batch_estimate_tempo(nc_windows, log=_log, start_bpm=nc_start_bpm)

    # ── 6. consensus + classification (pipeline.py:188-197)
    _log("Computing consensus…")
    result = build_result(
        src_pitches, nc_pitches, src_tempos, nc_tempos,
        nc_duration=nc_duration, src_duration=src_duration,
    )
    result.intro_offset_sec = intro_offset_sec
    result.pitch_method = pitch_method

    # ── 7. IBI ratio pass (pipeline.py:199-213)
    _log("Computing IBI ratio (high-precision beat timestamps, hop=64)…")
    nc_ibis = estimate_ibis_global(nc_audio, sr, start_bpm=nc_start_bpm)
    src_ibis = estimate_ibis_global(src_audio, sr)
    if (nc_ibis is not None and len(nc_ibis) >= 4 and
            src_ibis is not None and len(src_ibis) >= 4):
        ibi_r, ibi_c = compute_ibi_ratio(nc_ibis, src_ibis)
        result.ibi_ratio = ibi_r
        result.ibi_ci = ibi_c
        _log(f"  IBI ratio: {ibi_r:.6f}×  95% CI [{ibi_c[0]:.6f}, {ibi_c[1]:.6f}]")
    else:
        _log("  IBI ratio: insufficient beats — skipped")

    _log("Done.")
    return result


def run(
    nightcore_path: str,
    source_path: str,
    *,
    window_sec: float = WINDOW_SEC,
    hop_sec: float = HOP_SEC,
    energy_gate_db: float = ENERGY_GATE_DB,
    silence_strip_db: Optional[float] = SILENCE_STRIP_DB,
    src_trim_sec: float = 0.0,
    auto_align: bool = False,
    compute_pitch: bool = True,
    log: Optional[Callable[[str], None]] = print,
) -> AnalysisResult:
    """pipeline.py:23-216 — analyse the tempo and pitch relationship between a nightcore track and
    its source.  Files are ``.npy`` (float mono at 22 050 Hz) or PCM ``.wav`` (see io.load_audio)."""
    def _log(msg: str) -> None:
        if log is not None:
            log(msg)

    _log("Loading nightcore audio…")
    nc_audio, sr = load_audio(nightcore_path)
    _log(f"  {len(nc_audio) / sr:.1f} s  ({len(nc_audio):,} samples @ {sr} Hz)")
    _log("Loading source audio…")
    src_audio, _ = load_audio(source_path, sr=sr)
    _log(f"  {len(src_audio) / sr:.1f} s  ({len(src_audio):,} samples @ {sr} Hz)")
    return run_arrays(nc_audio, src_audio, sr, window_sec=window_sec, hop_sec=hop_sec, energy_gate_db=energy_gate_db,
                      silence_strip_db=silence_strip_db, src_trim_sec=src_trim_sec, auto_align=auto_align,
                      compute_pitch=compute_pitch, log=log)


def run_batch(
    pairs,
    sr: int = SAMPLE_RATE,
    *,
    window_sec: float = WINDOW_SEC,
    hop_sec: float = HOP_SEC,
    energy_gate_db: float = ENERGY_GATE_DB,
    silence_strip_db: Optional[float] = SILENCE_STRIP_DB,
    compute_pitch: bool = True,
    compute_ibi: bool = True,
    sub_batch: Optional[int] = None,
    workers: int = 2,
    stats: Optional[dict] = None,
) -> List[Union[AnalysisResult, Exception]]:
    """[(nc_audio, src_audio), ...] → one AnalysisResult (or the exception ``run`` would raise) per pair,
    computed stage by stage over sub-batches that stream through HBM (batch.analyse_batch): staging, host→device
    copies and analysis of different sub-batches overlap.  ``pairs`` is a sequence of numpy array pairs or a
    ``batch.PinnedBatch`` (``batch.pin_pairs``: the same pairs already laid out in pinned host memory, which skips the
    staging copy).  ``stats`` (optional dict) receives counters: windows, tracks, h2d_bytes, d2h_bytes, hop64_frames."""
    from . import batch
    return batch.analyse_batch(pairs, sr, sub_batch=sub_batch or batch.SUB_BATCH_PAIRS, workers=workers, stats=stats,
                               window_sec=window_sec, hop_sec=hop_sec, energy_gate_db=energy_gate_db,
                               silence_strip_db=silence_strip_db, compute_pitch=compute_pitch, compute_ibi=compute_ibi)
