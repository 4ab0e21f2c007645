"""
Full windowed consensus pipeline — drop-in for the reference's ``nightcore_analyzer/pipeline.py``
(pipeline.py:23-216: same ``run`` signature, stage order, log lines, exceptions and result object).

``run`` follows the reference stage by stage through the drop-in modules (io → pitch → tempo →
consensus → IBI), each of which batches its work onto the GPU.  ``run_arrays`` is the same pipeline
from already-loaded arrays; ``run_batch`` analyses many pairs in one data-parallel pass
(``batch.analyse_staged``) and is what ``bench.py`` measures.
"""
from __future__ import annotations

from dataclasses import dataclass
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


@dataclass
class _Pair:
    """The two tracks of one analysis as they move through the stages (each stage may shorten them)."""
    nc: np.ndarray
    src: np.ndarray
    sr: int
    say: Callable[[str], None]
    intro_offset_sec: Optional[float] = None

    @property
    def nc_sec(self) -> float:
        return len(self.nc) / self.sr

    @property
    def src_sec(self) -> float:
        return len(self.src) / self.sr


def _silent(_msg: str) -> None:
    pass


def _trim_silence(p: _Pair, top_db: Optional[float]) -> None:
    """pipeline.py:91-104 — leading / trailing silence off both tracks."""
    if top_db is None:
        return
    p.say(f"Stripping silence (top_db={top_db} dB)…")
    p.nc, nc_lead, nc_trail = strip_silence(p.nc, p.sr, top_db)
    p.src, src_lead, src_trail = strip_silence(p.src, p.sr, top_db)
    for label, lead, trail, sec in (("nightcore:", nc_lead, nc_trail, p.nc_sec), ("source:   ", src_lead, src_trail, p.src_sec)):
        p.say(f"  {label} −{lead:.2f}s leading, −{trail:.2f}s trailing  →  {sec:.1f} s")


def _skip_intro(p: _Pair, src_trim_sec: float, auto_align: bool) -> None:
    """pipeline.py:106-125 — drop an intro the source has and the nightcore does not: a manual trim wins over the
    RMS-envelope alignment, which is only trusted from ALIGN_MIN_OFFSET seconds up."""
    if src_trim_sec > 0.0:
        p.src = p.src[int(src_trim_sec * p.sr):]
        p.intro_offset_sec = src_trim_sec
        p.say(f"Manual source trim: skipping {src_trim_sec:.2f}s from source start")
        return
    if not auto_align:
        return
    p.say("Detecting intro offset (RMS envelope alignment)…")
    offset, speed_hint = find_content_offset(p.src, p.nc, p.sr)
    if offset < ALIGN_MIN_OFFSET:
        p.say(f"  No significant intro offset detected  (raw: {offset:.2f}s < {ALIGN_MIN_OFFSET:.1f}s threshold)")
        return
    p.src = p.src[int(offset * p.sr):]
    p.intro_offset_sec = offset
    p.say(f"  Intro detected — trimming {offset:.2f}s from source start  (speed hint: {speed_hint:.4f}×)")


def _gated_windows(p: _Pair, window_sec: float, hop_sec: float, gate_db: float):
    """pipeline.py:127-146 — overlapping windows of both tracks, then the energy gate; an emptied side is fatal."""
    p.say(f"Slicing into {window_sec:.0f} s windows (hop {hop_sec:.0f} s)…")
    sliced = [slice_windows(track, p.sr, window_sec, hop_sec) for track in (p.nc, p.src)]
    p.say(f"  nightcore: {len(sliced[0])} windows  |  source: {len(sliced[1])} windows")
    p.say(f"Energy gating (threshold {gate_db} dB below peak)…")
    nc_windows, src_windows = (energy_gate(ws, gate_db) for ws in sliced)
    p.say(f"  after gating — nightcore: {len(nc_windows)} windows  |  source: {len(src_windows)} windows")
    if not nc_windows or not src_windows:
        raise RuntimeError(
            "All windows were discarded by the energy gate.  "
            "Try raising --energy-gate (e.g. --energy-gate -60)."
        )
    return nc_windows, src_windows


def _pitches(p: _Pair, compute_pitch: bool):
    """pipeline.py:148-159 → (source Hz per chunk, nightcore Hz per chunk, method name | None)."""
    if not compute_pitch:
        p.say("Skipping pitch estimation.")
        return [], [], None
    p.say("Estimating pitch (chromagram cross-correlation)…")
    src_hz, nc_hz, method = estimate_pitch_combined(p.src, p.nc, p.sr, log=p.say)
    p.say(f"  Pitch method: {method}")
    return src_hz, nc_hz, method


def _tempos(p: _Pair, nc_windows, src_windows):
    """pipeline.py:161-186 — source windows first (120 BPM prior); their median, scaled by the duration ratio, is the
    nightcore windows' prior.  → (source tempos, nightcore tempos, nightcore prior)."""
    p.say("Estimating tempo (librosa)…")
    p.say("  ← source →")
    src_tempos = batch_estimate_tempo(src_windows, log=p.say)
    nc_prior = 120.0
    confident = [t for t in src_tempos if t is not None]
    if confident and p.nc_sec > 0 and p.src_sec > 0:
        src_median = float(np.median(confident))
        stretch = p.src_sec / p.nc_sec
        nc_prior = src_median * stretch
        p.say(f"  NC tempo prior: {nc_prior:.1f} BPM  (src median {src_median:.1f} BPM × dur ratio {stretch:.4f})")
    p.say("  ← nightcore →")
    return src_tempos, batch_estimate_tempo(nc_windows, log=p.say, start_bpm=nc_prior), nc_prior


def _attach_ibi(p: _Pair, result: AnalysisResult, nc_prior: float) -> None:
    """pipeline.py:199-213 — whole-track beat timing at hop 64; the ratio is attached only when both sides yield at
    least four inter-beat intervals."""
    p.say("Computing IBI ratio (high-precision beat timestamps, hop=64)…")
    nc_ibis = estimate_ibis_global(p.nc, p.sr, start_bpm=nc_prior)
    src_ibis = estimate_ibis_global(p.src, p.sr)
    if any(x is None or len(x) < 4 for x in (nc_ibis, src_ibis)):
        p.say("  IBI ratio: insufficient beats — skipped")
        return
    result.ibi_ratio, result.ibi_ci = compute_ibi_ratio(nc_ibis, src_ibis)
    p.say(f"  IBI ratio: {result.ibi_ratio:.6f}×  95% CI [{result.ibi_ci[0]:.6f}, {result.ibi_ci[1]:.6f}]")


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
    """pipeline.py:91-216 — everything of ``run`` after the two files are loaded: the stages above in the reference's
    order, with the reference's log lines."""
    p = _Pair(nc_audio, src_audio, sr, log if log is not None else _silent)
    _trim_silence(p, silence_strip_db)
    _skip_intro(p, src_trim_sec, auto_align)
    nc_windows, src_windows = _gated_windows(p, window_sec, hop_sec, energy_gate_db)
    src_pitches, nc_pitches, pitch_method = _pitches(p, compute_pitch)
    src_tempos, nc_tempos, nc_prior = _tempos(p, nc_windows, src_windows)
    p.say("Computing consensus…")
    result = build_result(src_pitches, nc_pitches, src_tempos, nc_tempos, nc_duration=p.nc_sec, src_duration=p.src_sec)
    result.intro_offset_sec = p.intro_offset_sec
    result.pitch_method = pitch_method
    _attach_ibi(p, result, nc_prior)
    p.say("Done.")
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
    say = log if log is not None else _silent
    sr = None
    tracks = []
    for label, path in (("nightcore", nightcore_path), ("source", source_path)):     # pipeline.py:78-89
        say(f"Loading {label} audio…")
        audio, sr = load_audio(path) if sr is None else load_audio(path, sr=sr)
        say(f"  {len(audio) / sr:.1f} s  ({len(audio):,} samples @ {sr} Hz)")
        tracks.append(audio)
    nc_audio, src_audio = tracks
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
