"""
Whole-file spectral statistics — drop-in for the analysis half of the reference's ``nightcore_analyzer/spectral.py``
(spectral.py:21-103: ``SpectralStats`` and ``analyze``).  SURVEY.md §8(f) row 3.

The STFT-derived numbers (centroid, 85 % roll-off, five band means, per-bin mean dB for the effective bandwidth) come
from libncfa's ``ncfa_spectral_stats_batched`` (the un-reduced warp-FFT kernel with reduction epilogues); the framed RMS
from ``ncfa_rms_frames``.  The plain-English report of the reference (``compare_and_print`` and its transcode
heuristics, spectral.py:113-359) is presentation code and is not rebuilt.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _engine
from .io import load_audio


@dataclass
class SpectralStats:
    """spectral.py:21-34 — same fields, same order."""
    centroid: float
    rolloff: float
    rms_mean: float
    rms_variance: float
    sub_bass: float
    bass: float
    midrange: float
    presence: float
    brilliance: float
    decay_rate: float
    duration: float
    effective_bandwidth_hz: float


_BANDS = ((20, 80), (80, 250), (250, 2000), (2000, 6000), (6000, 20000))   # spectral.py:74-78


def analyze_arrays(tracks: Sequence[np.ndarray], sr: int) -> list:
    """spectral.analyze for already-loaded mono float32 arrays at rate *sr* → [SpectralStats, ...] (one batched pass)."""
    eng = _engine.get_engine()
    arrays = [np.ascontiguousarray(t, dtype=np.float32) for t in tracks]
    audio, off, ln = eng.pack(arrays)
    stats, bins = eng.spectral_stats_dev(audio, off, ln, sr)
    stats, bins = eng.to_host(stats), eng.to_host(bins)
    freqs = np.fft.rfftfreq(n=2048, d=1.0 / sr)                             # librosa.fft_frequencies
    out = []
    for i, y in enumerate(arrays):
        n_frames = stats[i, 7]
        band = []
        for b, (lo, hi) in enumerate(_BANDS):
            n_bins = int(np.count_nonzero((freqs >= lo) & (freqs < hi)))
            band.append(float(np.float32(stats[i, 2 + b] / (n_bins * n_frames))) if n_bins else 0.0)
        rms = eng.to_host(eng.rms_frames_dev(audio[int(off[i]):], len(y), 2048, 512))
        loud = rms[rms > np.percentile(rms, 75)]
        freq_avg_db = bins[i]
        significant = freq_avg_db > (np.max(freq_avg_db) - 60.0)
        bw = float(freqs[np.where(significant)[0][-1]]) if significant.any() else float(freqs[-1])
        out.append(SpectralStats(
            centroid=float(stats[i, 0] / n_frames), rolloff=float(stats[i, 1] / n_frames),
            rms_mean=float(np.mean(rms)), rms_variance=float(np.var(rms)),
            sub_bass=band[0], bass=band[1], midrange=band[2], presence=band[3], brilliance=band[4],
            decay_rate=float(np.mean(np.diff(loud))) if len(loud) > 1 else 0.0,
            duration=len(y) / float(sr), effective_bandwidth_hz=bw))
    return out


def analyze(path: str, label: Optional[str] = None) -> SpectralStats:
    """spectral.py:38-103 — load *path* at its native rate and return its spectral statistics."""
    if label:
        print(f"  Loading {label}…")
    y, sr = load_audio(path, sr=None)
    return analyze_arrays([y], sr)[0]
