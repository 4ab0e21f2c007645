"""
Ratio estimation, bootstrap confidence intervals, classification and Rubber Band parameters —
drop-in for the reference's ``nightcore_analyzer/consensus.py`` (consensus.py:52-608).

The 2000-resample bootstraps (consensus.py:255-266, :300-311) run on the GPU: libncfa's
``ncfa_bootstrap_ratio_batched`` reproduces numpy's ``Generator(PCG64).choice`` stream bit for
bit (resample indices, medians, ``np.percentile``).  The scalar decision logic
(classification thresholds, half-time flip, warning texts, rounding) stays on the host with the
reference's semantics; ``tests/test_consensus_host.py`` pins it to the reference's outputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _engine

# ── tunables (consensus.py:52-62) ────────────────────────────────────────────
N_BOOTSTRAP: int = 2000
CI_LEVEL: float = 0.95
PURE_NC_TOLERANCE: float = 0.02
MIN_VALID: int = 3

NIGHTCORE_RATIO_MIN: float = 1.05
NIGHTCORE_RATIO_MAX: float = 1.50
NEAR_UNITY_TOLERANCE: float = 0.05
WIDE_CI_RELATIVE: float = 2.0
DURATION_TEMPO_MISMATCH_TOLERANCE: float = 0.08

BOOTSTRAP_SEED = 42  # consensus.py:255,300


@dataclass
class AnalysisResult:
    """Full output of the windowed consensus pipeline (field-for-field consensus.py:66-162)."""

    tempo_ratio: float
    pitch_ratio: float
    tempo_ci: Tuple[float, float]
    pitch_ci: Tuple[float, float]
    classification: str
    n_source_pitch_windows: int
    n_nc_pitch_windows: int
    n_source_tempo_windows: int
    n_nc_tempo_windows: int
    rubberband: dict = field(default_factory=dict)
    src_pitches_raw: Optional[List[Optional[float]]] = None
    nc_pitches_raw: Optional[List[Optional[float]]] = None
    src_tempos_raw: Optional[List[Optional[float]]] = None
    nc_tempos_raw: Optional[List[Optional[float]]] = None
    nc_duration: Optional[float] = None
    src_duration: Optional[float] = None
    nc_median_bpm: Optional[float] = None
    src_median_bpm: Optional[float] = None
    warnings: List[str] = field(default_factory=list)
    pitch_method: Optional[str] = None
    ibi_ratio: Optional[float] = None
    ibi_ci: Optional[Tuple[float, float]] = None
    xcorr_ratio: Optional[float] = None
    xcorr_quality: Optional[float] = None
    intro_offset_sec: Optional[float] = None

    def __str__(self) -> str:  # same report as consensus.py:164-232
        rb = self.rubberband
        out: List[str] = [f"WARNING  : {w}" for w in self.warnings]
        if self.warnings:
            out.append("")
        out.append(f"Classification  : {self.classification}")

        tempo_line = (
            f"Tempo ratio     : {self.tempo_ratio:.6f}"
            f"  95% CI [{self.tempo_ci[0]:.6f}, {self.tempo_ci[1]:.6f}]"
            f"  (from {self.n_source_tempo_windows} src / {self.n_nc_tempo_windows} nc windows)"
        )
        if self.nc_duration and self.src_duration:
            tempo_line += (
                f"  |  duration ratio {self.src_duration / self.nc_duration:.6f}×"
                f" ({self.src_duration:.1f} s / {self.nc_duration:.1f} s)"
            )
        out.append(tempo_line)

        if self.n_source_pitch_windows > 0 or self.n_nc_pitch_windows > 0:
            out.append(
                f"Pitch ratio     : {self.pitch_ratio:.6f}"
                f"  95% CI [{self.pitch_ci[0]:.6f}, {self.pitch_ci[1]:.6f}]"
                f"  (from {self.n_source_pitch_windows} src / {self.n_nc_pitch_windows} nc samples)"
            )
            if self.pitch_method:
                out.append(f"Pitch method    : {self.pitch_method}")
        else:
            out.append("Pitch ratio     : not computed in this step")

        if self.tempo_ratio > 0:
            tr = self.tempo_ratio
            out += [
                "",
                f"Speed summary   : nightcore is {tr:.4f}× the source speed",
                f"                  to hear original tempo → play nightcore at {1.0 / tr:.4f}× speed",
                f"                  (source was sped up by {tr:.4f}× to create the nightcore)",
            ]

        if self.nc_median_bpm is not None and self.src_median_bpm is not None:
            out.append(
                f"Median BPMs     : nightcore {self.nc_median_bpm:.2f}  |"
                f"  source {self.src_median_bpm:.2f}"
                f"  (raw detected; ratio = {self.nc_median_bpm / self.src_median_bpm:.6f})"
            )

        out.append("")
        out.append(
            f"Rubber Band     : --time {rb.get('time_ratio', '?'):.6f}"
            f"  --pitch {rb.get('pitch_semitones', '?'):.4f} st"
            "  (beat-detected ratio)"
        )
        out.append(f"CLI (detected)  : {rb.get('cli_command', '')}")
        if rb.get("duration_time_ratio"):
            out.append(
                f"Duration-based  : --time {rb['duration_time_ratio']:.6f}"
                f"  --pitch {rb['duration_pitch_semitones']:.4f} st"
                "  (uses file-length ratio — prefer this when CI is degenerate)"
            )
            out.append(f"CLI (duration)  : {rb.get('duration_cli_command', '')}")
        return "\n".join(out)


# ── internals ────────────────────────────────────────────────────────────────
def _valid(values: Sequence[Optional[float]]) -> np.ndarray:
    """consensus.py:236-240 — drop None / non-finite / non-positive entries."""
    isfinite = math.isfinite      # same predicate as np.isfinite on a Python / numpy scalar, a tenth of the cost
    keep = [v for v in values if v is not None and isfinite(v) and v > 0]
    return np.array(keep, dtype=np.float64)


def _median_small(values) -> float:
    """np.median of a short 1-D sequence of finite float64 values, bit for bit (middle element, or the mean of the two
    middle elements computed as (a + b) / 2 like np.mean of two values) without numpy's per-call overhead — the batch
    path takes a few medians per pair, 150 µs of interpreter time each way through np.median."""
    v = sorted(values.tolist() if isinstance(values, np.ndarray) else [float(x) for x in values])
    n = len(v)
    h = n >> 1
    return v[h] if n & 1 else (v[h - 1] + v[h]) / 2.0


def _percentile_args(ci: float) -> Tuple[float, float]:
    alpha = (1.0 - ci) / 2.0
    return alpha * 100, (1.0 - alpha) * 100


def _device_bootstrap(first: np.ndarray, second: np.ndarray, n_boot: int, ci: float) -> Tuple[float, Tuple[float, float]]:
    """median(first)/median(second) with first resampled before second in every iteration."""
    first = np.asarray(first, dtype=np.float64)
    second = np.asarray(second, dtype=np.float64)
    if len(first) == 0 or len(second) == 0:
        raise ValueError("cannot bootstrap an empty array")   # numpy's choice raises on an empty population
    q_lo, q_hi = _percentile_args(ci)
    out, _, _ = _engine.get_engine().bootstrap([(first, second)], BOOTSTRAP_SEED, n_boot, q_lo, q_hi)
    return float(out[0, 0]), (float(out[0, 1]), float(out[0, 2]))


def _bootstrap_ratio(nc_vals: np.ndarray, src_vals: np.ndarray, n_boot: int = N_BOOTSTRAP, ci: float = CI_LEVEL,
                     ) -> Tuple[float, Tuple[float, float]]:
    """consensus.py:243-267 — median(nc)/median(src) and its bootstrap CI (nc drawn first)."""
    return _device_bootstrap(nc_vals, src_vals, n_boot, ci)


def compute_ibi_ratio(nc_ibis: np.ndarray, src_ibis: np.ndarray, n_boot: int = N_BOOTSTRAP, ci: float = CI_LEVEL,
                      ) -> Tuple[float, Tuple[float, float]]:
    """consensus.py:270-312 — median(src_ibis)/median(nc_ibis) and its CI (src drawn first)."""
    return _device_bootstrap(src_ibis, nc_ibis, n_boot, ci)


def _classify(tempo_ratio: float, pitch_ratio: float, tempo_ci: Tuple[float, float], pitch_ci: Tuple[float, float],
              tol: float = PURE_NC_TOLERANCE) -> str:
    """consensus.py:315-336."""
    diff = pitch_ratio - tempo_ratio
    overlapping = tempo_ci[0] <= pitch_ci[1] and pitch_ci[0] <= tempo_ci[1]
    if abs(diff) <= tol or (overlapping and abs(diff) <= 2 * tol):
        return "pure_nightcore"
    if diff > tol:
        return "independent_pitch_shift"
    if tempo_ratio > 1.0 + tol and diff < -tol:
        return "time_stretch_only"
    return "ambiguous"


def _rb_command(time_ratio: float, pitch_st: float) -> str:
    return f"rubberband --time {time_ratio:.6f} --pitch {pitch_st:.4f} nightcore.flac reconstructed.flac"


def _rubberband_params(tempo_ratio: float, pitch_ratio: float, nc_duration: Optional[float] = None,
                       src_duration: Optional[float] = None) -> dict:
    """consensus.py:339-381 — parameters that rebuild the original from the nightcore."""
    pitch_st = -12.0 * math.log2(pitch_ratio)
    params = {
        "time_ratio": round(tempo_ratio, 6),
        "pitch_semitones": round(pitch_st, 4),
        "nc_to_source_speed": round(1.0 / tempo_ratio, 6) if tempo_ratio != 0 else None,
        "cli_command": _rb_command(tempo_ratio, pitch_st),
    }
    if nc_duration and src_duration and nc_duration > 0:
        by_length = src_duration / nc_duration
        by_length_st = -12.0 * math.log2(by_length)
        params["duration_time_ratio"] = round(by_length, 6)
        params["duration_pitch_semitones"] = round(by_length_st, 4)
        params["duration_cli_command"] = _rb_command(by_length, by_length_st)
    return params


def _check_sanity(tempo_ratio: float, pitch_ratio: float, tempo_ci: Tuple[float, float],
                  pitch_ci: Tuple[float, float], nc_duration: Optional[float] = None,
                  src_duration: Optional[float] = None, tempo_was_corrected: bool = False) -> List[str]:
    """consensus.py:384-515 — human-readable warnings for suspicious results (same texts, same order)."""
    notes: List[str] = []
    have_durations = nc_duration is not None and src_duration is not None

    if tempo_was_corrected:
        notes.append(
            f"Beat-tracker half-time artefact corrected: librosa returned a raw tempo "
            f"ratio < 1 (nightcore beat-detected at half-time), but the nightcore file "
            f"({nc_duration:.1f} s) is shorter than the source ({src_duration:.1f} s), "
            "confirming the nightcore IS faster. The ratio has been inverted "
            f"to {tempo_ratio:.4f}× automatically. This is a known librosa artefact "
            "for high-BPM music (>~130 BPM)."
        )
    elif have_durations:
        if abs(nc_duration / src_duration - 1.0) < NEAR_UNITY_TOLERANCE:
            notes.append(
                f"Both files are nearly the same duration "
                f"({nc_duration:.1f} s vs {src_duration:.1f} s). "
                "Did you accidentally provide two nightcore files, or two originals? "
                "A real nightcore should be ~10–35 % shorter than the source."
            )
    elif abs(tempo_ratio - 1.0) < NEAR_UNITY_TOLERANCE:
        notes.append(
            f"Tempo ratio is {tempo_ratio:.4f} — both files appear to be at the "
            "same speed. Did you accidentally provide two nightcore files, or two "
            "originals? A real nightcore should be 1.05–1.50× faster than the source."
        )
    elif tempo_ratio < 1.0:
        inverse = round(1.0 / tempo_ratio, 4)
        notes.append(
            f"Tempo ratio is {tempo_ratio:.4f} < 1.0. Two possible causes: "
            "(1) librosa half-time detection artefact — the true ratio may be "
            f"{inverse:.4f}× (the inverse); or (2) the files are in the wrong order. "
            "Re-run with the correct original FLAC as --source to disambiguate."
        )
    elif tempo_ratio > NIGHTCORE_RATIO_MAX:
        notes.append(
            f"Tempo ratio is {tempo_ratio:.4f}, above the typical nightcore range "
            f"({NIGHTCORE_RATIO_MIN}–{NIGHTCORE_RATIO_MAX}×). Verify the input files."
        )

    if have_durations:
        length_ratio = src_duration / nc_duration
        gap = abs(length_ratio - tempo_ratio) / tempo_ratio
        if gap > DURATION_TEMPO_MISMATCH_TOLERANCE:
            notes.append(
                f"Duration ratio ({length_ratio:.4f}×) and detected tempo ratio "
                f"({tempo_ratio:.4f}×) differ by {gap * 100:.1f}%. For a pure "
                "speed-up these should be nearly equal. Most likely cause: the two files "
                "are different edits or versions of the same song (e.g. radio edit vs. "
                "extended mix). Find the exact version used to create the nightcore, or "
                f"use the duration ratio ({length_ratio:.4f}×) directly as the "
                "rubberband --time factor."
            )

    if abs(tempo_ci[1] - tempo_ci[0]) < 0.001:
        if have_durations and nc_duration > 0:
            length_ratio = src_duration / nc_duration
            gap = abs(tempo_ratio - length_ratio) / length_ratio
            if gap < DURATION_TEMPO_MISMATCH_TOLERANCE:
                notes.append(
                    f"Tempo CI is degenerate [lo = hi = {tempo_ci[0]:.6f}]: every "
                    "analysis window returned the same BPM. This is expected for "
                    "constant-tempo music (drum machine / eurodance). The detected "
                    f"ratio ({tempo_ratio:.4f}×) agrees with the duration ratio "
                    f"({length_ratio:.4f}×) — result is reliable."
                )
            else:
                notes.append(
                    f"Tempo CI is degenerate [lo = hi = {tempo_ci[0]:.6f}] and the "
                    f"detected ratio ({tempo_ratio:.4f}×) disagrees with the duration "
                    f"ratio ({length_ratio:.4f}×) by {gap * 100:.1f}%. "
                    "This is a librosa BPM quantisation artefact — the beat tracker "
                    "snapped all windows to the same wrong grid BPM. "
                    "Use the 'Duration-based' CLI command instead of 'CLI (detected)'."
                )
        else:
            notes.append(
                f"Tempo CI is degenerate [lo = hi = {tempo_ci[0]:.6f}]: every "
                "analysis window returned the same BPM from librosa. This may be a "
                "quantisation artefact (beat tracker snapped to a fixed grid BPM) or "
                "simply a constant-tempo track. Provide both file durations to "
                "distinguish the two cases."
            )

    if pitch_ratio > 0 and (pitch_ci[1] - pitch_ci[0]) > WIDE_CI_RELATIVE * pitch_ratio:
        notes.append(
            f"Pitch CI is very wide ({pitch_ci[0]:.3f}–{pitch_ci[1]:.3f}) relative "
            f"to the point estimate ({pitch_ratio:.4f}). The pitch estimator could "
            "not reliably determine a consistent pitch ratio — this is common with "
            "polyphonic or heavily processed audio. "
            "Trust the tempo ratio; treat the pitch ratio and classification as "
            "approximate."
        )
    return notes


def _assemble(src_pitches, nc_pitches, src_tempos, nc_tempos, src_t, nc_t, pitch_stats, tempo_stats, n_pitch,
              nc_duration, src_duration) -> AnalysisResult:
    """Everything of build_result after the bootstraps (consensus.py:562-608)."""
    pitch_ratio, pitch_ci = pitch_stats
    tempo_ratio, tempo_ci = tempo_stats
    flipped = False
    if (nc_duration is not None and src_duration is not None and nc_duration < src_duration * 0.99
            and tempo_ratio < 1.0):
        # half-time artefact: the shorter file cannot be the slower one
        tempo_ratio = 1.0 / tempo_ratio
        lo, hi = tempo_ci
        tempo_ci = (1.0 / hi, 1.0 / lo)
        flipped = True
    return AnalysisResult(
        tempo_ratio=tempo_ratio,
        pitch_ratio=pitch_ratio,
        tempo_ci=tempo_ci,
        pitch_ci=pitch_ci,
        classification=_classify(tempo_ratio, pitch_ratio, tempo_ci, pitch_ci),
        n_source_pitch_windows=n_pitch[0],
        n_nc_pitch_windows=n_pitch[1],
        n_source_tempo_windows=len(src_t),
        n_nc_tempo_windows=len(nc_t),
        rubberband=_rubberband_params(tempo_ratio, pitch_ratio, nc_duration, src_duration),
        nc_duration=nc_duration,
        src_duration=src_duration,
        nc_median_bpm=_median_small(nc_t) if len(nc_t) > 0 else None,
        src_median_bpm=_median_small(src_t) if len(src_t) > 0 else None,
        warnings=_check_sanity(tempo_ratio, pitch_ratio, tempo_ci, pitch_ci, nc_duration, src_duration, flipped),
        src_pitches_raw=list(src_pitches),
        nc_pitches_raw=list(nc_pitches),
        src_tempos_raw=list(src_tempos),
        nc_tempos_raw=list(nc_tempos),
    )


def _insufficient(src_t, nc_t) -> ValueError:
    return ValueError(
        f"Insufficient valid tempo windows (source: {len(src_t)}, "
        f"nightcore: {len(nc_t)}).  Need ≥ {MIN_VALID} each."
    )


# ── public API ───────────────────────────────────────────────────────────────
def build_result(src_pitches: List[Optional[float]], nc_pitches: List[Optional[float]],
                 src_tempos: List[Optional[float]], nc_tempos: List[Optional[float]], *,
                 nc_duration: Optional[float] = None, src_duration: Optional[float] = None) -> AnalysisResult:
    """consensus.py:519-608 — consensus step over per-window estimates."""
    src_p, nc_p = _valid(src_pitches), _valid(nc_pitches)
    src_t, nc_t = _valid(src_tempos), _valid(nc_tempos)
    if len(src_t) < MIN_VALID or len(nc_t) < MIN_VALID:
        raise _insufficient(src_t, nc_t)
    if len(src_p) >= MIN_VALID and len(nc_p) >= MIN_VALID:
        pitch_stats = _bootstrap_ratio(nc_p, src_p)
        n_pitch = (len(src_p), len(nc_p))
    else:
        pitch_stats = (1.0, (1.0, 1.0))
        n_pitch = (0, 0)
    tempo_stats = _bootstrap_ratio(nc_t, src_t)
    return _assemble(src_pitches, nc_pitches, src_tempos, nc_tempos, src_t, nc_t, pitch_stats, tempo_stats, n_pitch,
                     nc_duration, src_duration)
