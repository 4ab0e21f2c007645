"""
Pitch-shift estimation by CQT-chroma cyclic cross-correlation — drop-in for the reference's
``nightcore_analyzer/pitch.py`` (pitch.py:42-291: same names, signatures, constants, return tuples
and log lines).

``_mean_chroma`` runs on the GPU (libncfa: tuning estimate → 7-octave CQT contraction → chroma fold →
inf-norm → time mean) and so do the 12-lag cyclic cross-correlation and the seed-0 chunk bootstrap
(PCG64-exact).  As in the reference, ``chroma_cqt`` is called with ``bins_per_octave=36`` but
librosa's default ``n_chroma=12``, so the mean chroma has 12 bins and the lag is divided by 3
(pitch.py:58-64,95) — reproduced, not fixed.  MELODIA needs essentia, which is not
available here: ``estimate_pitch_melodia`` logs the reference's message and returns ``None``.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _engine

# ── tunables (pitch.py:42-50) ────────────────────────────────────────────────
CHROMA_BINS_PER_OCTAVE: int = 36
CHROMA_HOP_LENGTH: int = 512
CHUNK_SEC: float = 20.0
MIN_CHUNKS: int = 3
MELODIA_AGREE_ST: float = 1.5
MAX_MELODIA_FRAMES: int = 2000

_REF_HZ: float = 440.0
N_CHROMA: int = 12            # librosa.feature.chroma_cqt default (not passed at pitch.py:58-63)
PITCH_BOOTSTRAP_SEED = 0      # pitch.py:144
PITCH_N_BOOT = 2000           # pitch.py:147


# ── chroma helpers ───────────────────────────────────────────────────────────
def _mean_chroma(audio: np.ndarray, sr: int) -> np.ndarray:
    """pitch.py:55-64 — time-averaged CQT chroma vector, shape (12,), float64."""
    eng = _engine.get_engine()
    a = np.ascontiguousarray(audio, dtype=np.float32)
    d, off, ln = eng.pack([a])
    chroma, _ = eng.chroma_mean_dev(d, off, ln, sr)
    return eng.to_host(chroma)[0]


def _cyclic_xcorr_peak(src_chroma: np.ndarray, nc_chroma: np.ndarray) -> int:
    """pitch.py:67-85 — lag (in chroma bins) at which *nc_chroma* best aligns with *src_chroma*."""
    eng = _engine.get_engine()
    s = eng.to_dev(np.ascontiguousarray(src_chroma, dtype=np.float64)[None, :])
    n = eng.to_dev(np.ascontiguousarray(nc_chroma, dtype=np.float64)[None, :])
    return int(eng.to_host(eng.cyclic_xcorr_dev(s, n))[0])


def _chroma_shift_for_chunk(src_chunk: np.ndarray, nc_chunk: np.ndarray, sr: int) -> float:
    """pitch.py:88-95."""
    lag = _cyclic_xcorr_peak(_mean_chroma(src_chunk, sr), _mean_chroma(nc_chunk, sr))
    return lag / 3.0


def _chunk_bounds(n_src: int, n_nc: int, sr: int) -> List[Tuple[int, int, int, int]]:
    """pitch.py:120-138 → [(src_start, src_len, nc_start, nc_len)]: same-index 20 s chunks, or the
    whole files when either is shorter than one chunk."""
    chunk_n = int(CHUNK_SEC * sr)
    n_chunks = min(n_src // chunk_n, n_nc // chunk_n) if chunk_n > 0 else 0
    if n_chunks < 1:
        return [(0, n_src, 0, n_nc)]
    return [(i * chunk_n, chunk_n, i * chunk_n, chunk_n) for i in range(n_chunks)]


def chroma_shifts_staged(audio, jobs: Sequence[tuple], sr: int, return_lags: bool = False):
    """Batched core shared by ``estimate_pitch_chroma`` and ``batch.analyse_staged``.

    ``audio`` is one device buffer; each job is ``(tag, src_off, src_len, nc_off, nc_len)`` in samples.
    Every chunk of every job goes through one tuning/CQT pass and one cyclic-xcorr launch.  Returns
    one float64 array of per-chunk semitone shifts per job (and the integer lags when asked)."""
    eng = _engine.get_engine()
    seg_off, seg_len, owner = [], [], []
    for j, (_tag, s_off, s_len, n_off, n_len) in enumerate(jobs):
        for (ss, sl, ns, nl) in _chunk_bounds(int(s_len), int(n_len), sr):
            seg_off += [int(s_off) + ss, int(n_off) + ns]
            seg_len += [sl, nl]
            owner.append(j)
    if not owner:
        return ([], []) if return_lags else []
    seg_off = np.asarray(seg_off, dtype=np.int64)
    seg_len = np.asarray(seg_len, dtype=np.int32)
    # stable sort by length keeps sub-batches of equal-sized segments together (all 20 s chunks are equal)
    chroma, _ = eng.chroma_mean_dev(audio, seg_off, seg_len, sr)
    lags = eng.to_host(eng.cyclic_xcorr_dev(chroma[0::2], chroma[1::2])).astype(np.int64)
    owner = np.asarray(owner)
    shifts = [lags[owner == j] / 3.0 for j in range(len(jobs))]
    if return_lags:
        return shifts, [lags[owner == j] for j in range(len(jobs))]
    return shifts


def hz_lists(shift_sts: np.ndarray) -> Tuple[List[Optional[float]], List[Optional[float]]]:
    """pitch.py:159-163: implied Hz pairs for consensus._bootstrap_ratio."""
    src_hz: List[Optional[float]] = [_REF_HZ] * len(shift_sts)
    nc_hz: List[Optional[float]] = [_REF_HZ * (2.0 ** (st / 12.0)) for st in shift_sts]
    return src_hz, nc_hz


# ── stage 1: chroma cross-correlation ────────────────────────────────────────
def estimate_pitch_chroma(
    src_audio: np.ndarray,
    nc_audio: np.ndarray,
    sr: int,
    log: Optional[Callable[[str], None]] = None,
) -> Tuple[List[Optional[float]], List[Optional[float]], float, Tuple[float, float], int]:
    """pitch.py:100-173."""
    eng = _engine.get_engine()
    s = np.ascontiguousarray(src_audio, dtype=np.float32)
    n = np.ascontiguousarray(nc_audio, dtype=np.float32)
    audio, off, ln = eng.pack([s, n])
    shift_sts = np.asarray(chroma_shifts_staged(audio, [(0, off[0], ln[0], off[1], ln[1])], sr)[0], dtype=np.float64)
    n_chunks = len(shift_sts)
    point_st = float(np.median(shift_sts))

    if n_chunks >= MIN_CHUNKS:
        # pitch.py:143-150: median of rng.choice(shift_sts) 2000×, default_rng(0), percentiles 2.5 / 97.5
        out, _, _ = eng.bootstrap([(shift_sts, None)], PITCH_BOOTSTRAP_SEED, PITCH_N_BOOT, 2.5, 97.5)
        ci_lo_st, ci_hi_st = float(out[0][1]), float(out[0][2])
    else:
        ci_lo_st = ci_hi_st = point_st
        if log:
            log(
                f"    Only {n_chunks} chunk(s) available (need ≥ {MIN_CHUNKS}) — "
                "pitch CI is degenerate; estimate may be less reliable."
            )

    src_hz, nc_hz = hz_lists(shift_sts)
    if log:
        log(
            f"    Chroma xcorr: {point_st:+.3f} st"
            f"  95% CI [{ci_lo_st:+.3f}, {ci_hi_st:+.3f}] st"
            f"  ({n_chunks} chunk{'s' if n_chunks != 1 else ''})"
        )
    return src_hz, nc_hz, point_st, (ci_lo_st, ci_hi_st), n_chunks


# ── stage 2: MELODIA refinement (optional in the reference; essentia is out of scope here) ──
def _try_import_essentia():
    """pitch.py:178-184."""
    try:
        import essentia.standard as es  # type: ignore[import]
        return es
    except Exception:
        return None


def estimate_pitch_melodia(
    src_audio: np.ndarray,
    nc_audio: np.ndarray,
    sr: int,
    log: Optional[Callable[[str], None]] = None,
) -> Optional[Tuple[List[Optional[float]], List[Optional[float]]]]:
    """pitch.py:187-241.  The MELODIA branch is out of scope: without essentia
    the reference logs one line and returns None, and so does this."""
    if log:
        log("    essentia not available — skipping MELODIA refinement")
    return None


# ── public API ───────────────────────────────────────────────────────────────
def estimate_pitch_combined(
    src_audio: np.ndarray,
    nc_audio: np.ndarray,
    sr: int,
    log: Optional[Callable[[str], None]] = None,
) -> Tuple[List[Optional[float]], List[Optional[float]], str]:
    """pitch.py:246-291."""
    src_chroma_hz, nc_chroma_hz, chroma_st, _, _ = estimate_pitch_chroma(src_audio, nc_audio, sr, log=log)
    melodia_result = estimate_pitch_melodia(src_audio, nc_audio, sr, log=log)
    if melodia_result is not None:  # unreachable without essentia; kept for signature fidelity
        src_mel_hz, nc_mel_hz = melodia_result
        src_med = float(np.median([v for v in src_mel_hz if v is not None]))
        nc_med = float(np.median([v for v in nc_mel_hz if v is not None]))
        if src_med > 0 and nc_med > 0:
            melodia_st = 12.0 * math.log2(nc_med / src_med)
            if abs(melodia_st - chroma_st) <= MELODIA_AGREE_ST:
                return src_mel_hz, nc_mel_hz, "chroma+melodia"
            if log:
                log(
                    f"    MELODIA ({melodia_st:+.3f} st) disagrees with chroma"
                    f" ({chroma_st:+.3f} st) by"
                    f" {abs(melodia_st - chroma_st):.2f} st"
                    f" > {MELODIA_AGREE_ST} st threshold — using chroma only"
                )
    return src_chroma_hz, nc_chroma_hz, "chroma_xcorr"
