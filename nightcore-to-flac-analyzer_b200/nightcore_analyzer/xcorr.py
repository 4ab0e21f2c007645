"""
Windowed cross-correlation speed estimator and intro alignment — drop-in for the reference's
``nightcore_analyzer/xcorr.py`` (xcorr.py:33-268: same names, signatures, constants and sentinels).

``estimate_speed_xcorr`` keeps the reference's host logic (edge trim, window positions, search
bounds, ``np.polyfit`` on ≤ 20 correspondences, median quality) and runs the candidate loop —
20 windows × ~64 strided normalised dot products of 3 s of audio, xcorr.py:113-148 — as one batched
GPU launch (libncfa ``ncfa_xcorr_search_batched``).  The reference is a strided direct search, not an
FFT correlation; the observable result is the argmax on that candidate grid.
``find_content_offset`` (xcorr.py:165-259) evaluates its 30-speed envelope correlation on the GPU too
(``ncfa_align_search``).
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Sequence, Tuple, Union

import numpy as np

from . import _engine
from .io import load_audio

# ── tunables (xcorr.py:33-51) ────────────────────────────────────────────────
XCORR_SR: int = 22050
XCORR_N_WINDOWS: int = 20
XCORR_WINDOW_SEC: float = 3.0
XCORR_SEARCH_RANGE: float = 0.05
XCORR_SKIP_EDGES: float = 0.10
XCORR_RMS_GATE: float = 1e-3

XCORR_QUALITY_GOOD: float = 0.70
XCORR_QUALITY_FAIR: float = 0.40

ALIGN_SR: int = 11025
ALIGN_HOP: int = 512
ALIGN_SPEED_LO: float = 1.03
ALIGN_SPEED_HI: float = 1.50
ALIGN_N_SPEEDS: int = 30
ALIGN_MAX_OFFSET: float = 120.0
ALIGN_MIN_OFFSET: float = 1.0


def _search_tables(len_a: int, len_b: int, sr: int, n_windows: int, window_sec: float, search_range: float,
                   skip_edges: float):
    """Host part of xcorr.py:95-125 for one pair → (s, la, lb, win, stride, a_positions, lo_b, n_cand) or None
    when a trimmed file is shorter than one window (the reference returns (1.0, 0.0))."""
    min_len = min(len_a, len_b)
    s = int(min_len * skip_edges)
    e = int(min_len * (1.0 - skip_edges))
    la = max(0, min(len_a, e) - s)   # len(ya[s:e])
    lb = max(0, min(len_b, e) - s)
    win = int(window_sec * sr)
    search = int(search_range * lb)
    stride = max(1, win // 4)
    if la < win or lb < win:
        return None
    a_positions = np.linspace(0, la - win, n_windows).astype(int)
    lo = np.zeros(len(a_positions), dtype=np.int64)
    n_cand = np.zeros(len(a_positions), dtype=np.int32)
    for i, pa in enumerate(a_positions):
        expected_pb = int(pa * lb / la)
        lo_b = max(0, expected_pb - search)
        hi_b = min(lb - win, expected_pb + search)
        lo[i] = lo_b
        n_cand[i] = 0 if lo_b >= hi_b else len(range(lo_b, hi_b, stride))
    return s, la, lb, win, stride, a_positions.astype(np.int64), lo, n_cand


def _fit(a_positions: np.ndarray, lo: np.ndarray, stride: int, best_j: np.ndarray, best_c: np.ndarray
         ) -> Tuple[float, float]:
    """xcorr.py:146-162."""
    keep = best_j >= 0
    if int(keep.sum()) < 3:
        return 1.0, 0.0
    a_arr = a_positions[keep].astype(float)
    b_arr = (lo[keep] + best_j[keep].astype(np.int64) * stride).astype(float)
    slope = float(np.polyfit(a_arr, b_arr, 1)[0])
    quality = float(np.median(best_c[keep]))
    return slope, quality


def estimate_speed_xcorr_batch(
    pairs: Sequence[Tuple[np.ndarray, np.ndarray]],
    sr: int = XCORR_SR,
    n_windows: int = XCORR_N_WINDOWS,
    window_sec: float = XCORR_WINDOW_SEC,
    search_range: float = XCORR_SEARCH_RANGE,
    skip_edges: float = XCORR_SKIP_EDGES,
    return_indices: bool = False,
):
    """[(ya, yb), ...] already loaded at *sr* → [(ratio, quality), ...]; every pair's windows go through
    one launch.  With ``return_indices`` also the per-window (a_pos, best_pb | -1) tables."""
    eng = _engine.get_engine()
    arrays = [np.ascontiguousarray(t, dtype=np.float32) for p in pairs for t in p]
    audio, off, ln = eng.pack(arrays)
    out: List = [None] * len(pairs)
    tables = []
    groups: dict = {}
    for i in range(len(pairs)):
        tb = _search_tables(int(ln[2 * i]), int(ln[2 * i + 1]), sr, n_windows, window_sec, search_range, skip_edges)
        tables.append(tb)
        if tb is None:
            out[i] = ((1.0, 0.0), (np.zeros(0, np.int64), np.zeros(0, np.int64))) if return_indices else (1.0, 0.0)
        else:
            groups.setdefault((tb[3], tb[4]), []).append(i)
    for (win, stride), idxs in groups.items():
        a_pos = np.concatenate([off[2 * i] + tables[i][0] + tables[i][5] for i in idxs])
        b_lo = np.concatenate([off[2 * i + 1] + tables[i][0] + tables[i][6] for i in idxs])
        n_cand = np.concatenate([tables[i][7] for i in idxs])
        bj, bc = eng.xcorr_search_dev(audio, audio, a_pos, b_lo, n_cand, win, stride, XCORR_RMS_GATE)
        bj, bc = eng.to_host(bj), eng.to_host(bc)
        p = 0
        for i in idxs:
            tb = tables[i]
            k = len(tb[5])
            res = _fit(tb[5], tb[6], stride, bj[p : p + k], bc[p : p + k])
            if return_indices:
                best_pb = np.where(bj[p : p + k] >= 0, tb[6] + bj[p : p + k].astype(np.int64) * stride, -1)
                out[i] = (res, (tb[5].copy(), best_pb))
            else:
                out[i] = res
            p += k
    return out


def estimate_speed_xcorr_arrays(ya: np.ndarray, yb: np.ndarray, sr: int = XCORR_SR, n_windows: int = XCORR_N_WINDOWS,
                                window_sec: float = XCORR_WINDOW_SEC, search_range: float = XCORR_SEARCH_RANGE,
                                skip_edges: float = XCORR_SKIP_EDGES) -> Tuple[float, float]:
    """xcorr.py:95-162 on arrays already loaded at *sr*."""
    return estimate_speed_xcorr_batch([(ya, yb)], sr, n_windows, window_sec, search_range, skip_edges)[0]


def estimate_speed_xcorr(
    path_a: Union[str, Path],
    path_b: Union[str, Path],
    sr: int = XCORR_SR,
    n_windows: int = XCORR_N_WINDOWS,
    window_sec: float = XCORR_WINDOW_SEC,
    search_range: float = XCORR_SEARCH_RANGE,
    skip_edges: float = XCORR_SKIP_EDGES,
) -> Tuple[float, float]:
    """xcorr.py:54-162 — estimate speed_A / speed_B; ``(1.0, 0.0)`` with fewer than 3 correspondences."""
    ya, _ = load_audio(str(path_a), sr=sr)
    yb, _ = load_audio(str(path_b), sr=sr)
    return estimate_speed_xcorr_arrays(ya, yb, sr, n_windows, window_sec, search_range, skip_edges)


def find_content_offset(
    src_audio: np.ndarray,
    nc_audio: np.ndarray,
    sr: int,
    *,
    speed_lo: float = ALIGN_SPEED_LO,
    speed_hi: float = ALIGN_SPEED_HI,
    n_speeds: int = ALIGN_N_SPEEDS,
    max_offset_sec: float = ALIGN_MAX_OFFSET,
) -> Tuple[float, float]:
    """xcorr.py:165-259 — seconds of *src_audio* that precede the content matching the start of
    *nc_audio*, and the speed at which the envelopes matched best."""
    eng = _engine.get_engine()
    src_env = eng.align_envelope(np.ascontiguousarray(src_audio, dtype=np.float32), sr, ALIGN_SR, ALIGN_HOP)
    nc_env = eng.align_envelope(np.ascontiguousarray(nc_audio, dtype=np.float32), sr, ALIGN_SR, ALIGN_HOP)
    hop_sec = ALIGN_HOP / ALIGN_SR
    max_offset_frames = int(max_offset_sec / hop_sec)
    speeds = np.linspace(speed_lo, speed_hi, n_speeds)
    n_src, n_nc = int(src_env.numel()), int(nc_env.numel())
    # per-speed geometry (xcorr.py:221-240); speeds the reference skips get search_len = -1
    n_str = np.zeros(n_speeds, dtype=np.int32)
    n_lag = np.full(n_speeds, -1, dtype=np.int32)
    for i, speed in enumerate(speeds):
        ns = int(n_nc / speed)
        if ns < 4 or ns >= n_src:
            continue
        search_len = min(max_offset_frames, n_src - ns)
        if search_len <= 0:
            continue
        n_str[i] = ns
        n_lag[i] = search_len + 1          # np.correlate(valid)[:search_len + 1]
    if not (n_lag > 0).any():
        return 0.0, (speed_lo + speed_hi) / 2.0
    peak_idx, score = eng.align_search(src_env, nc_env, n_str, n_lag)
    best_score, best_offset, best_speed = -1.0, 0.0, (speed_lo + speed_hi) / 2.0
    for i, speed in enumerate(speeds):
        if n_lag[i] <= 0:
            continue
        if score[i] > best_score:
            best_score = float(score[i])
            best_offset = int(peak_idx[i]) * hop_sec
            best_speed = speed
    return best_offset, best_speed


def quality_label(quality: float) -> str:
    """xcorr.py:262-268."""
    if quality >= XCORR_QUALITY_GOOD:
        return "good match"
    if quality >= XCORR_QUALITY_FAIR:
        return "moderate match"
    return "poor match — possible content mismatch or heavy lossy artefacts"
