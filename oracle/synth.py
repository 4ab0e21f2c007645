"""Synthetic audio generator of SURVEY.md §8(d) (TEST / BENCH INFRASTRUCTURE).

``synth(seed, dur_s, sr, bpm, speed, pitch_mult)``: kick on every beat, closed hat on every
off-beat at -12 dB, a I-V-vi-IV triad loop (6 harmonics at 1/h, one chord per 2 beats,
10 ms attack / 200 ms release), white noise at -50 dBFS; float32 mono, peak 0.8.
"""
from __future__ import annotations

import numpy as np


def synth(seed: int, dur_s: float, sr: int = 22050, bpm=None, speed: float = 1.0, pitch_mult: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if bpm is None:
        bpm = float(rng.uniform(90.0, 140.0))
    else:
        rng.uniform(90.0, 140.0)  # keep the stream aligned
    n = int(round(dur_s * sr))
    y = np.zeros(n, dtype=np.float64)
    period = 60.0 / (bpm * speed)
    n_beats = int(np.ceil(dur_s / period)) + 1

    # kick: 60 Hz·pitch_mult sine × exp(-t/0.08) + 4 ms noise click
    klen = int(0.4 * sr)
    tk = np.arange(klen) / sr
    kick = np.sin(2 * np.pi * 60.0 * pitch_mult * tk) * np.exp(-tk / 0.08)
    clen = int(0.004 * sr)
    hlen = int(0.008 * sr)
    for b in range(n_beats):
        s = int(round(b * period * sr))
        if s >= n:
            break
        e = min(n, s + klen)
        y[s:e] += kick[: e - s]
        click = rng.standard_normal(clen) * 0.5
        e = min(n, s + clen)
        y[s:e] += click[: e - s]
        # hat on the off-beat, -12 dB, first-difference high-pass
        sh = int(round((b + 0.5) * period * sr))
        if sh < n:
            hat = np.diff(rng.standard_normal(hlen + 1)) * 0.25
            e = min(n, sh + hlen)
            y[sh:e] += hat[: e - sh]

    # chords: I-V-vi-IV in C, triads, 6 harmonics at 1/h, one chord per 2 beats
    c4 = 261.6255653005986
    semis = [(0, 4, 7), (7, 11, 14), (9, 12, 16), (5, 9, 12)]
    chord_len = 2 * period
    n_chords = int(np.ceil(dur_s / chord_len))
    att, rel = 0.010, 0.200
    for c in range(n_chords):
        s = int(round(c * chord_len * sr))
        if s >= n:
            break
        e = min(n, int(round((c + 1) * chord_len * sr)))
        t = np.arange(e - s) / sr
        env = np.minimum(1.0, t / att) * np.minimum(1.0, np.maximum(0.0, (chord_len - t) / rel))
        seg = np.zeros(e - s)
        for st in semis[c % 4]:
            f0 = c4 * 2.0 ** (st / 12.0) * pitch_mult
            for h in range(1, 7):
                f = f0 * h
                if f < sr / 2:
                    seg += np.sin(2 * np.pi * f * t) / h
        y[s:e] += 0.15 * env * seg

    y += rng.standard_normal(n) * 10 ** (-50 / 20)
    y *= 0.8 / np.max(np.abs(y))
    return y.astype(np.float32)


def make_pair(seed: int, dur_s: float = 180.0, sr: int = 22050, up: int = 4, down: int = 5):
    """(src, nc): nc = resample_poly(src, 4, 5) — a pure 1.25x speed-up (config 1 / 5 of BASELINE.json)."""
    import scipy.signal
    src = synth(seed, dur_s, sr)
    nc = scipy.signal.resample_poly(src, up, down).astype(np.float32)
    return src, nc


# ---------------------------------------------------------------------------------------------- stress corpus
def _place(y: np.ndarray, times_s: np.ndarray, burst: np.ndarray, sr: int, gains=None) -> None:
    """Add `burst` at each onset time (vectorised over onsets)."""
    n = len(y)
    starts = np.round(np.asarray(times_s) * sr).astype(np.int64)
    starts = starts[(starts >= 0) & (starts < n)]
    g = np.ones(len(starts)) if gains is None else np.asarray(gains)[: len(starts)]
    for k in range(len(burst)):
        idx = starts + k
        ok = idx < n
        np.add.at(y, idx[ok], burst[k] * g[ok])


def stress(seed: int, dur_s: float, sr: int = 22050) -> np.ndarray:
    """Weakly periodic / noisy / tempo-drifting material for the parity stress corpus (tests/golden/
    make_fullsize_golden.py, tests/test_gpu_stress.py).  Eight families, picked by ``seed % 8``; everything is drawn
    from ``default_rng(seed)`` so the GPU box regenerates the same samples.  float32 mono, peak 0.8."""
    rng = np.random.default_rng(seed)
    kind = seed % 8
    n = int(round(dur_s * sr))
    t = np.arange(n) / sr
    y = np.zeros(n, dtype=np.float64)
    click = rng.standard_normal(int(0.006 * sr)) * np.exp(-np.arange(int(0.006 * sr)) / (0.0015 * sr))
    thump = np.sin(2 * np.pi * 70.0 * np.arange(int(0.12 * sr)) / sr) * np.exp(-np.arange(int(0.12 * sr)) / (0.03 * sr))
    bpm = float(rng.uniform(70.0, 180.0))
    period = 60.0 / bpm
    if kind == 0:      # noise with a slow, shallow amplitude modulation: almost no periodicity
        y = rng.standard_normal(n) * (1.0 + 0.15 * np.sin(2 * np.pi * t / period + rng.uniform(0, 6.28)))
    elif kind == 1:    # red-ish noise with faint clicks 20 dB below it
        w = rng.standard_normal(n)
        y = np.convolve(w, np.ones(8) / 8.0, mode="same") * 2.0
        _place(y, np.arange(0.0, dur_s, period), click * 0.1, sr)
    elif kind == 2:    # tempo drift: the beat period changes linearly by up to ±15 % over the signal
        drift = float(rng.uniform(-0.15, 0.15))
        times, cur = [], float(rng.uniform(0, period))
        while cur < dur_s:
            times.append(cur)
            cur += period * (1.0 + drift * cur / dur_s)
        _place(y, np.array(times), thump, sr)
        _place(y, np.array(times), click * 0.5, sr)
        y += rng.standard_normal(n) * 0.02
    elif kind == 3:    # two competing click trains of nearly equal strength (3:2, 4:3 or a 5 % detune)
        ratio = [1.5, 4.0 / 3.0, 1.05][int(rng.integers(0, 3))]
        _place(y, np.arange(rng.uniform(0, period), dur_s, period), thump, sr)
        _place(y, np.arange(rng.uniform(0, period), dur_s, period / ratio), thump * float(rng.uniform(0.8, 1.1)), sr)
        y += rng.standard_normal(n) * 0.01
    elif kind == 4:    # jittered beats (sigma 30 ms) with random accents
        times = np.arange(0.0, dur_s, period) + rng.normal(0.0, 0.03, int(np.ceil(dur_s / period)))
        _place(y, np.sort(times), thump, sr, gains=rng.uniform(0.3, 1.0, len(times)))
        y += rng.standard_normal(n) * 0.03
    elif kind == 5:    # sustained chord with tremolo, no percussive onsets
        f0 = float(rng.uniform(110.0, 440.0))
        for mult in (1.0, 1.26, 1.498, 2.0):
            for h in range(1, 5):
                y += np.sin(2 * np.pi * f0 * mult * h * t + rng.uniform(0, 6.28)) / h
        y *= 1.0 + 0.3 * np.sin(2 * np.pi * float(rng.uniform(3.0, 8.0)) * t)
        y += rng.standard_normal(n) * 0.005
    elif kind == 6:    # bursts of beats separated by stretches of exact digital silence
        times = np.arange(0.0, dur_s, period)
        gate = (np.floor(times / (4 * period)) % 2) == 0
        _place(y, times[gate], thump, sr)
        _place(y, times[gate] + period / 2, click * 0.4, sr)
    else:              # syncopated pattern + loud hiss (-20 dB)
        pattern = np.array([0.0, 0.75, 1.5, 2.0, 2.75, 3.5]) * period
        bars = np.arange(0.0, dur_s, 4 * period)
        times = (bars[:, None] + pattern[None, :]).ravel()
        _place(y, times, thump, sr, gains=rng.uniform(0.5, 1.0, len(times)))
        y += rng.standard_normal(n) * 0.1
    peak = float(np.max(np.abs(y)))
    if peak > 0:
        y *= 0.8 / peak
    return y.astype(np.float32)
