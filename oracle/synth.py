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
