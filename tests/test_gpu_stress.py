"""GPU: the parity STRESS corpus — weakly periodic, noisy, tempo-drifting, polyrhythmic and tonal material
(oracle/synth.stress, eight families) through the FULL CUDA path, compared with the FULL oracle path whose answers
tests/golden/make_fullsize_golden.py committed (the oracle needs minutes of CPU for this corpus; the GPU box only
regenerates the inputs from their seeds).

  512 ten-second windows, varying priors   tempogram lag and every beat frame     (tempo.py:27-77)
  24 whole tracks at hop 64                lag and every beat frame               (tempo.py:120-173)
  48 twenty-second chunks                  tuning histogram bin, mean chroma, 24 cyclic lags   (pitch.py:55-95)

Every integer must be identical.  The GPU onset envelope differs from the oracle's in the last float32 bits (stated
tolerance 1e-4 of its maximum), so a decision between two nearly tied candidates may in principle flip; such a case is
accepted ONLY when the oracle, given the GPU's own envelope, reproduces the GPU's answer exactly (the decision logic is
identical; the difference is upstream rounding inside the tolerance) — and the number of such cases is bounded by
ALLOWED_UPSTREAM_FLIPS and reported (DESIGN.md "Exactness policy").
"""
import json
import os

import numpy as np
import pytest

from oracle import librosa_restated as lr
from oracle import synth

pytestmark = pytest.mark.gpu
SR = 22050
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# measured on B200: number of windows / tracks whose lag or beats differ from the stored oracle answer
ALLOWED_UPSTREAM_FLIPS = {"windows": 0, "tracks": 0, "chunks": 0}


@pytest.fixture(scope="module")
def g():
    with open(os.path.join(ROOT, "tests", "golden", "fullsize_golden.json")) as f:
        return json.load(f)["stress"]


def window_prior(seed: int) -> float:
    return 120.0 if seed % 2 == 0 else float(np.random.default_rng(10_000 + seed).uniform(60.0, 200.0))


def track_spec(j: int):
    rng = np.random.default_rng(20_000 + j)
    return 1000 + j, float(rng.uniform(40.0, 90.0)), (120.0 if j % 3 == 0 else float(rng.uniform(70.0, 190.0)))


def report(name, payload):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, f"stress_{name}.json"), "w") as f:
            json.dump(payload, f)


def certify(engine, y, hop, prior, got_lag, got_beats):
    """The oracle on the GPU's own envelope must reproduce the GPU's lag and beats; the envelope itself must be within
    the stated tolerance of the oracle's."""
    env_gpu = engine.onset_strength([y], hop=hop, sr=SR)[0]
    env_cpu = lr.onset_strength(y, SR, hop)
    tol = 1e-4 * max(1e-12, float(np.max(np.abs(env_cpu))))
    assert float(np.max(np.abs(env_gpu - env_cpu))) <= tol
    lag = lr.tempo_lag(env_gpu, SR, hop, prior) if env_gpu.any() else 0
    _, beats = lr.beat_track(env_gpu, SR, hop, prior)
    return int(lag) == int(got_lag) and [int(b) for b in beats] == [int(b) for b in got_beats]


def compare_tempo(engine, cases, signals, priors, hop, name):
    lags, beats = engine.tempo_and_beats(signals, priors, hop, SR)
    flips, wrong = [], []
    for c, y, p, lag, b in zip(cases, signals, priors, lags, beats):
        if int(lag) == c["lag"] and b.tolist() == c["beats"]:
            continue
        (flips if certify(engine, y, hop, p, lag, b) else wrong).append(c["seed"])
    report(name, {"cases": len(cases), "upstream_rounding_flips": flips, "wrong": wrong})
    assert not wrong, f"{len(wrong)} of {len(cases)} {name} differ from the oracle beyond upstream rounding: {wrong[:10]}"
    assert len(flips) <= ALLOWED_UPSTREAM_FLIPS[name], f"{len(flips)} {name} flipped by upstream rounding: {flips[:10]}"


def test_stress_windows_full_path(engine, g):
    cases = g["windows"]
    assert len(cases) >= 500
    signals = [synth.stress(c["seed"], 10.0, SR) for c in cases]
    priors = [window_prior(c["seed"]) for c in cases]
    compare_tempo(engine, cases, signals, priors, 512, "windows")


def test_stress_hop64_tracks_full_path(engine, g):
    cases = g["tracks"]
    assert len(cases) >= 20
    specs = [track_spec(j) for j in range(len(cases))]
    assert [s[0] for s in specs] == [c["seed"] for c in cases]
    signals = [synth.stress(sd, dur, SR) for sd, dur, _ in specs]
    compare_tempo(engine, cases, signals, [p for _, _, p in specs], 64, "tracks")


def test_stress_chunks_tuning_chroma_and_lags(engine, g):
    import torch
    cases = g["chunks"]
    chunks = [synth.stress(c["seed"], 20.0, SR) for c in cases]
    audio, off, ln = engine.pack(chunks)
    chroma, tun = engine.chroma_mean_dev(audio, off, ln, SR)
    chroma, tun = chroma.cpu().numpy(), tun.cpu().numpy()
    bad_tuning = [c["seed"] for c, t in zip(cases, tun) if int(t) != c["tuning_idx"]]
    worst = 0.0
    for c, got, t in zip(cases, chroma, tun):
        if int(t) != c["tuning_idx"]:
            continue
        want = np.array([float.fromhex(v) for v in c["chroma"]])
        worst = max(worst, float(np.max(np.abs(got - want)) / np.max(np.abs(want))))
    half = len(cases) // 2
    lags = engine.cyclic_xcorr_dev(torch.from_numpy(chroma[:half]).cuda(), torch.from_numpy(chroma[half:]).cuda())
    lags = lags.cpu().numpy()
    bad_lags = [j for j, (a, b) in enumerate(zip(lags.tolist(), g["chunk_pair_lags"])) if a != b]
    report("chunks", {"cases": len(cases), "tuning_mismatch": bad_tuning, "lag_mismatch": bad_lags, "chroma_worst_rel": worst})
    assert len(bad_tuning) <= ALLOWED_UPSTREAM_FLIPS["chunks"], bad_tuning
    assert worst <= 1e-4, worst                      # mean chroma within 1e-4 of its maximum (split-TF32 contraction)
    assert not bad_lags or bad_tuning, bad_lags
