"""GPU: the BASELINE.json configurations at FULL size, checked through size-independent properties (the CPU oracle
needs minutes to hours at these sizes) and, where the oracle is cheap, exactly.

  config 2  60-min track, hop-64 onset + beat tracking      → beats strictly increasing, IBI median = beat period
  config 4  10-min pair at 44.1 kHz, waveform xcorr          → identical picks / slope to the CPU port (fast on CPU)
  config 5  pair batch                                       → tiling idempotence: equal pairs give bit-equal results,
                                                               batch order does not matter"""
import json
import os

import numpy as np
import pytest
import scipy.signal

from oracle import pipeline_port as port
from oracle import synth

pytestmark = pytest.mark.gpu
SR = 22050
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fh = lambda v: None if v is None else float.fromhex(v)


@pytest.fixture(scope="module")
def full():
    """Oracle answers at full BASELINE sizes (tests/golden/make_fullsize_golden.py — config 1 through the reference's own
    pipeline.run over the librosa shim)."""
    with open(os.path.join(ROOT, "tests", "golden", "fullsize_golden.json")) as f:
        return json.load(f)


def click_track(dur_s, bpm, sr=SR, seed=0):
    """Cheap long input: decaying noise bursts on every beat over a quiet noise floor (vectorised)."""
    rng = np.random.default_rng(seed)
    n = int(dur_s * sr)
    y = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    period = 60.0 / bpm
    starts = (np.arange(int(dur_s / period)) * period * sr).astype(np.int64)
    burst = (rng.standard_normal(400) * np.exp(-np.arange(400) / 80.0)).astype(np.float32)
    for k in range(400):
        idx = starts + k
        idx = idx[idx < n]
        y[idx] += 0.5 * burst[k]
    return y


def test_config2_sixty_minute_hop64_pass(engine):
    from nightcore_analyzer import tempo as ntempo
    bpm = 123.0
    y = click_track(3600.0, bpm)
    assert len(y) == 79_380_000
    ibis = ntempo.estimate_ibis_global(y, SR)
    assert ibis is not None and len(ibis) > 0.9 * 3600 * bpm / 60
    assert np.all(ibis > 0.05)                                   # tempo.py:170 filter
    period = 60.0 / bpm
    assert abs(np.median(ibis) - period) < 64 / SR               # within one hop-64 frame of the true beat period
    frames = np.round(np.cumsum(ibis) * SR / 64)
    assert np.all(np.diff(frames) > 0)                           # strictly increasing beat frames
    # the first three minutes alone give the same beat period (the pass does not degrade with length)
    ibis3 = ntempo.estimate_ibis_global(y[: 180 * SR], SR)
    assert abs(np.median(ibis3) - np.median(ibis)) < 64 / SR


def test_config4_ten_minute_44k_xcorr_matches_port(engine):
    from nightcore_analyzer import xcorr as nx
    sr = 44100
    a = np.tile(synth.synth(4000, 60.0, sr, bpm=124.0), 10)       # 10 minutes = 26 460 000 samples
    assert len(a) == 26_460_000
    b = scipy.signal.resample_poly(a, 1000, 1003).astype(np.float32)
    b = (b + np.random.default_rng(1).standard_normal(len(b)).astype(np.float32) * 0.01).astype(np.float32)
    (slope, quality), (pos, pb) = nx.estimate_speed_xcorr_batch([(a, b)], sr, return_indices=True)[0]
    (w_slope, w_quality), (w_pos, w_pb) = port.speed_xcorr_arrays(a, b, sr, return_indices=True)
    assert pos.tolist() == w_pos.tolist() and pb.tolist() == w_pb.tolist()
    assert slope == w_slope and abs(quality - w_quality) <= 1e-5
    # xcorr.py:57,92: the path-based reference call would load at 22 050 Hz — same check on 2:1 decimated copies
    a2, b2 = a[::2].copy(), b[::2].copy()
    got = nx.estimate_speed_xcorr_batch([(a2, b2)], 22050, return_indices=True)[0]
    want = port.speed_xcorr_arrays(a2, b2, 22050, return_indices=True)
    assert got[1][1].tolist() == want[1][1].tolist() and got[0][0] == want[0][0]


def test_config5_batch_tiling_idempotence(engine):
    import nightcore_analyzer as na
    distinct = []
    for i in range(3):
        src, nc = synth.make_pair(5000 + i, 60.0, SR)
        distinct.append((nc, src))
    order = [0, 1, 2, 1, 0, 2, 2, 0]
    out = na.run_batch([distinct[i] for i in order], SR)
    first = {}
    for i, r in zip(order, out):
        assert not isinstance(r, Exception), r
        key = (r.tempo_ratio, r.tempo_ci, r.pitch_ratio, r.pitch_ci, r.ibi_ratio, r.ibi_ci, tuple(r.nc_tempos_raw),
               tuple(r.src_tempos_raw), tuple(r.nc_pitches_raw), r.classification)
        if i in first:
            assert key == first[i]                               # bit-equal regardless of position in the batch
        first[i] = key
    # and equal to analysing the pair alone
    alone = na.run_batch([distinct[1]], SR)[0]
    assert (alone.tempo_ratio, alone.ibi_ratio, tuple(alone.nc_pitches_raw)) == (first[1][0], first[1][4], first[1][8])


# ---------------------------------------------------------------------------------------------- oracle-checked, full size
def check_config1(res, c):
    assert res.src_tempos_raw == [fh(v) for v in c["src_tempos"]]
    assert res.nc_tempos_raw == [fh(v) for v in c["nc_tempos"]]
    assert res.src_pitches_raw == [fh(v) for v in c["src_hz"]] and res.nc_pitches_raw == [fh(v) for v in c["nc_hz"]]
    assert (res.tempo_ratio, list(res.tempo_ci)) == (fh(c["tempo_ratio"]), [fh(v) for v in c["tempo_ci"]])
    assert (res.pitch_ratio, list(res.pitch_ci)) == (fh(c["pitch_ratio"]), [fh(v) for v in c["pitch_ci"]])
    assert (res.ibi_ratio, list(res.ibi_ci)) == (fh(c["ibi_ratio"]), [fh(v) for v in c["ibi_ci"]])
    assert res.classification == c["classification"] and res.warnings == c["warnings"] and res.rubberband == c["rubberband"]
    assert [res.n_source_pitch_windows, res.n_nc_pitch_windows, res.n_source_tempo_windows, res.n_nc_tempo_windows] == c["n"]


def test_config1_full_size_pair_matches_reference_flow(engine, full):
    """BASELINE config 1 at its real size (180 s source, 35 + 27 windows, 7 chunk pairs, 62 k / 50 k hop-64 frames):
    run_arrays AND run_batch against the reference's own pipeline.run (pipeline.py:23-216), every list, CI, string and
    log line; then the hop-64 lags and every beat frame of both tracks."""
    import nightcore_analyzer as na
    from nightcore_analyzer import io as nio
    c = full["config1"]
    src, nc = synth.make_pair(1000, 180.0, SR)
    logs = []
    res = na.run_arrays(nc, src, SR, log=logs.append)
    check_config1(res, c)
    assert str(res) == c["str"]
    assert logs == [m for m in c["logs"] if not m.startswith("Loading") and "samples @" not in m]
    check_config1(na.run_batch([(nc, src)], SR)[0], c)
    nc_s, _, _ = nio.strip_silence(nc, SR, 60.0)
    src_s, _, _ = nio.strip_silence(src, SR, 60.0)
    lags, beats = engine.tempo_and_beats([nc_s, src_s], [fh(c["nc_start_bpm"]), 120.0], 64, SR)
    for which, y, lag, b in zip(("nc", "src"), (nc_s, src_s), lags, beats):
        assert int(lag) == c["hop64_" + which]["lag"] and b.tolist() == c["hop64_" + which]["beats"]
        assert 1 + len(y) // 64 == c["hop64_" + which]["n_frames"]


def test_config3_full_size_chroma_pitch_matches_reference_flow(engine, full):
    """BASELINE config 3 at its real size: 180 s vs 144 s, independent +1.5 st, 7 chunk pairs (pitch.py:100-173):
    tuning bins, chunk lags, Hz lists, seed-0 bootstrap CI and log line identical; mean chroma within 1e-4 of its max."""
    from nightcore_analyzer import consensus as ncons, pitch as npitch
    c = full["config3"]
    src3 = synth.synth(3000, 180.0, SR)
    nc3 = synth.synth(3000, 144.0, SR, speed=1.25, pitch_mult=1.25 * 2.0 ** (1.5 / 12))
    logs = []
    s_hz, n_hz, point, ci, n_chunks = npitch.estimate_pitch_chroma(src3, nc3, SR, log=logs.append)
    assert n_chunks == c["n_chunks"] == 7 and logs == c["logs"]
    assert s_hz == [fh(v) for v in c["src_hz"]] and n_hz == [fh(v) for v in c["nc_hz"]]
    assert point == fh(c["point_st"]) and list(ci) == [fh(v) for v in c["ci_st"]]
    ratio, rci = ncons._bootstrap_ratio(ncons._valid(n_hz), ncons._valid(s_hz))
    assert ratio == fh(c["pitch_ratio"]) and list(rci) == [fh(v) for v in c["pitch_ci"]]
    chunks = [y for s, n in port.chunk_pairs(src3, nc3, SR) for y in (s, n)]
    audio, off, ln = engine.pack(chunks)
    chroma, tun = engine.chroma_mean_dev(audio, off, ln, SR)
    assert tun.cpu().numpy().tolist() == c["tuning_idx"]
    for got, want in zip(chroma.cpu().numpy(), c["chroma"]):
        want = np.array([fh(v) for v in want])
        assert np.max(np.abs(got - want)) <= 1e-4 * np.max(np.abs(want))


def test_config2_ten_minute_hop64_matches_oracle(engine, full):
    """BASELINE config 2 at the length BASELINE.md §5 plans for the faithful CPU form (10 minutes, 206 720 frames):
    tempogram lag and every one of the ~1200 beat frames identical to the streamed oracle (tempo.py:158-173)."""
    c = full["config2_10min"]
    y = synth.synth(2000, 600.0, SR)
    lags, beats = engine.tempo_and_beats([y], [120.0], 64, SR)
    assert 1 + len(y) // 64 == c["n_frames"]
    assert int(lags[0]) == c["lag"]
    assert beats[0].tolist() == c["beats"]


def test_two_rank_gather_is_identical_to_one_rank():
    """the table gathered from N ranks (NCCL all_gather of per-pair records) equals, bit for bit, the
    records of the same pairs analysed alone on one rank.  Needs two GPUs (skipped on a one-GPU box; bench.py performs
    the same check in every run and prints `gather_identical`)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--pairs", "48", "--pair-sec", "60",
           "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-pageable"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["gather_identical"] is True
    assert line["schedule"]["pairs_ok"] == 48
