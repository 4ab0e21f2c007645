"""Generate tests/golden/fullsize_golden.json: oracle answers at the FULL sizes of BASELINE.json's configurations and
for the parity stress corpus, so the `-m gpu` tests can compare the CUDA path with the oracle without spending
minutes of CPU on the GPU box (tests/test_gpu_fullsize.py, tests/test_gpu_stress.py).

  config1   180 s source + 1.25x nightcore through the REFERENCE's own pipeline.run (unmodified modules over the
            librosa shim, /root/reference/nightcore_analyzer/pipeline.py:23-216), every list, CI and log line; the
            CPU port must agree; plus the hop-64 lags / beat frames of both tracks
  config3   180 s vs 144 s (+1.5 st, 1.25x) through the reference's estimate_pitch_chroma (pitch.py:100-173):
            7 chunk pairs, per-chunk tuning bins, mean chroma vectors, lags, seed-0 bootstrap
  config2   10-minute track at hop 64 (tempo.py:158-173): tempogram lag and every beat frame
  stress    512 ten-second windows (64 per family of oracle/synth.stress) with varying priors, 24 hop-64 tracks,
            48 twenty-second chunks (tuning bin, mean chroma) paired into 24 cyclic lags — full ORACLE path

Inputs are regenerated from seeds on the GPU box; only integers and a few floats are stored.
Run in the build container (needs /root/reference):   python tests/golden/make_fullsize_golden.py [-j 8]
"""
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import librosa_restated as lr, pipeline_port as port, reference_shim, synth  # noqa: E402

SR = 22050
hx = lambda v: None if v is None else float(v).hex()

N_STRESS_WINDOWS = 512
N_STRESS_TRACKS = 24
N_STRESS_CHUNKS = 48


def stress_window_prior(seed: int) -> float:
    """Every other window keeps the default 120 BPM prior; the rest use an off-centre one (like pipeline.py:174-178)."""
    return 120.0 if seed % 2 == 0 else float(np.random.default_rng(10_000 + seed).uniform(60.0, 200.0))


def stress_track_spec(j: int):
    rng = np.random.default_rng(20_000 + j)
    return 1000 + j, float(rng.uniform(40.0, 90.0)), (120.0 if j % 3 == 0 else float(rng.uniform(70.0, 190.0)))


def _tempo_case(args):
    seed, dur, hop, prior = args
    y = synth.stress(seed, dur, SR)
    env = lr.onset_strength(y, SR, hop)
    if not env.any():
        return dict(seed=seed, lag=0, beats=[])
    lag = lr.tempo_lag(env, SR, hop, prior)
    bpm, beats = lr.beat_track(env, SR, hop, prior)
    return dict(seed=seed, lag=int(lag), beats=[int(b) for b in beats])


def _chunk_case(seed):
    y = synth.stress(seed, 20.0, SR)
    _, idx = lr.estimate_tuning(y, SR, 36, return_index=True)
    c = lr.chroma_cqt(y, SR, 512, 36).mean(axis=1)
    return dict(seed=seed, tuning_idx=int(idx), chroma=[hx(v) for v in c])


def _hop64(args):
    name, y, prior = args
    env = lr.onset_strength(y, SR, 64)
    lag = lr.tempo_lag(env, SR, 64, prior)
    _, beats = lr.beat_track(env, SR, 64, prior)
    return name, dict(lag=int(lag), beats=[int(b) for b in beats], n_frames=len(env))


def main():
    jobs = int(sys.argv[sys.argv.index("-j") + 1]) if "-j" in sys.argv else (os.cpu_count() or 1)
    t0 = time.time()
    out = {"note": "oracle answers at full BASELINE sizes + stress corpus (parity unpinned for librosa-backed numbers: "
                   "the arithmetic underneath is oracle/librosa_restated.py)", "numpy": np.__version__}
    pool = mp.get_context("fork").Pool(jobs)

    # ---- stress corpus first (asynchronous; the pool works while the serial reference runs below)
    w_async = pool.map_async(_tempo_case, [(s, 10.0, 512, stress_window_prior(s)) for s in range(N_STRESS_WINDOWS)],
                             chunksize=8)
    t_async = pool.map_async(_tempo_case, [(sd, dur, 64, pr) for sd, dur, pr in map(stress_track_spec,
                                                                                    range(N_STRESS_TRACKS))], chunksize=1)
    c_async = pool.map_async(_chunk_case, [3000 + j for j in range(N_STRESS_CHUNKS)], chunksize=2)

    # ---- config 1 at full size through the reference's own pipeline.run
    ref = reference_shim.load_reference()
    src, nc = synth.make_pair(1000, 180.0, SR)
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "src.npy"), src)
        np.save(os.path.join(td, "nc.npy"), nc)
        logs = []
        res = ref.pipeline.run(os.path.join(td, "nc.npy"), os.path.join(td, "src.npy"), log=logs.append)
    want = port.run_arrays(nc, src, SR, faithful_cost=False)
    assert res.src_tempos_raw == want["src_tempos"] and res.nc_tempos_raw == want["nc_tempos"]
    assert res.nc_pitches_raw == want["nc_hz"] and (res.tempo_ratio, res.tempo_ci) == want["tempo"]
    assert (res.ibi_ratio, res.ibi_ci) == want["ibi"] and (res.pitch_ratio, res.pitch_ci) == want["pitch"]
    out["config1"] = {
        "inputs": "synth.make_pair(1000, 180.0, 22050)",
        "src_tempos": [hx(v) for v in res.src_tempos_raw], "nc_tempos": [hx(v) for v in res.nc_tempos_raw],
        "src_hz": [hx(v) for v in res.src_pitches_raw], "nc_hz": [hx(v) for v in res.nc_pitches_raw],
        "tempo_ratio": hx(res.tempo_ratio), "tempo_ci": [hx(v) for v in res.tempo_ci],
        "pitch_ratio": hx(res.pitch_ratio), "pitch_ci": [hx(v) for v in res.pitch_ci],
        "ibi_ratio": hx(res.ibi_ratio), "ibi_ci": None if res.ibi_ci is None else [hx(v) for v in res.ibi_ci],
        "classification": res.classification, "warnings": res.warnings, "rubberband": res.rubberband, "str": str(res),
        "n": [res.n_source_pitch_windows, res.n_nc_pitch_windows, res.n_source_tempo_windows, res.n_nc_tempo_windows],
        "nc_start_bpm": hx(want["nc_start_bpm"]), "chunk_lags": [int(v) for v in want["chunk_lags"]], "logs": logs,
    }
    # hop-64 pass of both (stripped) tracks: lag + beat frames
    nc_s, _, _ = port.strip_silence(nc, SR, 60.0)
    src_s, _, _ = port.strip_silence(src, SR, 60.0)
    h_async = pool.map_async(_hop64, [("nc", nc_s, want["nc_start_bpm"]), ("src", src_s, 120.0)], chunksize=1)

    # ---- config 3 at full size through the reference's estimate_pitch_chroma
    src3 = synth.synth(3000, 180.0, SR)
    nc3 = synth.synth(3000, 144.0, SR, speed=1.25, pitch_mult=1.25 * 2.0 ** (1.5 / 12))
    logs = []
    s_hz, n_hz, point, ci, n_chunks = ref.pitch.estimate_pitch_chroma(src3, nc3, SR, log=logs.append)
    pairs = port.chunk_pairs(src3, nc3, SR)
    tun, chroma = [], []
    for s, n in pairs:
        for y in (s, n):
            tun.append(int(lr.estimate_tuning(y, SR, 36, return_index=True)[1]))
            chroma.append([hx(v) for v in lr.chroma_cqt(y, SR, 512, 36).mean(axis=1)])
    p_src, p_nc = port.valid(s_hz), port.valid(n_hz)
    ratio = port.bootstrap_ratio(p_nc, p_src)
    out["config3"] = {
        "inputs": "synth(3000,180) vs synth(3000,144,speed=1.25,pitch_mult=1.25*2^(1.5/12))",
        "src_hz": [hx(v) for v in s_hz], "nc_hz": [hx(v) for v in n_hz], "point_st": hx(point),
        "ci_st": [hx(v) for v in ci], "n_chunks": n_chunks, "logs": logs, "tuning_idx": tun, "chroma": chroma,
        "pitch_ratio": hx(ratio[0]), "pitch_ci": [hx(v) for v in ratio[1]],
    }

    # ---- config 2 at 10 minutes (the faithful form fits at this length)
    y2 = synth.synth(2000, 600.0, SR)
    _, c2 = _hop64(("c2", y2, 120.0))
    out["config2_10min"] = dict(inputs="synth(2000, 600.0, 22050)", **c2)

    for name, r in h_async.get():
        out["config1"]["hop64_" + name] = r
    out["stress"] = {
        "windows": w_async.get(), "tracks": t_async.get(), "chunks": c_async.get(),
        "window_priors": "120 for even seeds, default_rng(10000+seed).uniform(60,200) for odd",
    }
    ch = out["stress"]["chunks"]
    lags = []
    for j in range(N_STRESS_CHUNKS // 2):
        a = np.array([float.fromhex(v) for v in ch[j]["chroma"]])
        b = np.array([float.fromhex(v) for v in ch[j + N_STRESS_CHUNKS // 2]["chroma"]])
        lags.append(port.cyclic_xcorr_peak(a, b))
    out["stress"]["chunk_pair_lags"] = lags
    pool.close()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True, ensure_ascii=False)
    print(f"wrote fullsize_golden.json in {time.time() - t0:.0f}s")
    print(out["config1"]["classification"], out["config1"]["n"], out["config3"]["n_chunks"], out["config2_10min"]["lag"],
          len(out["config2_10min"]["beats"]))


if __name__ == "__main__":
    main()
