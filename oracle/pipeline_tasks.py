"""
Task-parallel execution of the CPU port (TEST / BASELINE INFRASTRUCTURE — see oracle/__init__.py).

``bench.py --impl reference`` and the ``cpu_baseline`` leg must time the reference's CPU path on the configuration the
metric is quoted on (180 s source + 1.25x nightcore), yet finish K timed steps within minutes.  One pair costs ~30 s
of one core, two thirds of it in the two hop-64 whole-track passes, so "one pair per process" would make a step half a
minute long.  This module runs exactly the computations of ``pipeline_port.run_arrays`` (same functions, same
arguments, same order of the per-pair decisions) cut into independent units —

    prep          strip_silence + slice_windows + energy_gate of both tracks   (io.py:58-126)
    tempo group   estimate_tempo of a few windows                               (tempo.py:27-77)
    chroma chunk  both mean chromas of one chunk pair + cyclic lag              (pitch.py:55-95)
    hop-64 pass   estimate_ibis_global of one whole track (tempo.py:120-173), itself cut into the onset envelope,
                  the frame blocks of the streamed tempogram mean (summed by the scheduler in the port's order, so the
                  float64 result is bit-identical) and the beat tracker
    finish        bootstraps                                                     (consensus.py:243-312)

— and deals the units of a SAMPLE of pairs to a pool of worker processes, honouring the reference's only serial
dependency (the nightcore prior needs the source windows' tempos, pipeline.py:174-186).  The results are identical to
``pipeline_port.run_arrays`` (tests/test_pipeline_tasks.py).
"""
from __future__ import annotations

import concurrent.futures as cf
import multiprocessing as mp
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import pipeline_port as port

_PAIRS: Sequence[Tuple[np.ndarray, np.ndarray]] = ()     # (nc, src), inherited by the forked workers
_SR = 22050
_FAITHFUL = True
TEMPO_GROUP = 6            # windows per tempo task


def _limit_threads():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        pass


def _track(i: int, which: int, bounds) -> np.ndarray:
    a = _PAIRS[i][which]
    return a if bounds is None else a[bounds[0] : bounds[1]]


def _prep(i: int, window_sec: float, hop_sec: float, gate_db: float, strip_db: Optional[float]):
    out = []
    for which in (0, 1):
        a = _PAIRS[i][which]
        if strip_db is not None:
            _, (s, e) = port.lr.trim(a, top_db=strip_db)
        else:
            s, e = 0, len(a)
        wins = port.energy_gate(port.slice_windows(a[s:e], _SR, window_sec, hop_sec), gate_db)
        out.append(((int(s), int(e)), [int(round(w[1] * _SR)) for w in wins], int(window_sec * _SR)))
    return out


def _tempo_group(i: int, which: int, bounds, starts: List[int], win_n: int, start_bpm: float):
    y = _track(i, which, bounds)
    return [port.estimate_tempo(y[s : s + win_n], _SR, start_bpm, _FAITHFUL) for s in starts]


def _chroma_chunk(i: int, b_nc, b_src, c: int):
    src, nc = _track(i, 1, b_src), _track(i, 0, b_nc)
    s, n = port.chunk_pairs(src, nc, _SR)[c]
    return port.cyclic_xcorr_peak(port.mean_chroma(s, _SR), port.mean_chroma(n, _SR))


TG_BLOCK = 2048            # frames per tempogram task = lr.tempogram_mean's streaming block


def _onset64(i: int, which: int, bounds):
    return port.lr.onset_strength(_track(i, which, bounds), _SR, port.IBI_HOP_LENGTH)


def _tg_block(env: np.ndarray, win_length: int, s: int, e: int):
    return port.lr.tempogram(env, win_length, s, e).sum(axis=1)


def _beats64(env: np.ndarray, lag: int, win_length: int):
    bpm = port.lr.tempo_frequencies(win_length, port.IBI_HOP_LENGTH, _SR)[lag]
    beats = port.lr.beat_track_frames(env, float(bpm), sr=_SR, hop_length=port.IBI_HOP_LENGTH)
    return port.ibis_from_beats(beats, _SR, port.IBI_HOP_LENGTH, 4)


def _init_worker():
    _limit_threads()


class TaskRunner:
    """A pool of `procs` forked workers over a fixed list of pairs [(nc, src), ...]."""

    def __init__(self, pairs, sr: int = 22050, procs: Optional[int] = None, faithful_cost: bool = True):
        global _PAIRS, _SR, _FAITHFUL
        _PAIRS, _SR, _FAITHFUL = list(pairs), sr, faithful_cost
        self.procs = procs or (os.cpu_count() or 1)
        self.pool = cf.ProcessPoolExecutor(self.procs, mp_context=mp.get_context("fork"), initializer=_init_worker)

    def close(self):
        self.pool.shutdown(wait=True, cancel_futures=True)

    def run(self, pair_ids: Sequence[int], *, window_sec: float = 10.0, hop_sec: float = 5.0,
            energy_gate_db: float = -40.0, silence_strip_db: Optional[float] = 60.0, compute_pitch: bool = True,
            compute_ibi: bool = True):
        """Analyse the pairs `pair_ids` (indices into the runner's list) → list of result dicts (or exceptions), as
        pipeline_port.run_arrays returns them, plus the number of gated windows."""
        sr = _SR
        win64 = int(np.floor(8.0 * sr / port.IBI_HOP_LENGTH))
        st = {k: dict(i=i) for k, i in enumerate(pair_ids)}
        pending = {}

        def submit(kind, k, fn, *args):
            pending[self.pool.submit(fn, *args)] = (kind, k)

        for k, i in enumerate(pair_ids):
            submit("prep", k, _prep, i, window_sec, hop_sec, energy_gate_db, silence_strip_db)
        results: list = [None] * len(pair_ids)
        windows = 0

        def launch_tempo(k, which, prior):
            s = st[k]
            bounds, starts, win_n = s["prep"][which]
            groups = [starts[a : a + TEMPO_GROUP] for a in range(0, len(starts), TEMPO_GROUP)]
            s[("tempo", which)] = [None] * len(groups)
            for g, grp in enumerate(groups):
                submit(("tempo", which, g), k, _tempo_group, s["i"], which, bounds, grp, win_n, prior)

        def maybe_finish(k):
            s = st[k]
            if results[k] is not None or s.get("failed"):
                return
            need = [("tempo", 1), ("tempo", 0)]
            for key in need:
                if key not in s or any(v is None for v in s[key]):
                    return
            if compute_pitch and any(v is None for v in s["lags"]):
                return
            if compute_ibi and ("ibi0" not in s or "ibi1" not in s):
                return
            results[k] = self._finish(s, compute_pitch, compute_ibi)

        while pending:
            done, _ = cf.wait(list(pending), return_when=cf.FIRST_COMPLETED)
            for f in done:
                kind, k = pending.pop(f)
                s = st[k]
                val = f.result()
                if kind == "prep":
                    s["prep"] = val
                    (b_nc, w_nc, _), (b_src, w_src, _) = val
                    if not w_nc or not w_src:
                        s["failed"] = True
                        results[k] = RuntimeError("All windows were discarded by the energy gate.  "
                                                  "Try raising --energy-gate (e.g. --energy-gate -60).")
                        continue
                    windows += len(w_nc) + len(w_src)
                    s["dur"] = ((b_nc[1] - b_nc[0]) / sr, (b_src[1] - b_src[0]) / sr)
                    launch_tempo(k, 1, 120.0)                       # source windows first: the prior waits for them
                    if compute_ibi:
                        s["bpm1"] = 120.0
                        submit(("onset64", 1), k, _onset64, s["i"], 1, b_src)
                    if compute_pitch:
                        chunk_n = int(port.CHUNK_SEC * sr)
                        n_chunks = max(1, min((b_src[1] - b_src[0]) // chunk_n, (b_nc[1] - b_nc[0]) // chunk_n))
                        s["lags"] = [None] * n_chunks
                        for c in range(n_chunks):
                            submit(("chroma", c), k, _chroma_chunk, s["i"], b_nc, b_src, c)
                    else:
                        s["lags"] = []
                elif kind[0] == "onset64":           # lr.beat_track: all-zero envelope → no beats
                    which = kind[1]
                    if not val.any():
                        s[f"ibi{which}"] = (None,)
                    else:
                        s[("env", which)] = val
                        n = len(val)
                        blocks = [(a, min(n, a + TG_BLOCK)) for a in range(0, n, TG_BLOCK)]
                        s[("tg", which)] = [None] * len(blocks)
                        for b, (a, e) in enumerate(blocks):
                            submit(("tg", which, b), k, _tg_block, val, win64, a, e)
                elif kind[0] == "tg":
                    _, which, b = kind
                    parts = s[("tg", which)]
                    parts[b] = val
                    if all(v is not None for v in parts):
                        acc = np.zeros(win64, dtype=np.float64)
                        for v in parts:                   # lr.tempogram_mean: acc += block sum, in block order
                            acc += v
                        env = s[("env", which)]
                        lag = port.lr.tempo_lag_from_mean(acc / len(env), sr, port.IBI_HOP_LENGTH, s[f"bpm{which}"])
                        submit(("beats64", which), k, _beats64, env, lag, win64)
                elif kind[0] == "beats64":
                    s[f"ibi{kind[1]}"] = (val,)
                elif kind[0] == "chroma":
                    s["lags"][kind[1]] = val
                elif kind[0] == "tempo":
                    _, which, g = kind
                    s[("tempo", which)][g] = val
                    if which == 1 and all(v is not None for v in s[("tempo", 1)]) and ("tempo", 0) not in s:
                        src_tempos = [t for grp in s[("tempo", 1)] for t in grp]
                        prior = 120.0
                        valid_src = [t for t in src_tempos if t is not None]
                        nc_d, src_d = s["dur"]
                        if valid_src and nc_d > 0 and src_d > 0:                   # pipeline.py:174-178
                            prior = float(np.median(valid_src)) * (src_d / nc_d)
                        s["prior"] = prior
                        if compute_ibi:
                            s["bpm0"] = prior
                            submit(("onset64", 0), k, _onset64, s["i"], 0, s["prep"][0][0])
                        launch_tempo(k, 0, prior)
                maybe_finish(k)
        return results, windows

    @staticmethod
    def _finish(s, compute_pitch: bool, compute_ibi: bool):
        """The tail of pipeline_port.run_arrays: validity gates and the three bootstraps."""
        src_tempos = [t for grp in s[("tempo", 1)] for t in grp]
        nc_tempos = [t for grp in s[("tempo", 0)] for t in grp]
        out = dict(src_tempos=src_tempos, nc_tempos=nc_tempos, nc_start_bpm=s["prior"], nc_duration=s["dur"][0],
                   src_duration=s["dur"][1], chunk_lags=list(s["lags"]))
        if compute_pitch:
            shift = np.array([lag / 3.0 for lag in s["lags"]])
            out["src_hz"] = [440.0] * len(shift)
            out["nc_hz"] = [440.0 * (2.0 ** (v / 12.0)) for v in shift]
            out["pitch_point_st"] = float(np.median(shift))
            if len(shift) >= port.MIN_CHUNKS:
                boots = port.native.bootstrap(shift, None, 2000, 0)
                out["pitch_ci_st"] = (float(np.percentile(boots, 2.5)), float(np.percentile(boots, 97.5)))
            else:
                out["pitch_ci_st"] = (out["pitch_point_st"],) * 2
        else:
            out["src_hz"], out["nc_hz"] = [], []
        src_t, nc_t = port.valid(src_tempos), port.valid(nc_tempos)
        if len(src_t) < port.MIN_VALID or len(nc_t) < port.MIN_VALID:
            return ValueError(f"Insufficient valid tempo windows (source: {len(src_t)}, "
                              f"nightcore: {len(nc_t)}).  Need ≥ {port.MIN_VALID} each.")
        src_p, nc_p = port.valid(out["src_hz"]), port.valid(out["nc_hz"])
        ok_p = len(src_p) >= port.MIN_VALID and len(nc_p) >= port.MIN_VALID
        out["pitch"] = port.bootstrap_ratio(nc_p, src_p) if ok_p else None
        out["tempo"] = port.bootstrap_ratio(nc_t, src_t)
        out["ibi"] = None
        if compute_ibi:
            nc_ibis, src_ibis = s["ibi0"][0], s["ibi1"][0]
            if nc_ibis is not None and len(nc_ibis) >= 4 and src_ibis is not None and len(src_ibis) >= 4:
                out["ibi"] = port.bootstrap_ratio(src_ibis, nc_ibis)
                out["n_ibis"] = (len(nc_ibis), len(src_ibis))
        return out
