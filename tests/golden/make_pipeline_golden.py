"""Generate tests/golden/pipeline_golden.json by running the REFERENCE's own modules, unmodified
(/root/reference/nightcore_analyzer/{io,tempo,pitch,xcorr,consensus,pipeline}.py), over the librosa shim
(oracle/reference_shim.py: our numpy/scipy restatement injected as `librosa`).  This pins the reference's
CONTROL FLOW (stage order, priors, gates, sentinels, bootstraps, log lines) — the librosa arithmetic
underneath is the restatement, so the file says "parity unpinned" for those numbers.

Run in the build container only (needs /root/reference):   python tests/golden/make_pipeline_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np
import scipy.signal

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim, synth  # noqa: E402

SR = 22050
ref = reference_shim.load_reference()
hx = lambda v: None if v is None else float(v).hex()
out = {"note": "reference control flow over restated librosa (parity unpinned for librosa-backed numbers)",
       "numpy": np.__version__}

with tempfile.TemporaryDirectory() as td:
    # ---- case A: BASELINE config 1 in miniature (45 s source, 1.25x nightcore)
    src, nc = synth.make_pair(1000, 45.0, SR)
    np.save(os.path.join(td, "src.npy"), src)
    np.save(os.path.join(td, "nc.npy"), nc)
    logs = []
    res = ref.pipeline.run(os.path.join(td, "nc.npy"), os.path.join(td, "src.npy"), log=logs.append)
    out["pipeline_A"] = {
        "inputs": "synth.make_pair(1000, 45.0, 22050)",
        "src_tempos": [hx(v) for v in res.src_tempos_raw], "nc_tempos": [hx(v) for v in res.nc_tempos_raw],
        "src_hz": [hx(v) for v in res.src_pitches_raw], "nc_hz": [hx(v) for v in res.nc_pitches_raw],
        "tempo_ratio": hx(res.tempo_ratio), "tempo_ci": [hx(v) for v in res.tempo_ci],
        "pitch_ratio": hx(res.pitch_ratio), "pitch_ci": [hx(v) for v in res.pitch_ci],
        "ibi_ratio": hx(res.ibi_ratio), "ibi_ci": None if res.ibi_ci is None else [hx(v) for v in res.ibi_ci],
        "classification": res.classification, "pitch_method": res.pitch_method, "warnings": res.warnings,
        "rubberband": res.rubberband, "str": str(res),
        "n": [res.n_source_pitch_windows, res.n_nc_pitch_windows, res.n_source_tempo_windows, res.n_nc_tempo_windows],
        "logs": logs,
    }
    # ---- case B: three 20 s chunk pairs with an independent +1.5 st shift (BASELINE config 3 in miniature)
    src3 = synth.synth(3000, 75.0, SR, bpm=112.0)
    nc3 = synth.synth(3000, 60.0, SR, bpm=112.0, speed=1.25, pitch_mult=1.25 * 2.0 ** (1.5 / 12))
    logs = []
    s_hz, n_hz, point, ci, n_chunks = ref.pitch.estimate_pitch_chroma(src3, nc3, SR, log=logs.append)
    out["pitch_B"] = {"inputs": "synth(3000,75,bpm=112) vs synth(3000,60,bpm=112,speed=1.25,pitch_mult=1.25*2^(1.5/12))",
                      "src_hz": [hx(v) for v in s_hz], "nc_hz": [hx(v) for v in n_hz], "point_st": hx(point),
                      "ci_st": [hx(v) for v in ci], "n_chunks": n_chunks, "logs": logs}
    # ---- case C: waveform xcorr (xcorr.py:54-162) on a 120 s pair, 0.3 % speed difference + noise
    a = synth.synth(4000, 120.0, SR, bpm=124.0)
    b = scipy.signal.resample_poly(a, 1000, 1003).astype(np.float32)
    b = (b + np.random.default_rng(4000).standard_normal(len(b)).astype(np.float32) * 0.01).astype(np.float32)
    np.save(os.path.join(td, "a.npy"), a)
    np.save(os.path.join(td, "b.npy"), b)
    ratio, quality = ref.xcorr.estimate_speed_xcorr(os.path.join(td, "a.npy"), os.path.join(td, "b.npy"))
    out["xcorr_C"] = {"inputs": "synth(4000,120,bpm=124) vs resample_poly(1000,1003)+0.01*N(0,1) rng(4000)",
                      "ratio": hx(ratio), "quality": hx(quality), "label": ref.xcorr.quality_label(quality)}
    # ---- case D: find_content_offset (xcorr.py:165-259)
    body = synth.synth(77, 60.0, SR, bpm=110.0)
    intro = synth.synth(78, 12.0, SR, bpm=90.0) * 0.3
    srcD = np.concatenate([intro, body]).astype(np.float32)
    ncD = scipy.signal.resample_poly(body, 4, 5).astype(np.float32)
    off, speed = ref.xcorr.find_content_offset(srcD, ncD, SR)
    out["align_D"] = {"offset_sec": hx(off), "speed": hx(speed)}

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pipeline_golden.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True, ensure_ascii=False)
print("wrote pipeline_golden.json")
print({k: out["pipeline_A"][k] for k in ("classification", "n")}, float.fromhex(out["pipeline_A"]["tempo_ratio"]))
print(out["xcorr_C"], out["align_D"])
