"""Generate tests/golden/golden.json by running the REFERENCE's own consensus.py / pitch.py
(loaded standalone from /root/reference — they import only math/numpy) on fully specified inputs
(SURVEY.md §4, GV1–GV7) plus numpy PCG64 known answers.  Run in the build container only:

    python tests/golden/make_golden.py

Floats are stored as hex strings (float.hex) so the comparison is bit-exact.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.reference_shim import load_reference_standalone  # noqa: E402

cons = load_reference_standalone("consensus")
pitch = load_reference_standalone("pitch")

hx = lambda v: float(v).hex()
g = 60.0 * 22050 / 512
out = {"numpy": np.__version__}

# ---- inputs (also re-created by the tests from the same recipe)
def gv1_inputs():
    src = [g / 22] * 20 + [g / 21] * 9 + [g / 43] * 4 + [g / 23] * 2
    nc = [g / 17] * 15 + [g / 18] * 8 + [g / 35] * 3 + [g / 16] * 1
    return np.array(nc), np.array(src)

def gv2_inputs():
    r = np.random.default_rng(1234)
    a = 120 + r.normal(0, 2, 35)
    b = 150 + r.normal(0, 3, 27)
    return b, a

def gv3_inputs():
    r = np.random.default_rng(99)
    s = 0.5 + r.normal(0, .004, 360)
    n = 0.4 + r.normal(0, .004, 361)
    return n, s

def gv4_inputs():
    r = np.random.default_rng(7)
    s = 0.5 + r.normal(0, .004, 7200)
    n = 0.4 + r.normal(0, .004, 7201)
    return n, s

def pack(point, ci):
    return {"point": hx(point), "lo": hx(ci[0]), "hi": hx(ci[1])}

nc, src = gv1_inputs()
out["GV1"] = pack(*cons._bootstrap_ratio(nc, src))
b, a = gv2_inputs()
out["GV2"] = pack(*cons._bootstrap_ratio(b, a))
n, s = gv3_inputs()
out["GV3"] = pack(*cons.compute_ibi_ratio(n, s))
n, s = gv4_inputs()
out["GV4"] = pack(*cons.compute_ibi_ratio(n, s))

# GV5: pitch-chunk bootstrap, pitch.py:143-150 (restated call sequence on the reference's numpy)
shift = np.array([4, 4, 3, 4, 5, 4, 4]) / 3.0
rng = np.random.default_rng(0)
boots = np.array([float(np.median(rng.choice(shift, size=7, replace=True))) for _ in range(2000)])
out["GV5"] = {"point": hx(np.median(shift)), "lo": hx(np.percentile(boots, 2.5)), "hi": hx(np.percentile(boots, 97.5)),
              "boots_head": [hx(v) for v in boots[:16]]}

# GV6: build_result
nc_t, src_t = gv1_inputs()
src_hz = [440.0] * 7
nc_hz = [440.0 * 2.0 ** (st / 12.0) for st in shift]
res = cons.build_result(src_hz, nc_hz, list(src_t), list(nc_t), nc_duration=144.0, src_duration=180.0)
out["GV6"] = {
    "classification": res.classification, "tempo_ratio": hx(res.tempo_ratio), "pitch_ratio": hx(res.pitch_ratio),
    "tempo_ci": [hx(v) for v in res.tempo_ci], "pitch_ci": [hx(v) for v in res.pitch_ci],
    "rubberband": res.rubberband, "warnings": res.warnings, "str": str(res),
    "n": [res.n_source_pitch_windows, res.n_nc_pitch_windows, res.n_source_tempo_windows, res.n_nc_tempo_windows],
    "nc_median_bpm": hx(res.nc_median_bpm), "src_median_bpm": hx(res.src_median_bpm),
}
# a second build_result exercising the half-time flip and warnings
res2 = cons.build_result([None] * 3, [None] * 3, list(nc_t[:10]), list(src_t[:10]), nc_duration=144.0, src_duration=180.0)
out["GV6b"] = {"classification": res2.classification, "tempo_ratio": hx(res2.tempo_ratio),
               "tempo_ci": [hx(v) for v in res2.tempo_ci], "warnings": res2.warnings, "str": str(res2),
               "rubberband": res2.rubberband}
res3 = cons.build_result(src_hz, nc_hz, list(src_t), list(nc_t))
out["GV6c"] = {"classification": res3.classification, "warnings": res3.warnings, "str": str(res3),
               "rubberband": res3.rubberband}

# GV7: cyclic xcorr peak
a7 = np.array([.9, .1, .3, .2, .8, .1, .05, .7, .1, .2, .1, .3])
out["GV7"] = {str(k): int(pitch._cyclic_xcorr_peak(a7, np.roll(a7, k))) for k in (0, 1, 4, 6, 7, 11)}

# RNG known answers straight from numpy
def pcg(seed):
    st = np.random.PCG64(seed).state["state"]
    return {"state": hex(st["state"]), "inc": hex(st["inc"]),
            "raw": [hex(int(v)) for v in np.random.PCG64(seed).random_raw(4)]}
out["PCG64_42"] = pcg(42)
out["PCG64_0"] = pcg(0)
r = np.random.default_rng(42)
out["rng42_int27"] = r.integers(0, 27, 27).tolist()
out["rng42_int35"] = r.integers(0, 35, 35).tolist()
r = np.random.default_rng(0)
out["rng0_int7x3"] = [r.integers(0, 7, 7).tolist() for _ in range(3)]
# resample indices of the first 3 iterations of GV4 (exercises odd/even buffering)
r = np.random.default_rng(42)
idx = []
for _ in range(3):
    idx.append(r.integers(0, 7200, 7200)[:8].tolist())
    idx.append(r.integers(0, 7201, 7201)[:8].tolist())
out["gv4_idx_heads"] = idx

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True, ensure_ascii=False)
print("wrote golden.json")
for k in ("GV1", "GV2", "GV3", "GV4"):
    print(k, {kk: float.fromhex(v) for kk, v in out[k].items()})
print("GV6", out["GV6"]["classification"], float.fromhex(out["GV6"]["pitch_ratio"]))
print("GV7", out["GV7"])
