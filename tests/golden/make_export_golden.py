"""Generate tests/golden/export_golden.json with the REFERENCE's own export.py (to_dict / export_json / export_csv) and
the JSON object of its cli.py (cli.py:171-186) on the GV6 AnalysisResult built by the reference's consensus.build_result.
Build container only:   python tests/golden/make_export_golden.py"""
import importlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402

ref = reference_shim.load_reference()
exp = importlib.import_module("nc_reference.export")
g = 60.0 * 22050 / 512
src = [g / 22] * 20 + [g / 21] * 9 + [g / 43] * 4 + [g / 23] * 2
nc = [g / 17] * 15 + [g / 18] * 8 + [g / 35] * 3 + [g / 16] * 1
shift = np.array([4, 4, 3, 4, 5, 4, 4]) / 3.0
res = ref.consensus.build_result([440.0] * 7, [440.0 * 2.0 ** (st / 12.0) for st in shift], src, nc,
                                 nc_duration=144.0, src_duration=180.0)
res2 = ref.consensus.build_result([], [], src, nc)   # no durations, no pitch
out = {"to_dict_json": json.dumps(exp.to_dict(res), indent=2), "to_dict_json_nodur": json.dumps(exp.to_dict(res2), indent=2)}
with tempfile.TemporaryDirectory() as td:
    exp.export_csv(res, os.path.join(td, "r.csv"))
    out["csv"] = open(os.path.join(td, "r.csv"), newline="").read()
    exp.export_json(res, os.path.join(td, "r.json"))
    out["json_file"] = open(os.path.join(td, "r.json")).read()
    exp.export_csv(res2, os.path.join(td, "r2.csv"))
    out["csv_nodur"] = open(os.path.join(td, "r2.csv"), newline="").read()
out["cli_json"] = json.dumps({
    "classification": res.classification, "tempo_ratio": round(res.tempo_ratio, 8), "pitch_ratio": round(res.pitch_ratio, 8),
    "tempo_ci_95": [round(res.tempo_ci[0], 8), round(res.tempo_ci[1], 8)],
    "pitch_ci_95": [round(res.pitch_ci[0], 8), round(res.pitch_ci[1], 8)],
    "windows_used": {"source_pitch": res.n_source_pitch_windows, "nightcore_pitch": res.n_nc_pitch_windows,
                     "source_tempo": res.n_source_tempo_windows, "nightcore_tempo": res.n_nc_tempo_windows},
    "rubberband": res.rubberband}, indent=2)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "export_golden.json"), "w") as f:
    json.dump(out, f, indent=1, ensure_ascii=False)
print("wrote export_golden.json")
