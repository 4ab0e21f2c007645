"""
Serialisation of ``AnalysisResult`` — drop-in for the reference's ``nightcore_analyzer/export.py``
(export.py:20-98: ``to_dict``, ``export_json``, ``export_csv`` with the same keys, rounding and column order).
SURVEY.md §8(f) row 2: a caller-side module next to the hot path; pure host code, no arithmetic beyond rounding.
"""
from __future__ import annotations

import csv
import json
from pathlib import Path
from typing import Union

from .consensus import AnalysisResult

PathLike = Union[str, Path]


def _r(value, digits):
    """round(value, digits) when the value is truthy, else None (export.py:38-49 use `if x else None`)."""
    return round(value, digits) if value else None


def _duration_ratio(result: AnalysisResult):
    if result.nc_duration and result.src_duration:
        return round(result.src_duration / result.nc_duration, 8)
    return None


def cli_dict(result: AnalysisResult) -> dict:
    """The JSON object the reference CLI prints (cli.py:171-186): the first seven keys of ``to_dict``."""
    return {
        "classification": result.classification,
        "tempo_ratio": round(result.tempo_ratio, 8),
        "pitch_ratio": round(result.pitch_ratio, 8),
        "tempo_ci_95": [round(v, 8) for v in result.tempo_ci],
        "pitch_ci_95": [round(v, 8) for v in result.pitch_ci],
        "windows_used": {
            "source_pitch": result.n_source_pitch_windows,
            "nightcore_pitch": result.n_nc_pitch_windows,
            "source_tempo": result.n_source_tempo_windows,
            "nightcore_tempo": result.n_nc_tempo_windows,
        },
        "rubberband": result.rubberband,
    }


def to_dict(result: AnalysisResult) -> dict:
    """export.py:20-51 — JSON-serialisable dict, key order included."""
    base = cli_dict(result)
    out = {"classification": base.pop("classification"), "warnings": result.warnings}
    out.update(base)
    out["durations"] = {
        "nightcore_sec": _r(result.nc_duration, 3),
        "source_sec": _r(result.src_duration, 3),
        "duration_ratio": _duration_ratio(result),
    }
    out["median_bpms"] = {"nightcore": _r(result.nc_median_bpm, 2), "source": _r(result.src_median_bpm, 2)}
    return out


def export_json(result: AnalysisResult, path: PathLike) -> None:
    """export.py:54-56."""
    Path(path).write_text(json.dumps(to_dict(result), indent=2), encoding="utf-8")


# (column, getter) in the reference's column order (export.py:68-97); missing optional values become ""
def _csv_row(result: AnalysisResult) -> dict:
    rb = result.rubberband
    blank = lambda v: "" if v is None else v  # noqa: E731
    cols = [
        ("classification", result.classification),
        ("tempo_ratio", round(result.tempo_ratio, 8)),
        ("pitch_ratio", round(result.pitch_ratio, 8)),
        ("tempo_ci_95_lo", round(result.tempo_ci[0], 8)),
        ("tempo_ci_95_hi", round(result.tempo_ci[1], 8)),
        ("pitch_ci_95_lo", round(result.pitch_ci[0], 8)),
        ("pitch_ci_95_hi", round(result.pitch_ci[1], 8)),
        ("source_pitch_windows", result.n_source_pitch_windows),
        ("nightcore_pitch_windows", result.n_nc_pitch_windows),
        ("source_tempo_windows", result.n_source_tempo_windows),
        ("nightcore_tempo_windows", result.n_nc_tempo_windows),
    ]
    for col, key in (("rb_time_ratio", "time_ratio"), ("rb_pitch_semitones", "pitch_semitones"),
                     ("rb_nc_to_source_speed", "nc_to_source_speed"), ("rb_cli_command", "cli_command"),
                     ("rb_dur_time_ratio", "duration_time_ratio"), ("rb_dur_pitch_semitones", "duration_pitch_semitones"),
                     ("rb_dur_cli_command", "duration_cli_command")):
        cols.append((col, rb.get(key, "")))
    cols += [
        ("nc_median_bpm", blank(_r(result.nc_median_bpm, 2))),
        ("src_median_bpm", blank(_r(result.src_median_bpm, 2))),
        ("nc_duration_sec", blank(_r(result.nc_duration, 3))),
        ("src_duration_sec", blank(_r(result.src_duration, 3))),
        ("duration_ratio", blank(_duration_ratio(result))),
        ("warnings", " | ".join(result.warnings)),
    ]
    return dict(cols)


def export_csv(result: AnalysisResult, path: PathLike) -> None:
    """export.py:59-98 — header row + one data row."""
    row = _csv_row(result)
    with open(path, "w", newline="", encoding="utf-8") as fh:
        writer = csv.DictWriter(fh, fieldnames=list(row))
        writer.writeheader()
        writer.writerow(row)
