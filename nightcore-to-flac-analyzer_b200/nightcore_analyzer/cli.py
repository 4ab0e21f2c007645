"""
Command-line front-end — drop-in for the reference's ``nightcore_analyzer/cli.py`` (cli.py:25-206): the same flags,
validation messages, exit codes (2 = bad arguments, 1 = analysis failed, 0 = ok), JSON object and trailing
human-readable report.  SURVEY.md §8(f) row 2.  Added beside the reference surface: ``--batch MANIFEST``, a text file
with one ``nightcore<TAB or comma>source`` pair per line, analysed in one data-parallel pass (``pipeline.run_batch``)
and emitted as a JSON list.

    python -m nightcore_analyzer.cli --nightcore nc.npy --source src.npy --output results.json
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path
from typing import List, Optional

from . import pipeline
from .export import cli_dict
from .io import ENERGY_GATE_DB, HOP_SEC, SILENCE_STRIP_DB, WINDOW_SEC, load_audio

_FLOAT_OPTS = (  # flag, default, metavar, help  (cli.py:59-111)
    ("--window", WINDOW_SEC, "SEC", "Analysis window duration in seconds"),
    ("--hop", HOP_SEC, "SEC", "Hop between consecutive windows in seconds (< --window for overlap)"),
    ("--energy-gate", ENERGY_GATE_DB, "DB", "Discard windows whose RMS energy is below peak + ENERGY_GATE dB"),
    ("--silence-strip-db", SILENCE_STRIP_DB, "DB", "Top-dB threshold for trimming leading/trailing silence"),
    ("--src-trim-sec", 0.0, "SEC", "Trim this many seconds from the start of the source before analysis"),
)
_SWITCHES = (  # flag(s), help
    (("--no-silence-strip",), "Disable leading/trailing silence stripping entirely."),
    (("--auto-align",), "Attempt automatic intro-offset detection via RMS envelope correlation."),
    (("--quiet", "-q"), "Suppress progress output (errors still go to stderr)"),
)


def _build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(
        prog="python -m nightcore_analyzer.cli",
        description="Tempo ratio and pitch ratio between a nightcore track and its source, plus the Rubber Band "
                    "parameters that reconstruct the original (B200 build).",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
    )
    p.add_argument("--nightcore", "-n", metavar="FILE", help="Nightcore audio file (.npy float mono at 22050 Hz, or PCM .wav)")
    p.add_argument("--source", "-s", metavar="FILE", help="Source audio file the nightcore was derived from")
    p.add_argument("--batch", metavar="MANIFEST", help="Text file of 'nightcore,source' lines analysed in one GPU pass")
    p.add_argument("--output", "-o", metavar="FILE", help="Write JSON results to this file (default: print to stdout)")
    for flag, default, metavar, text in _FLOAT_OPTS:
        p.add_argument(flag, type=float, default=default, metavar=metavar, help=text)
    for flags, text in _SWITCHES:
        p.add_argument(*flags, action="store_true", default=False, help=text)
    return p


def _emit(json_text: str, args, tail=None) -> None:
    if args.output:
        out_path = Path(args.output)
        out_path.write_text(json_text, encoding="utf-8")
        if not args.quiet:
            print(f"\nResults written to: {out_path}")
    else:
        print()
        print(json_text)
    if tail is not None and not args.quiet:
        print()
        print(tail)


def _run_batch(args, silence_strip_db) -> int:
    pairs, names = [], []
    for line in Path(args.batch).read_text(encoding="utf-8").splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        parts = [s.strip() for s in (line.split("\t") if "\t" in line else line.split(","))]
        if len(parts) != 2:
            print(f"ERROR: malformed manifest line: {line!r}", file=sys.stderr)
            return 2
        names.append(parts)
    load_errors = {}
    for k, (nc_path, src_path) in enumerate(names):
        try:                                  # an unreadable entry becomes that entry's "error", not a traceback
            nc, sr = load_audio(nc_path)
            src, _ = load_audio(src_path, sr=sr)
        except Exception as exc:  # noqa: BLE001
            load_errors[k] = exc
            continue
        pairs.append((nc, src))
    analysed = iter(pipeline.run_batch(pairs, window_sec=args.window, hop_sec=args.hop, energy_gate_db=args.energy_gate,
                                       silence_strip_db=silence_strip_db) if pairs else [])
    results = [load_errors[k] if k in load_errors else next(analysed) for k in range(len(names))]
    out = []
    for (nc_path, src_path), res in zip(names, results):
        entry = {"nightcore": nc_path, "source": src_path}
        entry.update({"error": str(res)} if isinstance(res, Exception) else cli_dict(res))
        out.append(entry)
    _emit(json.dumps(out, indent=2), args)
    return 1 if any(isinstance(r, Exception) for r in results) else 0


def main(argv: Optional[List[str]] = None) -> int:
    args = _build_parser().parse_args(argv)
    silence_strip_db = None if args.no_silence_strip else args.silence_strip_db

    problems = []
    if args.batch:
        if not Path(args.batch).exists():
            problems.append(f"Manifest file not found: {args.batch}")
        if args.src_trim_sec != 0.0 or args.auto_align:
            # run_batch has no per-pair intro trim / alignment stage (pipeline.py:91-139 runs only in the single-pair path)
            problems.append("--src-trim-sec / --auto-align apply to a single pair and cannot be combined with --batch")
    else:
        if not args.nightcore or not args.source:
            problems.append("--nightcore and --source are required (or --batch MANIFEST)")
        else:
            if not Path(args.nightcore).exists():
                problems.append(f"Nightcore file not found: {Path(args.nightcore)}")
            if not Path(args.source).exists():
                problems.append(f"Source file not found:    {Path(args.source)}")
    if args.hop >= args.window:
        problems.append("--hop must be less than --window for overlapping windows")
    if problems:
        for msg in problems:
            print(f"ERROR: {msg}", file=sys.stderr)
        return 2
    if args.batch:
        return _run_batch(args, silence_strip_db)

    try:
        result = pipeline.run(
            str(Path(args.nightcore)), str(Path(args.source)),
            window_sec=args.window, hop_sec=args.hop, energy_gate_db=args.energy_gate,
            silence_strip_db=silence_strip_db, src_trim_sec=args.src_trim_sec,
            auto_align=args.auto_align and args.src_trim_sec == 0.0,
            log=None if args.quiet else print,
        )
    except Exception as exc:  # the reference reports any pipeline failure as exit code 1 (cli.py:166-168)
        print(f"\nERROR: {exc}", file=sys.stderr)
        return 1
    _emit(json.dumps(cli_dict(result), indent=2), args, tail=result)
    return 0


if __name__ == "__main__":
    sys.exit(main())
