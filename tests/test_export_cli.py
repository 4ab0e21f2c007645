"""Caller-side modules next to the hot path: export.py and cli.py of the drop-in against the
reference's own export.py / cli.py output on the same AnalysisResult (tests/golden/export_golden.json, produced by
tests/golden/make_export_golden.py).  CPU tests build the result from golden bootstrap numbers; the GPU test drives the
whole CLI."""
import json
import os

import numpy as np
import pytest

import golden_inputs as gi
from conftest import fromhex

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eg():
    with open(os.path.join(ROOT, "tests", "golden", "export_golden.json")) as f:
        return json.load(f)


def gv6_result(golden, with_durations=True):
    from nightcore_analyzer import consensus as c
    nc_t, src_t = gi.gv1()
    g1, g6 = golden["GV1"], golden["GV6"]
    tempo = (fromhex(g1["point"]), (fromhex(g1["lo"]), fromhex(g1["hi"])))
    if with_durations:
        src_hz = [440.0] * 7
        nc_hz = [440.0 * 2.0 ** (st / 12.0) for st in gi.GV5_SHIFT]
        pitch = (fromhex(g6["pitch_ratio"]), tuple(fromhex(v) for v in g6["pitch_ci"]))
        return c._assemble(src_hz, nc_hz, list(src_t), list(nc_t), c._valid(list(src_t)), c._valid(list(nc_t)), pitch,
                           tempo, (7, 7), 144.0, 180.0)
    return c._assemble([], [], list(src_t), list(nc_t), c._valid(list(src_t)), c._valid(list(nc_t)), (1.0, (1.0, 1.0)),
                       tempo, (0, 0), None, None)


def test_to_dict_and_files_byte_identical(golden, eg, tmp_path):
    from nightcore_analyzer import export
    res = gv6_result(golden)
    assert json.dumps(export.to_dict(res), indent=2) == eg["to_dict_json"]
    export.export_json(res, tmp_path / "r.json")
    assert (tmp_path / "r.json").read_text(encoding="utf-8") == eg["json_file"]
    export.export_csv(res, tmp_path / "r.csv")
    assert open(tmp_path / "r.csv", newline="").read() == eg["csv"]
    assert json.dumps(export.cli_dict(res), indent=2) == eg["cli_json"]


def test_missing_durations_and_pitch(golden, eg, tmp_path):
    from nightcore_analyzer import export
    res = gv6_result(golden, with_durations=False)
    assert json.dumps(export.to_dict(res), indent=2) == eg["to_dict_json_nodur"]
    export.export_csv(res, tmp_path / "r.csv")
    assert open(tmp_path / "r.csv", newline="").read() == eg["csv_nodur"]


def test_cli_argument_validation(tmp_path, capsys):
    from nightcore_analyzer import cli
    rc = cli.main(["--nightcore", str(tmp_path / "missing_nc.npy"), "--source", str(tmp_path / "missing_src.npy"),
                   "--hop", "10", "--window", "10"])
    err = capsys.readouterr().err
    assert rc == 2
    assert "ERROR: Nightcore file not found:" in err and "ERROR: Source file not found:" in err
    assert "ERROR: --hop must be less than --window for overlapping windows" in err
    assert cli.main(["--batch", str(tmp_path / "nope.txt")]) == 2
    # single-pair stages cannot be combined with a manifest (run_batch has no intro trim / alignment stage)
    (tmp_path / "m.txt").write_text("a.npy,b.npy\n")
    capsys.readouterr()
    assert cli.main(["--batch", str(tmp_path / "m.txt"), "--auto-align"]) == 2
    assert "cannot be combined with --batch" in capsys.readouterr().err
    assert cli.main(["--batch", str(tmp_path / "m.txt"), "--src-trim-sec", "1.5"]) == 2


def test_cli_batch_reports_unreadable_entries_per_entry(tmp_path):
    """A manifest whose files cannot be read yields one "error" entry per line and exit code 1 — no traceback, and no
    device work (nothing is left to analyse)."""
    from nightcore_analyzer import cli
    (tmp_path / "m.txt").write_text(f"{tmp_path/'no_nc.npy'},{tmp_path/'no_src.npy'}\n")
    rc = cli.main(["--batch", str(tmp_path / "m.txt"), "-o", str(tmp_path / "b.json"), "-q"])
    rows = json.loads((tmp_path / "b.json").read_text())
    assert rc == 1 and len(rows) == 1 and "error" in rows[0] and rows[0]["nightcore"].endswith("no_nc.npy")


@pytest.mark.gpu
def test_cli_end_to_end(engine, tmp_path, capsys):
    from nightcore_analyzer import cli, export
    import nightcore_analyzer as na
    from oracle import synth
    src, nc = synth.make_pair(1000, 45.0, 22050)
    np.save(tmp_path / "nc.npy", nc)
    np.save(tmp_path / "src.npy", src)
    out = tmp_path / "res.json"
    rc = cli.main(["-n", str(tmp_path / "nc.npy"), "-s", str(tmp_path / "src.npy"), "-o", str(out), "--quiet"])
    assert rc == 0 and capsys.readouterr().out == ""
    want = na.run_arrays(nc, src, 22050, log=None)
    assert out.read_text(encoding="utf-8") == json.dumps(export.cli_dict(want), indent=2)
    # a pair the energy gate empties → exit code 1 with the reference's message on stderr
    np.save(tmp_path / "z.npy", np.zeros(1000, np.float32))
    rc = cli.main(["-n", str(tmp_path / "z.npy"), "-s", str(tmp_path / "z.npy"), "-q"])
    assert rc == 1 and "ERROR: All windows were discarded by the energy gate." in capsys.readouterr().err
    # batch manifest
    (tmp_path / "m.txt").write_text(f"{tmp_path/'nc.npy'},{tmp_path/'src.npy'}\n# comment\n{tmp_path/'z.npy'}\t{tmp_path/'z.npy'}\n")
    rc = cli.main(["--batch", str(tmp_path / "m.txt"), "-o", str(tmp_path / "b.json"), "-q"])
    rows = json.loads((tmp_path / "b.json").read_text())
    assert rc == 1 and len(rows) == 2 and rows[0]["tempo_ratio"] == round(want.tempo_ratio, 8) and "error" in rows[1]
