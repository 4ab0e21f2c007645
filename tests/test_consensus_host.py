"""CPU: host-side decision logic of the drop-in consensus module (classification, half-time flip,
rubberband parameters, warnings, report string) against the reference's own outputs (golden GV6*,
produced by tests/golden/make_golden.py from /root/reference/nightcore_analyzer/consensus.py).
The bootstrap numbers are fed in from the golden file — the kernels that produce them are GPU tests."""
import numpy as np

import golden_inputs as gi
from conftest import fromhex


def _stats(golden, name):
    g = golden[name]
    return fromhex(g["point"]), (fromhex(g["lo"]), fromhex(g["hi"]))


def test_build_result_host_logic_gv6(golden):
    from nightcore_analyzer import consensus as c
    nc_t, src_t = gi.gv1()
    src_hz = [440.0] * 7
    nc_hz = [440.0 * 2.0 ** (st / 12.0) for st in gi.GV5_SHIFT]
    g6 = golden["GV6"]
    tempo_stats = _stats(golden, "GV1")
    pitch_stats = (fromhex(g6["pitch_ratio"]), tuple(fromhex(v) for v in g6["pitch_ci"]))
    res = c._assemble(src_hz, nc_hz, list(src_t), list(nc_t), c._valid(list(src_t)), c._valid(list(nc_t)), pitch_stats,
                      tempo_stats, (7, 7), 144.0, 180.0)
    assert res.classification == g6["classification"]
    assert res.tempo_ratio == fromhex(g6["tempo_ratio"]) and res.pitch_ratio == fromhex(g6["pitch_ratio"])
    assert res.rubberband == g6["rubberband"]
    assert res.warnings == g6["warnings"]
    assert str(res) == g6["str"]
    assert [res.n_source_pitch_windows, res.n_nc_pitch_windows, res.n_source_tempo_windows,
            res.n_nc_tempo_windows] == g6["n"]
    assert res.nc_median_bpm == fromhex(g6["nc_median_bpm"]) and res.src_median_bpm == fromhex(g6["src_median_bpm"])


def test_valid_and_percentile_args():
    from nightcore_analyzer import consensus as c
    v = c._valid([None, 1.0, float("nan"), -2.0, 0.0, float("inf"), 3.5])
    assert v.tolist() == [1.0, 3.5] and v.dtype == np.float64
    lo, hi = c._percentile_args(0.95)
    assert lo == (1.0 - 0.95) / 2.0 * 100 and hi == (1.0 - (1.0 - 0.95) / 2.0) * 100   # 2.500000000000002 / 97.5


def test_insufficient_windows_message():
    from nightcore_analyzer import consensus as c
    import pytest
    with pytest.raises(ValueError, match=r"Insufficient valid tempo windows \(source: 2, nightcore: 0\)"):
        c.build_result([], [], [120.0, 121.0, None], [None], nc_duration=1.0, src_duration=1.0)


def test_classify_thresholds():
    from nightcore_analyzer import consensus as c
    assert c._classify(1.25, 1.25, (1.24, 1.26), (1.24, 1.26)) == "pure_nightcore"
    assert c._classify(1.25, 1.0, (1.24, 1.26), (0.99, 1.01)) == "time_stretch_only"
    assert c._classify(1.0, 1.1, (0.99, 1.01), (1.09, 1.11)) == "independent_pitch_shift"
    assert c._classify(1.0, 0.9, (0.99, 1.01), (0.89, 0.91)) == "ambiguous"
    assert c._classify(1.25, 1.28, (1.2, 1.3), (1.2, 1.3)) == "pure_nightcore"      # overlapping CIs, diff <= 2·tol


def test_small_median_is_np_median_bit_for_bit():
    """batch.analyse_staged and consensus._assemble take their per-pair medians through consensus._median_small."""
    import numpy as np
    from nightcore_analyzer import consensus
    rng = np.random.default_rng(3)
    for _ in range(5000):
        a = rng.uniform(40.0, 260.0, int(rng.integers(1, 48)))
        if rng.random() < 0.4:
            a = np.round(a, 1)          # ties, as quantised BPM values produce them
        assert consensus._median_small(a) == float(np.median(a))
    assert consensus._median_small([3.0, 1.0]) == 2.0
