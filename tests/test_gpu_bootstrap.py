"""GPU: PCG64-exact bootstrap kernels vs the reference's golden vectors and the C oracle (bit-exact)."""
import numpy as np
import pytest

from oracle import native
import golden_inputs as gi

pytestmark = pytest.mark.gpu
Q = (2.500000000000002, 97.5)   # consensus.py:264-266: alpha·100 and (1−alpha)·100


def fx(v):
    return float.fromhex(v)


@pytest.mark.parametrize("name,fn,order", [("GV1", gi.gv1, "nc_first"), ("GV2", gi.gv2, "nc_first"),
                                           ("GV3", gi.gv3, "src_first"), ("GV4", gi.gv4, "src_first")])
def test_bootstrap_golden(engine, golden, name, fn, order):
    nc, src = fn()
    a, b = (nc, src) if order == "nc_first" else (src, nc)
    out, boot, _ = engine.bootstrap([(a, b)], 42, 2000, Q[0], Q[1], want_boot=True)
    assert out[0, 0] == fx(golden[name]["point"])
    assert out[0, 1] == fx(golden[name]["lo"])
    assert out[0, 2] == fx(golden[name]["hi"])
    want = native.bootstrap(a, b, 2000, 42)
    assert np.array_equal(boot[0], want)


def test_pitch_bootstrap_golden(engine, golden):
    out, boot, idx = engine.bootstrap([(gi.GV5_SHIFT, None)], 0, 2000, 2.5, 97.5, want_boot=True, want_idx=True)
    assert out[0, 0] == fx(golden["GV5"]["point"]) and out[0, 1] == fx(golden["GV5"]["lo"])
    assert out[0, 2] == fx(golden["GV5"]["hi"])
    assert [float(v).hex() for v in boot[0, :16]] == golden["GV5"]["boots_head"]
    assert idx.reshape(2000, 7)[:3].tolist() == golden["rng0_int7x3"]


def test_resample_indices_bit_exact_with_rejections(engine, golden):
    a, b = np.arange(7200, dtype=float) + 0.5, np.arange(7201, dtype=float) + 0.25
    n_boot = 400
    out, boot, idx = engine.bootstrap([(a, b)], 42, n_boot, Q[0], Q[1], want_boot=True, want_idx=True)
    wboot, widx = native.bootstrap(a, b, n_boot, 42, return_indices=True)
    assert np.array_equal(idx, widx)
    assert np.array_equal(boot[0], wboot)
    heads = idx.reshape(n_boot, 14401)
    assert heads[0, :8].tolist() == golden["gv4_idx_heads"][0]
    assert heads[0, 7200:7208].tolist() == golden["gv4_idx_heads"][1]
    assert heads[2, 7200:7208].tolist() == golden["gv4_idx_heads"][5]


def test_small_kat_indices(engine, golden):
    a, b = np.arange(27, dtype=float), np.arange(35, dtype=float)
    _, _, idx = engine.bootstrap([(a, b)], 42, 1, 50.0, 50.0, want_idx=True)
    assert idx[:27].tolist() == golden["rng42_int27"] and idx[27:].tolist() == golden["rng42_int35"]


def test_batched_jobs_ragged(engine):
    rng = np.random.default_rng(5)
    jobs = []
    for n_a, n_b in [(3, 3), (27, 35), (7, 9), (1, 4), (64, 1), (361, 360), (2, 2)]:
        jobs.append((100 + rng.normal(0, 3, n_a), 80 + rng.normal(0, 3, n_b)))
    out, boot, _ = engine.bootstrap(jobs, 42, 500, Q[0], Q[1], want_boot=True)
    for j, (a, b) in enumerate(jobs):
        want = native.bootstrap(a, b, 500, 42)
        assert np.array_equal(boot[j], want), j
        assert out[j, 0] == float(np.median(a) / np.median(b))
        assert out[j, 1] == float(np.percentile(want, Q[0])) and out[j, 2] == float(np.percentile(want, Q[1]))


def test_ties_and_constant_arrays(engine):
    a = np.array([2.0] * 9 + [3.0] * 4)
    b = np.array([1.5] * 6)
    out, boot, _ = engine.bootstrap([(a, b)], 42, 300, Q[0], Q[1], want_boot=True)
    assert np.array_equal(boot[0], native.bootstrap(a, b, 300, 42))


def test_too_long_arrays_fail_loudly(engine):
    from nightcore_analyzer._native import NcfaError
    with pytest.raises(NcfaError):
        engine.bootstrap([(np.ones(40000), np.ones(40000))], 42, 10, 2.5, 97.5)
