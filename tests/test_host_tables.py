"""CPU: the host-side constant tables of libncfa (no GPU needed) against the CPU restatement —
the folded CQT contraction matrix K (basis ∘ DFT / sqrt(length)) and the half-band decimator."""
import ctypes

import numpy as np
import pytest

from oracle import librosa_restated as lr


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    import os
    if not os.path.exists(ge.LIB):
        ge._build_native()
    from nightcore_analyzer import _native
    return _native.lib


def test_halfband_taps_match_oracle(lib):
    h = np.zeros(127)
    assert lib.ncfa_host_halfband_taps(h.ctypes.data) == 0
    want = lr.halfband_fir()
    assert np.max(np.abs(h - want)) < 1e-15
    assert abs(h.sum() - 1.0) < 1e-14 and np.allclose(h, h[::-1], atol=1e-17)


@pytest.mark.parametrize("j", [0, 37, 50, 63, 99])
def test_cqt_matrix_equals_basis_times_dft(lib, j):
    """K·frame must equal fft_basis · rfft(frame) / sqrt(length) for every tuning (Appendix A.8)."""
    K = np.zeros((1024, 72), np.float32)
    assert lib.ncfa_host_cqt_matrix(22050, j, K.ctypes.data) == 0
    tuning = float(np.linspace(-0.5, 0.5, 101)[j])
    fmin = lr.note_c1_hz() * 2.0 ** (tuning / 36)
    basis, n_fft, lengths = lr.cqt_octave_basis(22050, fmin)
    assert n_fft == 1024
    x = np.random.default_rng(j).standard_normal(1024)
    want = basis.astype(np.complex128) @ np.fft.rfft(x) / np.sqrt(lengths[-36:])
    got = K[:, :36].astype(np.float64).T @ x + 1j * (K[:, 36:].astype(np.float64).T @ x)
    assert np.max(np.abs(got - want)) <= 2e-7 * np.max(np.abs(want))


def test_cqt_matrix_rejects_bad_arguments(lib):
    K = np.zeros((1024, 72), np.float32)
    assert lib.ncfa_host_cqt_matrix(22050, 100, K.ctypes.data) < 0
    assert lib.ncfa_host_cqt_matrix(22050, 0, None) < 0
    assert lib.ncfa_host_cqt_matrix(96000, 50, K.ctypes.data) < 0      # top-octave filters no longer fit n_fft 1024
    from nightcore_analyzer import _native
    assert "1024" in _native.last_error()


@pytest.mark.parametrize("sr", [22050, 44100, 11025, 16000])
def test_mel_lane_table_is_the_slaney_bank_and_conflict_free(lib, sr):
    """The lane-transposed mel bank of the STFT kernels: (1) scattering it back gives the restated
    librosa.filters.mel matrix exactly (float32); (2) within a group the 32 lanes start at distinct
    residues mod 32, so the lock-step reads of the power spectrum never share a shared-memory bank."""
    lb = np.zeros(128, np.int32)
    qw = np.zeros(4, np.int32)
    wt = np.zeros((128, 32), np.float32)
    assert lib.ncfa_host_mel_lanes(sr, lb.ctypes.data, qw.ctypes.data, wt.ctypes.data) == 0
    want = lr.mel_filter(sr, 2048, 128).astype(np.float32)
    got = np.zeros((128, 1025 + 256), np.float32)
    off = 0
    for q in range(4):
        starts = lb[q * 32:(q + 1) * 32]
        assert starts.min() >= 0
        assert len(set(int(b) % 32 for b in starts)) == 32, (sr, q, sorted(int(b) % 32 for b in starts))
        for lane in range(32):
            band = [lane, 63 - lane, 64 + lane, 127 - lane][q]
            b0 = int(starts[lane])
            got[band, b0:b0 + qw[q]] = wt[off:off + qw[q], lane]
        off += int(qw[q])
    assert off <= 128
    assert np.all(got[:, 1025:] == 0)
    assert np.array_equal(got[:, :1025], want)
