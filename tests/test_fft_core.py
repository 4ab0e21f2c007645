"""CPU: the butterfly network of csrc/fft_core.cuh (the 32-point in-register FFT both STFT kernels are built from),
compiled as plain host C++ — scalar code with the same products as the packed-FP32 device path — against numpy."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nightcore-to-flac-analyzer_b200", "csrc")

SHIM = r"""
#include "fft_core.cuh"
extern "C" void fft32(float *xy) {            // 32 interleaved complex values, in place, natural output order
    ncfa::cf v[32];
    for (int i = 0; i < 32; ++i) v[i] = ncfa::cf{xy[2 * i], xy[2 * i + 1]};
    ncfa::fft32_dif(v);
    for (int k = 0; k < 32; ++k) { xy[2 * k] = v[ncfa::br5(k)].x; xy[2 * k + 1] = v[ncfa::br5(k)].y; }
}
extern "C" void twiddle32(int t, float *xy) {  // xy *= W_32^t, t in [0, 16)
    ncfa::cf d{xy[0], xy[1]}, r = d;
    switch (t) {
#define C(T) case T: r = ncfa::mul_w32<T>(d); break;
        C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
#undef C
    }
    xy[0] = r.x; xy[1] = r.y;
}
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("fft_core")
    src = d / "shim.cpp"
    src.write_text(SHIM)
    so = d / "libfftcore.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", CSRC, str(src),
                           "-o", str(so)])
    return ctypes.CDLL(str(so))


def test_fft32_matches_numpy(lib):
    rng = np.random.default_rng(0)
    for _ in range(20):
        z = (rng.standard_normal(32) + 1j * rng.standard_normal(32)).astype(np.complex64)
        buf = z.view(np.float32).copy()
        lib.fft32(buf.ctypes.data_as(ctypes.c_void_p))
        got = buf.view(np.complex64)
        want = np.fft.fft(z.astype(np.complex128))
        assert np.max(np.abs(got - want)) <= 2e-6 * np.max(np.abs(want))


def test_fft32_of_an_impulse_and_a_tone_is_exact(lib):
    z = np.zeros(32, np.complex64)
    z[0] = 1.0
    buf = z.view(np.float32).copy()
    lib.fft32(buf.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(buf.view(np.complex64), np.ones(32, np.complex64))
    z = np.ones(32, np.complex64)
    buf = z.view(np.float32).copy()
    lib.fft32(buf.ctypes.data_as(ctypes.c_void_p))
    want = np.zeros(32, np.complex64)
    want[0] = 32.0
    assert np.array_equal(buf.view(np.complex64), want)


@pytest.mark.parametrize("t", range(16))
def test_twiddle_table(lib, t):
    xy = np.array([0.3, -1.7], np.float32)
    lib.twiddle32(t, xy.ctypes.data_as(ctypes.c_void_p))
    want = complex(0.3, -1.7) * np.exp(-2j * np.pi * t / 32)
    assert abs(complex(xy[0], xy[1]) - want) < 3e-7
