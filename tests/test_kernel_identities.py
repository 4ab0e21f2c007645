"""CPU: the algebraic identities the CUDA kernels are built on, checked in numpy against the plain definitions
(the kernels themselves are compared with the oracle on the GPU; these pin the mathematics they restate)."""
import numpy as np


def test_real_fft_unpack_identity():
    """stft_core.cuh PostStage: a 2048-point real FFT as a packed 1024-point complex FFT.  With z[n] = x[2n] + i x[2n+1],
    Z = FFT_1024(z), E = (Z[k] + conj Z[N-k]) / 2, O = (Z[k] - conj Z[N-k]) / (2i), W = exp(-2 pi i k / 2048):
    X[k] = E + W O and X[1024-k] = conj(E - W O) for k = 0..511, X[512] = conj(Z[512])."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(2048)
    X = np.fft.rfft(x)
    Z = np.fft.fft(x[0::2] + 1j * x[1::2])
    k = np.arange(512)
    Zc = np.conj(Z[(1024 - k) % 1024])
    E = 0.5 * (Z[k] + Zc)
    O = (Z[k] - Zc) / 2j
    W = np.exp(-2j * np.pi * k / 2048)
    assert np.max(np.abs(E + W * O - X[k])) < 1e-10
    assert np.max(np.abs(np.conj(E - W * O) - X[1024 - k])) < 1e-10
    assert abs(np.conj(Z[512]) - X[512]) < 1e-10
    # the kernel stores |.|^2 only, so conj / no conj is immaterial there
    assert abs(abs(Z[512]) ** 2 - abs(X[512]) ** 2) < 1e-9


def test_four_step_fft_identity():
    """stft_core.cuh: 1024 = 32 x 32.  Lane n2 holds z[32 n1 + n2]; FFT_32 over n1, multiply by W_1024^(k1 n2),
    transpose, FFT_32 over n2 gives Z[k1 + 32 k2]."""
    rng = np.random.default_rng(2)
    z = rng.standard_normal(1024) + 1j * rng.standard_normal(1024)
    a = z.reshape(32, 32)                       # a[n1, n2]
    b = np.fft.fft(a, axis=0)                   # b[k1, n2]
    b = b * np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 1024)
    c = np.fft.fft(b, axis=1)                   # c[k1, k2]
    Z = np.fft.fft(z)
    k1, k2 = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
    assert np.max(np.abs(c - Z[k1 + 32 * k2])) < 1e-9


def test_block_partial_initial_sums_identity():
    """tempo.cu tg_r0_kernel: S(f) = sum_{J<W} z[f+J] e^{i theta J} assembled from 32-sample block partials
    P[b] = sum_{j<32} z[32 b + j] e^{i theta j}:  S(32 t) = sum_{m < W//32} e^{i theta 32 m} P[t + m] + tail."""
    rng = np.random.default_rng(3)
    W = 2756
    z = rng.exponential(1.0, 32 * 40 + W + 64)
    theta = 2 * np.pi / W
    nb = len(z) // 32
    j = np.arange(32)
    for harm in (1, 2):
        P = np.array([np.sum(z[32 * b : 32 * b + 32] * np.exp(1j * harm * theta * j)) for b in range(nb)])
        for t in (0, 3, 39):
            f = 32 * t
            direct = np.sum(z[f : f + W] * np.exp(1j * harm * theta * np.arange(W)))
            M = W // 32
            s = sum(np.exp(1j * harm * theta * 32 * m) * P[t + m] for m in range(M))
            J = np.arange(32 * M, W)
            s += np.sum(z[f + J] * np.exp(1j * harm * theta * J))
            assert abs(s - direct) < 1e-9 * abs(direct)


def test_halfband_even_taps_hit_odd_samples():
    """chroma.cu decimate2_kernel: a 127-tap half-band FIR has zeros at every even offset from the centre (index 63),
    i.e. only h[63] and the even indices h[0], h[2], ... are non-zero; out[t] = sum_i h[i] x[2t + 63 - i] then reads
    the even input sample 2t once (centre tap) and odd input samples otherwise."""
    from oracle import librosa_restated as lr
    h = lr.halfband_fir()
    assert len(h) == 127
    odd_idx = np.arange(1, 127, 2)
    odd_idx = odd_idx[odd_idx != 63]
    assert np.max(np.abs(h[odd_idx])) < 1e-15 * np.max(np.abs(h)) + 1e-17
    x = np.random.default_rng(4).standard_normal(4000)
    full = np.convolve(x, h)[63::2][:1900]                 # centred 2:1 decimation
    t = np.arange(200, 1700)
    he = h[0::2]                                            # 64 even-index taps
    mine = h[63] * x[2 * t] + np.array([np.sum(he * x[2 * tt + 63 - 2 * np.arange(64)]) for tt in t])
    assert np.max(np.abs(mine - full[t])) < 1e-12
