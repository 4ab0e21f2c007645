"""CPU: pin the oracle against the reference's golden vectors and numpy known answers."""
import numpy as np
import pytest

from oracle import native
from oracle import librosa_restated as lr
import golden_inputs as gi


def fx(v):
    return float.fromhex(v)


def oracle_ratio(a_first, b_second, seed=42, n_boot=2000, q=(2.500000000000002, 97.5)):
    boot = native.bootstrap(a_first, b_second, n_boot, seed)
    point = float(np.median(a_first) / np.median(b_second)) if b_second is not None else float(np.median(a_first))
    return point, float(np.percentile(boot, q[0])), float(np.percentile(boot, q[1])), boot


def test_pcg64_raw_kat(golden):
    for seed, key in ((42, "PCG64_42"), (0, "PCG64_0")):
        raw = native.pcg64_raw(seed, 4)
        assert [hex(int(v)) for v in raw] == golden[key]["raw"]
        hi, lo, ihi, ilo = native.pcg64_state(seed)
        assert hex((hi << 64) | lo) == golden[key]["state"]
        assert hex((ihi << 64) | ilo) == golden[key]["inc"]


def test_bounded_draws_kat(golden):
    # default_rng(42): integers(0,27,27) then integers(0,35,35) — second call starts on a buffered half
    a, b = np.arange(27, dtype=float), np.arange(35, dtype=float)
    _, idx = native.bootstrap(a, b, 1, 42, return_indices=True)
    assert idx[:27].tolist() == golden["rng42_int27"]
    assert idx[27:].tolist() == golden["rng42_int35"]
    _, idx = native.bootstrap(np.arange(7, dtype=float), None, 3, 0, return_indices=True)
    assert idx.reshape(3, 7).tolist() == golden["rng0_int7x3"]


def test_lemire_rejections_match_numpy():
    # 7201/7200-sized draws hit real rejections within a few hundred iterations
    r = np.random.default_rng(42)
    want = np.concatenate([np.concatenate([r.integers(0, 7200, 7200), r.integers(0, 7201, 7201)]) for _ in range(300)])
    _, idx = native.bootstrap(np.arange(7200, dtype=float), np.arange(7201, dtype=float), 300, 42, return_indices=True)
    assert np.array_equal(idx, want.astype(np.int32))


@pytest.mark.parametrize("name,fn,order", [("GV1", gi.gv1, "nc_first"), ("GV2", gi.gv2, "nc_first"),
                                           ("GV3", gi.gv3, "src_first"), ("GV4", gi.gv4, "src_first")])
def test_bootstrap_golden(golden, name, fn, order):
    nc, src = fn()
    a, b = (nc, src) if order == "nc_first" else (src, nc)
    point, lo, hi, _ = oracle_ratio(a, b)
    assert point == fx(golden[name]["point"])
    assert lo == fx(golden[name]["lo"])
    assert hi == fx(golden[name]["hi"])


def test_pitch_bootstrap_golden(golden):
    point, lo, hi, boot = oracle_ratio(gi.GV5_SHIFT, None, seed=0, q=(2.5, 97.5))
    assert point == fx(golden["GV5"]["point"]) and lo == fx(golden["GV5"]["lo"]) and hi == fx(golden["GV5"]["hi"])
    assert [float(v).hex() for v in boot[:16]] == golden["GV5"]["boots_head"]


def test_hann_and_mel_against_torch():
    import torch
    w = lr.hann_periodic(2048)
    assert np.max(np.abs(w - torch.hann_window(2048, periodic=True, dtype=torch.float64).numpy())) < 1e-15
    try:
        import torchaudio
    except Exception:
        pytest.skip("torchaudio not installed")
    fb = torchaudio.functional.melscale_fbanks(1025, 0.0, 11025.0, 128, 22050, norm="slaney", mel_scale="slaney").numpy().T
    assert np.max(np.abs(fb - lr.mel_filter(22050, 2048, 128))) < 1e-6


def test_stft_against_torch():
    import torch
    rng = np.random.default_rng(3)
    y = rng.standard_normal(22050).astype(np.float32)
    D = lr.stft(y, 2048, 512)
    T = torch.stft(torch.from_numpy(y).double(), 2048, 512, window=torch.hann_window(2048, dtype=torch.float64),
                   center=True, pad_mode="constant", return_complex=True).numpy()
    assert D.shape == T.shape == (1025, 1 + 22050 // 512)
    assert np.max(np.abs(D - T)) < 1e-3 * np.max(np.abs(T))


def test_logmel_db_against_torchaudio():
    """The front half of onset_strength (melspectrogram(power=2) -> power_to_db(top_db=80)) against torchaudio's
    independent implementation of the same librosa semantics (MelSpectrogram with Slaney scale/norm, zero-padded
    centred frames, AmplitudeToDB('power', top_db=80)) on synthetic music, at both hops the path uses."""
    import torch
    try:
        import torchaudio
    except Exception:
        pytest.skip("torchaudio not installed")
    from oracle import synth
    y = synth.synth(17, 4.0, 22050, bpm=124.0)
    for hop in (512, 64):
        mel = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=2048, hop_length=hop, f_min=0.0, f_max=11025.0,
                                                   n_mels=128, power=2.0, center=True, pad_mode="constant", norm="slaney",
                                                   mel_scale="slaney").double()
        S = mel(torch.from_numpy(y).double())
        want = torchaudio.transforms.AmplitudeToDB(stype="power", top_db=80.0)(S).numpy()
        got = lr.logmel_db(y, 22050, hop)
        assert got.shape == want.shape == (128, 1 + len(y) // hop)
        assert np.max(np.abs(got - want)) < 2e-3          # dB; float32 mel bank on one side, float64 on the other
        # and the flux built on it (the onset envelope itself) agrees to 1e-4 of its maximum
        env_t = lr.onset_from_logmel(want.astype(np.float32), 2048, hop)
        env_o = lr.onset_strength(y, 22050, hop)
        assert np.max(np.abs(env_t - env_o)) < 1e-4 * float(env_o.max())


def test_onset_shape_and_padding():
    from oracle import synth
    y = synth.synth(11, 10.0, 22050, bpm=120.0)
    env = lr.onset_strength(y, 22050, 512)
    assert env.shape == (431,) and env.dtype == np.float32
    assert np.all(env[:3] == 0) and env.max() > 0
    env64 = lr.onset_strength(y[:22050], 22050, 64)
    assert env64.shape == (1 + 22050 // 64,) and np.all(env64[:17] == 0)


def test_tempo_and_beats_on_click_track():
    """Metamorphic: a 120 BPM synthetic track is tracked at lag 22 (117.45 BPM grid point nearest
    the true tempo given the prior) or 21/22; beats are ~fpb apart."""
    from oracle import synth
    y = synth.synth(5, 10.0, 22050, bpm=120.0)
    env = lr.onset_strength(y, 22050, 512)
    lag = lr.tempo_lag(env, 22050, 512, 120.0)
    assert lag in (21, 22)
    bpm, beats = lr.beat_track(env, 22050, 512, 120.0)
    assert len(beats) >= 15
    d = np.diff(beats)
    assert abs(np.median(d) - 60.0 * 22050 / 512 / 120.0) <= 1.0


def test_tempogram_against_its_definition_and_the_sliding_sum_identity():
    """(1) The FFT-route tempogram of the restatement equals the direct definition
    R_t[k] = sum_j w[j] w[j+k] x[t+j] x[t+j+k] / max_k R_t[k] on a small case.  (2) The identity behind the CUDA
    kernel (tempo.cu): with w the periodic Hann window, w[j]·w[j+k] is a three-term trigonometric polynomial in j, so
    R_t[k] = 1/4 [(1 + c/2) S0 - (1 + c) Re S1 + s Im S1 + c/2 Re S2 - s/2 Im S2] with S_m = sum_{j<W-k} z_j e^{i m theta j},
    z_j = x[t+j] x[t+j+k], theta = 2 pi / W, c = cos(theta k), s = sin(theta k)."""
    rng = np.random.default_rng(9)
    W = 48
    env = rng.exponential(1.0, 200).astype(np.float32)
    tg = lr.tempogram(env, W)
    x = lr.tempogram_frames(env, W).astype(np.float64)
    w = lr.hann_periodic(W)
    theta = 2.0 * np.pi / W
    for t in (0, 7, 101, 199):
        direct = np.array([np.sum(w[: W - k] * w[k:] * x[t : t + W - k] * x[t + k : t + W]) for k in range(W)])
        want = direct / np.max(np.abs(direct))
        assert np.max(np.abs(tg[:, t] - want)) < 1e-12
        for k in (1, 5, 17, 40):
            j = np.arange(W - k)
            z = x[t + j] * x[t + j + k]
            S0, S1, S2 = z.sum(), np.sum(z * np.exp(1j * theta * j)), np.sum(z * np.exp(2j * theta * j))
            c, s_ = np.cos(theta * k), np.sin(theta * k)
            R = 0.25 * ((1 + 0.5 * c) * S0 - (1 + c) * S1.real + s_ * S1.imag + 0.5 * c * S2.real - 0.5 * s_ * S2.imag)
            assert abs(R - direct[k]) < 1e-12 * max(1.0, abs(direct[k]))
