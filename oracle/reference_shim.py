"""
Run the reference's own modules UNMODIFIED on top of the restated librosa (TEST / BASELINE INFRASTRUCTURE).

``install_librosa_shim()`` puts a minimal ``librosa`` module (our numpy/scipy restatement,
oracle/librosa_restated.py) into ``sys.modules``; ``load_reference()`` then imports
``/root/reference/nightcore_analyzer`` under the alias ``nc_reference`` so that its io.py /
tempo.py / pitch.py / xcorr.py / consensus.py / pipeline.py control flow executes verbatim.
``/root/reference`` exists only in the build container — callers must check
``reference_available()`` and skip otherwise (the GPU box has only the committed golden vectors).

``load_reference_standalone(name)`` loads consensus.py / pitch.py alone (they need only numpy),
which is how the golden vectors GV1-GV7 were produced.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
import wave

import numpy as np

from . import librosa_restated as lr

REFERENCE_ROOT = os.environ.get("NCFA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nightcore_analyzer"))


def _load_file(path, sr=22050, mono=True, dtype=np.float32):
    """librosa.load stand-in: .npy (already at the target rate) or PCM .wav; resamples with our FIR."""
    path = str(path)
    if path.endswith(".npy"):
        y = np.load(path).astype(dtype)
        return y, sr
    with wave.open(path, "rb") as w:
        n, ch, sw, fr = w.getnframes(), w.getnchannels(), w.getsampwidth(), w.getframerate()
        raw = w.readframes(n)
    if sw == 2:
        y = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        y = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    else:
        raise ValueError("unsupported WAV sample width")
    if ch > 1:
        y = y.reshape(-1, ch).mean(axis=1)
    if sr is not None and fr != sr:
        y = lr.resample(y, fr, sr)
    return y.astype(dtype), (sr or fr)


def install_librosa_shim() -> types.ModuleType:
    if "librosa" in sys.modules and getattr(sys.modules["librosa"], "__ncfa_shim__", False):
        return sys.modules["librosa"]
    m = types.ModuleType("librosa")
    m.__ncfa_shim__ = True
    m.__version__ = "0.11.0-restated"
    m.load = _load_file
    m.resample = lambda y, *, orig_sr, target_sr, **kw: lr.resample(y, orig_sr, target_sr, scale=kw.get("scale", False))
    m.frames_to_time = lambda frames, *, sr=22050, hop_length=512, **kw: lr.frames_to_time(frames, sr, hop_length)
    m.stft = lambda y, **kw: lr.stft(y, **kw)

    effects = types.ModuleType("librosa.effects")
    effects.trim = lambda y, *, top_db=60, **kw: lr.trim(y, top_db=top_db)
    onset = types.ModuleType("librosa.onset")
    onset.onset_strength = lambda *, y=None, sr=22050, hop_length=512, **kw: lr.onset_strength(y, sr, hop_length)
    beat = types.ModuleType("librosa.beat")
    beat.beat_track = lambda *, onset_envelope=None, sr=22050, hop_length=512, start_bpm=120.0, **kw: lr.beat_track(
        onset_envelope, sr, hop_length, start_bpm)
    feature = types.ModuleType("librosa.feature")
    feature.tempo = lambda *, onset_envelope=None, sr=22050, hop_length=512, start_bpm=120.0, **kw: lr.tempo(
        onset_envelope, sr, hop_length, start_bpm)
    feature.tempogram = lambda *, onset_envelope=None, sr=22050, hop_length=512, win_length=384, **kw: lr.tempogram(
        onset_envelope, win_length)
    feature.rms = lambda *, y=None, frame_length=2048, hop_length=512, **kw: lr.rms(y, frame_length, hop_length)[None, :]
    feature.chroma_cqt = lambda *, y=None, sr=22050, bins_per_octave=36, hop_length=512, **kw: lr.chroma_cqt(
        y, sr, hop_length, bins_per_octave)
    m.effects, m.onset, m.beat, m.feature = effects, onset, beat, feature
    sys.modules["librosa"] = m
    for sub in (effects, onset, beat, feature):
        sys.modules[sub.__name__] = sub
    return m


def load_reference_standalone(name: str):
    """Import /root/reference/nightcore_analyzer/<name>.py by file path (consensus, pitch)."""
    path = os.path.join(REFERENCE_ROOT, "nightcore_analyzer", name + ".py")
    spec = importlib.util.spec_from_file_location("nc_reference_standalone_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Import the reference package (io, tempo, pitch, xcorr, consensus, pipeline) over the librosa shim,
    as ``nc_reference`` — without executing its __init__ (which would pull in export/session only)."""
    if "nc_reference" in sys.modules:
        return sys.modules["nc_reference"]
    install_librosa_shim()
    pkg_dir = os.path.join(REFERENCE_ROOT, "nightcore_analyzer")
    pkg = types.ModuleType("nc_reference")
    pkg.__path__ = [pkg_dir]
    sys.modules["nc_reference"] = pkg
    for name in ("io", "consensus", "tempo", "pitch", "xcorr", "pipeline"):
        setattr(pkg, name, importlib.import_module("nc_reference." + name))
    return pkg
