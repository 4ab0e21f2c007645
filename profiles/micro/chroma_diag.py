"""Dump the mean chroma of the 48 stress chunks (tests/golden/fullsize_golden.json) from the GPU path to
gpurun_out/chroma_diag_<impl>.npy; run with NCFA_CQT_IMPL=simt for the CUDA-core contraction."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nightcore-to-flac-analyzer_b200"))
from oracle import synth
from nightcore_analyzer import _engine
g = json.load(open(os.path.join(ROOT, "tests/golden/fullsize_golden.json")))["stress"]["chunks"]
eng = _engine.get_engine()
chunks = [synth.stress(c["seed"], 20.0, 22050) for c in g]
audio, off, ln = eng.pack(chunks)
chroma, tun = eng.chroma_mean_dev(audio, off, ln, 22050)
impl = os.environ.get("NCFA_CQT_IMPL", "tc")
np.save(os.path.join(ROOT, "gpurun_out", f"chroma_diag_{impl}.npy"), chroma.cpu().numpy())
want = np.array([[float.fromhex(v) for v in c["chroma"]] for c in g])
err = np.abs(chroma.cpu().numpy() - want).max(axis=1) / np.abs(want).max(axis=1)
for c, e in zip(g, err):
    print(impl, c["seed"], c["seed"] % 8, f"{e:.2e}")
