"""
The verification step of the reference's workflow as a library call.

``workflow.run_full_suite`` (workflow.py:778-833) checks a freshly created HQNC against the original nightcore by
running ``pipeline.run(nightcore=NCOG, source=HQNC, compute_pitch=False)``, attaching
``xcorr.estimate_speed_xcorr(HQNC, NCOG)`` and applying the tolerances of ``_print_verification_result``
(workflow.py:160-163,299-400): IBI ratio within 0.5 % of 1 when available, else windowed BPM ratio within 2 %; pitch
ratio within 2 %; cross-correlation discarded below quality 0.30; length ratio flagged beyond 0.5 %.  The interactive
prompts, sox/rubberband subprocesses and printing stay out of scope; this returns the numbers and verdicts.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import pipeline, xcorr
from .consensus import AnalysisResult
from .io import SAMPLE_RATE, load_audio

NEAR_UNITY = 0.02            # workflow.py:160
XCORR_QUALITY_GATE = 0.30    # workflow.py:162
LEN_RATIO_WARN = 0.005       # workflow.py:163
IBI_TOLERANCE = 0.005        # workflow.py:381


@dataclass
class Verification:
    result: AnalysisResult           # carries xcorr_ratio / xcorr_quality like the workflow's result2
    best_ratio: float                # IBI ratio when available, else the windowed BPM ratio (workflow.py:380)
    estimator: str                   # "IBI" | "BPM"
    tempo_ok: bool
    pitch_ok: bool
    xcorr_discarded: bool            # quality below the 0.30 gate
    xcorr_label: Optional[str]
    length_ratio: Optional[float]    # NCOG ÷ HQNC duration after silence trim
    length_warn: bool
    corrected_speed_factor: float    # multiply the sox speed by this for a re-run (workflow.py:808-811)


def verify_arrays(hqnc: np.ndarray, ncog: np.ndarray, sr: int = SAMPLE_RATE, log=None) -> Verification:
    """HQNC-vs-NCOG verification for arrays already loaded at *sr* (the xcorr runs at the same rate)."""
    res = pipeline.run_arrays(ncog, hqnc, sr, compute_pitch=False, log=log)
    ratio, quality = xcorr.estimate_speed_xcorr_arrays(hqnc, ncog, sr)
    res.xcorr_ratio, res.xcorr_quality = ratio, quality
    q = quality or 0.0
    discarded = q < XCORR_QUALITY_GATE
    if res.ibi_ratio is not None:
        best, estimator = res.ibi_ratio, "IBI"
        tempo_ok = abs(best - 1.0) < IBI_TOLERANCE
    else:
        best, estimator = res.tempo_ratio, "BPM"
        tempo_ok = abs(res.tempo_ratio - 1.0) < NEAR_UNITY
    pitch_ok = abs(res.pitch_ratio - 1.0) < NEAR_UNITY
    length_ratio, length_warn = None, False
    if res.nc_duration and res.src_duration:
        length_ratio = res.nc_duration / res.src_duration
        length_warn = abs(length_ratio - 1.0) > LEN_RATIO_WARN
    return Verification(result=res, best_ratio=best, estimator=estimator, tempo_ok=tempo_ok, pitch_ok=pitch_ok,
                        xcorr_discarded=discarded, xcorr_label=None if discarded else xcorr.quality_label(q),
                        length_ratio=length_ratio, length_warn=length_warn, corrected_speed_factor=best)


def verify(hqnc_path: str, ncog_path: str, log=None) -> Verification:
    """File form: both files are loaded at 22 050 Hz, which is also what ``estimate_speed_xcorr`` does (xcorr.py:57,92)."""
    hqnc, sr = load_audio(str(hqnc_path))
    ncog, _ = load_audio(str(ncog_path), sr=sr)
    return verify_arrays(hqnc, ncog, sr, log=log)
