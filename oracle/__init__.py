"""
CPU oracle for the nightcore front-end hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``nightcore_analyzer``) never imports it and has no CPU fallback.

Parity status: **parity unpinned** for every librosa-backed function
(librosa is not installed and not installable here; the reference has no
tests or golden vectors).  ``librosa_restated`` restates librosa 0.10.2+/0.11
semantics from SURVEY.md Appendix A.  What *is* pinned:

* ``consensus.py`` / ``pitch.py`` of the reference load standalone and are run
  verbatim (golden vectors GV1-GV7 in ``tests/golden``; generator script
  ``tests/golden/make_golden.py``).
* PCG64 / bounded-integer known-answer tests against numpy itself.
* mel filter bank vs torchaudio, periodic Hann vs torch, STFT vs torch.stft.
"""
