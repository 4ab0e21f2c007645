import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_PARENT = os.path.join(ROOT, "nightcore-to-flac-analyzer_b200")
for p in (ROOT, PKG_PARENT):
    if p not in sys.path:
        sys.path.insert(0, p)
# an operator-provided install of the reference's dependency (librosa) under baseline/_ref switches on the
# real-librosa tier (tests/test_real_librosa_tier.py); it goes LAST so that nothing else is shadowed
_REF_SITE = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF_SITE) and _REF_SITE not in sys.path:
    sys.path.append(_REF_SITE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from nightcore_analyzer import _engine
    return _engine.get_engine()


def fromhex(v):
    return float.fromhex(v)
