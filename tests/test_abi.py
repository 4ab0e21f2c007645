"""CPU: the C-ABI library loads and exports every symbol include/ncfa.h declares (no compute)."""
import os


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge._build_native()
    from nightcore_analyzer import _native
    names = _native.check_symbols()
    assert "ncfa_onset_strength_batched" in names and "ncfa_bootstrap_ratio_batched" in names
    assert _native.lib.ncfa_version() >= 100
    # every declared symbol has a ctypes signature
    assert set(names) <= set(_native._SIGNATURES), set(names) - set(_native._SIGNATURES)


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nightcore_analyzer import _engine, _native
    with pytest.raises(_native.NcfaError):
        _engine.get_engine()


def test_product_never_imports_oracle():
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nightcore-to-flac-analyzer_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
