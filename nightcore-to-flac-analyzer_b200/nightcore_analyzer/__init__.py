"""nightcore_analyzer — B200-native drop-in for the windowed spectral front-end of
Tealdragon204/nightcore-to-flac-analyzer (reference __init__.py:20-26)."""
from . import _native  # noqa: F401  (fails loudly when libncfa.so is missing)

__version__ = "0.3.0"
