"""nightcore_analyzer — B200-native drop-in for the windowed spectral front-end of
Tealdragon204/nightcore-to-flac-analyzer (reference __init__.py:20-26).

``run`` / ``AnalysisResult`` are the reference's package exports; ``run_arrays`` and ``run_batch`` are the
array-level and data-parallel entry points added beside them.  Importing the package loads libncfa.so and
fails loudly when it is missing — there is no CPU fallback."""
from . import _native  # noqa: F401  (fails loudly when libncfa.so is missing)
from . import export  # noqa: F401  (reference __init__.py:22 re-exports the module)
from .consensus import AnalysisResult
from .pipeline import run, run_arrays, run_batch

__version__ = "0.3.0"
__all__ = ["run", "run_arrays", "run_batch", "AnalysisResult", "export", "__version__"]
