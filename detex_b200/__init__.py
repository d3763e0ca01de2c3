"""detex_b200 -- B200-native (sm_100a) implementation of Detex's data-parallel hot path:
subspace detection statistic, pairwise CCX matrix and FAS null-space statistics, behind
the reference's own callables.  See DESIGN.md / INTEGRATION.md.

Importing the package does not touch the GPU; the CUDA library is loaded (and built with
nvcc if missing) on first use.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from .engine import DtxError, Engine, ShortChunk  # noqa: F401
