"""bayesfmmm_b200 -- B200 (sm_100a) engine for BayesFMMM's per-iteration Gibbs/Metropolis sampler.

The product is the C-ABI shared library ``libbfmmm_b200.so`` (include/bfmmm.h) built from
``bayesfmmm_b200/csrc``; this package is the thin host-side mirror used by the tests and bench.py.
There is no CPU fallback: importing works anywhere, creating an Engine needs the CUDA library.
"""
from ._lib import load_library, library_path, build_library  # noqa: F401
from .engine import Engine, EngineError  # noqa: F401
from .sampler import Sampler, default_hyper, SWEEP_THETA, SWEEP_NU_Z, SWEEP_FULL  # noqa: F401

__all__ = ["Engine", "EngineError", "Sampler", "default_hyper", "SWEEP_THETA", "SWEEP_NU_Z", "SWEEP_FULL", "load_library", "library_path", "build_library"]
