"""ntg_b200 -- B200-native batched evaluator for NTG's collocation hot path.

The product is the C ABI in include/ntg_b200.h (libntg_b200.so + callback
packs, hand-written CUDA for sm_100a).  This Python package is the thin host
mirror used by tests and bench: problem descriptions, ctypes bindings, build
recipe.  It contains no evaluation code of its own and no CPU fallback.
"""
from .abi import JAC_BAND, JAC_DENSE, JAC_NONE, ProblemSpec, linspace  # noqa: F401
from . import configs  # noqa: F401

__all__ = ["ProblemSpec", "linspace", "configs", "JAC_BAND", "JAC_DENSE", "JAC_NONE", "Problem"]


def __getattr__(name):
    if name in ("Problem", "NtgError", "core", "load_pack"):
        from . import problem
        return getattr(problem, name)
    raise AttributeError(name)
