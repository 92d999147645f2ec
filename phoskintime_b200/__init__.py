"""phoskintime_b200 — B200-native ensemble ODE engine behind PhosKinTime's solve_ode surface.

Hot path only: batched `solve_ode` for the distributive / successive / random
phosphorylation models with fused loss and Morris epilogues, launched through the C ABI in
include/phoskin_b200.h.  No CPU fallback: importing works anywhere, computing needs a B200.
"""
from ._lib import PhoskinError, LIB_PATH  # noqa: F401
from .engine import Engine, get_engine, local_dims, DEFAULT_RTOL, DEFAULT_ATOL  # noqa: F401

__all__ = ["Engine", "get_engine", "local_dims", "PhoskinError", "LIB_PATH"]
