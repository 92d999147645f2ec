"""Shared implementation of the drop-in `solve_ode` modules (models/distmod.py etc.)."""
import numpy as np

from ..engine import get_engine

# NORMALIZE_MODEL_OUTPUT of the reference (config/constants.py:73, config.toml:208 default false)
NORMALIZE_MODEL_OUTPUT = False


def solve_ode_single(model, params, init_cond, num_psites, t, **kw):
    """Reference signature: returns (sol[T,n], flat[L]) as fresh numpy arrays.
    Accepts list/tuple/ndarray params like the reference's callers pass
    (sensitivity/analysis.py:188-193 builds a tuple)."""
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    res = get_engine().solve_local_batch(model, np.asarray(params, dtype=np.float64).reshape(1, -1),
                                         np.asarray(init_cond, dtype=np.float64), num_psites, t,
                                         want=("sol", "flat"), normalize=NORMALIZE_MODEL_OUTPUT, **kw)
    return res["sol"][0], res["flat"][0]


def solve_ode_batch(model, params, init_cond, num_psites, t, want=("sol", "flat"), **kw):
    kw.setdefault("normalize", NORMALIZE_MODEL_OUTPUT)
    return get_engine().solve_local_batch(model, params, init_cond, num_psites, t, want=want, **kw)
