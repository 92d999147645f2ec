"""Drop-in for the reference's `models/succmod.py` (successive phosphorylation model).

`solve_ode(params, init_cond, num_psites, t) -> (sol, flat)` keeps the reference signature
and slicing (models/succmod.py:114-152); the work is one launch of the sm_100a kernel through the C ABI.
`solve_ode_batch` is the batched form the GPU is built for.
"""
from ._common import solve_ode_batch as _batch, solve_ode_single as _single

MODEL = "succmod"


def solve_ode(params, init_cond, num_psites, t, **kw):
    return _single(MODEL, params, init_cond, num_psites, t, **kw)


def solve_ode_batch(params, init_cond, num_psites, t, want=("sol", "flat"), **kw):
    return _batch(MODEL, params, init_cond, num_psites, t, want=want, **kw)
