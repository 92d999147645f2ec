"""Drop-in for the reference's `models/randmod.py` (random (2^n-state) phosphorylation model).

`solve_ode(params, init_cond, num_psites, t) -> (sol, flat)` keeps the reference signature
and slicing (models/randmod.py:249-305); the work is one launch of the sm_100a kernel through the C ABI.
`solve_ode_batch` is the batched form the GPU is built for.
"""
from ._common import solve_ode_batch as _batch, solve_ode_single as _single

MODEL = "randmod"


def solve_ode(params, init_cond, num_psites, t, **kw):
    return _single(MODEL, params, init_cond, num_psites, t, **kw)


def solve_ode_batch(params, init_cond, num_psites, t, want=("sol", "flat"), **kw):
    return _batch(MODEL, params, init_cond, num_psites, t, want=want, **kw)
