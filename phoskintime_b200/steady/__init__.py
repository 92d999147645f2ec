"""Steady-state initial conditions — drop-in for the reference's `steady` package.

The reference finds the all-ones-parameter steady state with SLSQP on a constant objective
(steady/initdist.py:9-50, initsucc.py:9-55, initrand.py:10-77), i.e. it solves a linear
system iteratively.  It runs once per (model, num_psites) on the host and is not per-sample
work, so it stays on the host here too — solved directly:
  * distmod / succmod: closed form R=1, P=1/(1+n/2), P_i=P/2.  (initsucc.py:38-41 uses the
    *distributive* equations, so both share it.)
  * randmod: linear solve in initrand's own state order (subsets by size, then lexicographic —
    NOT the bitmask order models/randmod.py uses; reproduced as is).
"""
from itertools import combinations

import numpy as np

from ..models import ODE_MODEL as _DEFAULT_MODEL


def initial_condition(num_psites: int, model: str = None) -> list:
    model = model or _current_model()
    if num_psites < 1:
        raise ValueError("num_psites must be >= 1")
    if model in ("distmod", "succmod"):
        P = 1.0 / (1.0 + num_psites / 2.0)
        return [1.0, P] + [P / 2.0] * num_psites
    if model != "randmod":
        raise ValueError(f"Unsupported ODE_MODEL: {model}")
    subsets = [c for k in range(1, num_psites + 1) for c in combinations(range(num_psites), k)]
    pos = {s: 2 + i for i, s in enumerate(subsets)}
    n = 2 + len(subsets)
    M = np.zeros((n, n))
    b = np.zeros(n)
    M[0, 0], b[0] = -1.0, -1.0                      # A - B R = 0
    M[1, 0], M[1, 1] = 1.0, -(1.0 + num_psites)     # C R - D P - sum(S) P + singles
    for sub, row in pos.items():
        k = len(sub)
        if k == 1:
            M[1, row] += 1.0
            M[row, 1] += 1.0
        else:
            for site in sub:
                M[row, pos[tuple(x for x in sub if x != site)]] += 1.0
        M[row, row] -= (num_psites - k) + k + 1.0
        for site in range(num_psites):
            if site not in sub:
                M[row, pos[tuple(sorted(sub + (site,)))]] += 1.0
    return np.linalg.solve(M, b).tolist()


def _current_model():
    from .. import models
    return models.ODE_MODEL or _DEFAULT_MODEL
